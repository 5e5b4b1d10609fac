"""oracle/compat/gsl_compat.c stands in for GSL when the reference's WHOLE driver is built
(oracle/Makefile `driver`).  It is test infrastructure; this checks it against scipy so that the
set-up stages it serves (mass tables, temperature and potential integrals, Eddington inversion)
compute what they are meant to."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

scipy_integrate = pytest.importorskip("scipy.integrate")
scipy_interpolate = pytest.importorskip("scipy.interpolate")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "oracle", "compat")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("gsl") / "libgslcompat.so"
    subprocess.run(["gcc", "-std=c99", "-O2", "-fPIC", "-shared", "-I", COMPAT,
                    os.path.join(COMPAT, "gsl_compat.c"), "-o", str(so), "-lm"], check=True)
    L = C.CDLL(str(so))
    L.gsl_integration_workspace_alloc.restype = C.c_void_p
    L.gsl_spline_alloc.restype = C.c_void_p
    L.gsl_spline_alloc.argtypes = [C.c_void_p, C.c_size_t]
    L.gsl_spline_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    for f in (L.gsl_spline_eval, L.gsl_spline_eval_deriv2):
        f.restype = C.c_double
        f.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    return L


FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


class GF(C.Structure):
    _fields_ = [("function", FN), ("params", C.c_void_p)]


def _integrate(L, f, a, b, rel, key=None):
    w = C.c_void_p(L.gsl_integration_workspace_alloc(4096))
    g = GF(FN(lambda x, p: f(x)), None)
    r, e = C.c_double(), C.c_double()
    if key is None:
        L.gsl_integration_qags.argtypes = [C.POINTER(GF), C.c_double, C.c_double, C.c_double, C.c_double,
                                           C.c_size_t, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.gsl_integration_qags(C.byref(g), a, b, 0.0, rel, 4096, w, C.byref(r), C.byref(e))
    else:
        L.gsl_integration_qag.argtypes = [C.POINTER(GF), C.c_double, C.c_double, C.c_double, C.c_double,
                                          C.c_size_t, C.c_int, C.c_void_p, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double)]
        L.gsl_integration_qag(C.byref(g), a, b, 0.0, rel, 4096, key, w, C.byref(r), C.byref(e))
    return r.value


def test_gauss_kronrod_tables_are_rules():
    """Weights sum to 2 and the rules integrate polynomials of their degree exactly."""
    import re
    text = open(os.path.join(COMPAT, "gk_tables.h")).read()
    for n in (21, 41, 61):
        x = np.array([float(v) for v in re.search(rf"xgk{n}\[\d+\] = \{{(.*?)\}};", text, re.S).group(1).split(",")])
        w = np.array([float(v) for v in re.search(rf"wgk{n}\[\d+\] = \{{(.*?)\}};", text, re.S).group(1).split(",")])
        assert abs(2 * w[:-1].sum() + w[-1] - 2) < 1e-14
        for k in (2, 8, (3 * (n // 2) + 1) // 2 * 2):       # even powers up to the rule's degree
            got = 2 * (w[:-1] * x[:-1] ** k).sum() + (w[-1] if k == 0 else 0.0)
            assert abs(got - 2 / (k + 1)) < 1e-13, (n, k)


@pytest.mark.parametrize("key", [2, 4, 6])
def test_qag_mass_integrand(lib, key):
    beta = 0.54
    f = lambda r: 4 * math.pi * r * r * (1 + (r / 250.) ** 2) ** (-1.5 * beta) / (1 + (r / 2500.) ** 4)   # setup.c:640
    want = scipy_integrate.quad(f, 0, 3000., epsrel=1e-12)[0]
    assert abs(_integrate(lib, f, 0, 3000., 1e-6, key) - want) <= 1e-6 * want


def test_qags_endpoint_singularity(lib):
    f = lambda x: math.exp(x) / math.sqrt(1 - x) if x < 1 else 0.0      # like velocities.c:318
    want = scipy_integrate.quad(f, 0, 1, epsrel=1e-12)[0]
    assert abs(_integrate(lib, f, 0, 1, 1e-4) - want) <= 1e-4 * want


def test_natural_cubic_spline(lib):
    x = np.sort(np.random.default_rng(0).uniform(0, 10, 40))
    y = np.sin(x)
    s = C.c_void_p(lib.gsl_spline_alloc(None, len(x)))
    lib.gsl_spline_init(s, x.ctypes.data, y.ctypes.data, len(x))
    cs = scipy_interpolate.CubicSpline(x, y, bc_type="natural")
    for t in np.linspace(x[0], x[-1] - 1e-9, 300):
        assert abs(lib.gsl_spline_eval(s, float(t), None) - cs(t)) < 1e-12
        assert abs(lib.gsl_spline_eval_deriv2(s, float(t), None) - cs(t, 2)) < 1e-10
