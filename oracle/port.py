"""ctypes wrapper around ``oracle/libtoyoracle.so`` (``oracle/toy_oracle.c``), the plain-C
restatement of the reference's SPH/WVT hot path -- TEST INFRASTRUCTURE ONLY.

The control flow that the reference keeps in ``Find_sph_quantities`` (sph.c:13-75) and
``Regularise_sph_particles`` (wvt_relax.c:25-225) is restated here in Python on top of the C
pieces, so a test can stop after any iteration and look at every intermediate array.
Nothing under ``toycluster_b200/`` may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libtoyoracle.so")
NGBMAX = 2360

_lib = None


class _Sys(C.Structure):
    _fields_ = [("n", C.c_int), ("box", C.c_double), ("mpart", C.c_double),
                ("mtotal", C.c_double), ("nhalos", C.c_int), ("halos", C.POINTER(C.c_double))]


def build():
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(
                os.path.join(HERE, "toy_oracle.c")):
            build()
        _lib = C.CDLL(LIB)
        _lib.to_find_ngb.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_float,
                                     C.c_void_p]
        _lib.to_find_ngb_simple.argtypes = _lib.to_find_ngb.argtypes
        _lib.to_tree_displaced.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_void_p,
                                           C.c_void_p, C.c_int]
        _lib.to_sort.argtypes = [C.c_int, C.c_void_p, C.c_double] + [C.c_void_p] * 3
        _lib.to_peano_key.argtypes = [C.c_double] * 3 + [C.c_int, C.POINTER(C.c_uint64),
                                                         C.POINTER(C.c_uint64)]
        _lib.to_guess_hsml.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_void_p]
        _lib.to_density.argtypes = [C.POINTER(_Sys)] + [C.c_void_p] * 5
        _lib.to_density_model.argtypes = [C.POINTER(_Sys), C.c_void_p, C.c_void_p]
        _lib.to_wvt_displace.argtypes = [C.POINTER(_Sys), C.c_void_p, C.c_void_p, C.c_double,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)] + \
            [C.c_void_p] * 4
        _lib.to_bfld_from_rotA.argtypes = [C.POINTER(_Sys)] + [C.c_void_p] * 6
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class System:
    """Param.{Npart[0], Boxsize, Mpart[0], Mtotal} + the halo table."""

    def __init__(self, w, n=None):
        self.n = int(n if n is not None else w.n_gas)
        self.box, self.mpart, self.mtotal = float(w.boxsize), float(w.mpart_gas), float(w.mtotal)
        self._halos = np.ascontiguousarray(w.halo_table(), dtype=np.float64)
        self.c = _Sys(self.n, self.box, self.mpart, self.mtotal, len(self._halos),
                      self._halos.ctypes.data_as(C.POINTER(C.c_double)))


def peano_key(x, y, z, reversed_=False):
    hi, lo = C.c_uint64(), C.c_uint64()
    lib().to_peano_key(x, y, z, int(reversed_), C.byref(hi), C.byref(lo))
    return hi.value, lo.value


def sort(pos, box):
    """peano.c:46-81 -> (perm, key_hi, key_lo, n_duplicate_keys), keys in sorted order."""
    pos = np.ascontiguousarray(pos, np.float32)
    n = len(pos)
    perm = np.empty(n, np.int32)
    hi, lo = np.empty(n, np.uint64), np.empty(n, np.uint64)
    dup = lib().to_sort(n, _p(pos), float(box), _p(perm), _p(hi), _p(lo))
    return perm, hi, lo, dup


def find_ngb(pos, box, i, h):
    pos = np.ascontiguousarray(pos, np.float32)
    out = np.empty(NGBMAX, np.int32)
    cnt = lib().to_find_ngb(len(pos), _p(pos), float(box), int(i), float(h), _p(out))
    return out[:cnt].copy()


def find_ngb_simple(pos, box, i, h):
    """wvt_relax.c:296-340 (Find_ngb_simple): the predicate without the tree."""
    pos = np.ascontiguousarray(pos, np.float32)
    out = np.empty(NGBMAX, np.int32)
    cnt = lib().to_find_ngb_simple(len(pos), _p(pos), float(box), int(i), float(h), _p(out))
    return out[:cnt].copy()


def tree_displaced(pos_sorted, box, cap=4096):
    """Leaves of the reference octree whose particles lie outside the node's cube (centre
    displaced by the sign test of tree.c:298-302) -> (first particle, count) per leaf."""
    pos_sorted = np.ascontiguousarray(pos_sorted, np.float32)
    first, count = np.empty(cap, np.int32), np.empty(cap, np.int32)
    m = lib().to_tree_displaced(len(pos_sorted), _p(pos_sorted), float(box), _p(first),
                                _p(count), cap)
    m = min(m, cap)
    return first[:m].copy(), count[:m].copy()


def guess_hsml(pos_sorted, box):
    pos_sorted = np.ascontiguousarray(pos_sorted, np.float32)
    out = np.empty(len(pos_sorted), np.float32)
    lib().to_guess_hsml(len(pos_sorted), _p(pos_sorted), float(box), _p(out))
    return out


def density_model(w, pos):
    pos = np.ascontiguousarray(pos, np.float32)
    s = System(w, len(pos))
    out = np.empty(len(pos), np.float32)
    lib().to_density_model(C.byref(s.c), _p(pos), _p(out))
    return out


def find_sph_quantities(w, pos, hsml=None, ids=None):
    """sph.c:13-75: sort, reorder, density.  Returns the state in the new Peano order."""
    pos = np.ascontiguousarray(pos, np.float32)
    n = len(pos)
    s = System(w, n)
    perm, hi, lo, dup = sort(pos, s.box)
    pos_s = np.ascontiguousarray(pos[perm])
    h = np.zeros(n, np.float32) if hsml is None else np.ascontiguousarray(hsml, np.float32)[perm]
    h = np.ascontiguousarray(h)
    ids = perm.copy() if ids is None else np.ascontiguousarray(ids)[perm]
    rho, var = np.empty(n, np.float32), np.empty(n, np.float32)
    stats = np.zeros(3, np.int64)
    bad = lib().to_density(C.byref(s.c), _p(pos_s), _p(h), _p(rho), _p(var), _p(stats))
    if bad:
        raise RuntimeError(f"{bad} particles did not converge")
    return dict(pos=pos_s, id=ids.astype(np.int32), hsml=h, rho=rho, varhsml=var, key_hi=hi,
                key_lo=lo, duplicates=dup, pair_evals=int(stats[0]), searches=int(stats[1]),
                hsml_iters=int(stats[2]))


def wvt_iteration(w, pos, hsml, step, ids=None, dens=None):
    """One pass of wvt_relax.c:66-214 with the step its displacement uses.  ``dens`` may carry
    the density pass of this iteration when the caller has already run it."""
    st = dict(dens) if dens is not None else find_sph_quantities(w, pos, hsml, ids)
    n = len(st["pos"])
    s = System(w, n)
    emax, emean = C.c_double(), C.c_double()
    rm, hw = np.empty(n, np.float32), np.empty(n, np.float32)
    delta = np.empty((n, 3), np.float32)
    new_pos = st["pos"].copy()
    stats = np.zeros(2, np.int64)
    lib().to_wvt_displace(C.byref(s.c), _p(new_pos), _p(st["rho"]), float(step), C.byref(emax),
                          C.byref(emean), _p(rm), _p(hw), _p(delta), _p(stats))
    st.update(pos_before=st["pos"], pos=new_pos, rho_model=rm, hw=hw, delta=delta,
              err_max=emax.value, err_mean=emean.value, wvt_pairs=int(stats[0]))
    return st


def regularise(w, pos, max_iters=1 << 30, keep=False):
    """wvt_relax.c:25-225.  Returns (log rows, final state, per-iteration states if keep)."""
    step = 0.0085
    if w.mtotal < 1e5:
        step /= 2
    err_last = err_diff_last = sys.float_info.max
    it, rows, states = -1, [], []
    state = dict(pos=np.ascontiguousarray(pos, np.float32), hsml=None, id=None)
    while True:
        it += 1
        if it - 1 >= 64 or it >= max_iters:          # `if (it++ >= NUMITER) break;`
            break
        dens = find_sph_quantities(w, state["pos"], state["hsml"], state["id"])
        rm = density_model(w, dens["pos"])
        err = (np.abs((dens["rho"] - rm).astype(np.float64)) / rm.astype(np.float64)).astype(np.float32)
        err_max, err_mean = float(err.max()), float(err.astype(np.float64).sum() / len(err))
        err_diff = (err_last - err_mean) / err_mean
        rows.append(dict(it=it, max=err_max, mean=err_mean, diff=err_diff, step=step))
        if (err_diff < 0.01 and it > 25) or (err_diff < 0 and err_diff_last < 0 and it > 10):
            state = dens
            break
        if err_diff < 0.01 and it > 1:
            step *= 0.8
        err_last, err_diff_last = err_mean, err_diff
        state = wvt_iteration(w, None, None, step, dens=dens)
        if keep:
            states.append(state)
    return rows, state, states


def bfld_from_rotA(w, st, apot):
    n = len(st["pos"])
    s = System(w, n)
    apot = np.ascontiguousarray(apot, np.float32)
    out = np.empty((n, 3), np.float32)
    lib().to_bfld_from_rotA(C.byref(s.c), _p(st["pos"]), _p(st["hsml"]), _p(st["rho"]),
                            _p(st["varhsml"]), _p(apot), _p(out))
    return out
