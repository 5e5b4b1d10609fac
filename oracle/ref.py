"""ctypes wrapper around ``oracle/_ref/libtoyref.so`` -- TEST INFRASTRUCTURE ONLY.

``libtoyref.so`` is the reference's own hot path (peano.c sort.c tree.c sph.c wvt_relax.c
aux.c) compiled unmodified by ``oracle/Makefile`` plus ``oracle/ref_harness.c``.  It is
used to pin the CPU restatement (``oracle/port.py``), to generate the fixtures under
``tests/golden/`` and as the ``cpu_baseline`` / ``--impl reference`` arm of ``bench.py``.
Nothing under ``toycluster_b200/`` may import this module.

The reference keeps per-process statics sized by the first problem it sees
(peano.c:53-61, tree.c:343-346), so every :class:`Ref` loads a private copy of the library.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libtoyref.so")
# Same harness API, but the three operators are toycluster_b200/host/gpu_shim.c + libtoygpu.so
SHIM = os.path.join(HERE, "_ref", "libtoyshim.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_ITER_CB = C.CFUNCTYPE(C.c_int, C.c_int)


def available() -> bool:
    return os.path.exists(LIB)


def build() -> bool:
    """(Re)build from /root/reference when it is present; keep the prebuilt file otherwise."""
    import subprocess
    ref = os.environ.get("TOY_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "src")):
        return available()
    subprocess.run(["make", "-C", HERE, "ref", f"REF={ref}"], check=True,
                   stdout=subprocess.DEVNULL)
    return available()


class Ref:
    NGBMAX = 2360

    def __init__(self, n_gas, boxsize, mpart, mtotal, halo_table, nthreads=0, shim=False):
        src = SHIM if shim else LIB
        if not os.path.exists(src):
            raise RuntimeError(f"{src} missing: run `make -C oracle ref shim`")
        if shim:   # resolve the shim's DT_NEEDED libtoygpu.so by soname
            gpu = os.path.join(os.path.dirname(HERE), "toycluster_b200", "libtoygpu.so")
            self._gpu = C.CDLL(gpu, mode=C.RTLD_GLOBAL)
        self._tmp = tempfile.NamedTemporaryFile(suffix=".so", delete=False)
        self._tmp.close()
        shutil.copyfile(src, self._tmp.name)
        lib = self.lib = C.CDLL(self._tmp.name)
        os.unlink(self._tmp.name)
        self.n = int(n_gas)
        halo_table = np.ascontiguousarray(halo_table, dtype=np.float64).reshape(-1, 9)
        lib.ref_setup.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                  np.ctypeslib.ndpointer(np.float64), C.c_int]
        if not shim:
            lib.ref_guess_hsml.restype = C.c_float
            lib.ref_guess_hsml.argtypes = [C.c_int]
            lib.ref_find_ngb_tree.argtypes = [C.c_int, C.c_float, _i32p]
            lib.ref_find_ngb_simple.argtypes = [C.c_int, C.c_float, _i32p]
            lib.ref_peano_key.argtypes = [C.c_double, C.c_double, C.c_double,
                                          C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong),
                                          C.c_int]
        lib.ref_global_density_model.restype = C.c_float
        lib.ref_log.restype = C.c_char_p
        lib.ref_time_density.restype = C.c_double
        lib.ref_wvt_scratch.restype = C.POINTER(C.c_float)
        lib.ref_regularise.argtypes = [C.c_int, C.c_void_p, C.c_int]
        rc = lib.ref_setup(self.n, boxsize, mpart, mtotal, len(halo_table), halo_table,
                           int(nthreads))
        if rc != 0:
            raise RuntimeError(f"ref_setup failed: {rc}")
        self.nthreads = lib.ref_nthreads()

    # -- data in / out -----------------------------------------------------------------
    def load(self, pos, hsml=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        assert pos.shape == (self.n, 3)
        hp = None
        if hsml is not None:
            hsml = np.ascontiguousarray(hsml, dtype=np.float32)
            hp = hsml.ctypes.data_as(C.c_void_p)
        self.lib.ref_load(pos.ctypes.data_as(C.c_void_p), hp)

    def set_apot(self, apot):
        apot = np.ascontiguousarray(apot, dtype=np.float32)
        self.lib.ref_set_apot(apot.ctypes.data_as(C.c_void_p))

    def read(self):
        """All per-particle state in the current (Peano-sorted) order."""
        n = self.n
        out = dict(pos=np.empty((n, 3), np.float32), id=np.empty(n, np.int32),
                   hsml=np.empty(n, np.float32), rho=np.empty(n, np.float32),
                   varhsml=np.empty(n, np.float32), rho_model=np.empty(n, np.float32),
                   bfld=np.empty((n, 3), np.float32), key_hi=np.empty(n, np.uint64),
                   key_lo=np.empty(n, np.uint64), tree_parent=np.empty(n, np.int32))
        order = ["pos", "id", "hsml", "rho", "varhsml", "rho_model", "bfld", "key_hi",
                 "key_lo", "tree_parent"]
        self.lib.ref_read(*[out[k].ctypes.data_as(C.c_void_p) for k in order])
        return out

    # -- pieces of the path -------------------------------------------------------------
    def peano_key(self, x, y, z, reversed_=False):
        hi, lo = C.c_ulonglong(), C.c_ulonglong()
        self.lib.ref_peano_key(x, y, z, C.byref(hi), C.byref(lo), int(reversed_))
        return hi.value, lo.value

    def sort(self):
        self.lib.ref_sort()

    def build_tree(self):
        self.lib.ref_build_tree()

    def find_ngb_tree(self, i, h):
        buf = np.zeros(self.NGBMAX, np.int32)
        cnt = self.lib.ref_find_ngb_tree(int(i), float(h), buf)
        return buf[:cnt].copy()

    def find_ngb_simple(self, i, h):
        buf = np.zeros(self.NGBMAX, np.int32)
        cnt = self.lib.ref_find_ngb_simple(int(i), float(h), buf)
        return buf[:cnt].copy()

    def guess_hsml(self, i):
        return np.float32(self.lib.ref_guess_hsml(int(i)))

    def find_sph_quantities(self):
        self.lib.ref_find_sph_quantities()

    def bfld_from_rotA(self):
        self.lib.ref_bfld_from_rotA()

    def regularise(self, max_iters=1 << 30, callback=None, echo=False):
        """Run wvt_relax.c:25; ``callback(it)`` fires before the density pass of iteration
        ``it`` (return non-zero to stop).  Returns the number of density passes started."""
        cb = _ITER_CB(callback) if callback is not None else None
        ptr = C.cast(cb, C.c_void_p) if cb is not None else None
        return self.lib.ref_regularise(int(max_iters), ptr, int(echo))

    def wvt_scratch(self):
        """(hsml_wvt, delta[n,3]) of the last completed iteration, previous-sort order."""
        bufs = []
        for k in range(4):
            p = self.lib.ref_wvt_scratch(k)
            if not p:
                return None
            bufs.append(np.ctypeslib.as_array(p, shape=(self.n,)).copy())
        return bufs[0], np.stack(bufs[1:], axis=1)

    def log(self) -> str:
        return self.lib.ref_log().decode()

    def time_density(self) -> float:
        return self.lib.ref_time_density()


def parse_log(text: str):
    """The '#NN: Err max=… mean=… diff=… step=…' lines of wvt_relax.c:91-92."""
    rows = []
    for line in text.splitlines():
        line = line.strip()
        if not line.startswith("#"):
            continue
        it, rest = line[1:].split(":", 1)
        kv = dict(tok.split("=") for tok in rest.replace("Err", "").split())
        rows.append(dict(it=int(it), max=float(kv["max"]), mean=float(kv["mean"]),
                         diff=float(kv["diff"]), step=float(kv["step"])))
    return rows
