"""Host-side mirror of the reference's operator interface for the SPH/WVT hot path.

The reference exposes the path as three global-state procedures:

    Regularise_sph_particles()   wvt_relax.c:25
    Find_sph_quantities()        sph.c:13
    Bfld_from_rotA_SPH()         sph.c:216

:class:`HotPath` binds ``libtoygpu.so`` (C ABI in ``include/toygpu.h``) with ctypes and
offers the same three operators under the same names (snake-cased), on the same state the
driver holds: positions in ``[0, Boxsize]``, ``SphP.Hsml`` as warm start (0 = cold), the halo
table and ``Param.{Boxsize, Mpart[0], Mtotal}``.  There is no CPU fallback: if the CUDA
library or a GPU is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtoygpu.so")

DESNNGB = 295      # globals.h:48
NGBMAX = 2360      # globals.h:50
NUMITER = 64       # wvt_relax.c:7

WVT_SEQUENTIAL = 1  # tg_config.flags
FAST = 4  # tg_config.flags: FP32 kernel arithmetic in the warm sweep (tile_fast.cuh)
EXACT_NEIGHBOURS = 2  # tg_config.flags: exact predicate sets, not the reference tree's (include/toygpu.h)


class ToyGpuError(RuntimeError):
    pass


class _Halo(C.Structure):
    _fields_ = [("dcom", C.c_double * 3), ("rho0", C.c_double), ("beta", C.c_double),
                ("rcore", C.c_double), ("rcut", C.c_double), ("cuspy", C.c_int),
                ("mass_gas", C.c_double)]


class _Config(C.Structure):
    _fields_ = [("device", C.c_int), ("n_gas", C.c_int), ("boxsize", C.c_double),
                ("mpart_gas", C.c_double), ("mtotal", C.c_double), ("flags", C.c_uint),
                ("rank", C.c_int), ("nranks", C.c_int), ("stream", C.c_void_p),
                ("ngpus", C.c_int), ("devices", C.POINTER(C.c_int)),
                ("rho0_fac", C.c_double), ("rc_fac", C.c_double)]


class _BField(C.Structure):
    _fields_ = [("bfld_norm", C.c_double), ("bfld_eta", C.c_double), ("bmax_main", C.c_double),
                ("bmax_sub", C.c_double), ("sub_first", C.c_int),
                ("r_sample_gas", C.POINTER(C.c_double)), ("r_sample_dm", C.POINTER(C.c_double)),
                ("is_stripped", C.POINTER(C.c_int))]


class Stats(C.Structure):
    _fields_ = [("pair_evals", C.c_ulonglong), ("gathered", C.c_ulonglong),
                ("searches", C.c_ulonglong), ("hsml_iters", C.c_ulonglong),
                ("kernels", C.c_ulonglong), ("sweep_ms", C.c_double), ("step_ms", C.c_double),
                ("handed_back", C.c_ulonglong), ("displaced_nodes", C.c_ulonglong),
                ("displaced_particles", C.c_ulonglong), ("displaced_overflow", C.c_ulonglong),
                ("handback_why", C.c_ulonglong * 5),
                ("index_ms", C.c_double), ("tail_ms", C.c_double)]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k == "handback_why" else getattr(self, k))
                for k, _ in self._fields_}


class _Exchange(C.Structure):
    _fields_ = [("pos_hsml_dev", C.c_void_p), ("rho_dev", C.c_void_p),
                ("varhsml_dev", C.c_void_p), ("delta_dev", C.c_void_p),
                ("err_dev", C.c_void_p), ("lo", C.c_int), ("hi", C.c_int), ("chunk", C.c_int)]


_LOG_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                      C.c_void_p)

EXPORTS = [
    "tg_create", "tg_destroy", "tg_last_error", "tg_set_halos", "tg_upload", "tg_upload_soa",
    "tg_upload_soa_slice", "tg_set_cold", "tg_download_soa_slice", "tg_set_apot", "tg_download", "tg_download_soa", "tg_find_sph_quantities", "tg_regularise",
    "tg_bfld_from_rotA", "tg_wvt_iteration", "tg_wvt_begin", "tg_wvt_finish", "tg_wvt_scratch", "tg_get_stats", "tg_peano_keys",
    "tg_sort", "tg_find_ngb", "tg_guess_hsml", "tg_get_exchange",
    "tg_make_magnetic_field", "tg_get_apot", "tg_pin_host", "tg_unpin_host",
    "tg_comm_id", "tg_comm_init", "tg_halo_ids", "tg_sync_results",
    "tg_set_output_order", "tg_fill_block",
]

_lib = None


def build(verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a into toycluster_b200/libtoygpu.so (nvcc cross-compiles)."""
    out = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=True,
                         text=True)
    if verbose or out.returncode:
        print(out.stdout, out.stderr)
    if out.returncode:
        raise ToyGpuError("building libtoygpu.so failed")
    return LIB_PATH


def load():
    """dlopen libtoygpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("TOYGPU_LIB", LIB_PATH)      # developer knob: A/B builds of the kernels
    if not os.path.exists(path):
        raise ToyGpuError(f"{path} missing: run toycluster_b200.build() "
                          "(make -C toycluster_b200/csrc)")
    lib = C.CDLL(path)
    lib.tg_last_error.restype = C.c_char_p
    lib.tg_last_error.argtypes = [C.c_void_p]
    lib.tg_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Config)]
    lib.tg_destroy.argtypes = [C.c_void_p]
    lib.tg_set_halos.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Halo)]
    lib.tg_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    lib.tg_download.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    lib.tg_upload_soa.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tg_set_apot.argtypes = [C.c_void_p, C.c_void_p]
    lib.tg_upload_soa_slice.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.tg_set_cold.argtypes = [C.c_void_p, C.c_int]
    lib.tg_download_soa_slice.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tg_download_soa.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    lib.tg_find_sph_quantities.argtypes = [C.c_void_p]
    lib.tg_regularise.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_int)]
    lib.tg_bfld_from_rotA.argtypes = [C.c_void_p]
    lib.tg_wvt_iteration.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]
    lib.tg_wvt_begin.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.tg_wvt_finish.argtypes = [C.c_void_p, C.c_double]
    lib.tg_sync_results.argtypes = [C.c_void_p]
    lib.tg_wvt_scratch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tg_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.tg_peano_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tg_sort.argtypes = [C.c_void_p, C.c_void_p]
    lib.tg_find_ngb.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.POINTER(C.c_int)]
    lib.tg_guess_hsml.argtypes = [C.c_void_p, C.c_void_p]
    lib.tg_get_exchange.argtypes = [C.c_void_p, C.POINTER(_Exchange)]
    lib.tg_make_magnetic_field.argtypes = [C.c_void_p, C.POINTER(_BField), C.POINTER(C.c_double),
                                           C.POINTER(C.c_int)]
    lib.tg_get_apot.argtypes = [C.c_void_p, C.c_void_p]
    lib.tg_pin_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.tg_unpin_host.argtypes = [C.c_void_p, C.c_void_p]
    lib.tg_halo_ids.argtypes = [C.c_void_p, C.POINTER(_BField), C.c_void_p, C.c_void_p]
    lib.tg_set_output_order.argtypes = [C.c_void_p, C.c_void_p]
    lib.tg_fill_block.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.tg_comm_id.argtypes = [C.c_void_p]
    lib.tg_comm_init.argtypes = [C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class HotPath:
    """One device context == the global state the reference's path works on."""

    def __init__(self, n_gas, boxsize, mpart_gas, mtotal, halo_table, device=0, flags=0,
                 rank=0, nranks=1, stream=None, ngpus=0, devices=None, rho0_fac=0.0, rc_fac=0.0):
        self.lib = load()
        self.n = int(n_gas)
        self._ctx = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices) if devices else None
        cfg = _Config(int(device), self.n, float(boxsize), float(mpart_gas), float(mtotal),
                      int(flags), int(rank), int(nranks), C.c_void_p(stream or None),
                      int(ngpus), devs, float(rho0_fac), float(rc_fac))
        rc = self.lib.tg_create(C.byref(self._ctx), C.byref(cfg))
        if rc != 0:
            msg = self.lib.tg_last_error(None).decode()
            self._ctx = C.c_void_p()
            raise ToyGpuError(f"tg_create failed ({rc}): {msg}")
        rows = np.asarray(halo_table, dtype=np.float64).reshape(-1, 9)
        halos = (_Halo * len(rows))()
        for h, r in zip(halos, rows):
            h.dcom[0], h.dcom[1], h.dcom[2] = r[0], r[1], r[2]
            h.rho0, h.beta, h.rcore, h.rcut = r[3], r[4], r[5], r[6]
            h.cuspy, h.mass_gas = int(r[7]), r[8]
        self._check(self.lib.tg_set_halos(self._ctx, len(rows), halos))
        self.nhalos = len(rows)

    @classmethod
    def from_workload(cls, w, **kw):
        return cls(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), **kw)

    def _check(self, rc):
        if rc != 0:
            raise ToyGpuError(f"libtoygpu error {rc}: "
                              f"{self.lib.tg_last_error(self._ctx).decode()}")

    # ---- multi-GPU inside the library (one process per GPU) ------------------------------
    @staticmethod
    def comm_id() -> bytes:
        """128-byte NCCL id (rank 0); ship it to the other ranks, then comm_init everywhere."""
        buf = (C.c_ubyte * 128)()
        lib = load()
        if lib.tg_comm_id(buf) != 0:
            raise ToyGpuError("tg_comm_id: " + lib.tg_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, id128: bytes):
        buf = (C.c_ubyte * 128).from_buffer_copy(id128)
        self._check(self.lib.tg_comm_init(self._ctx, buf))

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.tg_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data in / out ------------------------------------------------------------------
    def upload(self, pos, hsml=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        assert pos.shape == (self.n, 3)
        if hsml is not None:
            hsml = np.ascontiguousarray(hsml, dtype=np.float32)
            assert hsml.shape == (self.n,)
        self._check(self.lib.tg_upload_soa(self._ctx, _ptr(pos), _ptr(hsml)))

    def upload_slice(self, pos, hsml=None) -> bool:
        """Multi-rank upload of this rank's slice only; returns the slice's cold flag.  The
        caller all-gathers the state buffer and calls set_cold(any rank cold)."""
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        if hsml is not None:
            hsml = np.ascontiguousarray(hsml, dtype=np.float32)
        cold = C.c_int(0)
        self._check(self.lib.tg_upload_soa_slice(self._ctx, _ptr(pos), _ptr(hsml), C.byref(cold)))
        return bool(cold.value)

    def set_cold(self, any_cold: bool):
        self._check(self.lib.tg_set_cold(self._ctx, int(bool(any_cold))))

    def download_slice(self, pos, hsml):
        """Write this rank's slice of (pos, Hsml) into the full host arrays."""
        self._check(self.lib.tg_download_soa_slice(self._ctx, _ptr(pos), _ptr(hsml)))

    def upload_records(self, P, SphP):
        """AoS records as the C driver holds them (numpy structured or raw byte arrays)."""
        self._check(self.lib.tg_upload(self._ctx, _ptr(P), P.strides[0], _ptr(SphP),
                                       SphP.strides[0]))

    def download_records(self, P, SphP):
        self._check(self.lib.tg_download(self._ctx, _ptr(P), P.strides[0], _ptr(SphP),
                                         SphP.strides[0]))

    def pin_host(self, arr):
        """Page-lock a numpy array that will be passed to upload_records / download_records
        repeatedly (the C driver's P / SphP).  Keep ``arr`` alive until unpin_host / close."""
        self._check(self.lib.tg_pin_host(self._ctx, _ptr(arr), arr.nbytes))

    def unpin_host(self, arr):
        self._check(self.lib.tg_unpin_host(self._ctx, _ptr(arr)))

    def set_apot(self, apot):
        apot = np.ascontiguousarray(apot, dtype=np.float32)
        assert apot.shape == (self.n, 3)
        self._check(self.lib.tg_set_apot(self._ctx, _ptr(apot)))

    def download(self, bfld=False):
        n = self.n
        out = dict(pos=np.empty((n, 3), np.float32), id=np.empty(n, np.int32),
                   hsml=np.empty(n, np.float32), rho=np.empty(n, np.float32),
                   varhsml=np.empty(n, np.float32), rho_model=np.empty(n, np.float32))
        b = np.empty((n, 3), np.float32) if bfld else None
        self._check(self.lib.tg_download_soa(self._ctx, _ptr(out["pos"]), _ptr(out["id"]),
                                             _ptr(out["hsml"]), _ptr(out["rho"]),
                                             _ptr(out["varhsml"]), _ptr(out["rho_model"]),
                                             _ptr(b)))
        if bfld:
            out["bfld"] = b
        return out

    # ---- the three operators ------------------------------------------------------------
    def find_sph_quantities(self):
        self._check(self.lib.tg_find_sph_quantities(self._ctx))

    def regularise_sph_particles(self, max_iters=1 << 30, log=None):
        """wvt_relax.c:25.  ``log(it, err_max, err_mean, err_diff, step)`` is called where the
        reference prints its '#NN: Err max=…' line; a truthy return stops the loop."""
        rows = []

        def _cb(it, emax, emean, ediff, step, _user):
            rows.append(dict(it=it, max=emax, mean=emean, diff=ediff, step=step))
            return int(bool(log(it, emax, emean, ediff, step))) if log else 0

        cb = _LOG_FN(_cb)
        done = C.c_int(0)
        self._check(self.lib.tg_regularise(self._ctx, int(min(max_iters, 1 << 30)),
                                           C.cast(cb, C.c_void_p), None, C.byref(done)))
        return done.value, rows

    def bfld_from_rotA_sph(self):
        self._check(self.lib.tg_bfld_from_rotA(self._ctx))

    def make_magnetic_field(self, bfld_norm, bfld_eta, r_sample_gas=None, r_sample_dm=None,
                            is_stripped=None, sub_first=None, bmax_main=18e-6, bmax_sub=2e-6):
        """Make_magnetic_field() (magnetic_field.c:12-131) -> (norm, particles capped); B and A
        stay on the device (download(bfld=True), get_apot())."""
        nh = self.nhalos
        gas = np.ascontiguousarray(np.zeros(nh) if r_sample_gas is None else r_sample_gas, np.float64)
        dm = np.ascontiguousarray(np.zeros(nh) if r_sample_dm is None else r_sample_dm, np.float64)
        st = np.ascontiguousarray(np.zeros(nh) if is_stripped is None else is_stripped, np.int32)
        par = _BField(bfld_norm, bfld_eta, bmax_main, bmax_sub,
                      nh if sub_first is None else int(sub_first),
                      gas.ctypes.data_as(C.POINTER(C.c_double)), dm.ctypes.data_as(C.POINTER(C.c_double)),
                      st.ctypes.data_as(C.POINTER(C.c_int)))
        norm, cnt = C.c_double(), C.c_int()
        self._check(self.lib.tg_make_magnetic_field(self._ctx, C.byref(par), C.byref(norm), C.byref(cnt)))
        return norm.value, cnt.value

    def halo_ids(self, r_sample_gas, is_stripped=None, sub_first=None):
        """positions.c:264-283: (haloID per gas particle, particles per halo) of the current state."""
        nh = self.nhalos
        gas = np.ascontiguousarray(r_sample_gas, np.float64)
        dm = np.zeros(nh)
        st = np.ascontiguousarray(np.zeros(nh) if is_stripped is None else is_stripped, np.int32)
        par = _BField(0.0, 0.0, 0.0, 0.0, nh if sub_first is None else int(sub_first),
                      gas.ctypes.data_as(C.POINTER(C.c_double)), dm.ctypes.data_as(C.POINTER(C.c_double)),
                      st.ctypes.data_as(C.POINTER(C.c_int)))
        ids = np.empty(self.n, np.int32)
        cnt = np.zeros(nh, np.int64)
        self._check(self.lib.tg_halo_ids(self._ctx, C.byref(par), _ptr(ids), _ptr(cnt)))
        return ids, cnt

    BLOCKS = {"POS": (0, 3), "RHO": (1, 1), "HSML": (2, 1), "BFLD": (3, 3), "RHOM": (4, 1)}

    def set_output_order(self, order=None):
        """io.c:85-133 / positions.c:405-443: order[k] = current device index of the particle
        the file holds at position k (None: the device order)."""
        if order is not None:
            order = np.ascontiguousarray(order, np.uint64)
            assert order.shape == (self.n,)
        self._check(self.lib.tg_set_output_order(self._ctx, _ptr(order)))

    def fill_block(self, label):
        """The gas part of the write buffer of one Gadget block, from the device arrays."""
        code, vals = self.BLOCKS[label]
        out = np.empty((self.n, vals) if vals > 1 else self.n, np.float32)
        self._check(self.lib.tg_fill_block(self._ctx, code, _ptr(out)))
        return out

    def get_apot(self):
        out = np.empty((self.n, 3), np.float32)
        self._check(self.lib.tg_get_apot(self._ctx, _ptr(out)))
        return out

    def wvt_iteration(self, step):
        emax, emean = C.c_double(), C.c_double()
        self._check(self.lib.tg_wvt_iteration(self._ctx, float(step), C.byref(emax),
                                              C.byref(emean)))
        return emax.value, emean.value

    def wvt_begin(self, step_guess):
        """-> (err_sum, err_max, count) of this rank's slice (wvt_relax.c:73-87)."""
        s, m, n = C.c_double(), C.c_double(), C.c_int()
        self._check(self.lib.tg_wvt_begin(self._ctx, float(step_guess), C.byref(s), C.byref(m),
                                          C.byref(n)))
        return s.value, m.value, n.value

    def wvt_finish(self, step_final):
        self._check(self.lib.tg_wvt_finish(self._ctx, float(step_final)))

    def sync_results(self):
        """Multi-rank: gather Rho / VarHsmlFac of every slice (after wvt_begin / wvt_finish)."""
        self._check(self.lib.tg_sync_results(self._ctx))

    def wvt_scratch(self):
        h = np.empty(self.n, np.float32)
        d = np.empty((self.n, 3), np.float32)
        self._check(self.lib.tg_wvt_scratch(self._ctx, _ptr(h), _ptr(d)))
        return h, d

    def stats(self) -> dict:
        s = Stats()
        self._check(self.lib.tg_get_stats(self._ctx, C.byref(s)))
        return s.as_dict()

    # ---- test hooks ---------------------------------------------------------------------
    def peano_keys(self):
        hi = np.empty(self.n, np.uint64)
        lo = np.empty(self.n, np.uint64)
        self._check(self.lib.tg_peano_keys(self._ctx, _ptr(hi), _ptr(lo)))
        return hi, lo

    def sort(self):
        perm = np.empty(self.n, np.int32)
        self._check(self.lib.tg_sort(self._ctx, _ptr(perm)))
        return perm

    def find_ngb(self, i, h):
        buf = np.empty(NGBMAX, np.int32)
        cnt = C.c_int(0)
        self._check(self.lib.tg_find_ngb(self._ctx, int(i), float(h), _ptr(buf), C.byref(cnt)))
        return buf[:cnt.value].copy()

    def guess_hsml(self):
        out = np.empty(self.n, np.float32)
        self._check(self.lib.tg_guess_hsml(self._ctx, _ptr(out)))
        return out

    def exchange(self):
        ex = _Exchange()
        self._check(self.lib.tg_get_exchange(self._ctx, C.byref(ex)))
        return ex
