"""Synthetic ``cluster.par`` workloads for the SPH/WVT hot path.

The full reference driver cannot run here (GSL is absent), so the inputs of the path --
gas positions in ``[0, Boxsize]``, the halo table ``Global_density_model`` reads and the
scalars ``Param.{Boxsize, Mpart[0], Mtotal}`` -- are produced by restating the closed-form
parts of the reference set-up:

* cosmology and Delta_c              cosmo.c:8-92
* R200, c_NFW (Duffy08), a_Hernquist setup.c:53-62, setup.c:503-521
* Boxsize, R_Sample, Rcut, rc        setup.c:65-77, setup.c:582-589
* Rho0 from M_gas(<R200)             setup.c:93-97 (trapezoid table instead of GSL QAG)
* particle numbers / masses          setup.c:189-218
* D_CoM of the two clusters          setup.c:274-293
* gas sampling and rejection         positions.c:90-133, positions.c:363-385
* shift into the box                 setup.c:473-497

Random numbers come from numpy (``data: synthetic``), not from the reference's per-thread
``erand48`` streams, so particle sets are statistically -- not bitwise -- those of the driver.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

# globals.h:66-78, compat constants (CGS)
_GRAV = 6.673e-8
_MSOL = 1.98892e33
_KPC = 3.08568025e21
_FOURPITHIRD = 4.18879032135009765
_R200_TO_RMAX_RATIO = 3.75

# cluster.par:25-27
_UNIT_LENGTH = 3.085678e21
_UNIT_MASS = 1.989e43

_PIERPAOLI = np.array(  # cosmo.c:69-74
    [[546.67, -137.82, 94.083, -204.68, 111.51],
     [-1745.6, 627.22, -1175.2, 2445.7, -1341.7],
     [3928.8, -1519.3, 4015.8, -8415.3, 4642.1],
     [-4384.8, 1748.7, -5362.1, 11257.0, -6218.2],
     [1842.3, -765.53, 2507.7, -5210.7, 2867.5]])


@dataclasses.dataclass
class HaloRow:
    """One row of the table ``Global_density_model`` walks (wvt_relax.c:235-253)."""
    dcom: tuple
    rho0: float
    beta: float
    rcore: float
    rcut: float
    cuspy: int
    mass_gas: float
    r200: float = 0.0
    r_sample_gas: float = 0.0
    npart_gas: int = 0

    def as_row9(self):
        return [self.dcom[0], self.dcom[1], self.dcom[2], self.rho0, self.beta,
                self.rcore, self.rcut, float(self.cuspy), self.mass_gas]


@dataclasses.dataclass
class Workload:
    name: str
    n_gas: int
    boxsize: float
    mpart_gas: float
    mtotal: float           # Param.Mtotal (sum of sampled halo masses, setup.c:152)
    halos: list
    pos: np.ndarray | None = None     # (n_gas, 3) float32 in [0, Boxsize]

    def halo_table(self) -> np.ndarray:
        return np.array([h.as_row9() for h in self.halos], dtype=np.float64)


def gas_density_profile(r, rho0, beta, rc, rcut):
    """setup.c:598-615, default build (no DOUBLE_BETA_COOL_CORES)."""
    r = np.asarray(r, dtype=np.float64)
    return rho0 * np.power(1 + (r / rc) ** 2, -1.5 * beta) / (1 + (r / rcut) ** 4)


def _mass_table(rho0, beta, rc, rcut, rmax, ntab=1024, sub=64):
    """M(<r) on the reference's 1024-point log grid from 0.1 kpc to rmax (setup.c:655-683)."""
    rmin = 0.1
    r_tab = rmin * np.power(10.0, math.log10(rmax / rmin) / (ntab - 1) * np.arange(ntab))
    r_tab[0] = 0.0                               # the reference leaves entry 0 at r = 0, M = 0
    m_tab = np.zeros(ntab)
    # composite Simpson between table nodes; the first interval starts at 0
    acc = 0.0
    lo = 0.0
    for i in range(1, ntab):
        hi = r_tab[i]
        x = np.linspace(lo, hi, sub + 1)
        y = 4 * math.pi * x * x * gas_density_profile(x, rho0, beta, rc, rcut)
        h = (hi - lo) / sub
        acc += h / 3 * (y[0] + y[-1] + 4 * y[1:-1:2].sum() + 2 * y[2:-1:2].sum())
        m_tab[i] = acc
        lo = hi
    return r_tab, m_tab


def derive(n_gas: int, mass_ratio: float, mtot200: float = 1e5, redshift: float = 0.87,
           bf: float = 0.17, beta: float = 0.54, impact_param: float = 50.0,
           name: str = "", n_sub: int = 0) -> Workload:
    """Halo table and scalars for a one- or two-cluster system, without positions.

    ``n_sub`` > 0 appends that many subhalo rows to the table (BASELINE config 5).  They are a
    synthetic stand-in for substructure.c:116-183 (Giocoli mass-function sampling, which is
    host set-up outside the hot path): the -DADD_THIRD_SUBHALO halo of cluster.par:46
    (1e12 Msol at the origin) followed by masses log-uniform in [1e12, 2e13] Msol placed
    uniformly inside R200 of the main halo, each with the main halos' closed-form profile
    parameters.  What the hot path sees is what matters: a Global_density_model table of
    ~70 rows and strong density contrast."""
    h100, om, ol = 0.7, 0.3, 0.7                                   # cosmo.c:11-13
    h0_cgs = 100 * h100 * 1e5 / 1000 / _KPC
    ez = math.sqrt(ol + (1 - om - ol) * (1 + redshift) ** 2 + om * (1 + redshift) ** 3)
    rho_crit = 3 * (h0_cgs * ez) ** 2 / (8 * math.pi * _GRAV)
    x, y = om - 0.2, ol
    delta = om * sum(_PIERPAOLI[i, j] * x ** i * y ** j for i in range(5) for j in range(5))

    nh = 1 if mass_ratio == 0 else 2
    m200 = [mtot200 / (1 + mass_ratio), 0.0]
    m200[1] = mtot200 - m200[0]
    n_main = nh
    sub_pos = []
    if n_sub > 0:
        rng = np.random.default_rng(68)                       # substructure.c:127: <= 68 subhalos
        m200 = m200[:nh] + [100.0] + list(np.exp(rng.uniform(math.log(100.0), math.log(2000.0),
                                                            n_sub - 1)))
        nh += n_sub

    r200, a_hq, rs = [], [], []
    for i in range(nh):
        mass_msol = m200[i] * _UNIT_MASS / _MSOL
        c_nfw = 5.74 * (mass_msol / (2e12 / h100)) ** -0.097 * (1 + redshift) ** -0.47
        r = (m200[i] * _UNIT_MASS / (delta * rho_crit * _FOURPITHIRD)) ** (1 / 3) / _UNIT_LENGTH
        r200.append(r)
        rs.append(r / c_nfw)
        a_hq.append(rs[i] * math.sqrt(2 * (math.log(1 + c_nfw) - c_nfw / (1 + c_nfw))))

    boxsize = math.floor(2 * _R200_TO_RMAX_RATIO * r200[0])

    halos, mass_dm = [], []
    for i in range(nh):
        rs_gas = 1.8 * r200[i] if i else math.sqrt(3) * boxsize / 2
        rs_dm = 1.8 * r200[i] if i else boxsize / 2
        rcut = 1.4 * r200[i]
        rc = rs[i] / 3
        mgas200 = m200[i] - m200[i] / (1 + bf)
        r_tab, m_tab = _mass_table(1.0, beta, rc, rcut, 1.1 * rs_gas)
        rho0 = mgas200 / np.interp(r200[i], r_tab, m_tab)
        mass_gas = rho0 * np.interp(min(rs_gas, r_tab[-1]), r_tab, m_tab)
        a = a_hq[i]
        corr = 1 / (1 + 2 * a / rs_dm + (a / rs_dm) ** 2)
        mass_dm.append(m200[i] / (1 + bf) * (1 + 2 * a / r200[i] + (a / r200[i]) ** 2) * corr)
        halos.append(HaloRow(dcom=(0.0, 0.0, 0.0), rho0=float(rho0), beta=beta, rcore=rc,
                             rcut=rcut, cuspy=0, mass_gas=float(mass_gas), r200=r200[i],
                             r_sample_gas=rs_gas))

    if n_sub > 0:
        c0 = np.array(halos[0].dcom)
        for k in range(n_sub):
            if k == 0:
                off = np.zeros(3)                              # SubFirstPos = 0 (cluster.par:48-50)
            else:
                v = rng.normal(size=3)
                off = v / np.linalg.norm(v) * r200[0] * rng.random() ** (1 / 3)
            sub_pos.append(off)
    if n_main == 2:
        d = 0.9 * (r200[0] + r200[1])
        x0 = -m200[1] * d / mtot200
        y0 = -m200[1] * impact_param / mtot200
        halos[0].dcom = (x0, y0, 0.0)
        halos[1].dcom = (d + x0, impact_param + y0, 0.0)
    for k, off in enumerate(sub_pos):                          # subhalos ride with the main halo
        halos[n_main + k].dcom = tuple(float(v) for v in np.array(halos[0].dcom) + off)

    mgas_tot = sum(h.mass_gas for h in halos)
    mpart = mgas_tot / n_gas
    for h in halos:
        h.npart_gas = int(round(h.mass_gas / mpart))
    halos[0].npart_gas += n_gas - sum(h.npart_gas for h in halos)   # keep the total exact

    return Workload(name=name, n_gas=n_gas, boxsize=float(boxsize), mpart_gas=float(mpart),
                    mtotal=float(mgas_tot + sum(mass_dm)), halos=halos)


def _halo_containing(halos, xyz):
    """positions.c:363-385 for gas: arg-max of the model density inside each sampling radius."""
    best = np.zeros(len(xyz), dtype=np.int64)
    rho_max = np.zeros(len(xyz))
    for j, h in enumerate(halos):
        d = xyz - np.asarray(h.dcom)
        r = np.sqrt((d * d).sum(axis=1)).astype(np.float32).astype(np.float64)
        rho = gas_density_profile(r, h.rho0, h.beta, h.rcore, h.rcut)
        take = (rho > rho_max) & (r < h.r_sample_gas)
        best[take] = j
        rho_max[take] = rho[take]
    return best


def sample_positions(w: Workload, seed: int = 14041981) -> np.ndarray:
    """Gas positions as Make_positions + Shift_Origin leave them (float32, [0, Boxsize])."""
    boxhalf = w.boxsize / 2
    out = np.empty((w.n_gas, 3), dtype=np.float32)
    start = 0
    for i, h in enumerate(w.halos):
        rng = np.random.default_rng(seed * (i + 1))
        r_tab, m_tab = _mass_table(h.rho0, h.beta, h.rcore, h.rcut, 1.1 * h.r_sample_gas)
        need = h.npart_gas
        got = []
        while need > 0:
            n = int(need * 1.3) + 1024
            theta = np.arccos(2 * rng.random(n) - 1)
            phi = 2 * math.pi * rng.random(n)
            r = np.interp(rng.random(n) * h.mass_gas, m_tab, r_tab)
            xyz = np.stack([r * np.sin(theta) * np.cos(phi),
                            r * np.sin(theta) * np.sin(phi),
                            r * np.cos(theta)], axis=1)
            ok = _halo_containing(w.halos, (xyz + np.asarray(h.dcom)).astype(np.float32)
                                  .astype(np.float64)) == i
            ok &= (np.abs(xyz) <= boxhalf).all(axis=1)
            xyz = xyz[ok][:need]
            got.append(xyz)
            need -= len(xyz)
        xyz = np.concatenate(got).astype(np.float32)
        # Shift_Origin: += D_CoM (as float), += boxhalf, wrap into [0, Boxsize]
        xyz += np.asarray(h.dcom, dtype=np.float32)
        out[start:start + h.npart_gas] = xyz
        start += h.npart_gas
    box = np.float32(w.boxsize)
    out += np.float32(box / 2)
    out = np.where(out > box, out - box, out)
    out = np.where(out < 0, out + box, out)
    return np.ascontiguousarray(out, dtype=np.float32)


# BASELINE.json configs. n_gas = Ntotal / 2.
CONFIGS = {
    "single_1e5": dict(n_gas=100_000, mass_ratio=0.0),     # configs[0]: shipped cluster.par
    "merger_1e6": dict(n_gas=1_000_000, mass_ratio=0.3125),  # configs[1]
    "merger_1e7": dict(n_gas=10_000_000, mass_ratio=0.3125),  # configs[2]
    "merger_sub_1e7": dict(n_gas=10_000_000, mass_ratio=0.3125, n_sub=68),  # configs[4]
}


def make(name: str, n_gas: int | None = None, seed: int = 14041981,
         with_positions: bool = True) -> Workload:
    """Build a named workload; ``n_gas`` overrides the particle count (reduced-size tests)."""
    cfg = dict(CONFIGS[name])
    if n_gas is not None:
        cfg["n_gas"] = int(n_gas)
    w = derive(name=name, **cfg)
    if with_positions:
        w.pos = sample_positions(w, seed)
    return w


def snap_to_cell_planes(pos: np.ndarray, boxsize: float, count: int, seed: int = 7,
                        levels=(3, 4, 5, 6, 7)) -> np.ndarray:
    """Stress input for the reference octree's node placement (tree.c:298-310): move one
    coordinate of ``count`` particles exactly onto the centre plane of the level-(L-1) cell
    they lie in.  Such a particle belongs to the upper level-L cell by its key, while the sign
    test ``Pos > centre`` says lower, so whenever it is the first particle of its cell the
    reference displaces that node and its whole subtree by one cell size.  Production inputs
    do this by chance (~8 nodes per 1e6 particles); this makes it common at test sizes."""
    rng = np.random.default_rng(seed)
    out = np.array(pos, dtype=np.float32, copy=True)
    pick = rng.choice(len(out), size=count, replace=False)
    for i in pick:
        L = int(rng.choice(levels))
        d = int(rng.integers(3))
        cells = 1 << (L - 1)
        k = min(int(np.float64(out[i, d]) / boxsize * cells), cells - 1)
        out[i, d] = np.float32(boxsize * (k + 0.5) / cells)
    return out
