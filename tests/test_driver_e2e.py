"""the WHOLE reference program -- main.c, cluster.par parser, set-up, sampling,
WVT relaxation, magnetic field, temperatures, velocities, Gadget writer, all unmodified
reference sources -- built twice by oracle/Makefile (`make driver`): with its own
tree.o/sph.o/wvt_relax.o/peano.o, and with toycluster_b200/host/gpu_shim.c + libtoygpu.so in
their place.  Same parameter file in, Gadget IC file out; the files must agree block by block.
(GSL is absent here; both variants link the same stand-in, oracle/compat/gsl_compat.c.)"""
import os
import struct
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
CPU, GPU = os.path.join(REFDIR, "Toycluster_cpu"), os.path.join(REFDIR, "Toycluster_gpu")
GPU_B = os.path.join(REFDIR, "Toycluster_gpu_b")      # ... with Make_magnetic_field on the device too

PAR = """Output_file ./{out}
Ntotal      {ntotal}
Mtotal      1e5
Mass_Ratio  {mass_ratio}
ImpactParam 50
ZeroEOrbitFrac 0.8
Cuspy       0
Redshift	0.87
Bfld_Norm   {bnorm}
Bfld_Eta    0.5
Bfld_Scale  100
bf          0.17
h_100       0.7
UnitLength_in_cm 			3.085678e21
UnitMass_in_g 				1.989e43
UnitVelocity_in_cm_per_s 	1e5
"""


def read_gadget2(path):
    """Format-2 snapshot as io.c:41-133 writes it -> {label: raw bytes}."""
    blocks = {}
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    while off < len(data):
        (sz,) = struct.unpack_from("<i", data, off)
        assert sz == 8
        label = data[off + 4:off + 8].decode()
        off += 4 + 8 + 4
        (n,) = struct.unpack_from("<i", data, off)
        blocks[label] = data[off + 4:off + 4 + n]
        off += 4 + n + 4
    return blocks


def run(exe, par, cwd, env_extra):
    env = dict(os.environ, OMP_NUM_THREADS="1", **env_extra)
    r = subprocess.run([exe, par], cwd=cwd, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


needs_drivers = pytest.mark.skipif(not (os.path.exists(CPU) and os.path.exists(GPU)),
                                   reason="oracle/_ref drivers not built (make -C oracle driver)")


@needs_drivers
def test_cpu_driver_runs_and_writes_all_blocks(tmp_path):
    (tmp_path / "a.par").write_text(PAR.format(out="IC_a", ntotal=6000, mass_ratio=0, bnorm="20e-6"))
    out = run(CPU, "a.par", tmp_path, {})
    assert "Starting iterative SPH regularisation" in out and "#00: Err max=" in out
    b = read_gadget2(tmp_path / "IC_a")
    assert [k.strip() for k in b] == ["HEAD", "POS", "VEL", "ID", "U", "RHO", "HSML", "BFLD", "RHOM"]
    assert len(b["POS "]) == 6000 * 12 and len(b["RHO "]) == 3000 * 4


@pytest.mark.gpu
@needs_drivers
@pytest.mark.parametrize("mass_ratio,ntotal,exe,bnorm", [
    (0, 20000, GPU, "20e-6"), (0.3125, 24000, GPU, "20e-6"),
    (0.3125, 24000, GPU_B, "20e-6"),
    (0, 20000, GPU_B, "60e-6")])        # strong field: the cap of magnetic_field.c:115-125 bites
def test_gpu_driver_writes_the_same_gadget_file(tmp_path, mass_ratio, ntotal, exe, bnorm):
    (tmp_path / "c.par").write_text(PAR.format(out="IC_c", ntotal=ntotal, mass_ratio=mass_ratio, bnorm=bnorm))
    (tmp_path / "g.par").write_text(PAR.format(out="IC_g", ntotal=ntotal, mass_ratio=mass_ratio, bnorm=bnorm))
    out_c = run(CPU, "c.par", tmp_path, {})
    out_g = run(exe, "g.par", tmp_path, {"TOYGPU_FLAGS": "1"})      # TG_WVT_SEQUENTIAL
    # the per-iteration log lines of wvt_relax.c:91-92, as printed
    it_c = [l for l in out_c.splitlines() if l.lstrip().startswith("#")]
    it_g = [l for l in out_g.splitlines() if l.lstrip().startswith("#")]
    assert len(it_c) >= 3 and it_c == it_g
    # magnetic_field.c:90,129 print the normalisation and the number of capped particles
    for key in ("Bfld Norm =", "Bfld of "):
        lc = [l for l in out_c.splitlines() if l.startswith(key)]
        lg = [l for l in out_g.splitlines() if l.startswith(key)]
        assert len(lc) == 1 and len(lg) == 1
        if key == "Bfld of ":
            nc, ng_ = int(lc[0].split()[2]), int(lg[0].split()[2])
            assert abs(nc - ng_) <= max(2, nc // 200)       # B within 1e-5: a few sit on the cap
            if bnorm != "20e-6":
                assert nc > 10
        else:
            vc, vg = float(lc[0].split("=")[1]), float(lg[0].split("=")[1])
            assert abs(vc - vg) <= 2e-5 * abs(vc)
    c, g = read_gadget2(tmp_path / "IC_c"), read_gadget2(tmp_path / "IC_g")
    assert list(c) == list(g)
    for label in c:
        if label == "BFLD":
            continue
        assert c[label] == g[label], label
    # rot(A) sums in FP64 trees instead of serially: 1e-5 of the field scale
    bc = np.frombuffer(c["BFLD"], np.float32).reshape(-1, 3)
    bg = np.frombuffer(g["BFLD"], np.float32).reshape(-1, 3)
    scale = np.abs(bc).max(axis=1, keepdims=True) + 1e-30
    assert (np.abs(bg - bc) / scale).max() < 1e-5


# ---- compile-time variants of the reference (Makefile:4-25) ---------------------------------
# Each pair is the WHOLE reference program built with the option, once with its own hot path
# and once with gpu_shim.c on libtoygpu (the cubic-spline pair links libtoygpu_m4.so, the
# library built with -DTG_CUBIC_SPLINE).

def _variant_pair(tmp_path, suffix, extra_par, ntotal=20000, mass_ratio=0.0):
    cpu, gpu = CPU + suffix, GPU + suffix
    if not (os.path.exists(cpu) and os.path.exists(gpu)):
        pytest.skip(f"oracle/_ref drivers{suffix} not built (make -C oracle driver)")
    for tag in ("c", "g"):
        (tmp_path / f"{tag}.par").write_text(
            PAR.format(out=f"IC_{tag}", ntotal=ntotal, mass_ratio=mass_ratio, bnorm="20e-6") + extra_par)
    out_c = run(cpu, "c.par", tmp_path, {})
    out_g = run(gpu, "g.par", tmp_path, {"TOYGPU_FLAGS": "1"})      # TG_WVT_SEQUENTIAL
    it_c = [l for l in out_c.splitlines() if l.lstrip().startswith("#")]
    it_g = [l for l in out_g.splitlines() if l.lstrip().startswith("#")]
    assert len(it_c) >= 3 and it_c == it_g, (it_c[:3], it_g[:3])
    c, g = read_gadget2(tmp_path / "IC_c"), read_gadget2(tmp_path / "IC_g")
    assert list(c) == list(g)
    for label in c:
        if label == "BFLD":
            bc = np.frombuffer(c["BFLD"], np.float32).reshape(-1, 3)
            bg = np.frombuffer(g["BFLD"], np.float32).reshape(-1, 3)
            scale = np.abs(bc).max(axis=1, keepdims=True) + 1e-30
            assert (np.abs(bg - bc) / scale).max() < 1e-5
        else:
            assert c[label] == g[label], label
    return out_c


@pytest.mark.gpu
def test_cubic_spline_build_writes_the_same_gadget_file(tmp_path):
    """-DSPH_CUBIC_SPLINE (sph.c:140-146,201,442-466; globals.h:40-52; wvt_relax.c:48-49): M4
    kernel, 50 neighbours, NGBMAX 400, no bias correction, VarHsmlFac = 1, step 0.035."""
    out = _variant_pair(tmp_path, "_m4", "")
    assert "step=0.035" in out


@pytest.mark.gpu
def test_cool_core_build_writes_the_same_gadget_file(tmp_path):
    """-DDOUBLE_BETA_COOL_CORES (setup.c:604-612): halo 0 is cuspy (Cuspy bit 0) and gets the
    second beta component rho0*Rho0_Fac / (1 + (r Rc_Fac / rc)^2) in Global_density_model."""
    par = PAR.replace("Cuspy       0", "Cuspy       1") + "Rho0_Fac    50\nRc_Fac      40\n"
    cpu, gpu = CPU + "_cc", GPU + "_cc"
    if not (os.path.exists(cpu) and os.path.exists(gpu)):
        pytest.skip("oracle/_ref drivers_cc not built (make -C oracle driver)")
    for tag in ("c", "g"):
        (tmp_path / f"{tag}.par").write_text(par.format(out=f"IC_{tag}", ntotal=20000, mass_ratio=0, bnorm="20e-6"))
    out_c = run(cpu, "c.par", tmp_path, {})
    out_g = run(gpu, "g.par", tmp_path, {"TOYGPU_FLAGS": "1"})
    it_c = [l for l in out_c.splitlines() if l.lstrip().startswith("#")]
    it_g = [l for l in out_g.splitlines() if l.lstrip().startswith("#")]
    assert len(it_c) >= 3 and it_c == it_g
    c, g = read_gadget2(tmp_path / "IC_c"), read_gadget2(tmp_path / "IC_g")
    for label in c:
        if label != "BFLD":
            assert c[label] == g[label], label
    # and the option really changed the model: the default build on the same file differs
    (tmp_path / "d.par").write_text(par.format(out="IC_d", ntotal=20000, mass_ratio=0, bnorm="20e-6"))
    run(CPU, "d.par", tmp_path, {})
    d = read_gadget2(tmp_path / "IC_d")
    assert d["RHOM"] != c["RHOM"]


@pytest.mark.gpu
def test_reassign_on_the_device_writes_the_same_gadget_file(tmp_path):
    """Reassign_particles_to_halos() from the shim -- Halo_containing per particle on
    the device (tg_halo_ids), the reference's own index heapsort on those ids -- leaves the
    particle order of the Gadget file unchanged, two halos and all."""
    gpu_r = GPU + "_r"
    if not (os.path.exists(CPU) and os.path.exists(gpu_r)):
        pytest.skip("oracle/_ref/Toycluster_gpu_r not built (make -C oracle driver)")
    for tag in ("c", "g"):
        (tmp_path / f"{tag}.par").write_text(PAR.format(out=f"IC_{tag}", ntotal=30000, mass_ratio=0.3125, bnorm="20e-6"))
    out_c = run(CPU, "c.par", tmp_path, {})
    out_g = run(gpu_r, "g.par", tmp_path, {"TOYGPU_FLAGS": "1"})
    dist_c = [l for l in out_c.splitlines() if l.startswith("   Main ") or l.startswith("   Bullet ")]
    dist_g = [l for l in out_g.splitlines() if l.startswith("   Main ") or l.startswith("   Bullet ")]
    assert len(dist_c) >= 2 and dist_c == dist_g          # "Particle Distribution after Relaxation"
    c, g = read_gadget2(tmp_path / "IC_c"), read_gadget2(tmp_path / "IC_g")
    for label in c:
        if label != "BFLD":
            assert c[label] == g[label], label


@pytest.mark.gpu
def test_blocks_from_the_device_write_the_same_gadget_file(tmp_path):
    """add_block() (io.c:85-133) from the shim -- the gas range of POS and the blocks
    RHO, HSML, BFLD, RHOM are filled on the device from the SoA arrays in the file order
    Reassign_particles_to_halos() left (tg_set_output_order / tg_fill_block); everything else of
    the writer is io.c's own.  Two halos, so the file order differs from the device's."""
    gpu_w = GPU + "_w"
    if not (os.path.exists(CPU) and os.path.exists(gpu_w)):
        pytest.skip("oracle/_ref/Toycluster_gpu_w not built (make -C oracle driver)")
    for tag in ("c", "g"):
        (tmp_path / f"{tag}.par").write_text(PAR.format(out=f"IC_{tag}", ntotal=30000, mass_ratio=0.3125, bnorm="60e-6"))
    out_c = run(CPU, "c.par", tmp_path, {})
    # TOYSHIM_POISON_RECORDS: the shim overwrites SphP.Rho/Hsml/Bfld/Rho_Model and the gas P.Pos of
    # the driver's records with NaN just before the writer runs, so a block that were still
    # filled from the records could not pass
    out_g = run(gpu_w, "g.par", tmp_path, {"TOYGPU_FLAGS": "1", "TOYSHIM_POISON_RECORDS": "1"})
    blk_c = [l for l in out_c.splitlines() if l.startswith("   Block ")]
    blk_g = [l for l in out_g.splitlines() if l.startswith("   Block ")]
    assert len(blk_c) == 8 and blk_c == blk_g
    c, g = read_gadget2(tmp_path / "IC_c"), read_gadget2(tmp_path / "IC_g")
    assert list(c) == list(g)
    for label in c:
        if label != "BFLD":
            assert c[label] == g[label], label
    bc = np.frombuffer(c["BFLD"], np.float32).reshape(-1, 3)
    bg = np.frombuffer(g["BFLD"], np.float32).reshape(-1, 3)
    scale = np.abs(bc).max(axis=1, keepdims=True) + 1e-30
    assert (np.abs(bg - bc) / scale).max() < 1e-5
