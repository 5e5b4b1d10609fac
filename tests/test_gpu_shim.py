"""The drop-in boundary end to end: the reference-facing C shim
(toycluster_b200/host/gpu_shim.c, compiled against the reference's own headers) drives
libtoygpu.so from the driver's global AoS state, and must leave that state exactly as the
reference's own Regularise_sph_particles / Find_sph_quantities / Bfld_from_rotA_SPH do."""
import numpy as np
import pytest

from oracle import ref
from toycluster_b200 import workloads

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ref.available() and __import__("os").path.exists(ref.SHIM)),
                                 reason="oracle/_ref libraries not built")]


def test_driver_sequence_through_the_shim():
    w = workloads.make("merger_1e6", n_gas=16384)
    args = (w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), 8)
    a = ref.Ref(*args)                   # the reference's own code
    b = ref.Ref(*args, shim=True)        # same harness, operators from gpu_shim.c + libtoygpu
    b.lib.toyshim_set_flags(1)           # TG_WVT_SEQUENTIAL: bit-comparable displacement
    for r in (a, b):
        r.load(w.pos)
        r.regularise(6)                  # main.c:52 (cut to 6 iterations)
        r.find_sph_quantities()          # main.c:54
    la, lb = ref.parse_log(a.log()), ref.parse_log(b.log())
    assert la == lb and len(la) == 6      # the printed '#NN: Err ...' lines are identical
    sa, sb = a.read(), b.read()
    for k in ("id", "pos", "hsml", "rho", "varhsml", "rho_model", "key_hi", "key_lo"):
        assert np.array_equal(sa[k], sb[k]), k
    apot = np.repeat(np.sqrt(sa["rho_model"] / sa["rho_model"].max())[:, None], 3, 1).astype(np.float32)
    apot[:, 1] *= 0.5
    for r in (a, b):
        r.set_apot(apot)                 # magnetic_field.c:33-69 writes Apot in the current order
        r.bfld_from_rotA()               # magnetic_field.c:21
    ba, bb = a.read()["bfld"], b.read()["bfld"]
    scale = np.abs(ba).max()
    assert np.abs(bb.astype(np.float64) - ba).max() <= 1e-5 * scale
    assert (ba == bb).mean() > 0.99
