"""CPU restatements of two arithmetic short-cuts of find_hsml_fast (toycluster_b200/csrc/
tile_fast.cuh), to pin the error bounds DESIGN.md quotes for them.  No GPU, no library: numpy
reproduces the device arithmetic step by step (float32 lane sums, round-to-nearest conversion,
int32 warp sum; float32 reciprocal + one FP64 Newton step)."""
import numpy as np


def _w(u):          # WC6 shape (1 - u)^8 (1 + 8u + 25u^2 + 32u^3), 0 <= w <= 1
    t = 1.0 - u
    return t ** 8 * (1 + 8 * u + 25 * u * u + 32 * u ** 3)


def test_fixed_point_lane_sum_is_below_the_float_partial_sums():
    """Sw = sum over 32 lanes of float32 partial sums, reduced as int32 after scaling by
    2^wshift with wshift = clz(cnt) - 1 (cnt << wshift < 2^31): order independent, and its
    rounding (<= 0.5 * 2^-wshift per lane) is below the float32 rounding already in the lanes."""
    rng = np.random.default_rng(3)
    for cnt in (295, 300, 511, 512, 549, 768, 2368):
        wshift = (32 - int(cnt).bit_length()) - 1
        assert (cnt << wshift) < 2 ** 31 <= (cnt << (wshift + 2))
        u = rng.uniform(0, 1, cnt)
        w = _w(u).astype(np.float32)
        pad = (-cnt) % 64
        w = np.concatenate([w, np.zeros(pad, np.float32)])
        lanes = w.reshape(-1, 32)                      # entry k belongs to lane k % 32
        part = np.zeros(32, np.float32)
        for row in lanes:                              # float32 accumulation per lane
            part = (part + row).astype(np.float32)
        exact = float(w.astype(np.float64).sum())
        fixed = np.rint(part.astype(np.float64) * 2.0 ** wshift).astype(np.int64)
        assert np.abs(fixed).max() < 2 ** 31 and fixed.sum() < 2 ** 31     # no overflow, per lane or total
        total = float(fixed.sum()) * 2.0 ** -wshift
        tree = float(part.astype(np.float64).sum())    # what the FP64 tree over the lanes gave
        assert abs(total - tree) <= 16 * 2.0 ** -wshift            # 32 lanes x half a unit
        assert abs(total - tree) < 2e-7 * exact                    # far below the 1e-5 tolerance
        # ... and no worse than the float32 partial sums are themselves
        assert abs(total - exact) <= abs(tree - exact) + 16 * 2.0 ** -wshift


def test_reciprocal_with_one_newton_step():
    """rcp_fast(x): float32 reciprocal (MUFU.RCP, ~1e-7) + one Newton step in double -> ~1e-14."""
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(1e-3, 1e9, 2000), 10.0 ** rng.uniform(-9, 12, 2000)])
    y0 = (np.float32(1) / x.astype(np.float32)).astype(np.float64) * (1 + rng.uniform(-2e-7, 2e-7, x.size))
    y = y0 + y0 * (1.0 - x * y0)
    assert np.abs(y * x - 1).max() < 1e-13
