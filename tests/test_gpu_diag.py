"""GPU parity of the flow-derived diagnostics (ripcurrents_module.cpp:900-1138):
bit-exact against the oracle (which tests/test_oracle_diag.py pins to cv2 / libm / a numpy restatement)."""
import os

import numpy as np
import pytest

from util import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ripcurrents_b200 import Context
    c = Context(0)
    yield c
    c.close()


def flow_field(h, w, seed, amp=1.0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    fl = np.stack([np.sin(xx / 17) * 2 + rng.standard_normal((h, w)) * 0.3, np.cos(yy / 11) * 1.5 + rng.standard_normal((h, w)) * 0.3], -1)
    return (fl * amp).astype(np.float32)


@pytest.mark.parametrize("shape", [(1080, 1920), (97, 131), (21, 21)])
def test_vector_to_color(ctx, oracle, shape):
    fl = flow_field(*shape, seed=1)
    fl[0, :6] = [(0, 0), (1, 0), (-1, 0), (-1, -0.0), (0, 1), (0, -1)]
    fl[1, 0] = (np.nan, 1); fl[1, 1] = (np.inf, 1); fl[1, 2] = (1e-30, -1e-38); fl[1, 3] = (-3e38, 3e38)
    maxd = 0.0
    for it in range(3):                                       # the maximum of frame t normalises frame t+1
        with np.errstate(over="ignore"):
            f = fl * np.float32(1 + 0.4 * it)
        _, ref, ref_max = oracle.vector_to_color(f, maxd)
        got, got_max = ctx.vector_to_color(f, maxd)
        assert np.array_equal(got, ref), it
        assert got_max == ref_max
        maxd = 3.0 if it == 0 else got_max
    assert np.array_equal(ctx.vector_to_color(fl, 2.0, flags=2)[0], oracle.vector_to_color(fl, 2.0, fma=False)[1])


def test_vector_to_color_all_directions(ctx, oracle):
    """Every hue boundary: 1.6 M directions around the circle at several radii (atan2f must match libm bit for bit)."""
    ang = np.linspace(-np.pi, np.pi, 1600 * 1024, dtype=np.float64)
    r = np.repeat(np.array([1e-3, 0.7, 1.0, 3.3], np.float64), ang.size // 4)
    fl = np.stack([r * np.cos(ang), r * np.sin(ang)], -1).astype(np.float32).reshape(1600, 1024, 2)
    _, ref, _ = oracle.vector_to_color(fl, 3.3)
    assert np.array_equal(ctx.vector_to_color(fl, 3.3)[0], ref)


@pytest.mark.parametrize("shape", [(1080, 1920), (97, 131), (20, 40)])
def test_shear_rate_to_color(ctx, oracle, shape):
    fl = flow_field(*shape, seed=2)
    rng = np.random.default_rng(3)
    img_o = rng.integers(0, 256, shape + (3,), dtype=np.uint8); img_g = img_o.copy()
    maxf = 0.0
    for it in range(3):
        ref_max = oracle.shear_to_color(fl * np.float32(1 + it), img_o, maxf)
        got_max = ctx.shear_rate_to_color(fl * np.float32(1 + it), img_g, maxf)
        assert np.array_equal(img_g, img_o), it
        assert got_max == ref_max
        maxf = got_max if it else 2.5


def test_context_state_mirrors_the_statics(oracle):
    """max pointer NULL: the previous maxima live in the context, starting at 0 like the reference's statics."""
    from ripcurrents_b200 import Context
    c = Context(0)
    fl = flow_field(64, 96, seed=4)
    vmax = smax = 0.0
    img_o = np.zeros((64, 96, 3), np.uint8); img_g = img_o.copy()
    for it in range(3):
        f = fl * np.float32(1 + 0.3 * it)
        _, ref, vmax = oracle.vector_to_color(f, vmax)
        assert np.array_equal(c.vector_to_color(f, None)[0], ref), it
        smax = oracle.shear_to_color(f, img_o, smax)
        c.shear_rate_to_color(f, img_g, None)
        assert np.array_equal(img_g, img_o), it
    c.close()


def test_subtract_mean_magnitude(ctx, oracle):
    for shape in [(1080, 1920), (97, 131)]:
        fl = flow_field(*shape, seed=5); fl[0, 0] = 0
        ref = fl.copy(); mv = oracle.subtract_mean_magnitude(ref)
        strict = fl.copy()
        assert ctx.subtract_mean_magnitude(strict, flags=1) == mv              # sequential fp32 sum: bit-exact
        assert np.array_equal(strict, ref)
        fast = fl.copy()
        mv_fast = ctx.subtract_mean_magnitude(fast)                            # fp64 tree sum: the better-rounded mean
        exact = np.hypot(fl[..., 0].astype(np.float64), fl[..., 1].astype(np.float64)).mean()
        assert abs(mv_fast - exact) <= 1e-6 * exact                            # tolerance: 1e-6 relative on the mean
        assert abs(mv_fast - mv) <= 2e-3 * exact                               # the reference's own fp32 drift
        assert np.abs(fast - ref).max() <= abs(mv_fast - mv) * 1.0001 + 1e-6


def test_hsv2bgr_golden_through_shear_border(ctx):
    """An image too small to have an interior goes through the HSV->BGR conversion only: cv2 fixtures."""
    z = np.load(os.path.join(GOLDEN, "hsv2bgr.npz"))
    for k in ("plane", "rnd"):
        hsv = z[k + "_hsv"]
        rows = hsv.reshape(-1, 3)
        n = rows.shape[0] // 16 * 16
        img = rows[:n].reshape(16, n // 16, 3).copy()                           # 16 rows <= 2*offset: no interior
        ctx.shear_rate_to_color(np.zeros((16, n // 16, 2), np.float32), img, 1.0)
        assert np.array_equal(img.reshape(-1, 3), z[k + "_bgr"].reshape(-1, 3)[:n]), k


def test_error_paths(ctx):
    """Bad arguments come back as negative codes (no crash, no CPU fallback): NULL pointers, bad geometry, missing state."""
    import ctypes as C
    from ripcurrents_b200 import Context
    lib = ctx.lib
    fl = np.zeros((8, 8, 2), np.float32); img = np.zeros((8, 8, 3), np.uint8)
    p = lambda a: C.c_void_p(a.ctypes.data)
    none = C.c_void_p(0)
    assert lib.rc_vector_to_color(ctx.h, p(fl), C.c_size_t(64), C.c_int(8), C.c_int(8), none, C.c_size_t(24), None, C.c_int(0)) == -1
    assert lib.rc_vector_to_color(ctx.h, p(fl), C.c_size_t(8), C.c_int(8), C.c_int(8), p(img), C.c_size_t(24), None, C.c_int(0)) == -1
    assert lib.rc_shear_rate_to_color(ctx.h, p(fl), C.c_size_t(64), C.c_int(8), C.c_int(8), p(img), C.c_size_t(8), None, C.c_int(0)) == -1
    assert lib.rc_subtract_mean_magnitude(ctx.h, none, C.c_size_t(64), C.c_int(8), C.c_int(8), C.c_int(0), None) == -1
    assert lib.rc_particle_fields(ctx.h, none, none, C.c_int(8), C.c_int(8), C.c_int(0), none, none, none, none, none, None) == -1
    assert lib.rc_particle_fields(ctx.h, p(fl), none, C.c_int(8), C.c_int(8), C.c_int(0), none, none, p(img), none, none, None) == -1
    assert lib.rc_normalize_jet(ctx.h, none, C.c_size_t(4), none, none, None) == -1
    assert lib.rc_streamline_positions(ctx.h, p(fl), C.c_int(0), C.c_int(8), p(fl), C.c_int(0)) == -1
    fresh = Context(0)                                   # flow == NULL means "the context's last flow": none yet
    assert lib.rc_vector_to_color(fresh.h, none, C.c_size_t(0), C.c_int(0), C.c_int(0), p(img), C.c_size_t(24), None, C.c_int(0)) == -4
    fresh.close()
    assert lib.rc_vector_to_color(None, p(fl), C.c_size_t(64), C.c_int(8), C.c_int(8), p(img), C.c_size_t(24), None, C.c_int(0)) == -1
