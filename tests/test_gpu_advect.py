"""GPU parity of A7 (particle advection, streaklines) through the C ABI: bit-exact against the CPU oracle
(pathlines.cpp:9-46, ripcurrents_module.cpp:486-679, ripcurrents.cpp:611-698, Streakline.cpp:22-48)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ripcurrents_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _field(rng, h, w, s=1.5):
    return rng.normal(0, s, (h, w, 2)).astype(np.float32)


@pytest.mark.parametrize("variant,dt,it,upper", [(0, 2.0, 3, 0.0), (1, 2.0, 1, 2.5), (2, 0.1, 100, 45.0),
                                                 (2, 0.1, 20, 1.5), (3, 0.3, 7, 0.0), (4, 0.0, 0, 0.0),
                                                 (6, 2.0, 5, 1.2)])
def test_variants_bit_exact(ctx, oracle, variant, dt, it, upper):
    rng = np.random.default_rng(variant + 10)
    h, w = 120, 200
    flow = _field(rng, h, w)
    seeds = (rng.random((20000, 2)) * [w + 6, h + 6] - 3).astype(np.float32)     # some start outside
    home = None
    if variant == 6:
        home = rng.integers(0, [w, h], (20000, 2)).astype(np.int32)
        seeds = rng.normal(0, 2, (20000, 2)).astype(np.float32)
    ref = seeds.copy()
    oracle.advect(flow, ref, dt, it, upper, variant, home=home)
    ctx.advect(flow, seeds, dt, it, upper, variant, home=home)
    assert np.array_equal(seeds.view(np.uint32), ref.view(np.uint32))


def test_field_dense_bit_exact(ctx, oracle):
    # streamline_field over every pixel for several frames (ripcurrents.cpp:229-231: dt=2, iterations=1)
    rng = np.random.default_rng(3)
    h, w = 96, 160
    disp = np.zeros((h * w, 2), np.float32); dist = np.zeros(h * w, np.float32)
    rdisp = disp.copy(); rdist = dist.copy()
    for t in range(5):
        flow = _field(rng, h, w, 0.8)
        oracle.advect(flow, rdisp, 2.0, 1, 1.5, oracle.ADV_FIELD, dist=rdist)
        ctx.advect(flow, disp, 2.0, 1, 1.5, 5, dist=dist)
        assert np.array_equal(disp.view(np.uint32), rdisp.view(np.uint32))
        assert np.array_equal(dist.view(np.uint32), rdist.view(np.uint32))


def test_rotation_field_invariant_1m_seeds(ctx):
    """Size-independent property at BASELINE size (1M seeds, 1080p): on the analytic rotation field of
    validate_streamlines (main.cpp:375-380) one Euler step multiplies b(x-cx)^2 + a(y-cy)^2 by (1 + a b dt^2)."""
    h, w = 1080, 1920
    rr, cc = np.mgrid[0:h, 0:w].astype(np.float64)
    a, b, dt, steps = 100 / h, 100 / w, 0.03, 50
    flow = np.stack([-(rr - h / 2) * a, (cc - w / 2) * b], -1).astype(np.float32)
    rng = np.random.default_rng(1)
    n = 1 << 20
    seeds = (rng.random((n, 2)) * [w * 0.5, h * 0.5] + [w * 0.25, h * 0.25]).astype(np.float32)
    q0 = b * (seeds[:, 0].astype(np.float64) - w / 2) ** 2 + a * (seeds[:, 1].astype(np.float64) - h / 2) ** 2
    ctx.advect(flow, seeds, dt, steps, 1e9, 2)
    q1 = b * (seeds[:, 0].astype(np.float64) - w / 2) ** 2 + a * (seeds[:, 1].astype(np.float64) - h / 2) ** 2
    ok = q0 > 1.0
    assert np.abs(q1[ok] / q0[ok] - (1 + a * b * dt * dt) ** steps).max() < 2e-3


def test_streakline_bit_exact(ctx, oracle):
    rng = np.random.default_rng(5)
    h, w, E, cap = 100, 150, 37, 24
    em = (rng.random((E, 2)) * [w - 4, h - 4] + 2).astype(np.float32)
    verts = np.zeros((E, cap, 2), np.float32); cnt = np.ones(E, np.int32); verts[:, 0] = em
    rverts = verts.copy(); rcnt = cnt.copy()
    for t in range(30):                         # runs past cap: growth must stop, motion continues
        flow = _field(rng, h, w, 2.0 if t % 7 else 12.0)      # every 7th frame has jumps > 10 % of the frame
        oracle.streakline_step(flow, em, rverts, rcnt)
        ctx.streakline_step(flow, em, verts, cnt)
        assert np.array_equal(cnt, rcnt), t
        for e in range(E):
            assert np.array_equal(verts[e, :cnt[e]].view(np.uint32), rverts[e, :rcnt[e]].view(np.uint32)), (t, e)
    assert cnt.max() == cap


def test_advect_on_context_flow(ctx, oracle):
    """flow == NULL: seeds ride the flow the context just computed (device-resident pipeline)."""
    from ripcurrents_b200 import synth
    fr = synth.clip(256, 192, 2, seed=2)
    flow = ctx.farneback(fr[0], fr[1], 0.5, 2, 3, 2, 15, 1.2, 0)
    rng = np.random.default_rng(0)
    seeds = (rng.random((5000, 2)) * [254, 190] + 1).astype(np.float32)
    ref = seeds.copy()
    oracle.advect(flow, ref, 1.0, 1, 0.0, oracle.ADV_PATHLINE)
    ctx.advect(None, seeds, 1.0, 1, 0.0, 0)
    assert np.array_equal(seeds.view(np.uint32), ref.view(np.uint32))


def test_empty_and_degenerate_inputs(ctx, oracle):
    """No seeds, no emitters, zero iterations, a streakline that is already full: calls succeed and leave data alone."""
    flow = (np.random.default_rng(0).standard_normal((32, 48, 2)) * 2).astype(np.float32)
    ctx.advect(flow, np.empty((0, 2), np.float32), 1.0, 1, 0.0, 0)                       # nothing to move
    seeds = np.array([[5.5, 6.5], [40.0, 20.0]], np.float32); keep = seeds.copy()
    ctx.advect(flow, seeds, 1.0, 0, 0.0, 0)                                               # zero iterations
    assert np.array_equal(seeds, keep)
    ctx.streakline_step(flow, np.empty((0, 2), np.float32), np.empty((0, 4, 2), np.float32), np.empty(0, np.int32))
    em = np.array([[10.0, 10.0]], np.float32)
    verts = np.zeros((1, 3, 2), np.float32); verts[0, :, :] = [[10, 10], [11, 11], [12, 12]]
    cnt = np.array([3], np.int32)                                                         # capacity reached
    rv, rc = verts.copy(), cnt.copy()
    oracle.streakline_step(flow, em, rv, rc)
    ctx.streakline_step(flow, em, verts, cnt)
    assert np.array_equal(cnt, rc) and np.array_equal(verts[0, :cnt[0]], rv[0, :rc[0]])
