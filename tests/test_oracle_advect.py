"""Pins oracle/advect_oracle.c: literal numpy re-reading of the Euler/bilinear step and the analytic
rotation-field property of validate_streamlines (RipCurrents_main/main.cpp:372-431)."""
import numpy as np


def _step_np(flow, x, y):
    f = np.float32
    xi, yi = int(np.floor(x)), int(np.floor(y))
    h, w, _ = flow.shape
    if xi < 1 or yi < 1 or xi + 2 > w or yi + 2 > h:
        return None
    xr, yr = f(x - f(xi)), f(y - f(yi))
    ax, ay = f(f(1) - xr), f(f(1) - yr)
    t = [f(f(flow[yi, xi] * ax) * ay), f(f(flow[yi, xi + 1] * xr) * ay),
         f(f(flow[yi + 1, xi] * ax) * yr), f(f(flow[yi + 1, xi + 1] * xr) * yr)]
    return f(f(f(t[0] + t[1]) + t[2]) + t[3])


def test_variants_against_numpy(oracle):
    rng = np.random.default_rng(2)
    h, w = 48, 64
    flow = rng.normal(0, 1.5, (h, w, 2)).astype(np.float32)
    seeds0 = (rng.random((300, 2)) * [w + 4, h + 4] - 2).astype(np.float32)
    f = np.float32
    for variant, dt, it, upper in [(oracle.ADV_PATHLINE, 2.0, 3, 0.0), (oracle.ADV_LEGACY, 2.0, 1, 2.5),
                                   (oracle.ADV_MODULE, 0.1, 20, 2.5), (oracle.ADV_CUT5, 0.3, 7, 0.0),
                                   (oracle.ADV_FIXED100, 0.0, 0, 0.0)]:
        seeds = seeds0.copy()
        oracle.advect(flow, seeds, dt, it, upper, variant)
        nit = 100 if variant == oracle.ADV_FIXED100 else it
        for s in range(len(seeds0)):
            p = seeds0[s].copy()
            for _ in range(nit):
                d = _step_np(flow, p[0], p[1])
                if d is None:
                    break
                r = np.sqrt(f(f(d[0] * d[0]) + f(d[1] * d[1])))
                if variant in (oracle.ADV_LEGACY, oracle.ADV_MODULE) and r > f(upper):
                    break
                if variant == oracle.ADV_CUT5 and r > 5:
                    break
                if variant in (oracle.ADV_PATHLINE, oracle.ADV_LEGACY):
                    p = f(p + f(f(d * f(dt)) / f(it)))
                elif variant == oracle.ADV_FIXED100:
                    p = f(p + (d.astype(np.float64) * 0.1).astype(np.float32))
                else:
                    p = f(p + f(d * f(dt)))
            assert np.array_equal(p, seeds[s]), (variant, s)


def test_field_and_get_delta(oracle):
    rng = np.random.default_rng(4)
    h, w = 20, 30
    flow = rng.normal(0, 1.0, (h, w, 2)).astype(np.float32)
    disp = np.zeros((h * w, 2), np.float32); dist = np.zeros(h * w, np.float32)
    oracle.advect(flow, disp, 2.0, 1, 1.2, oracle.ADV_FIELD, dist=dist)
    mag = np.sqrt(flow[..., 0] * flow[..., 0] + flow[..., 1] * flow[..., 1]).ravel()
    inner = np.zeros((h, w), bool); inner[1:h - 1, 1:w - 1] = True
    moved = inner.ravel() & (mag <= np.float32(1.2))
    assert np.array_equal(disp[moved], (flow.reshape(-1, 2)[moved] * np.float32(2.0)) / np.float32(1))
    assert np.all(disp[~moved] == 0) and np.array_equal(dist[moved], mag[moved]) and np.all(dist[~moved] == 0)
    d2 = np.zeros((h * w, 2), np.float32)
    oracle.advect(flow, d2, 2.0, 5, 1.2, oracle.ADV_GET_DELTA)
    assert np.array_equal(d2, disp)


def test_rotation_field_invariant(oracle):
    # main.cpp:375-380: flow(row,col) = (-(row-240)/480*100, (col-320)/640*100); an Euler step multiplies
    # b(x-cx)^2 + a(y-cy)^2 by exactly (1 + a b dt^2), a = 100/480, b = 100/640.
    h, w = 480, 640
    rr, cc = np.mgrid[0:h, 0:w].astype(np.float64)
    flow = np.stack([-(rr - 240) / 480 * 100, (cc - 320) / 640 * 100], -1).astype(np.float32)
    a, b, dt = 100 / 480, 100 / 640, 0.03
    p = np.array([[200.0, 200.0]], np.float32)
    q0 = b * (200 - 320) ** 2 + a * (200 - 240) ** 2
    steps = 400
    oracle.advect(flow, p, dt, steps, 1e9, oracle.ADV_MODULE)
    q1 = b * (p[0, 0] - 320) ** 2 + a * (p[0, 1] - 240) ** 2
    assert abs(q1 / q0 - (1 + a * b * dt * dt) ** steps) < 1e-3


def test_streakline_lifecycle(oracle):
    h, w = 40, 60
    flow = np.zeros((h, w, 2), np.float32); flow[..., 0] = 1.5; flow[..., 1] = 0.25
    em = np.array([[10.0, 10.0], [30.5, 20.5]], np.float32)
    cap = 8
    verts = np.zeros((2, cap, 2), np.float32); cnt = np.ones(2, np.int32)
    verts[:, 0] = em                                   # Streakline ctor: first vertex = generation point
    for t in range(5):
        oracle.streakline_step(flow, em, verts, cnt)
    assert list(cnt) == [6, 6]
    for e in range(2):
        for i in range(6):                             # index 0 newest: vertex i has been moved i times
            assert np.allclose(verts[e, i], em[e] + i * np.array([1.5, 0.25]), atol=1e-5)
    # big jump (> 10 % of the frame) is rejected, vertex stays
    flow[...] = 0; flow[..., 0] = 7.0                 # 7 > 0.1*60
    before = verts.copy()
    oracle.streakline_step(flow, em, verts, cnt)
    assert np.array_equal(verts[:, 1:7], before[:, 0:6])
