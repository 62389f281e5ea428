"""GPU parity of the derived particle fields (ripcurrents.cpp:231-279 ==
ripcurrents_module.cpp:13-59): bit-exact against the oracle, the committed cv2 fixture and live cv2."""
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


def particle_state(h, w, seed, still_rows=3):
    """A displacement / path-length pair as streamline_field leaves them: |disp| <= dist, some particles never moved."""
    rng = np.random.default_rng(seed)
    field = (rng.standard_normal((h, w, 2)) * 6).astype(np.float32)
    field[:still_rows] = 0
    dist = (np.hypot(field[..., 0], field[..., 1]) + np.abs(rng.standard_normal((h, w)) * 5)).astype(np.float32)
    dist[:still_rows] = 0
    return field, dist


def oracle_fields(oracle, field, dist, div0_zero=False):
    sf = oracle.field_magnitude(field)
    with np.errstate(all="ignore"):
        ratio = oracle.divide(sf, dist, div0_zero)
        out = {"streamfield": sf, "ratio": ratio}
        for name, src in (("disp", sf), ("motion", dist), ("ratio", ratio)):
            out[name + "_max"], out[name + "_gray"], out[name + "_bgr"] = oracle.normalize_jet(src)
    out["density"] = oracle.positions(field)
    return out


def check(got, ref):
    assert np.array_equal(got["streamfield"], ref["streamfield"], equal_nan=True)
    for k, name in enumerate(("disp", "motion", "ratio")):
        assert got["max"][k] == ref[name + "_max"] or (np.isnan(got["max"][k]) and np.isnan(ref[name + "_max"])), name
        assert np.array_equal(got[name + "_bgr"], ref[name + "_bgr"]), name
    assert np.array_equal(got["density"], ref["density"])


def test_fields_golden(ctx):
    z = np.load(os.path.join(GOLDEN, "fields.npz"))
    got = ctx.particle_fields(z["field"], z["dist"])
    assert np.array_equal(got["streamfield"], z["streamfield"])
    for k, name in enumerate(("disp", "motion", "ratio")):
        assert got["max"][k] == float(z[name + "_max"])
        assert np.array_equal(got[name + "_bgr"], z[name + "_bgr"]), name
    for name, src in (("disp", z["streamfield"]), ("motion", z["dist"])):
        mx, gray, bgr = ctx.normalize_jet(src)
        assert mx == float(z[name + "_max"]) and np.array_equal(gray, z[name + "_gray"]) and np.array_equal(bgr, z[name + "_bgr"])
    mx, ratio, gray, bgr = ctx.ratio_jet(z["streamfield"], z["dist"])
    assert np.array_equal(ratio, z["ratio"]) and np.array_equal(gray, z["ratio_gray"]) and np.array_equal(bgr, z["ratio_bgr"])
    # every level of the colour map
    ramp = np.arange(256, dtype=np.float32)
    assert np.array_equal(ctx.normalize_jet(ramp)[2], z["jet_lut"])


@pytest.mark.parametrize("shape", [(1080, 1920), (97, 131), (5, 7), (1, 1)])
@pytest.mark.parametrize("flags", [0, 1])
def test_fields_vs_oracle(ctx, oracle, shape, flags):
    field, dist = particle_state(*shape, seed=shape[0] + flags, still_rows=min(3, shape[0] - 1))
    if shape[0] > 4:
        field[4, 1] = (np.nan, 1); field[4, 2] = (np.inf, 0); field[4, 3] = (-np.inf, 0); field[4, 4] = (1e20, -1e20)
    check(ctx.particle_fields(field, dist, flags=flags), oracle_fields(oracle, field, dist, div0_zero=bool(flags)))


def test_fields_device_pointers_and_unaligned(ctx, oracle):
    import ctypes as C
    import torch
    h, w = 270, 481                                               # odd pixel count: vector body + scalar tail
    field, dist = particle_state(h, w, seed=5)
    ref = oracle_fields(oracle, field, dist)
    dev = torch.device("cuda", 0)
    n = h * w
    for off in (0, 1):                                            # off = 1: 4-byte aligned only -> scalar kernels
        fbuf = torch.zeros(2 * n + 8, device=dev); dbuf = torch.zeros(n + 8, device=dev)
        fbuf[2 * off:2 * off + 2 * n] = torch.from_numpy(field.ravel()).to(dev); dbuf[off:off + n] = torch.from_numpy(dist.ravel()).to(dev)
        sf = torch.empty(n + 8, device=dev); imgs = [torch.empty(3 * n + 16, dtype=torch.uint8, device=dev) for _ in range(3)]
        dens = torch.full((3 * n,), 7.0, device=dev)
        mx = (C.c_double * 3)()
        p = lambda t, o=0: C.c_void_p(t.data_ptr() + o)
        ctx._chk(ctx.lib.rc_particle_fields(ctx.h, p(fbuf, 8 * off), p(dbuf, 4 * off), C.c_int(w), C.c_int(h), C.c_int(0),
                                            p(sf, 4 * off), p(imgs[0], off), p(imgs[1], off), p(imgs[2], off), p(dens), mx))
        ctx.synchronize()
        assert np.array_equal(sf[off:off + n].cpu().numpy().reshape(h, w), ref["streamfield"])
        for k, name in enumerate(("disp", "motion", "ratio")):
            assert mx[k] == ref[name + "_max"]
            assert np.array_equal(imgs[k][off:off + 3 * n].cpu().numpy().reshape(h, w, 3), ref[name + "_bgr"]), (off, name)
        assert np.array_equal(dens.cpu().numpy().reshape(h, w, 3), ref["density"])


def test_pieces(ctx, oracle):
    field, dist = particle_state(120, 160, seed=8)
    sf = ctx.field_magnitude(field)
    assert np.array_equal(sf, oracle.field_magnitude(field))
    keep = np.full((120, 160, 3), 0.25, np.float32)
    ref = oracle.positions(field, keep.copy())
    assert np.array_equal(ctx.streamline_positions(field, keep), ref) and (ref == 0.25).any() and (ref == 1).any()
    assert np.array_equal(ctx.streamline_positions(field), oracle.positions(field))
    neg = -np.abs(sf) - 1                                              # all-negative input: max < 0, everything saturates to 0
    mx, gray, _ = ctx.normalize_jet(neg)
    assert mx == oracle.fmax(neg) and np.array_equal(gray, oracle.normalize_jet(neg)[1])
    mx, gray, bgr = ctx.normalize_jet(np.zeros((4, 9), np.float32))     # max 0 -> alpha inf -> 0 * inf = NaN -> 0
    assert mx == 0.0 and not gray.any() and np.array_equal(bgr, oracle.normalize_jet(np.zeros((4, 9), np.float32))[2])
    assert np.isnan(ctx.normalize_jet(np.full(16, np.nan, np.float32))[0])


def test_fields_live_cv2(ctx):
    cv2 = pytest.importorskip("cv2")
    field, dist = particle_state(480, 640, seed=2, still_rows=0)
    dist += 0.5                                                        # NaN-free: cv2's maxima are defined
    got = ctx.particle_fields(field, dist)
    cv2.setUseOptimized(False)
    try:
        sf = cv2.magnitude(field[..., 0].copy(), field[..., 1].copy())
        assert np.array_equal(got["streamfield"], sf)
        for k, src in enumerate((sf, dist, cv2.divide(sf, dist))):
            mx = cv2.minMaxLoc(src)[1]
            ref = cv2.applyColorMap(cv2.convertScaleAbs(src, alpha=255 / mx), cv2.COLORMAP_JET)
            assert got["max"][k] == mx
            assert np.array_equal(got[("disp_bgr", "motion_bgr", "ratio_bgr")[k]], ref), k
    finally:
        cv2.setUseOptimized(True)
