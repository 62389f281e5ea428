"""ctypes front-end of the CPU oracle (oracle/*.c) -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs; the product
package (ripcurrents_b200) never imports this module.

Function <-> reference map (details in the C headers of each file):
  farneback()           cv::calcOpticalFlowFarneback as called at RipCurrents_main/ripcurrents.cpp:215,
                        main.cpp:264,609,742,961,1119,1481 (algorithm: SURVEY.md Appendix A)
  histogram()/thresholds()/classify_accumulate()   ripcurrents.cpp:319-439
  window_update()       main.cpp:1143-1153
  advect()/streakline_step()   pathlines.cpp:9-46, ripcurrents_module.cpp:486-679, Streakline.cpp:22-48
  subtract_mean_magnitude()/vector_to_color()/shear_to_color()/hsv2bgr()   ripcurrents_module.cpp:900-1138
  field_magnitude()/divide()/fmax()/normalize_jet()/positions()   ripcurrents.cpp:231-279, ripcurrents_module.cpp:13-59
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librc_oracle.so")

HIST_BINS, HIST_DIRECTIONS, HIST_RESOLUTION, HIST_ROWS = 50, 36, 20, 37
FARNEBACK_GAUSSIAN = 256

ADV_PATHLINE, ADV_LEGACY, ADV_MODULE, ADV_CUT5, ADV_FIXED100, ADV_FIELD, ADV_GET_DELTA = range(7)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("farneback_oracle.c", "aggregate_oracle.c", "advect_oracle.c", "ingest_oracle.c", "fields_oracle.c", "diag_oracle.c")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs if os.path.exists(s))):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.rc_oracle_farneback.restype = C.c_int
        _lib.rc_oracle_layers.restype = C.c_int
        _lib.rc_oracle_pyr_layer.restype = C.c_int
        _lib.rc_oracle_max.restype = C.c_double
        _lib.rc_oracle_subtract_mean_magnitude.restype = C.c_float
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def layers(w, h, pyr_scale, levels):
    lw = (C.c_int * 16)()
    lh = (C.c_int * 16)()
    n = lib().rc_oracle_layers(C.c_int(w), C.c_int(h), C.c_double(pyr_scale), C.c_int(levels), lw, lh)
    return [(lw[i], lh[i]) for i in range(n)]


def farneback(prev, nxt, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags, gauss_det_mode=None):
    prev = np.ascontiguousarray(prev, np.uint8)
    nxt = np.ascontiguousarray(nxt, np.uint8)
    h, w = prev.shape
    flow = np.empty((h, w, 2), np.float32)
    if gauss_det_mode is not None:
        lib().rc_oracle_set_gauss_det_mode(C.c_int(gauss_det_mode))
    rc = lib().rc_oracle_farneback(_p(prev), C.c_size_t(w), _p(nxt), C.c_size_t(w), C.c_int(w), C.c_int(h), _p(flow),
                                   C.c_double(pyr_scale), C.c_int(levels), C.c_int(winsize), C.c_int(iterations),
                                   C.c_int(poly_n), C.c_double(poly_sigma), C.c_int(flags))
    if rc != 0:
        raise RuntimeError("rc_oracle_farneback failed: %d" % rc)
    return flow


def pyr_layer(img, pyr_scale, k, lw, lh):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((lh, lw), np.float32)
    lib().rc_oracle_pyr_layer(_p(img), C.c_int(w), C.c_int(h), C.c_size_t(w), C.c_double(pyr_scale), C.c_int(k),
                              C.c_int(lw), C.c_int(lh), _p(out))
    return out


def polyexp(I, n, sigma):
    I = np.ascontiguousarray(I, np.float32)
    h, w = I.shape
    out = np.empty((h, w, 5), np.float32)
    lib().rc_oracle_polyexp(_p(I), C.c_int(w), C.c_int(h), C.c_int(n), C.c_double(sigma), _p(out))
    return out


def poly_kernels(n, sigma):
    g = np.empty(n + 1, np.float32); xg = np.empty(n + 1, np.float32); xxg = np.empty(n + 1, np.float32)
    ig = np.empty(4, np.float64)
    lib().rc_oracle_poly_kernels(C.c_int(n), C.c_double(sigma), _p(g), _p(xg), _p(xxg), _p(ig))
    return g, xg, xxg, ig


def update_matrices(R0, R1, flow):
    R0 = np.ascontiguousarray(R0, np.float32); R1 = np.ascontiguousarray(R1, np.float32)
    flow = np.ascontiguousarray(flow, np.float32)
    h, w, _ = R0.shape
    M = np.empty((h, w, 5), np.float32)
    lib().rc_oracle_update_matrices(_p(R0), _p(R1), _p(flow), C.c_int(w), C.c_int(h), _p(M))
    return M


def update_flow(M, winsize, gaussian, det_mode=0):
    M = np.ascontiguousarray(M, np.float32)
    h, w, _ = M.shape
    flow = np.empty((h, w, 2), np.float32)
    if gaussian:
        lib().rc_oracle_update_flow_gauss(_p(M), C.c_int(w), C.c_int(h), C.c_int(winsize), C.c_int(det_mode), _p(flow))
    else:
        lib().rc_oracle_update_flow_box(_p(M), C.c_int(w), C.c_int(h), C.c_int(winsize), _p(flow))
    return flow


def upsample_flow(coarse, fw, fh, pyr_scale):
    coarse = np.ascontiguousarray(coarse, np.float32)
    ch, cw, _ = coarse.shape
    fine = np.empty((fh, fw, 2), np.float32)
    lib().rc_oracle_upsample_flow(_p(coarse), C.c_int(cw), C.c_int(ch), _p(fine), C.c_int(fw), C.c_int(fh),
                                  C.c_double(pyr_scale))
    return fine


def cart_to_polar(x, y):
    x = np.ascontiguousarray(x, np.float32).ravel(); y = np.ascontiguousarray(y, np.float32).ravel()
    mag = np.empty_like(x); ang = np.empty_like(x)
    lib().rc_oracle_cart_to_polar(_p(x), _p(y), C.c_size_t(x.size), _p(mag), _p(ang))
    return mag, ang


class HistState:
    """Cumulative counters of ripcurrents.cpp:147-153 (int64; 37 direction rows, see aggregate_oracle.c)."""

    def __init__(self):
        self.hist = np.zeros(HIST_BINS, np.int64)
        self.histsum = np.zeros(1, np.int64)
        self.hist2d = np.zeros((HIST_ROWS, HIST_BINS), np.int64)
        self.histsum2d = np.zeros(HIST_ROWS, np.int64)


def histogram(flow, st):
    flow = np.ascontiguousarray(flow, np.float32)
    lib().rc_oracle_histogram(_p(flow), C.c_size_t(flow.size // 2), _p(st.hist), _p(st.histsum), _p(st.hist2d),
                              _p(st.histsum2d))


def thresholds(st):
    upper = np.zeros(1, np.float32); upper2d = np.zeros(HIST_DIRECTIONS, np.float32)
    prop = np.zeros(HIST_DIRECTIONS, np.float32)
    lib().rc_oracle_thresholds(_p(st.hist), C.c_int64(int(st.histsum[0])), _p(st.hist2d), _p(st.histsum2d),
                               _p(upper), _p(upper2d), _p(prop))
    return float(upper[0]), upper2d, prop


def classify_accumulate(flow, upper, framecount, acc_x, mid=0.5, lower=0.2):
    flow = np.ascontiguousarray(flow, np.float32)
    n = flow.size // 2
    outmask = np.empty(n, np.uint8); waveclass = np.empty(n, np.uint8); waterclass = np.empty(n, np.uint8)
    assert acc_x.dtype == np.float32 and acc_x.size == n and acc_x.flags.c_contiguous
    lib().rc_oracle_classify_accumulate(_p(flow), C.c_size_t(n), C.c_float(upper), C.c_float(mid), C.c_float(lower),
                                        C.c_int(framecount), _p(acc_x), _p(outmask), _p(waveclass), _p(waterclass))
    shp = flow.shape[:2]
    return outmask.reshape(shp), waveclass.reshape(shp), waterclass.reshape(shp)


def window_update(avg, slot, flow, W):
    assert avg.dtype == slot.dtype == np.float32 and avg.flags.c_contiguous and slot.flags.c_contiguous
    flow = np.ascontiguousarray(flow, np.float32)
    lib().rc_oracle_window_update(_p(avg), _p(slot), _p(flow), C.c_size_t(flow.size), C.c_int(W))


def subtract_mean(flow):
    assert flow.dtype == np.float32 and flow.flags.c_contiguous
    mean = np.zeros(2, np.float64)
    lib().rc_oracle_subtract_mean(_p(flow), C.c_size_t(flow.size // 2), _p(mean))
    return mean


def advect(flow, seeds, dt, iterations, upper, variant, dist=None, home=None):
    flow = np.ascontiguousarray(flow, np.float32)
    h, w, _ = flow.shape
    assert seeds.dtype == np.float32 and seeds.flags.c_contiguous
    if home is not None:
        home = np.ascontiguousarray(home, np.int32)
    lib().rc_oracle_advect(_p(flow), C.c_int(w), C.c_int(h), _p(seeds), C.c_size_t(seeds.size // 2), C.c_float(dt),
                           C.c_int(iterations), C.c_float(upper), C.c_int(variant),
                           _p(dist) if dist is not None else None, _p(home) if home is not None else None)


def streakline_step(flow, emitters, vertices, count, dt=1.0):
    flow = np.ascontiguousarray(flow, np.float32)
    h, w, _ = flow.shape
    E, cap, _ = vertices.shape
    assert vertices.dtype == np.float32 and count.dtype == np.int32 and emitters.dtype == np.float32
    lib().rc_oracle_streakline_step(_p(flow), C.c_int(w), C.c_int(h), _p(emitters), C.c_int(E), _p(vertices),
                                    _p(count), C.c_int(cap), C.c_float(dt))


def ingest_bgr(bgr, dw, dh, legacy14=False):
    """resize(INTER_LINEAR) + cvtColor(BGR2GRAY) of ripcurrents.cpp:209-210 (oracle/ingest_oracle.c)."""
    bgr = np.ascontiguousarray(bgr, np.uint8)
    sh, sw, _ = bgr.shape
    out = np.empty((dh, dw), np.uint8)
    lib().rc_oracle_ingest_bgr(_p(bgr), C.c_size_t(sw * 3), C.c_int(sw), C.c_int(sh), _p(out), C.c_int(dw), C.c_int(dh),
                               C.c_int(1 if legacy14 else 0))
    return out


def ingest_bgr_area(bgr, dw, dh, legacy14=False):
    """resize(INTER_AREA) + cvtColor(BGR2GRAY) of the PRIMING frame (ripcurrents.cpp:186-187); downscaling only."""
    bgr = np.ascontiguousarray(bgr, np.uint8)
    sh, sw, _ = bgr.shape
    out = np.empty((dh, dw), np.uint8)
    rc = lib().rc_oracle_ingest_bgr_area(_p(bgr), C.c_size_t(sw * 3), C.c_int(sw), C.c_int(sh), _p(out), C.c_int(dw), C.c_int(dh),
                                         C.c_int(1 if legacy14 else 0))
    if rc != 0:
        raise ValueError("INTER_AREA ingest is restated for downscaling only")
    return out


def edges(mask):
    """create_edges (ripcurrents_module.cpp:216-220): 5x5 elliptical dilate + morphological gradient."""
    mask = np.ascontiguousarray(mask, np.uint8)
    h, w = mask.shape
    out = np.empty((h, w), np.uint8)
    lib().rc_oracle_edges(_p(mask), C.c_int(w), C.c_int(h), _p(out))
    return out


def jet_lut():
    lut = np.empty((256, 3), np.uint8)
    lib().rc_oracle_jet_lut(_p(lut))
    return lut


def field_magnitude(field):
    """split + magnitude of ripcurrents.cpp:232-233 on a (h, w, 2) displacement field."""
    field = np.ascontiguousarray(field, np.float32)
    out = np.empty(field.shape[:-1], np.float32)
    lib().rc_oracle_field_magnitude(_p(field), C.c_size_t(out.size), _p(out))
    return out


def divide(a, b, div0_zero=False):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    out = np.empty_like(a)
    lib().rc_oracle_divide(_p(a), _p(b), C.c_size_t(a.size), C.c_int(1 if div0_zero else 0), _p(out))
    return out


def fmax(src):
    """minMaxLoc(src, NULL, &max) (NaNs ignored, see fields_oracle.c)."""
    src = np.ascontiguousarray(src, np.float32)
    return float(lib().rc_oracle_max(_p(src), C.c_size_t(src.size)))


def normalize_jet(src, maxval=None):
    """module:13-29: convertTo(CV_8UC1, 255/max) + applyColorMap(JET) -> (max, gray, bgr)."""
    src = np.ascontiguousarray(src, np.float32)
    if maxval is None:
        maxval = fmax(src)
    gray = np.empty(src.shape, np.uint8); bgr = np.empty(src.shape + (3,), np.uint8)
    with np.errstate(all="ignore"):
        lib().rc_oracle_normalize_jet(_p(src), C.c_size_t(src.size), C.c_double(maxval), _p(gray), _p(bgr))
    return maxval, gray, bgr


def positions(field, density=None):
    """streamline_positions (module:44-59); density=None mirrors the Mat::zeros of ripcurrents.cpp:261."""
    field = np.ascontiguousarray(field, np.float32)
    h, w, _ = field.shape
    zero = density is None
    if zero:
        density = np.empty((h, w, 3), np.float32)
    assert density.dtype == np.float32 and density.flags.c_contiguous
    lib().rc_oracle_positions(_p(field), C.c_int(w), C.c_int(h), _p(density), C.c_int(1 if zero else 0))
    return density


def hsv2bgr(hsv, fma=True):
    """cvtColor(COLOR_HSV2BGR) on 8-bit pixels (cv2's block path, see diag_oracle.c)."""
    hsv = np.ascontiguousarray(hsv, np.uint8)
    out = np.empty_like(hsv)
    lib().rc_oracle_hsv2bgr(_p(hsv), C.c_size_t(hsv.size // 3), _p(out), C.c_int(1 if fma else 0))
    return out


def subtract_mean_magnitude(flow):
    """subtructMeanMagnitude (module:900-1015), in place; returns meanval."""
    assert flow.dtype == np.float32 and flow.flags.c_contiguous
    return float(lib().rc_oracle_subtract_mean_magnitude(_p(flow), C.c_size_t(flow.size // 2)))


def vector_to_color(flow, max_displacement, fma=True):
    """vectorToColor (module:1017-1057) -> (hsv, bgr, new max_displacement)."""
    flow = np.ascontiguousarray(flow, np.float32)
    h, w, _ = flow.shape
    hsv = np.empty((h, w, 3), np.uint8); bgr = np.empty((h, w, 3), np.uint8)
    m = C.c_float(max_displacement)
    lib().rc_oracle_vector_to_color(_p(flow), C.c_int(w), C.c_int(h), _p(hsv), _p(bgr), C.byref(m), C.c_int(1 if fma else 0))
    return hsv, bgr, m.value


def shear_to_color(flow, img, max_frobenius, fma=True):
    """shearRateToColor (module:1059-1138); img (h,w,3) u8 is updated in place -> new max_frobeniusNorm."""
    flow = np.ascontiguousarray(flow, np.float32)
    h, w, _ = flow.shape
    assert img.dtype == np.uint8 and img.flags.c_contiguous and img.shape == (h, w, 3)
    m = C.c_float(max_frobenius)
    lib().rc_oracle_shear_to_color(_p(flow), C.c_int(w), C.c_int(h), _p(img), C.byref(m), C.c_int(1 if fma else 0))
    return m.value
