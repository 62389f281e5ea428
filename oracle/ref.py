"""ctypes front-end of oracle/_ref/librc_ref.so -- the REFERENCE'S OWN functions, compiled from /root/reference by
oracle/ref_build.py.  TEST INFRASTRUCTURE ONLY (tests/, bench.py's CPU legs); the product never imports this.

Same call shapes as oracle/oracle.py so that tests compare the C restatement with the compiled reference directly.
"""
import ctypes as C
import os

import numpy as np

from . import ref_build

HIST_BINS, HIST_DIRECTIONS, HIST_ROWS = 50, 36, 37
_lib = None


def available():
    return ref_build.build() is not None


def lib():
    global _lib
    if _lib is None:
        so = ref_build.build()
        if so is None:
            raise RuntimeError("oracle/_ref/librc_ref.so is absent and /root/reference is not present to build it")
        _lib = C.CDLL(so)
        _lib.rc_ref_legacy_new.restype = C.c_void_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def dims():
    return lib().rc_ref_xdim(), lib().rc_ref_ydim()


class HistState:
    """The reference's `int` counters (ripcurrents.cpp:147-153), over-allocated to 37 direction rows."""

    def __init__(self):
        self.hist = np.zeros(HIST_BINS, np.int32)
        self.histsum = np.zeros(1, np.int32)
        self.hist2d = np.zeros((HIST_ROWS, HIST_BINS), np.int32)
        self.histsum2d = np.zeros(HIST_ROWS, np.int32)


def create_histogram(polar3, st):
    """create_histogram (ripcurrents_module.cpp:89-144) on a (YDIM, w, 3) f32 polar image -> (UPPER, UPPER2d, prop)."""
    polar3 = np.ascontiguousarray(polar3, np.float32)
    h, w, _ = polar3.shape
    assert h == dims()[1], "the reference compiles YDIM into create_histogram's row loop"
    up = np.zeros(1, np.float32); up2 = np.zeros(HIST_DIRECTIONS, np.float32); prop = np.zeros(HIST_DIRECTIONS, np.float32)
    with np.errstate(all="ignore"):
        lib().rc_ref_create_histogram(_p(polar3), C.c_int(w), _p(st.hist), _p(st.histsum), _p(st.hist2d), _p(st.histsum2d),
                                      _p(up), _p(up2), _p(prop))
    return float(up[0]), up2, prop


def classify_accumulate(polar3, upper, framecount, acc_x, upper2d, mid=0.5, lower=0.2):
    """create_flow + create_accumulationbuffer (module:153-212) -> (outmask, waveclass, waterclass); acc_x in place."""
    polar3 = np.array(polar3, np.float32, order="C")       # create_flow rescales it for display
    h, w, _ = polar3.shape
    n = h * w
    assert acc_x.dtype == np.float32 and acc_x.size == n and acc_x.flags.c_contiguous
    upper2d = np.ascontiguousarray(upper2d, np.float32)
    mask = np.empty(n, np.uint8); wave = np.empty(n, np.uint8); water = np.empty(n, np.uint8)
    lib().rc_ref_classify_accumulate(_p(polar3), C.c_int(w), C.c_int(h), C.c_float(upper), C.c_float(mid), C.c_float(lower),
                                     _p(upper2d), C.c_int(framecount), _p(acc_x), _p(mask), _p(wave), _p(water))
    return mask.reshape(h, w), wave.reshape(h, w), water.reshape(h, w)


def advect(flow, seeds, dt, iterations, upper, variant, dist=None, home=None):
    flow = np.ascontiguousarray(flow, np.float32)
    h, w, _ = flow.shape
    assert seeds.dtype == np.float32 and seeds.flags.c_contiguous
    if home is not None:
        home = np.ascontiguousarray(home, np.int32)
    lib().rc_ref_advect(_p(flow), C.c_int(w), C.c_int(h), _p(seeds), C.c_size_t(seeds.size // 2), C.c_float(dt),
                        C.c_int(iterations), C.c_float(upper), C.c_int(variant),
                        _p(dist) if dist is not None else None, _p(home) if home is not None else None)


class Window:
    """The sliding-window block of main.cpp:1143-1153 with its state (:1084-1092)."""

    def __init__(self, w, h, W):
        self.w, self.h, self.W = w, h, W
        self.avg = np.zeros((h, w, 2), np.float32)
        self.ring = np.zeros((W, h, w, 2), np.float32)
        self.cur = np.zeros(1, np.int32)

    def update(self, flow):
        flow = np.ascontiguousarray(flow, np.float32)
        lib().rc_ref_window_update(_p(self.avg), _p(self.ring), _p(self.cur), _p(flow), C.c_int(self.w), C.c_int(self.h),
                                   C.c_int(self.W))


def average_vector(buffer_slot, current_flow, average, upper):
    """averageVector's window update (module:386-400); XDIM x YDIM; `average` in place."""
    buffer_slot = np.ascontiguousarray(buffer_slot, np.float32); current_flow = np.ascontiguousarray(current_flow, np.float32)
    assert average.dtype == np.float32 and average.flags.c_contiguous
    lib().rc_ref_average_vector(_p(buffer_slot), _p(current_flow), _p(average), C.c_float(upper))


def streakline_step(flow, emitters, vertices, count, dt=1.0):
    flow = np.ascontiguousarray(flow, np.float32)
    h, w, _ = flow.shape
    E, cap, _ = vertices.shape
    assert vertices.dtype == np.float32 and count.dtype == np.int32 and emitters.dtype == np.float32
    lib().rc_ref_streakline_step(_p(flow), C.c_int(w), C.c_int(h), _p(emitters), C.c_int(E), _p(vertices), _p(count),
                                 C.c_int(cap), C.c_float(dt))


class LegacyLoop:
    """ripcurrents.cpp's frame-loop aggregation (:305-439) with its cumulative state (:133-154), as written in main()."""

    def __init__(self):
        self.h = C.c_void_p(lib().rc_ref_legacy_new())
        self.w_img, self.h_img = dims()

    def frame(self, flow, framecount):
        flow = np.ascontiguousarray(flow, np.float32)
        assert flow.shape == (self.h_img, self.w_img, 2)
        n = self.w_img * self.h_img
        out = {"mask": np.empty((self.h_img, self.w_img), np.uint8), "UPPER": np.zeros(1, np.float32),
               "UPPER2d": np.zeros(HIST_DIRECTIONS, np.float32), "prop": np.zeros(HIST_DIRECTIONS, np.float32),
               "hist": np.zeros(HIST_BINS, np.int32), "histsum": np.zeros(1, np.int32),
               "hist2d": np.zeros((HIST_DIRECTIONS, HIST_BINS), np.int32), "histsum2d": np.zeros(HIST_DIRECTIONS, np.int32),
               "acc": np.empty(n, np.float32)}
        with np.errstate(all="ignore"):
            lib().rc_ref_legacy_frame(self.h, _p(flow), C.c_int(framecount), _p(out["mask"]), _p(out["UPPER"]),
                                      _p(out["UPPER2d"]), _p(out["prop"]), _p(out["hist"]), _p(out["histsum"]),
                                      _p(out["hist2d"]), _p(out["histsum2d"]), _p(out["acc"]))
        return out

    def close(self):
        if self.h:
            lib().rc_ref_legacy_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
