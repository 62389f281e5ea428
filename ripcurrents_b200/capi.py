"""ctypes binding of the C ABI (include/ripcurrents_b200.h).  Used by the tests and bench.py; the same symbols
are what a cgo / JNI / C++ binding would call.  There is no CPU fallback: if the shared
library is missing or no CUDA device is present, construction fails loudly."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("RC_B200_LIB") or os.path.join(_HERE, "lib", "libripcurrents_b200.so")   # env: A/B builds

HIST_BINS, HIST_DIRECTIONS, HIST_RESOLUTION, HIST_ROWS = 50, 36, 20, 37
FARNEBACK_GAUSSIAN = 256
FARNEBACK_STRICT = 0x10000
ADV_PATHLINE, ADV_LEGACY, ADV_MODULE, ADV_CUT5, ADV_FIXED100, ADV_FIELD, ADV_GET_DELTA = range(7)

# every symbol include/ripcurrents_b200.h declares (tests/test_abi.py checks header <-> library <-> this list)
SYMBOLS = [
    "rc_version", "rc_error_string", "rc_last_error", "rc_kernel_launches", "rc_create", "rc_destroy",
    "rc_set_stream", "rc_synchronize", "rc_profile_enable", "rc_profile_reset", "rc_profile_count", "rc_profile_get",
    "rc_farneback", "rc_flow_configure", "rc_flow_push", "rc_flow_device", "rc_flow_configure_batch",
    "rc_flow_push_batch", "rc_flow_device_at", "rc_process_frames", "rc_submit_frames", "rc_wait",
    "rc_hist_reset", "rc_polar_hist", "rc_hist_get", "rc_hist_add", "rc_hist_device", "rc_cart_to_polar",
    "rc_thresholds", "rc_accumulator_reset", "rc_classify_accumulate", "rc_accumulator_get",
    "rc_accumulator_device", "rc_window_configure", "rc_window_update", "rc_window_get", "rc_window_device",
    "rc_subtract_mean", "rc_average_vector", "rc_batch_hist", "rc_aggregate_last", "rc_accumulator_mask", "rc_mask_edges", "rc_ingest_bgr", "rc_submit_frames_bgr", "rc_hist_from_polar", "rc_create_flow", "rc_create_accumulationbuffer", "rc_advect", "rc_streakline_step", "rc_process_frame",
    "rc_particle_fields", "rc_normalize_jet", "rc_ratio_jet", "rc_field_magnitude", "rc_streamline_positions",
    "rc_subtract_mean_magnitude", "rc_vector_to_color", "rc_shear_rate_to_color",
    "rc_comm_unique_id", "rc_comm_init", "rc_comm_attach", "rc_comm_destroy", "rc_allreduce_accumulators",
    "rc_shard_configure", "rc_shard_step", "rc_shard_report", "rc_shard_window_get", "rc_set_mask_format",
]


class FrameResult(C.Structure):
    _fields_ = [("produced", C.c_int), ("UPPER", C.c_float), ("UPPER2d", C.c_float * HIST_DIRECTIONS),
                ("prop_above_upper", C.c_float * HIST_DIRECTIONS), ("histsum", C.c_int64)]


class RcError(RuntimeError):
    pass


_lib = None


def _prefer_bundled_nccl():
    """The library dlopens libnccl.so.2 at first multi-GPU use.  Inside a Python process that may import torch LATER, the
    system libnccl must not be mapped first (same SONAME, older version: torch's own import would then fail on missing
    symbols), so point RC_NCCL_LIB at the NCCL wheel torch itself uses when it is installed."""
    if os.environ.get("RC_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            p = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(p):
                os.environ["RC_NCCL_LIB"] = p
    except Exception:
        pass


def load():
    """Loads the shared library (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    _prefer_bundled_nccl()
    if not os.path.exists(SO_PATH):
        raise RcError("%s is missing: run `python -m ripcurrents_b200.build` (there is no CPU fallback)" % SO_PATH)
    lib = C.CDLL(SO_PATH)
    lib.rc_error_string.restype = C.c_char_p
    lib.rc_last_error.restype = C.c_char_p
    lib.rc_last_error.argtypes = [C.c_void_p]
    lib.rc_kernel_launches.restype = C.c_int64
    lib.rc_kernel_launches.argtypes = [C.c_void_p]
    lib.rc_destroy.restype = None
    lib.rc_destroy.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _ptr(a):
    """numpy array -> host pointer; int -> raw (device) pointer; None -> NULL; objects with data_ptr() (torch)."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError("unsupported buffer type %r" % type(a))


class Context:
    """One rc_ctx: a (GPU, camera stream) pair."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.rc_create(C.byref(h), C.c_int(device))
        if rc != 0:
            raise RcError("rc_create(device=%d) failed: %s (a CUDA device is required; no CPU fallback)"
                          % (device, self.lib.rc_error_string(rc).decode()))
        self.h = h
        self.w = self.h_img = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.rc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise RcError("%s: %s" % (self.lib.rc_error_string(rc).decode(), self.lib.rc_last_error(self.h).decode()))
        return rc

    # -- context -------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._chk(self.lib.rc_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._chk(self.lib.rc_synchronize(self.h))

    @property
    def kernel_launches(self):
        return int(self.lib.rc_kernel_launches(self.h))

    def profile_enable(self, on=True):
        self._chk(self.lib.rc_profile_enable(self.h, C.c_int(1 if on else 0)))

    def profile_reset(self):
        self._chk(self.lib.rc_profile_reset(self.h))

    def profile_read(self):
        """-> {kernel class: dict(ms=total ms, launches=n, bytes=total algorithmic bytes)} for classes that ran"""
        out = {}
        for i in range(self.lib.rc_profile_count()):
            name = C.c_char_p(); ms = C.c_double(); n = C.c_int64(); b = C.c_double()
            self._chk(self.lib.rc_profile_get(self.h, C.c_int(i), C.byref(name), C.byref(ms), C.byref(n), C.byref(b)))
            if n.value:
                out[name.value.decode()] = dict(ms=ms.value, launches=int(n.value), bytes=b.value)
        return out

    # -- A1 ------------------------------------------------------------------------------------------
    def farneback(self, prev, nxt, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags, out=None):
        prev = np.ascontiguousarray(prev, np.uint8); nxt = np.ascontiguousarray(nxt, np.uint8)
        h, w = prev.shape
        assert nxt.shape == prev.shape
        flow = out if out is not None else np.empty((h, w, 2), np.float32)
        self._chk(self.lib.rc_farneback(self.h, _ptr(prev), C.c_size_t(prev.strides[0]), _ptr(nxt),
                                        C.c_size_t(nxt.strides[0]), C.c_int(w), C.c_int(h), _ptr(flow),
                                        C.c_size_t(w * 8), C.c_double(pyr_scale), C.c_int(levels), C.c_int(winsize),
                                        C.c_int(iterations), C.c_int(poly_n), C.c_double(poly_sigma), C.c_int(flags)))
        self.w, self.h_img = w, h
        return flow

    def flow_configure(self, w, h, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags):
        self._chk(self.lib.rc_flow_configure(self.h, C.c_int(w), C.c_int(h), C.c_double(pyr_scale), C.c_int(levels),
                                             C.c_int(winsize), C.c_int(iterations), C.c_int(poly_n),
                                             C.c_double(poly_sigma), C.c_int(flags)))
        self.w, self.h_img = w, h

    def flow_configure_batch(self, w, h, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags, max_batch):
        self._chk(self.lib.rc_flow_configure_batch(self.h, C.c_int(w), C.c_int(h), C.c_double(pyr_scale),
                                                   C.c_int(levels), C.c_int(winsize), C.c_int(iterations),
                                                   C.c_int(poly_n), C.c_double(poly_sigma), C.c_int(flags),
                                                   C.c_int(max_batch)))
        self.w, self.h_img = w, h

    def flow_push_batch(self, frames, count=None, flows=None):
        """frames: numpy (count,h,w) u8 / device pointer to dense frames.  Returns the number of flows produced."""
        if isinstance(frames, np.ndarray):
            assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.flags.c_contiguous
            count = frames.shape[0]
        n = self.w * self.h_img
        return self._chk(self.lib.rc_flow_push_batch(self.h, _ptr(frames), C.c_size_t(self.w), C.c_size_t(n),
                                                     C.c_int(count), _ptr(flows), C.c_size_t(self.w * 8),
                                                     C.c_size_t(n * 8)))

    def flow_device_at(self, back):
        p = C.c_void_p()
        self._chk(self.lib.rc_flow_device_at(self.h, C.c_int(back), C.byref(p)))
        return p.value

    def flow_host_at(self, back):
        out = np.empty((self.h_img, self.w, 2), np.float32)
        self.synchronize()
        _memcpy_d2h(out, self.flow_device_at(back))
        return out

    def flow_push(self, frame, step=None, flow=None):
        """frame: numpy u8 (host) / torch tensor / raw device pointer.  Returns 1 when a flow was produced."""
        if isinstance(frame, np.ndarray):
            assert frame.dtype == np.uint8 and frame.ndim == 2
            step = frame.strides[0]
        return self._chk(self.lib.rc_flow_push(self.h, _ptr(frame), C.c_size_t(step or self.w), _ptr(flow),
                                               C.c_size_t(self.w * 8)))

    def flow_device(self):
        p = C.c_void_p(); w = C.c_int(); h = C.c_int()
        self._chk(self.lib.rc_flow_device(self.h, C.byref(p), C.byref(w), C.byref(h)))
        return p.value, w.value, h.value

    def flow_host(self):
        p, w, h = self.flow_device()
        out = np.empty((h, w, 2), np.float32)
        self.synchronize()
        _memcpy_d2h(out, p)
        return out

    # -- A2/A3 ---------------------------------------------------------------------------------------
    def hist_reset(self):
        self._chk(self.lib.rc_hist_reset(self.h))

    def polar_hist(self, flow=None, w=0, h=0):
        if isinstance(flow, np.ndarray):
            flow = np.ascontiguousarray(flow, np.float32)
            h, w = flow.shape[:2]
        self._chk(self.lib.rc_polar_hist(self.h, _ptr(flow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h)))

    def hist_get(self):
        hist = np.zeros(HIST_BINS, np.int64); histsum = C.c_int64()
        hist2d = np.zeros((HIST_ROWS, HIST_BINS), np.int64); histsum2d = np.zeros(HIST_ROWS, np.int64)
        self._chk(self.lib.rc_hist_get(self.h, _ptr(hist), C.byref(histsum), _ptr(hist2d), _ptr(histsum2d)))
        return hist, int(histsum.value), hist2d, histsum2d

    def hist_add(self, hist2d):
        hist2d = np.ascontiguousarray(hist2d, np.int64)
        assert hist2d.size == HIST_ROWS * HIST_BINS
        self._chk(self.lib.rc_hist_add(self.h, _ptr(hist2d)))

    def hist_device(self):
        p = C.c_void_p()
        self._chk(self.lib.rc_hist_device(self.h, C.byref(p)))
        return p.value

    def cart_to_polar(self, flow):
        flow = np.ascontiguousarray(flow, np.float32).reshape(-1, 2)
        n = flow.shape[0]
        mag = np.empty(n, np.float32); ang = np.empty(n, np.float32)
        self._chk(self.lib.rc_cart_to_polar(self.h, _ptr(flow), C.c_size_t(n), _ptr(mag), _ptr(ang)))
        return mag, ang

    # -- A4 ------------------------------------------------------------------------------------------
    def thresholds(self):
        up = C.c_float(); up2 = np.zeros(HIST_DIRECTIONS, np.float32); prop = np.zeros(HIST_DIRECTIONS, np.float32)
        self._chk(self.lib.rc_thresholds(self.h, C.byref(up), _ptr(up2), _ptr(prop)))
        return float(up.value), up2, prop

    # -- A5 ------------------------------------------------------------------------------------------
    def accumulator_reset(self):
        self._chk(self.lib.rc_accumulator_reset(self.h))

    def classify_accumulate(self, flow, upper, framecount, want=("mask", "wave", "water"), w=0, h=0):
        if isinstance(flow, np.ndarray):
            flow = np.ascontiguousarray(flow, np.float32)
            h, w = flow.shape[:2]
        elif flow is None:
            w, h = self.w, self.h_img
        outs = {k: (np.empty((h, w), np.uint8) if k in want else None) for k in ("mask", "wave", "water")}
        self._chk(self.lib.rc_classify_accumulate(self.h, _ptr(flow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h),
                                                  C.c_float(upper), C.c_int(framecount), _ptr(outs["mask"]),
                                                  _ptr(outs["wave"]), _ptr(outs["water"])))
        return outs["mask"], outs["wave"], outs["water"]

    def accumulator_get(self, w, h):
        acc = np.empty((h, w), np.float32)
        self._chk(self.lib.rc_accumulator_get(self.h, _ptr(acc)))
        return acc

    def accumulator_device(self):
        p = C.c_void_p(); w = C.c_int(); h = C.c_int()
        self._chk(self.lib.rc_accumulator_device(self.h, C.byref(p), C.byref(w), C.byref(h)))
        return p.value, w.value, h.value

    # -- A6 ------------------------------------------------------------------------------------------
    def window_configure(self, w, h, W):
        self._chk(self.lib.rc_window_configure(self.h, C.c_int(w), C.c_int(h), C.c_int(W)))
        self._win = (w, h)

    def window_update(self, flow=None):
        if isinstance(flow, np.ndarray):
            flow = np.ascontiguousarray(flow, np.float32)
        self._chk(self.lib.rc_window_update(self.h, _ptr(flow), C.c_size_t(self._win[0] * 8)))

    def window_get(self):
        w, h = self._win
        avg = np.empty((h, w, 2), np.float32)
        self._chk(self.lib.rc_window_get(self.h, _ptr(avg), C.c_size_t(w * 8)))
        return avg

    def average_vector(self, old_slot, flow, average, frames=300, dt=2.0, upper=0.0, want_new=False):
        """averageVector's window update (module:392-400); `average` (h,w,2) f32 numpy is updated in place."""
        flow = np.ascontiguousarray(flow, np.float32)
        h, w, _ = flow.shape
        assert average.dtype == np.float32 and average.flags.c_contiguous and average.shape == (h, w, 2)
        if old_slot is not None:
            old_slot = np.ascontiguousarray(old_slot, np.float32)
        new = np.empty((h, w, 2), np.float32) if want_new else None
        self._chk(self.lib.rc_average_vector(self.h, _ptr(old_slot), _ptr(flow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h),
                                             _ptr(average), _ptr(new), C.c_int(frames), C.c_float(dt), C.c_float(upper)))
        return new

    def subtract_mean(self, flow):
        assert isinstance(flow, np.ndarray) and flow.dtype == np.float32 and flow.flags.c_contiguous
        h, w = flow.shape[:2]
        mean = np.zeros(2, np.float64)
        self._chk(self.lib.rc_subtract_mean(self.h, _ptr(flow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h), _ptr(mean)))
        return mean

    # -- A7 ------------------------------------------------------------------------------------------
    def advect(self, flow, seeds, dt, iterations, upper, variant, dist=None, home=None, w=0, h=0, n=None):
        if isinstance(flow, np.ndarray):
            flow = np.ascontiguousarray(flow, np.float32)
            h, w = flow.shape[:2]
        if isinstance(seeds, np.ndarray):
            assert seeds.dtype == np.float32 and seeds.flags.c_contiguous
            n = seeds.size // 2
        if isinstance(home, np.ndarray):
            home = np.ascontiguousarray(home, np.int32)
        self._chk(self.lib.rc_advect(self.h, _ptr(flow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h), _ptr(seeds),
                                     C.c_size_t(n), C.c_float(dt), C.c_int(iterations), C.c_float(upper),
                                     C.c_int(variant), _ptr(dist), _ptr(home)))

    def streakline_step(self, flow, emitters, vertices, count, dt=1.0, w=0, h=0, E=None, cap=None):
        if isinstance(flow, np.ndarray):
            flow = np.ascontiguousarray(flow, np.float32)
            h, w = flow.shape[:2]
        if isinstance(vertices, np.ndarray):
            E, cap = vertices.shape[:2]
            assert vertices.dtype == np.float32 and count.dtype == np.int32 and emitters.dtype == np.float32
        self._chk(self.lib.rc_streakline_step(self.h, _ptr(flow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h),
                                              _ptr(emitters), C.c_int(E), _ptr(vertices), _ptr(count), C.c_int(cap),
                                              C.c_float(dt)))

    # -- fused per-frame step ------------------------------------------------------------------------
    def process_frame(self, frame, framecount, outmask=None, want_result=True, step=None):
        if isinstance(frame, np.ndarray):
            step = frame.strides[0]
        res = FrameResult() if want_result else None
        rc = self._chk(self.lib.rc_process_frame(self.h, _ptr(frame), C.c_size_t(step or self.w), C.c_int(framecount),
                                                 _ptr(outmask), C.byref(res) if res is not None else None))
        return rc, res


    def process_frames(self, frames, framecount0, outmasks=None, want_results=True, count=None, submit_only=False,
                       results=None):
        """frames: numpy (count,h,w) u8 (host) / tensor / device pointer to dense frames.
        Returns (flows produced, ctypes array of FrameResult or None)."""
        if isinstance(frames, np.ndarray):
            assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.flags.c_contiguous
            count = frames.shape[0]
        n = self.w * self.h_img
        if results is None and want_results:
            results = (FrameResult * count)()
        fn = self.lib.rc_submit_frames if submit_only else self.lib.rc_process_frames
        mstride = n // 8 if getattr(self, "_mask_packed", False) else n
        rc = self._chk(fn(self.h, _ptr(frames), C.c_size_t(self.w), C.c_size_t(n), C.c_int(count),
                          C.c_int(framecount0), _ptr(outmasks), C.c_size_t(mstride),
                          C.byref(results) if results is not None else None))
        return rc, results

    def batch_hist(self, nb):
        deltas = np.zeros((nb, HIST_ROWS, HIST_BINS), np.int64)
        self._chk(self.lib.rc_batch_hist(self.h, C.c_int(nb), _ptr(deltas)))
        return deltas

    def aggregate_last(self, nb, framecount0, want_results=True):
        res = (FrameResult * nb)() if want_results else None
        self._chk(self.lib.rc_aggregate_last(self.h, C.c_int(nb), C.c_int(framecount0),
                                             C.byref(res) if res is not None else None))
        return res

    def accumulator_mask(self, framecount):
        mask = np.empty((self.h_img, self.w), np.uint8)
        self._chk(self.lib.rc_accumulator_mask(self.h, C.c_int(framecount), _ptr(mask)))
        return mask

    def mask_edges(self, masks):
        """masks: (h,w) or (count,h,w) u8 -> edges of the same shape (create_edges)."""
        masks = np.ascontiguousarray(masks, np.uint8)
        shp = masks.shape
        m3 = masks.reshape((-1,) + shp[-2:])
        count, h, w = m3.shape
        out = np.empty_like(m3)
        self._chk(self.lib.rc_mask_edges(self.h, _ptr(m3), C.c_size_t(w), C.c_size_t(w * h), C.c_int(w), C.c_int(h),
                                         C.c_int(count), _ptr(out), C.c_size_t(w), C.c_size_t(w * h)))
        return out.reshape(shp)

    # ---- derived particle fields (ripcurrents.cpp:231-279, module:13-59) ----
    def particle_fields(self, field, dist, flags=0, density=True):
        """field (h,w,2) f32 displacements, dist (h,w) f32 path lengths -> dict with streamfield, the three JET images,
        the position scatter and the maxima."""
        field = np.ascontiguousarray(field, np.float32); dist = np.ascontiguousarray(dist, np.float32)
        h, w, _ = field.shape
        out = {"streamfield": np.empty((h, w), np.float32), "disp_bgr": np.empty((h, w, 3), np.uint8),
               "motion_bgr": np.empty((h, w, 3), np.uint8), "ratio_bgr": np.empty((h, w, 3), np.uint8)}
        dens = np.empty((h, w, 3), np.float32) if density else None
        mx = (C.c_double * 3)()
        self._chk(self.lib.rc_particle_fields(self.h, _ptr(field), _ptr(dist), C.c_int(w), C.c_int(h), C.c_int(flags),
                                              _ptr(out["streamfield"]), _ptr(out["disp_bgr"]), _ptr(out["motion_bgr"]),
                                              _ptr(out["ratio_bgr"]), _ptr(dens) if density else None, mx))
        out["density"] = dens
        out["max"] = (mx[0], mx[1], mx[2])
        return out

    def normalize_jet(self, src):
        src = np.ascontiguousarray(src, np.float32)
        gray = np.empty(src.shape, np.uint8); bgr = np.empty(src.shape + (3,), np.uint8)
        mx = C.c_double()
        self._chk(self.lib.rc_normalize_jet(self.h, _ptr(src), C.c_size_t(src.size), _ptr(gray), _ptr(bgr), C.byref(mx)))
        return mx.value, gray, bgr

    def ratio_jet(self, a, b, flags=0):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        ratio = np.empty(a.shape, np.float32); gray = np.empty(a.shape, np.uint8); bgr = np.empty(a.shape + (3,), np.uint8)
        mx = C.c_double()
        self._chk(self.lib.rc_ratio_jet(self.h, _ptr(a), _ptr(b), C.c_size_t(a.size), C.c_int(flags), _ptr(ratio), _ptr(gray),
                                        _ptr(bgr), C.byref(mx)))
        return mx.value, ratio, gray, bgr

    def field_magnitude(self, field):
        field = np.ascontiguousarray(field, np.float32)
        mag = np.empty(field.shape[:-1], np.float32)
        self._chk(self.lib.rc_field_magnitude(self.h, _ptr(field), C.c_size_t(mag.size), _ptr(mag)))
        return mag

    def streamline_positions(self, field, density=None):
        field = np.ascontiguousarray(field, np.float32)
        h, w, _ = field.shape
        keep = density is not None
        if not keep:
            density = np.empty((h, w, 3), np.float32)
        self._chk(self.lib.rc_streamline_positions(self.h, _ptr(field), C.c_int(w), C.c_int(h), _ptr(density),
                                                   C.c_int(2 if keep else 0)))
        return density

    # ---- flow-derived diagnostics (ripcurrents_module.cpp:900-1138) ----
    def subtract_mean_magnitude(self, flow, flags=0):
        """In place on a (h,w,2) f32 numpy flow; returns meanval."""
        assert isinstance(flow, np.ndarray) and flow.dtype == np.float32 and flow.flags.c_contiguous
        h, w, _ = flow.shape
        mv = C.c_float()
        self._chk(self.lib.rc_subtract_mean_magnitude(self.h, _ptr(flow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h),
                                                      C.c_int(flags), C.byref(mv)))
        return mv.value

    def vector_to_color(self, flow, max_displacement, flags=0):
        """-> (bgr, new max_displacement); max_displacement=None keeps the state in the context."""
        flow = np.ascontiguousarray(flow, np.float32)
        h, w, _ = flow.shape
        bgr = np.empty((h, w, 3), np.uint8)
        m = C.c_float(max_displacement if max_displacement is not None else 0.0)
        self._chk(self.lib.rc_vector_to_color(self.h, _ptr(flow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h), _ptr(bgr),
                                              C.c_size_t(w * 3), C.byref(m) if max_displacement is not None else None,
                                              C.c_int(flags)))
        return bgr, (m.value if max_displacement is not None else None)

    def shear_rate_to_color(self, flow, img, max_frobenius, flags=0):
        """img (h,w,3) u8 numpy is updated in place -> new max_frobeniusNorm (None keeps the state in the context)."""
        flow = np.ascontiguousarray(flow, np.float32)
        h, w, _ = flow.shape
        assert img.dtype == np.uint8 and img.flags.c_contiguous and img.shape == (h, w, 3)
        m = C.c_float(max_frobenius if max_frobenius is not None else 0.0)
        self._chk(self.lib.rc_shear_rate_to_color(self.h, _ptr(flow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h), _ptr(img),
                                                  C.c_size_t(w * 3), C.byref(m) if max_frobenius is not None else None,
                                                  C.c_int(flags)))
        return m.value if max_frobenius is not None else None

    def ingest_bgr(self, bgr, dw, dh, flags=0):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        sh, sw, _ = bgr.shape
        gray = np.empty((dh, dw), np.uint8)
        self._chk(self.lib.rc_ingest_bgr(self.h, _ptr(bgr), C.c_size_t(sw * 3), C.c_int(sw), C.c_int(sh), _ptr(gray),
                                         C.c_size_t(dw), C.c_int(dw), C.c_int(dh), C.c_int(flags)))
        return gray

    def submit_frames_bgr(self, bgr_frames, framecount0, outmasks=None, results=None, ingest_flags=0):
        """bgr_frames: numpy (count, sh, sw, 3) u8.  Asynchronous like process_frames(submit_only=True)."""
        assert isinstance(bgr_frames, np.ndarray) and bgr_frames.dtype == np.uint8 and bgr_frames.ndim == 4
        count, sh, sw, _ = bgr_frames.shape
        n = self.w * self.h_img
        rc = self._chk(self.lib.rc_submit_frames_bgr(self.h, _ptr(bgr_frames), C.c_size_t(sw * 3), C.c_size_t(sw * sh * 3),
                                                     C.c_int(sw), C.c_int(sh), C.c_int(count), C.c_int(framecount0),
                                                     C.c_int(ingest_flags), _ptr(outmasks), C.c_size_t(n),
                                                     C.byref(results) if results is not None else None))
        return rc

    def wait(self):
        self._chk(self.lib.rc_wait(self.h))

    def set_mask_format(self, packed):
        self._chk(self.lib.rc_set_mask_format(self.h, C.c_int(1 if packed else 0)))
        self._mask_packed = bool(packed)

    # ---- multi-GPU (NCCL inside the library) ----
    def comm_init(self, uid, rank, nranks):
        assert len(uid) == 128
        self._chk(self.lib.rc_comm_init(self.h, C.c_char_p(bytes(uid)), C.c_int(rank), C.c_int(nranks)))

    def comm_destroy(self):
        self._chk(self.lib.rc_comm_destroy(self.h))

    def allreduce_accumulators(self, shared_acc=None, shared_hist=None):
        """device pointers (ints) or None = in place"""
        self._chk(self.lib.rc_allreduce_accumulators(self.h, None, _ptr(shared_acc), _ptr(shared_hist)))

    def shard_configure(self, window_W, owner):
        self._chk(self.lib.rc_shard_configure(self.h, C.c_int(window_W), C.c_int(owner)))

    def shard_step(self, frames, framecount0, pairs_per_rank, count=None, want_results=True):
        """frames: numpy (pairs+1, h, w) u8 / device pointer (with count) / None for a rank without pairs."""
        if isinstance(frames, np.ndarray):
            assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.flags.c_contiguous
            count = frames.shape[0]
        count = count or 0
        ppr = (C.c_int * len(pairs_per_rank))(*pairs_per_rank)
        nb = max(count - 1, 0)
        res = (FrameResult * nb)() if (want_results and nb) else None
        n = self.w * self.h_img
        rc = self._chk(self.lib.rc_shard_step(self.h, _ptr(frames), C.c_size_t(self.w), C.c_size_t(n), C.c_int(count),
                                              C.c_int(framecount0), ppr, C.byref(res) if res is not None else None))
        return rc, res

    def shard_report(self, framecount, want_mask=True, want_acc=True, want_hist=True):
        mask = np.empty((self.h_img, self.w), np.uint8) if want_mask else None
        acc = np.empty((self.h_img, self.w), np.float32) if want_acc else None
        hist = np.empty((HIST_ROWS, HIST_BINS), np.int64) if want_hist else None
        self._chk(self.lib.rc_shard_report(self.h, C.c_int(framecount), _ptr(mask), _ptr(acc), _ptr(hist)))
        return mask, acc, hist

    def shard_window_get(self):
        avg = np.empty((self.h_img, self.w, 2), np.float32)
        self._chk(self.lib.rc_shard_window_get(self.h, _ptr(avg)))
        return avg


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, every rank passes it to Context.comm_init)."""
    buf = C.create_string_buffer(128)
    rc = load().rc_comm_unique_id(buf)
    if rc != 0:
        raise RcError("rc_comm_unique_id failed: %s" % load().rc_error_string(rc).decode())
    return buf.raw


_cudart = None


def _memcpy_h2d(dev_ptr, src):
    """Test helper: numpy -> device copy."""
    _load_cudart()
    rc = _cudart.cudaMemcpy(C.c_void_p(dev_ptr), C.c_void_p(src.ctypes.data), C.c_size_t(src.nbytes), C.c_int(1))
    if rc != 0:
        raise RcError("cudaMemcpy H2D failed: %d" % rc)


def _memcpy_d2h(dst, dev_ptr):
    """Test helper: device -> numpy copy through the CUDA runtime (torch-free)."""
    _load_cudart()
    rc = _cudart.cudaMemcpy(C.c_void_p(dst.ctypes.data), C.c_void_p(dev_ptr), C.c_size_t(dst.nbytes), C.c_int(2))
    if rc != 0:
        raise RcError("cudaMemcpy D2H failed: %d" % rc)


def _load_cudart():
    global _cudart
    if _cudart is None:
        for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                _cudart = C.CDLL(name)
                break
            except OSError:
                continue
        if _cudart is None:
            raise RcError("libcudart not found")
