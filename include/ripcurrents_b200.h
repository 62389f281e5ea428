/*
 * ripcurrents_b200.h -- C ABI of the B200-native ripcurrents hot path.
 *
 * Drop-in boundary for the per-frame path of borgor/ripcurrents: dense Farneback optical flow,
 * the temporal aggregation of ripcurrents.cpp (polar conversion, cumulative speed/direction
 * histograms, tail thresholds, classification, accumulation), the sliding-window flow mean of
 * main.cpp, and the Euler/bilinear pathline / streakline particle advection.
 *
 * The reference has no FFI layer of its own for this path: it calls OpenCV and its own C++ free
 * functions directly.  Each entry point below names the reference call
 * (file:line under /root/reference/RipCurrents_main unless stated) that a binding replaces; the
 * header-compatible C++ wrappers (ripcurrents.hpp / Streakline.hpp / pathlines.h signatures) live in
 * ripcurrents_b200/cpp and are thin callers of this ABI.  INTEGRATION.md shows the wiring.
 *
 * Conventions
 *   - plain C types only; every function returns RC_OK (0) or a negative RC_ERR_* code and never
 *     throws; rc_last_error(ctx) gives a message for the last failure on that context.
 *   - image / flow / seed pointers may be HOST or DEVICE pointers (detected with
 *     cudaPointerGetAttributes); host buffers are staged through pinned memory on the context's stream.
 *   - *_step arguments are row strides in BYTES (cv::Mat::step).
 *   - flow is CV_32FC2: interleaved (dx,dy) fp32, row-major.
 *   - one rc_ctx per (GPU, camera stream); a context is not thread-safe.
 *   - all work is enqueued on the context's CUDA stream (rc_set_stream; default: a private
 *     non-blocking stream).  Calls that return data to HOST memory synchronise that stream; calls whose
 *     outputs are device pointers do not.
 *   - there is NO CPU fallback: without a CUDA device rc_create fails with RC_ERR_CUDA.
 */
#ifndef RIPCURRENTS_B200_H
#define RIPCURRENTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RC_OK 0
#define RC_ERR_INVALID (-1)      /* bad argument */
#define RC_ERR_CUDA (-2)         /* CUDA runtime error (message in rc_last_error) */
#define RC_ERR_NOMEM (-3)        /* allocation failed */
#define RC_ERR_STATE (-4)        /* call order (e.g. no flow computed yet) */
#define RC_ERR_UNSUPPORTED (-5)  /* parameter outside the supported range */

/* ripcurrents.hpp:7-9 */
#define RC_HIST_BINS 50
#define RC_HIST_DIRECTIONS 36
#define RC_HIST_RESOLUTION 20
/* hist2d / histsum2d carry one extra row: direction index 36 (angle == 360.0f exactly), which is an
 * out-of-bounds write in the reference (ripcurrents.cpp:326); it is counted here so totals stay exact. */
#define RC_HIST_ROWS 37

/* cv::OPTFLOW_FARNEBACK_GAUSSIAN */
#define RC_FARNEBACK_GAUSSIAN 256
/* Arithmetic selection (ORed into `flags`; bits unused by OpenCV).
 * default: fp32 FMA accumulation in the horizontal polynomial-expansion pass, Gaussian tails of the
 *          expansion kernel below 1e-12 of the centre weight dropped.  Measured within the tolerance of the
 *          task (mean EPE <= 1e-3 px, max <= 1e-2 px against cv2) with >10x margin, see DESIGN.md.
 * RC_FARNEBACK_STRICT: every tap, fp64 accumulators exactly where OpenCV uses them, no FMA contraction. */
#define RC_FARNEBACK_STRICT 0x10000

/* particle-advection variants (the reference's seven copies of the Euler/bilinear step) */
#define RC_ADV_PATHLINE 0   /* pathlines.cpp:9-46          pt += delta*dt/iterations, no cut-off     */
#define RC_ADV_LEGACY 1     /* ripcurrents.cpp:656-698     cut-off r > UPPER, pt += delta*dt/iterations */
#define RC_ADV_MODULE 2     /* ripcurrents_module.cpp:486-528  cut-off r > UPPER, pt += delta*dt     */
#define RC_ADV_CUT5 3       /* ripcurrents_module.cpp:531-569  streamline_2: cut-off r > 5           */
#define RC_ADV_FIXED100 4   /* ripcurrents_module.cpp:572-606  streamline_3: 100 steps of delta*0.1  */
#define RC_ADV_FIELD 5      /* ripcurrents.cpp:611-651 == module:608-648  streamline_field           */
#define RC_ADV_GET_DELTA 6  /* ripcurrents_module.cpp:650-679  get_delta                             */

typedef struct rc_ctx rc_ctx;

/* per-frame record returned by the fused / batched aggregation calls */
typedef struct rc_frame_result {
    int produced;                            /* 1 if a flow was produced (0 for the priming frame) */
    float UPPER;
    float UPPER2d[RC_HIST_DIRECTIONS];
    float prop_above_upper[RC_HIST_DIRECTIONS];
    int64_t histsum;
} rc_frame_result;

/* ---- library / context ------------------------------------------------------------------------ */
int rc_version(void);
const char* rc_error_string(int code);
const char* rc_last_error(const rc_ctx* ctx);
/* number of CUDA kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t rc_kernel_launches(const rc_ctx* ctx);

int rc_create(rc_ctx** out, int device);
void rc_destroy(rc_ctx* ctx);
/* cuda_stream: a cudaStream_t.  NULL selects a private non-blocking stream (the default); to run on the legacy
 * default stream pass cudaStreamLegacy ((void*)1), for the per-thread default stream cudaStreamPerThread ((void*)2). */
int rc_set_stream(rc_ctx* ctx, void* cuda_stream);
int rc_synchronize(rc_ctx* ctx);

/* Per-kernel device timing for bench.py's roofline: when enabled every kernel launch is bracketed by two CUDA
 * events on the launching stream; rc_profile_get returns, per kernel class, the summed duration, the number of
 * launches and the summed ALGORITHMIC bytes (DESIGN.md lists the per-pixel figures).  Off by default. */
int rc_profile_enable(rc_ctx* ctx, int on);
int rc_profile_reset(rc_ctx* ctx);
int rc_profile_count(void);
int rc_profile_get(rc_ctx* ctx, int idx, const char** name, double* total_ms, int64_t* launches, double* alg_bytes);

/* ---- A1: dense optical flow ------------------------------------------------------------------- */
/* Replaces cv::calcOpticalFlowFarneback(prev, next, flow, pyr_scale, levels, winsize, iterations, poly_n,
 * poly_sigma, flags) as called at ripcurrents.cpp:215, main.cpp:264,609,742,961,1119,1481.
 * prev/next: 8-bit single-channel w x h.  flow: w x h x 2 fp32 out (may be NULL: result stays on the device,
 * see rc_flow_device).  OPTFLOW_USE_INITIAL_FLOW (4) is not supported (never used by the reference). */
int rc_farneback(rc_ctx* ctx, const uint8_t* prev, size_t prev_step, const uint8_t* next, size_t next_step, int w, int h,
                 float* flow, size_t flow_step, double pyr_scale, int levels, int winsize, int iterations, int poly_n,
                 double poly_sigma, int flags);

/* Streaming form of the same call for a VideoCapture loop (ripcurrents.cpp:194-217: `u_f1.copyTo(u_f2)`):
 * rc_flow_configure fixes geometry + parameters; each rc_flow_push consumes ONE new frame, reuses the cached
 * pyramid + polynomial expansion of the previous frame, and (from the second frame on) produces the flow
 * previous -> new.  Returns 1 when a flow was produced, 0 for the priming frame, <0 on error. */
int rc_flow_configure(rc_ctx* ctx, int w, int h, double pyr_scale, int levels, int winsize, int iterations, int poly_n,
                      double poly_sigma, int flags);
int rc_flow_push(rc_ctx* ctx, const uint8_t* frame, size_t step, float* flow, size_t flow_step);
/* device pointer to the most recent flow (w*h*2 fp32, dense rows), valid until the next flow call */
int rc_flow_device(rc_ctx* ctx, float** dev_flow, int* w, int* h);

/* Batched streaming form -- the B200-native way to run a recorded clip: up to max_batch consecutive frames go
 * through every kernel together (one launch per stage for the whole batch), which is what keeps 148 SMs busy on
 * the small pyramid layers.  Frame j of the call starts at frames + j*frame_stride.  flows (may be NULL) receives
 * the produced flows back to back (flow_stride bytes apart).  Returns the number of flows produced: count, or
 * count-1 when the first frame of the call primed an empty context.  Device-pointer input may exceed max_batch
 * (it is processed in sub-batches); host input is limited to max_batch frames per call. */
int rc_flow_configure_batch(rc_ctx* ctx, int w, int h, double pyr_scale, int levels, int winsize, int iterations,
                            int poly_n, double poly_sigma, int flags, int max_batch);
int rc_flow_push_batch(rc_ctx* ctx, const uint8_t* frames, size_t step, size_t frame_stride, int count, float* flows,
                       size_t flow_step, size_t flow_stride);
/* device pointer to the flow produced `back` pairs before the most recent one (0 = most recent); flows stay
 * resident for max_batch + window pairs */
int rc_flow_device_at(rc_ctx* ctx, int back, float** dev_flow);

/* ---- A2 + A3: polar conversion and cumulative histograms ---------------------------------------- */
/* Replaces ripcurrents.cpp:305-330 / create_histogram's counting loop (ripcurrents_module.cpp:94-107).
 * Counters live on the device in the context and are CUMULATIVE (ripcurrents.cpp:147-153) until rc_hist_reset.
 * flow == NULL means "the context's most recent flow". */
int rc_hist_reset(rc_ctx* ctx);
int rc_polar_hist(rc_ctx* ctx, const float* flow, size_t flow_step, int w, int h);
int rc_hist_get(rc_ctx* ctx, int64_t hist[RC_HIST_BINS], int64_t* histsum,
                int64_t hist2d[RC_HIST_ROWS * RC_HIST_BINS], int64_t histsum2d[RC_HIST_ROWS]);
/* adds external counts (e.g. the exclusive prefix of other ranks' frames) */
int rc_hist_add(rc_ctx* ctx, const int64_t hist2d[RC_HIST_ROWS * RC_HIST_BINS]);
/* device pointer to the counters: int64[RC_HIST_ROWS*RC_HIST_BINS] hist2d (hist/histsum/histsum2d are its
 * marginals) -- for NCCL all-gather / all-reduce by the caller */
int rc_hist_device(rc_ctx* ctx, int64_t** dev_hist2d);
/* cv::cartToPolar(x, y, mag, angle, true) (ripcurrents.cpp:308) on n points; bit-exact restatement */
int rc_cart_to_polar(rc_ctx* ctx, const float* flow, size_t n, float* mag, float* angle_deg);

/* ---- A4: thresholds ----------------------------------------------------------------------------- */
/* Replaces ripcurrents.cpp:333-366 / ripcurrents_module.cpp:110-143, evaluated on the device counters.
 * Any output pointer may be NULL.  The values also stay on the device for rc_classify_accumulate. */
int rc_thresholds(rc_ctx* ctx, float* UPPER, float UPPER2d[RC_HIST_DIRECTIONS],
                  float prop_above_upper[RC_HIST_DIRECTIONS]);

/* ---- A5: classify + accumulate + mask ----------------------------------------------------------- */
/* Replaces ripcurrents.cpp:376-439 / create_flow + create_accumulationbuffer (module:153-212).
 * upper: threshold to use; NaN = the value computed by the last rc_thresholds (kept on the device).
 * The accumulator (.x lane of the reference's CV_32FC3 `accumulator`) lives in the context.
 * outmask / waveclass / waterclass: w*h u8 outputs, each may be NULL (host or device pointers).
 *   outmask    255 = calm, 0 = wave                                   (ripcurrents.cpp:436)
 *   waveclass  0 calm, 1 = val < .2*framecount, 2 = otherwise         (ripcurrents.cpp:429-433)
 *   waterclass 3 > UPPER, 2 > MID(0.5), 1 > LOWER(0.2), 0 otherwise   (ripcurrents.cpp:384-390) */
int rc_accumulator_reset(rc_ctx* ctx);
int rc_classify_accumulate(rc_ctx* ctx, const float* flow, size_t flow_step, int w, int h, float upper, int framecount,
                           uint8_t* outmask, uint8_t* waveclass, uint8_t* waterclass);
int rc_accumulator_get(rc_ctx* ctx, float* acc_x /* w*h fp32 */);
int rc_accumulator_device(rc_ctx* ctx, float** dev_acc_x, int* w, int* h);

/* ---- A6: sliding-window flow mean ---------------------------------------------------------------- */
/* Replaces main.cpp:1084-1092 + 1143-1153 (W=10), :1446 + 1505-1515 (W=100), module:392-400 (W=300):
 * avg -= slot/W; slot = flow; avg += slot/W  in fp32, in order.  flow == NULL: the context's last flow. */
int rc_window_configure(rc_ctx* ctx, int w, int h, int W);
int rc_window_update(rc_ctx* ctx, const float* flow, size_t flow_step);
int rc_window_get(rc_ctx* ctx, float* avg, size_t avg_step);
int rc_window_device(rc_ctx* ctx, float** dev_avg);
/* averageVector's window update (ripcurrents_module.cpp:392-400, declared at ripcurrents.hpp:48), one pass per frame:
 *   average -= old_slot / frames;  new = the get_delta field of `flow` (every pixel: one step of dt from displacement 0
 *   at its own position, cut-off r > upper; module:396-398 calls get_delta(&pixel, x, y, current, 2, UPPER));
 *   average += new / frames.
 * old_slot: the buffer entry being replaced (w*h*2 fp32; NULL = zeros); average: w*h*2 fp32 in/out; new_slot (may be NULL)
 * receives the new field.  All dense; host or device pointers (all of one kind).  The reference passes its buffer vector BY
 * VALUE, so its caller's slot is never rewritten -- the header-compatible wrapper reproduces that. */
int rc_average_vector(rc_ctx* ctx, const float* old_slot, const float* flow, size_t flow_step, int w, int h, float* average,
                      float* new_slot, int frames, float dt, float upper);
/* subtructAverage (ripcurrents_module.cpp:810-863): flow -= mean(flow); in place; mean_xy[2] out (may be NULL) */
int rc_subtract_mean(rc_ctx* ctx, float* flow, size_t flow_step, int w, int h, double* mean_xy);

/* ---- A7: particle advection ----------------------------------------------------------------------- */
/* Replaces streamline / streamline_2 / streamline_3 / streamline_field / get_delta (see RC_ADV_*).
 * flow == NULL: the context's last flow (w,h ignored).  seeds: n x (x,y) fp32 updated in place.
 * dist: n fp32 path lengths (RC_ADV_FIELD) or NULL.  home: n x (x,y) int32 home pixels for RC_ADV_FIELD /
 * RC_ADV_GET_DELTA, NULL = seed i lives at pixel (i % w, i / w) (the reference's per-pixel field). */
int rc_advect(rc_ctx* ctx, const float* flow, size_t flow_step, int w, int h, float* seeds, size_t n, float dt,
              int iterations, float upper, int variant, float* dist, const int32_t* home);
/* One Streakline::runLK frame (Streakline.cpp:22-48) for E emitters, vertex motion taken from the dense flow
 * (variant RC_ADV_MODULE step, no cut-off).  vertices: E x cap x (x,y), reference order (index 0 = newest);
 * count[e] vertices valid; after the call count[e] grows by one (until cap). */
int rc_streakline_step(rc_ctx* ctx, const float* flow, size_t flow_step, int w, int h, const float* emitters, int E,
                       float* vertices, int32_t* count, int cap, float dt);

/* ---- one stream sharded by frame pair across ranks --------------------------------------
 * The flow of a pair depends only on its two frames, but thresholds at frame t need the counts of ALL frames <= t.
 * Each rank: rc_flow_push_batch(its block) -> rc_batch_hist (per-frame counts of its last nb flows) -> the caller
 * all-gathers the counts, adds those of all EARLIER frames with rc_hist_add -> rc_aggregate_last (exact per-frame
 * thresholds, classification into this rank's accumulator) -> all-reduce(SUM) of rc_accumulator_device (integer-
 * valued, exact in any order) -> rc_accumulator_mask at the reporting point.  The sliding-window mean needs the flows
 * in order and is not available in this mode. */
int rc_batch_hist(rc_ctx* ctx, int nb, int64_t* deltas /* nb * RC_HIST_ROWS * RC_HIST_BINS, host or device */);
int rc_aggregate_last(rc_ctx* ctx, int nb, int framecount0, rc_frame_result* results /* nb records or NULL */);
int rc_accumulator_mask(rc_ctx* ctx, int framecount, uint8_t* outmask /* w*h u8 */);

/* ---- multi-GPU over NCCL / NVLink -------------------------------------------------------------
 * The reference is a single process on one stream; these entry points have no reference counterpart.  They give a C / C++
 * host (main.cpp stays C++) both ways of using the 8 GPUs of a box without any Python:
 *   camera streams on different GPUs:  one context per (GPU, stream), no exchange in the flow;
 *                                      rc_allreduce_accumulators builds the shared wave-activity map;
 *   ONE stream split by frame pair:    rc_shard_configure / rc_shard_step / rc_shard_report (below).
 * NCCL (libnccl.so.2) is loaded at first use with dlopen; without it these calls fail with RC_ERR_UNSUPPORTED and
 * everything else works.  All collectives are enqueued on the context's stream. */
#define RC_COMM_ID_BYTES 128
/* rank 0: a fresh NCCL unique id; the host program hands the 128 bytes to every rank */
int rc_comm_unique_id(char id[RC_COMM_ID_BYTES]);
/* collective over all ranks: communicator of `nranks` ranks on this context's device (ncclCommInitRank) */
int rc_comm_init(rc_ctx* ctx, const char id[RC_COMM_ID_BYTES], int rank, int nranks);
/* or adopt a communicator the host program already has (nccl_comm is an ncclComm_t; not destroyed by the context) */
int rc_comm_attach(rc_ctx* ctx, void* nccl_comm, int rank, int nranks);
int rc_comm_destroy(rc_ctx* ctx);
/* all-reduce(SUM) of this context's accumulator.x (w*h fp32) and cumulative hist2d (RC_HIST_ROWS*RC_HIST_BINS int64) over
 * the ranks of nccl_comm (NULL = the context's communicator).  Results go to the device buffers shared_acc / shared_hist,
 * or in place when NULL.  Both are integer-valued sums: exact in any order. */
int rc_allreduce_accumulators(rc_ctx* ctx, void* nccl_comm, float* shared_acc, int64_t* shared_hist);

/* One stream sharded by frame pair.  The stream is consumed in SUPER-BLOCKS: rank r takes the r-th run of
 * pairs_per_rank[r] (<= max_batch) consecutive pairs of each super-block.  Every rank calls rc_shard_step once per
 * super-block (ranks with no pairs pass count = 0):
 *   frames        this rank's count = pairs + 1 frames; the first one is the last frame of the preceding run (one duplicated
 *                 frame per block edge)
 *   framecount0   the reference loop counter (ripcurrents.cpp:194) of the frame that completes this rank's first pair
 *   results       pairs records (exact per-frame thresholds: all-gather of per-frame counts + exclusive prefix) or NULL
 * rc_shard_configure(window_W, owner): window_W > 0 keeps the fp32 sliding-window mean (main.cpp:1143-1153).  Its rounding is
 * order-dependent PER PIXEL, so the state is sharded by pixel band: rank r owns rows [r h / N, (r+1) h / N) of the mean;
 * every step each rank hands band r of its flows to rank r (all-to-all over NVLink: (N-1)/N of a rank's own flows out, as
 * much in, independent of N) and applies the super-block's updates to its band in stream order -- bit-identical to the
 * sequential pipeline, no serial tail.  rc_shard_window_get is COLLECTIVE (every rank broadcasts its band) and leaves the
 * whole mean on every rank.  `owner` is reserved (pass 0).  rc_shard_report (collective): all-reduce of the ranks'
 * accumulators, then outmask (ripcurrents.cpp:424-439) / accumulator.x / the stream's cumulative hist2d, each optional,
 * host or device. */
int rc_shard_configure(rc_ctx* ctx, int window_W, int owner);
int rc_shard_step(rc_ctx* ctx, const uint8_t* frames, size_t step, size_t frame_stride, int count, int framecount0,
                  const int* pairs_per_rank, rc_frame_result* results);
int rc_shard_report(rc_ctx* ctx, int framecount, uint8_t* outmask, float* acc_x, int64_t* hist2d);
int rc_shard_window_get(rc_ctx* ctx, float* avg /* w*h*2 fp32 */);

/* ---- entry points on the reference's own intermediate formats (used by the header-compatible C++ wrappers) ---- */
/* counting loop of create_histogram (ripcurrents_module.cpp:94-107) on the merged polar image CV_32FC3
 * (angle deg, mag, mag); adds into the context's cumulative counters like rc_polar_hist */
int rc_hist_from_polar(rc_ctx* ctx, const float* polar3, size_t step, int w, int h);
/* create_flow (ripcurrents_module.cpp:153-182): all three images are CV_32FC3, updated in place */
int rc_create_flow(rc_ctx* ctx, float* current3, size_t cur_step, float* waterclass3, size_t wc_step,
                   float* accumulator2_3, size_t acc2_step, int w, int h, float UPPER, float MID, float LOWER,
                   const float UPPER2d[RC_HIST_DIRECTIONS]);
/* create_accumulationbuffer (ripcurrents_module.cpp:189-212): accumulator/accumulator2/out CV_32FC3, outmask CV_8UC1 */
int rc_create_accumulationbuffer(rc_ctx* ctx, float* accumulator3, size_t acc_step, const float* accumulator2_3,
                                 size_t acc2_step, float* out3, size_t out_step, uint8_t* outmask, size_t mask_step, int w,
                                 int h, int framecount);

/* ---- frame ingest ---------------------------------------------------------------
 * cv::resize(frame, subframe, Size(dst_w,dst_h), 0, 0, INTER_LINEAR) + cv::cvtColor(subframe, gray, COLOR_BGR2GRAY):
 * ripcurrents.cpp:209-210, main.cpp:258-259,1111-1112.  8-bit, OpenCV's fixed-point arithmetic, bit-exact against
 * cv2 4.13.0.  flags: RC_INGEST_GRAY14 selects OpenCV 3.4's 14-bit gray weights (1868/9617/4899) instead of the
 * 15-bit ones of OpenCV 4 (3735/19235/9798). */
#define RC_INGEST_GRAY14 1
/* RC_INGEST_AREA: cv::resize(..., INTER_AREA) instead of INTER_LINEAR -- what the reference applies to the PRIMING frame of
 * every loop (ripcurrents.cpp:186, main.cpp:223,570,704,936,1067,1429) before cvtColor; downscaling only (both ratios >= 1,
 * RC_ERR_UNSUPPORTED otherwise), bit-exact against cv2 4.13.0 on fractional, integer and 2x2 ratios. */
#define RC_INGEST_AREA 2
int rc_ingest_bgr(rc_ctx* ctx, const uint8_t* bgr, size_t step, int src_w, int src_h, uint8_t* gray, size_t gray_step,
                  int dst_w, int dst_h, int flags);
/* rc_submit_frames fed with BGR camera frames of any size: H2D of the BGR frames, ingest on the device to the
 * configured working size, then the same pipeline.  Same asynchronous contract as rc_submit_frames (rc_wait). */
int rc_submit_frames_bgr(rc_ctx* ctx, const uint8_t* bgr_frames, size_t step, size_t frame_stride, int src_w, int src_h,
                         int count, int framecount0, int ingest_flags, uint8_t* outmasks, size_t mask_stride,
                         rc_frame_result* results);

/* ---- mask clean-up ---------------------------------------------------------------
 * create_edges (ripcurrents_module.cpp:216-220 == ripcurrents.cpp:494-496): dilate(outmask, 5x5 ellipse) followed by
 * morphologyEx(MORPH_GRADIENT) with the same element; bit-exact against cv2.  `count` masks `mask_stride` bytes
 * apart are processed in one launch; edges may alias masks only when count == 1 and the pointers are on the host. */
int rc_mask_edges(rc_ctx* ctx, const uint8_t* masks, size_t mask_step, size_t mask_stride, int w, int h, int count,
                  uint8_t* edges, size_t edges_step, size_t edges_stride);

/* ---- derived fields of the per-pixel particle state -----------------------------
 * What ripcurrents.cpp:231-279 (== ripcurrents_module.cpp:13-59, called from main_old.cpp:373-386) computes after
 * streamline_field() with split / magnitude / minMaxLoc / convertTo / applyColorMap(JET) / divide and the position
 * scatter.  Bit-exact against cv2 4.13 (cv2.magnitude compared with cv2.setUseOptimized(False); the IPP-backed default
 * differs from OpenCV's own definition by <= 2 ulp).  All arrays are dense (no row padding); every pointer may be a host
 * or a device pointer; optional outputs may be NULL.  Calls with only device pointers and maxes == NULL are asynchronous
 * on the context's stream. */
#define RC_FIELDS_DIV_ZERO_IS_ZERO 1   /* cv::divide of OpenCV 3.x (the version the reference names): x / 0 -> 0;
                                          default is OpenCV 4.x: IEEE inf / NaN */
#define RC_FIELDS_KEEP_DENSITY 2       /* do not clear `density` first (module:44 takes the caller's Mat as it is;
                                          ripcurrents.cpp:261 starts from Mat::zeros, the default) */
/* One call for the whole block: field = streamlines_mat (w*h float2 displacements), dist = streamlines_distance.
 *   streamfield  magnitude(field)                                                       (ripcurrents.cpp:232-233)
 *   disp_bgr     JET(convertTo(streamfield, 255 / max))      "streamline displacement"  (:236-240, module:13-19)
 *   motion_bgr   JET(convertTo(dist, 255 / max))             "streamline total motion"  (:243-247, module:23-28)
 *   ratio_bgr    JET(convertTo(streamfield / dist, 255/max)) "displacement/motion ratio"(:250-256, module:34-40)
 *   density      CV_32FC3, (1,1,1) where a particle ends up                             (:261-276, module:44-59)
 *   maxes        {lenmax, distmax, ratiomax}; NaNs are ignored by the maxima (cv2 4.x returns a SIMD-lane dependent
 *                value when NaNs are present), NaN only if every element is NaN. */
int rc_particle_fields(rc_ctx* ctx, const float* field, const float* dist, int w, int h, int flags, float* streamfield,
                       uint8_t* disp_bgr, uint8_t* motion_bgr, uint8_t* ratio_bgr, float* density, double* maxes);
/* The pieces, as the module factors them.  streamline_displacement / streamline_total_motion (module:13-29):
 * minMaxLoc -> convertTo(CV_8UC1, 255/max) -> applyColorMap(JET) of n floats; gray (before the colour map) and bgr
 * are optional. */
int rc_normalize_jet(rc_ctx* ctx, const float* src, size_t n, uint8_t* gray, uint8_t* bgr, double* maxval);
/* streamline_ratio (module:34-40): divide(a, b) then the same normalisation; `ratio` (optional) receives a / b. */
int rc_ratio_jet(rc_ctx* ctx, const float* a, const float* b, size_t n, int flags, float* ratio, uint8_t* gray,
                 uint8_t* bgr, double* maxval);
/* split + magnitude (ripcurrents.cpp:232-233) of n interleaved (x, y) pairs */
int rc_field_magnitude(rc_ctx* ctx, const float* field, size_t n, float* mag);
/* streamline_positions (module:44-59) */
int rc_streamline_positions(rc_ctx* ctx, const float* field, int w, int h, float* density, int flags);

/* ---- flow-derived diagnostics -----------------------------------------------------
 * ripcurrents_module.cpp:900-1138.  Bit-exact against the CPU restatement of the reference as GCC/x86-64 + glibc 2.39
 * + cv2 4.13 execute it (float->uchar stores truncate and wrap; atan2f as libm computes it; 8-bit HSV->BGR as cv2's
 * 32-pixel block path does: truncating, with FMAs -- RC_DIAG_HSV_NOFMA selects cv2's setUseOptimized(False) variant).
 * The reference's function-local `static` maxima (each frame is normalised with the previous frame's maximum) are the
 * in/out argument: a host float (in = previous maximum, out = this frame's; the call synchronises), or NULL to keep
 * the state inside the context (starts at 0 like the reference's statics; the call stays asynchronous). */
#define RC_DIAG_SEQUENTIAL_SUM 1   /* subtructMeanMagnitude: accumulate the mean in fp32 in pixel order on one thread,
                                      as the reference's loop does (bit-exact, ~5 ms at 1080p); default: fp64 tree sum */
#define RC_DIAG_HSV_NOFMA 2
/* subtructMeanMagnitude (module:900-1015): v <- v/|v| * (|v| - mean|v|), in place (host or device flow); *meanval
 * (optional) receives the mean magnitude.  The reference's printf diagnostics are not reproduced. */
int rc_subtract_mean_magnitude(rc_ctx* ctx, float* flow, size_t flow_step, int w, int h, int flags, float* meanval);
/* vectorToColor (module:1017-1057): hue = direction/2, saturation 255, value = |v| * 255 / max, cvtColor(HSV2BGR).
 * flow == NULL: the context's last flow.  bgr: w*h*3 u8, host or device. */
int rc_vector_to_color(rc_ctx* ctx, const float* flow, size_t flow_step, int w, int h, uint8_t* bgr, size_t bgr_step,
                       float* max_displacement, int flags);
/* shearRateToColor (module:1059-1138): Frobenius norm of the velocity Jacobian (central differences at +-10 px) as
 * hue = 128 - norm * 128 / max on the interior; the border of the caller's image is left alone and then, like the
 * interior, converted HSV->BGR (img is in/out, as in the reference). */
int rc_shear_rate_to_color(rc_ctx* ctx, const float* flow, size_t flow_step, int w, int h, uint8_t* img, size_t img_step,
                           float* max_frobenius, int flags);

/* ---- fused per-frame step (what main()'s loop body does between video.read and imshow) ----------- */

/* rc_flow_push + rc_polar_hist + rc_thresholds + rc_classify_accumulate (+ rc_window_update when a window is
 * configured) for one new frame, on the device, in stream order.  outmask (w*h u8) and result may be NULL;
 * when both are NULL nothing is copied back and the call does not synchronise. */
int rc_process_frame(rc_ctx* ctx, const uint8_t* frame, size_t step, int framecount, uint8_t* outmask,
                     rc_frame_result* result);
/* The same for `count` (<= max_batch) consecutive frames: frame i of the call has loop counter framecount0 + i.
 * outmasks: count masks mask_stride bytes apart (entries of frames that produced no flow are left untouched);
 * results: count records.  Temporal state (cumulative histograms -> per-frame thresholds, accumulator, window
 * mean) advances frame by frame in the reference's order inside the batch. */
int rc_process_frames(rc_ctx* ctx, const uint8_t* frames, size_t step, size_t frame_stride, int count, int framecount0,
                      uint8_t* outmasks, size_t mask_stride, rc_frame_result* results);
/* Asynchronous form for PINNED host buffers: returns as soon as the work is enqueued.  Host->device copies run on
 * a copy stream into one of two staging slots, kernels on the context stream, device->host copies on a third
 * stream, so the transfers of batch s+1 / s-1 overlap the kernels of batch s.  Frames, outmasks and results of a
 * submit must stay untouched until rc_wait() -- or until the second-next rc_submit_frames returns, which waits for
 * the batch that used the same staging slot. */
int rc_submit_frames(rc_ctx* ctx, const uint8_t* frames, size_t step, size_t frame_stride, int count, int framecount0,
                     uint8_t* outmasks, size_t mask_stride, rc_frame_result* results);
int rc_wait(rc_ctx* ctx);
/* Format of the outmasks written by rc_process_frame(s) / rc_submit_frames(_bgr).  The reference's outmask holds only 0 and
 * 255 (ripcurrents.cpp:436), so RC_MASK_PACKED returns it as 1 bit per pixel -- bit (p & 7) of byte (p >> 3) is 1 where the
 * u8 mask is 255, pixels in row-major order; w must be a multiple of 8; one mask is w*h/8 bytes -- which cuts the
 * device->host traffic of a 1080p stream from 2.07 MB to 0.26 MB per frame.  Default RC_MASK_U8 (the reference's CV_8UC1). */
#define RC_MASK_U8 0
#define RC_MASK_PACKED 1
int rc_set_mask_format(rc_ctx* ctx, int format);

#ifdef __cplusplus
}
#endif
#endif /* RIPCURRENTS_B200_H */
