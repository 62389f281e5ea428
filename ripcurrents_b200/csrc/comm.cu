// Multi-GPU entry points of the C ABI: NCCL over NVLink 5 / NVSwitch.
//
//  * rc_comm_*                 one NCCL communicator per context (one rank per GPU); the unique id travels through whatever
//                              the host program uses to talk between its ranks (MPI, a socket, torch.distributed in bench.py)
//  * rc_allreduce_accumulators the only collective the stream-per-GPU split needs: all-reduce(SUM) of the per-stream
//                              accumulators and histograms into a shared wave-activity map (integer-valued: exact in any order)
//  * rc_shard_*                ONE stream split by frame pair: per super-block every rank computes the flows of its run of
//                              pairs (no exchange), then
//                                all-gather of per-frame counts + exclusive prefix over ranks -> exact per-frame thresholds
//                                (ripcurrents.cpp:147-153: the counters are cumulative over the clip),
//                                classify into the rank's own accumulator (all-reduced at reporting points),
//                                the order-dependent fp32 window mean (main.cpp:1143-1153) is order-dependent PER PIXEL, so
//                                its state is sharded by pixel band: rank r owns rows [r h / N, (r+1) h / N) of the mean, every
//                                rank hands band r of each of its flows to rank r (an all-to-all of ncclSend/ncclRecv: each
//                                rank moves (N-1)/N of ITS OWN flows out and as much in, however many ranks there are) and
//                                applies all updates of the super-block to its band in stream order -- no serial tail.
//
// NCCL is loaded with dlopen at first use (libnccl.so.2: the copy torch already mapped when the library runs inside a
// Python process, the system one otherwise); the library itself links only cudart, so single-GPU users never need NCCL.
#include <dlfcn.h>
#include <string.h>
#include <nccl.h>

#include "rc_internal.h"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        const char* names[] = {getenv("RC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) return a;
#define RC_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, name))
        RC_SYM(GetUniqueId, "ncclGetUniqueId"); RC_SYM(CommInitRank, "ncclCommInitRank"); RC_SYM(CommDestroy, "ncclCommDestroy");
        RC_SYM(AllReduce, "ncclAllReduce"); RC_SYM(AllGather, "ncclAllGather"); RC_SYM(Send, "ncclSend"); RC_SYM(Recv, "ncclRecv");
        RC_SYM(Broadcast, "ncclBroadcast"); RC_SYM(GroupStart, "ncclGroupStart"); RC_SYM(GroupEnd, "ncclGroupEnd"); RC_SYM(GetErrorString, "ncclGetErrorString");
#undef RC_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.Send && a.Recv && a.Broadcast && a.GroupStart &&
               a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}

#define NCCL_TRY(c, expr)                                                                                   \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess) return rc_fail((c), RC_ERR_CUDA, #expr ": %s", nccl().GetErrorString(_r));    \
    } while (0)
#define CU_TRY(c, expr)                                                                                     \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess) return rc_fail((c), RC_ERR_CUDA, #expr ": %s", cudaGetErrorString(_e));       \
    } while (0)

int need_nccl(rc_ctx* c)
{
    if (!nccl().ok) return rc_fail(c, RC_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded (set RC_NCCL_LIB)%s", "");
    return RC_OK;
}

size_t flow_floats(const rc_ctx* c) { return 2 * (size_t)c->prm.w * c->prm.h; }
// pixel band of rank r: rows [r h / N, (r+1) h / N)
int band_row0(const rc_ctx* c, int r) { return (int)((long long)r * c->prm.h / c->nranks); }
size_t band_off(const rc_ctx* c, int r) { return 2 * (size_t)c->prm.w * band_row0(c, r); }                       // floats
size_t band_floats(const rc_ctx* c, int r) { return 2 * (size_t)c->prm.w * (band_row0(c, r + 1) - band_row0(c, r)); }
// this rank's band of the flow of stream pair `pair` (ring of W + nranks * B slots)
float* shard_slot(rc_ctx* c, long long pair) { return c->d_shard_ring + (size_t)(pair % c->shard_slots) * band_floats(c, c->rank); }

void free_shard(rc_ctx* c)
{
    if (c->s_comm) cudaStreamSynchronize(c->s_comm);
    if (c->stream) cudaStreamSynchronize(c->stream);
    void* bufs[] = {c->d_gather, c->d_hist_global, c->d_acc_global, c->d_shard_ring, c->d_shard_avg, c->d_shard_flows};
    for (void* p : bufs) if (p) cudaFree(p);
    c->d_gather = nullptr; c->d_hist_global = nullptr; c->d_acc_global = nullptr; c->d_shard_ring = nullptr; c->d_shard_avg = nullptr;
    c->d_shard_flows = nullptr;
    for (int i = 0; i < 2; i++) {
        if (c->ev_flows[i]) { cudaEventDestroy(c->ev_flows[i]); c->ev_flows[i] = nullptr; }
        if (c->ev_comm[i]) { cudaEventDestroy(c->ev_comm[i]); c->ev_comm[i] = nullptr; }
    }
    c->shard_configured = false; c->shard_slots = 0; c->shard_pairs = 0; c->shard_steps = 0;
}

int ensure_comm_stream(rc_ctx* c)
{
    if (c->s_comm) return RC_OK;
    CU_TRY(c, cudaStreamCreateWithFlags(&c->s_comm, cudaStreamNonBlocking));
    CU_TRY(c, cudaEventCreateWithFlags(&c->ev_ar_snap, cudaEventDisableTiming));
    CU_TRY(c, cudaEventCreateWithFlags(&c->ev_ar_done, cudaEventDisableTiming));
    return RC_OK;
}

}  // namespace

// called by rc_synchronize / rc_wait / rc_destroy (api.cu): everything enqueued on the communication stream has finished
int rc_comm_fence(rc_ctx* c, bool destroy)
{
    if (!c->s_comm) return RC_OK;
    if (cudaStreamSynchronize(c->s_comm) != cudaSuccess) return RC_ERR_CUDA;
    if (destroy) {
        cudaEventDestroy(c->ev_ar_snap); cudaEventDestroy(c->ev_ar_done); cudaStreamDestroy(c->s_comm);
        c->ev_ar_snap = c->ev_ar_done = nullptr; c->s_comm = nullptr;
    }
    return RC_OK;
}

extern "C" {

int rc_comm_unique_id(char id[RC_COMM_ID_BYTES])
{
    if (!id) return RC_ERR_INVALID;
    if (!nccl().ok) return RC_ERR_UNSUPPORTED;
    static_assert(RC_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    ncclUniqueId u;
    if (nccl().GetUniqueId(&u) != ncclSuccess) return RC_ERR_CUDA;
    memcpy(id, u.internal, RC_COMM_ID_BYTES);
    return RC_OK;
}

int rc_comm_init(rc_ctx* c, const char id[RC_COMM_ID_BYTES], int rank, int nranks)
{
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return RC_ERR_INVALID;
    int rc = need_nccl(c); if (rc) return rc;
    cudaSetDevice(c->device);
    rc_comm_destroy(c);
    ncclUniqueId u;
    memcpy(u.internal, id, RC_COMM_ID_BYTES);
    ncclComm_t comm = nullptr;
    NCCL_TRY(c, nccl().CommInitRank(&comm, nranks, u, rank));
    c->nccl = comm; c->own_comm = true; c->rank = rank; c->nranks = nranks;
    return RC_OK;
}

int rc_comm_attach(rc_ctx* c, void* nccl_comm, int rank, int nranks)
{
    if (!c || !nccl_comm || nranks < 1 || rank < 0 || rank >= nranks) return RC_ERR_INVALID;
    int rc = need_nccl(c); if (rc) return rc;
    rc_comm_destroy(c);
    c->nccl = nccl_comm; c->own_comm = false; c->rank = rank; c->nranks = nranks;
    return RC_OK;
}

int rc_comm_destroy(rc_ctx* c)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    free_shard(c);
    if (c->nccl && c->own_comm) {
        if (c->stream) cudaStreamSynchronize(c->stream);
        nccl().CommDestroy(static_cast<ncclComm_t>(c->nccl));
    }
    c->nccl = nullptr; c->own_comm = false; c->rank = 0; c->nranks = 1;
    return RC_OK;
}

int rc_allreduce_accumulators(rc_ctx* c, void* nccl_comm, float* shared_acc, int64_t* shared_hist)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm ? nccl_comm : c->nccl);
    if (!c->d_acc || !c->d_hist2d) return rc_fail(c, RC_ERR_STATE, "no accumulator / histogram on this context yet%s", "");
    if ((shared_acc && !rc_is_device_ptr(shared_acc)) || (shared_hist && !rc_is_device_ptr(shared_hist)))
        return rc_fail(c, RC_ERR_INVALID, "shared_acc / shared_hist must be device pointers%s", "");
    const size_t n = (size_t)c->acc_w * c->acc_h;
    float* acc_out = shared_acc ? shared_acc : c->d_acc;
    void* hist_out = shared_hist ? static_cast<void*>(shared_hist) : static_cast<void*>(c->d_hist2d);
    if (!comm) {          // a single rank: the reduction is the identity
        if (shared_acc) CU_TRY(c, cudaMemcpyAsync(shared_acc, c->d_acc, n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        if (shared_hist) CU_TRY(c, cudaMemcpyAsync(shared_hist, c->d_hist2d, RC_HIST_CELLS * 8, cudaMemcpyDeviceToDevice, c->stream));
        return RC_OK;
    }
    int rc = need_nccl(c); if (rc) return rc;
    if (shared_acc && shared_hist) {
        // out of place: snapshot on the compute stream, reduce the snapshot on the communication stream -- the collective of
        // step s overlaps the kernels of step s+1 (the next snapshot waits for it); results are complete after rc_wait /
        // rc_synchronize
        rc = ensure_comm_stream(c); if (rc) return rc;
        CU_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_ar_done, 0));
        CU_TRY(c, cudaMemcpyAsync(shared_acc, c->d_acc, n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        CU_TRY(c, cudaMemcpyAsync(shared_hist, c->d_hist2d, RC_HIST_CELLS * 8, cudaMemcpyDeviceToDevice, c->stream));
        CU_TRY(c, cudaEventRecord(c->ev_ar_snap, c->stream));
        CU_TRY(c, cudaStreamWaitEvent(c->s_comm, c->ev_ar_snap, 0));
        NCCL_TRY(c, nccl().GroupStart());
        NCCL_TRY(c, nccl().AllReduce(shared_acc, shared_acc, n, ncclFloat, ncclSum, comm, c->s_comm));
        NCCL_TRY(c, nccl().AllReduce(shared_hist, shared_hist, RC_HIST_CELLS, ncclInt64, ncclSum, comm, c->s_comm));
        NCCL_TRY(c, nccl().GroupEnd());
        CU_TRY(c, cudaEventRecord(c->ev_ar_done, c->s_comm));
        c->launches += 2;
        return RC_OK;
    }
    NCCL_TRY(c, nccl().GroupStart());
    NCCL_TRY(c, nccl().AllReduce(c->d_acc, acc_out, n, ncclFloat, ncclSum, comm, c->stream));
    NCCL_TRY(c, nccl().AllReduce(c->d_hist2d, hist_out, RC_HIST_CELLS, ncclInt64, ncclSum, comm, c->stream));
    NCCL_TRY(c, nccl().GroupEnd());
    c->launches += 2;
    return RC_OK;
}

// ---- one stream sharded by frame pair ---------------------------------------------------------------------------
int rc_shard_configure(rc_ctx* c, int window_W, int owner)
{
    if (!c || window_W < 0) return RC_ERR_INVALID;
    if (!c->configured) return rc_fail(c, RC_ERR_STATE, "rc_flow_configure_batch has not been called%s", "");
    if (owner < 0 || owner >= c->nranks) return rc_fail(c, RC_ERR_INVALID, "owner must be a rank of the communicator%s", "");
    if (c->nranks > 1 && !c->nccl) return rc_fail(c, RC_ERR_STATE, "rc_comm_init has not been called%s", "");
    cudaSetDevice(c->device);
    int rc = rc_ensure_aggregate(c); if (rc) return rc;
    rc = rc_ensure_accumulator(c, c->prm.w, c->prm.h); if (rc) return rc;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    free_shard(c);
    const size_t n = (size_t)c->prm.w * c->prm.h, ff = flow_floats(c);
    CU_TRY(c, cudaMalloc((void**)&c->d_gather, sizeof(unsigned int) * RC_HIST_CELLS * (size_t)c->B * c->nranks));
    CU_TRY(c, cudaMalloc((void**)&c->d_hist_global, sizeof(unsigned long long) * RC_HIST_CELLS));
    CU_TRY(c, cudaMalloc((void**)&c->d_acc_global, sizeof(float) * (n + 4)));
    CU_TRY(c, cudaMemsetAsync(c->d_hist_global, 0, sizeof(unsigned long long) * RC_HIST_CELLS, c->stream));
    CU_TRY(c, cudaMemsetAsync(c->d_hist2d, 0, sizeof(unsigned long long) * RC_HIST_CELLS, c->stream));
    CU_TRY(c, cudaMemsetAsync(c->d_acc, 0, sizeof(float) * n, c->stream));
    c->shard_owner = owner; c->shard_W = window_W; c->shard_pairs = 0; c->shard_steps = 0;
    rc = ensure_comm_stream(c); if (rc) return rc;
    for (int i = 0; i < 2; i++) {
        CU_TRY(c, cudaEventCreateWithFlags(&c->ev_flows[i], cudaEventDisableTiming));
        CU_TRY(c, cudaEventCreateWithFlags(&c->ev_comm[i], cudaEventDisableTiming));
    }
    CU_TRY(c, cudaMalloc((void**)&c->d_shard_flows, sizeof(float) * ff * 2 * (size_t)c->B + 64));
    if (window_W > 0) {
        c->shard_slots = window_W + 2 * c->nranks * c->B;
        const size_t bf = band_floats(c, c->rank);
        if (cudaMalloc((void**)&c->d_shard_ring, sizeof(float) * bf * c->shard_slots + 64) != cudaSuccess ||
            cudaMalloc((void**)&c->d_shard_avg, sizeof(float) * ff + 64) != cudaSuccess) {
            cudaGetLastError(); free_shard(c);
            return rc_fail(c, RC_ERR_NOMEM, "cudaMalloc failed (band ring of the window mean)%s", "");
        }
        CU_TRY(c, cudaMemsetAsync(c->d_shard_avg, 0, sizeof(float) * ff, c->stream));     // full frame; this rank maintains its band
    }
    c->shard_configured = true;
    return RC_OK;
}

int rc_shard_step(rc_ctx* c, const uint8_t* frames, size_t step, size_t frame_stride, int count, int framecount0,
                  const int* pairs_per_rank, rc_frame_result* results)
{
    if (!c || !pairs_per_rank || count < 0 || count == 1) return RC_ERR_INVALID;
    if (!c->shard_configured) return rc_fail(c, RC_ERR_STATE, "rc_shard_configure has not been called%s", "");
    cudaSetDevice(c->device);
    const int w = c->prm.w, h = c->prm.h, B = c->B, nb = count ? count - 1 : 0, R = c->nranks, me = c->rank;
    const size_t n = (size_t)w * h, ff = flow_floats(c);
    if (nb > B) return rc_fail(c, RC_ERR_INVALID, "more pairs than max_batch%s", "");
    if (pairs_per_rank[me] != nb) return rc_fail(c, RC_ERR_INVALID, "pairs_per_rank[rank] must equal count - 1%s", "");
    int total = 0, before_me = 0;
    for (int r = 0; r < R; r++) {
        if (pairs_per_rank[r] < 0 || pairs_per_rank[r] > B) return rc_fail(c, RC_ERR_INVALID, "pairs_per_rank entries must be in [0, max_batch]%s", "");
        if (r < me) before_me += pairs_per_rank[r];
        total += pairs_per_rank[r];
    }
    if (nb && (!frames || step < (size_t)w || frame_stride < step * (size_t)(h - 1) + w))
        return rc_fail(c, RC_ERR_INVALID, "bad frame pointer / step / stride%s", "");
    ncclComm_t comm = static_cast<ncclComm_t>(c->nccl);
    const int half = (int)(c->shard_steps & 1);

    // 1. flows of this rank's pairs + their per-frame counts (rows >= nb of the delta buffer stay zero)
    CU_TRY(c, cudaMemsetAsync(c->d_hist_delta, 0, sizeof(unsigned int) * RC_HIST_CELLS * (size_t)B, c->stream));
    float* dst[RC_MAX_BATCH];
    if (nb) {
        // double-buffered: the previous super-block's flows may still be on their way to the band owners (comm stream)
        CU_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_comm[half], 0));
        for (int j = 0; j < nb; j++) dst[j] = c->d_shard_flows + ((size_t)half * B + j) * ff;
        // the block's first frame primes the pyramid / expansion cache (it is the previous block's last frame: one duplicated
        // frame per block edge); the remaining nb frames complete nb pairs in ONE batched pass
        const uint8_t* d0 = frames; const uint8_t* d1 = frames + frame_stride; size_t ds = step, dfs = frame_stride;
        if (!rc_is_device_ptr(frames)) {
            CU_TRY(c, cudaMemcpy2DAsync(c->d_frames[1], w, frames, step, w, h, cudaMemcpyHostToDevice, c->stream));
            for (int j = 0; j < nb; j++)
                CU_TRY(c, cudaMemcpy2DAsync(c->d_frames[0] + (size_t)j * n, w, frames + (size_t)(j + 1) * frame_stride, step, w, h,
                                            cudaMemcpyHostToDevice, c->stream));
            d0 = c->d_frames[1]; d1 = c->d_frames[0]; ds = w; dfs = n;
        }
        c->frames_seen = 0;
        int produced = rc_run_frames_hist(c, d0, ds, dfs, 1, nullptr);
        if (produced == 0) produced = rc_run_frames_hist(c, d1, ds, dfs, nb, dst);
        if (produced != nb) return rc_fail(c, RC_ERR_CUDA, "flow pass failed%s", "");
    }

    // 2. all-gather of the per-frame counts, exclusive prefix over ranks, next global counters
    if (R > 1) {
        int rc = need_nccl(c); if (rc) return rc;
        NCCL_TRY(c, nccl().AllGather(c->d_hist_delta, c->d_gather, (size_t)B * RC_HIST_CELLS, ncclUint32, comm, c->stream));
        c->launches += 1;
    } else {
        CU_TRY(c, cudaMemcpyAsync(c->d_gather, c->d_hist_delta, sizeof(unsigned int) * RC_HIST_CELLS * (size_t)B, cudaMemcpyDeviceToDevice, c->stream));
    }
    rc_launch_shard_prefix(c, c->d_gather, me, R, B, c->d_hist_global, c->d_hist2d);

    // 3. exact per-frame thresholds, classification into this rank's own accumulator
    if (nb) {
        rc_launch_thresholds_batch(c, c->d_hist2d, c->d_hist_delta, nb, c->d_thr_batch[0], c->d_thr);
        ClassifyBatch cb;
        cb.nb = nb;
        for (int j = 0; j < nb; j++) { cb.flow[j] = dst[j]; cb.old[j] = nullptr; }
        rc_launch_classify_batch(c, cb, w, h, c->d_thr_batch[0], framecount0, c->d_acc, nullptr, nullptr, 0);
    }

    // 4. window mean, sharded by pixel band: band r of every flow goes to rank r (all-to-all), which applies the super-block's
    //    updates to its band in stream order
    //    -- on the communication stream, overlapping the next super-block's flow kernels
    CU_TRY(c, cudaEventRecord(c->ev_flows[half], c->stream));
    CU_TRY(c, cudaStreamWaitEvent(c->s_comm, c->ev_flows[half], 0));
    if (c->shard_W > 0) {
        const size_t my_off = band_off(c, me), my_cnt = band_floats(c, me);
        if (R > 1) {
            NCCL_TRY(c, nccl().GroupStart());
            int off = 0;
            for (int r = 0; r < R; r++) {
                if (r != me) {
                    for (int j = 0; j < nb; j++) NCCL_TRY(c, nccl().Send(dst[j] + band_off(c, r), band_floats(c, r), ncclFloat, r, comm, c->s_comm));
                    for (int j = 0; j < pairs_per_rank[r]; j++)
                        NCCL_TRY(c, nccl().Recv(shard_slot(c, c->shard_pairs + off + j), my_cnt, ncclFloat, r, comm, c->s_comm));
                }
                off += pairs_per_rank[r];
            }
            NCCL_TRY(c, nccl().GroupEnd());
            c->launches += 1;
        }
        for (int j = 0; j < nb; j++)          // this rank's own band of its own flows
            CU_TRY(c, cudaMemcpyAsync(shard_slot(c, c->shard_pairs + before_me + j), dst[j] + my_off, my_cnt * sizeof(float),
                                      cudaMemcpyDeviceToDevice, c->s_comm));
        const int brows = band_row0(c, me + 1) - band_row0(c, me);
        for (int p0 = 0; p0 < total && brows > 0; p0 += RC_MAX_BATCH) {
            ClassifyBatch cb;
            cb.nb = total - p0 < RC_MAX_BATCH ? total - p0 : RC_MAX_BATCH;
            for (int j = 0; j < cb.nb; j++) {
                const long long p = c->shard_pairs + p0 + j;
                cb.flow[j] = shard_slot(c, p);
                cb.old[j] = p >= c->shard_W ? shard_slot(c, p - c->shard_W) : nullptr;
            }
            cudaStream_t keep = c->stream;
            c->stream = c->s_comm;            // the launcher (and its profiling events) follow the context's current stream
            rc_launch_window_batch(c, cb, w, brows, c->d_shard_avg + my_off, c->shard_W);
            c->stream = keep;
        }
    }
    CU_TRY(c, cudaEventRecord(c->ev_comm[half], c->s_comm));
    c->shard_steps++;
    c->shard_pairs += total;
    const cudaError_t le = cudaGetLastError();
    if (!c->launch_err.empty()) { const std::string m = c->launch_err; c->launch_err.clear(); return rc_fail(c, RC_ERR_CUDA, "kernel launch failed: %s", m.c_str()); }
    if (le != cudaSuccess) return rc_fail(c, RC_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(le));
    if (results && nb) {
        CU_TRY(c, cudaMemcpyAsync(c->h_thr[0], c->d_thr_batch[0], sizeof(float) * RC_THR_FLOATS * nb, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        rc_fill_results(c->h_thr[0], nb, 0, nb, results);
    }
    return nb;
}

int rc_shard_report(rc_ctx* c, int framecount, uint8_t* outmask, float* acc_x, int64_t* hist2d)
{
    if (!c) return RC_ERR_INVALID;
    if (!c->shard_configured) return rc_fail(c, RC_ERR_STATE, "rc_shard_configure has not been called%s", "");
    cudaSetDevice(c->device);
    const size_t n = (size_t)c->prm.w * c->prm.h;
    for (int i = 0; i < 2; i++) CU_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_comm[i], 0));
    if (c->nranks > 1) {
        int rc = need_nccl(c); if (rc) return rc;
        NCCL_TRY(c, nccl().AllReduce(c->d_acc, c->d_acc_global, n, ncclFloat, ncclSum, static_cast<ncclComm_t>(c->nccl), c->stream));
        c->launches += 1;
    } else {
        CU_TRY(c, cudaMemcpyAsync(c->d_acc_global, c->d_acc, n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    }
    bool host = false;
    if (outmask) {
        const bool dev = rc_is_device_ptr(outmask);
        uint8_t* d = dev ? outmask : c->d_cls;
        rc_launch_acc_mask(c, c->d_acc_global, n, framecount, d);
        if (!dev) { CU_TRY(c, cudaMemcpyAsync(outmask, d, n, cudaMemcpyDeviceToHost, c->stream)); host = true; }
    }
    if (acc_x) {
        const bool dev = rc_is_device_ptr(acc_x);
        CU_TRY(c, cudaMemcpyAsync(acc_x, c->d_acc_global, n * sizeof(float), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
        host = host || !dev;
    }
    if (hist2d) {          // the stream's cumulative counters (identical on every rank)
        const bool dev = rc_is_device_ptr(hist2d);
        CU_TRY(c, cudaMemcpyAsync(hist2d, c->d_hist_global, RC_HIST_CELLS * 8, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
        host = host || !dev;
    }
    if (host) CU_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_shard_window_get(rc_ctx* c, float* avg)
{
    if (!c || !avg) return RC_ERR_INVALID;
    if (!c->shard_configured || !c->d_shard_avg) return rc_fail(c, RC_ERR_STATE, "no window mean configured (rc_shard_configure)%s", "");
    cudaSetDevice(c->device);
    for (int i = 0; i < 2; i++) CU_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_comm[i], 0));
    if (c->nranks > 1) {          // collective: every rank broadcasts its band; all ranks end up with the whole mean
        int rc = need_nccl(c); if (rc) return rc;
        NCCL_TRY(c, nccl().GroupStart());
        for (int r = 0; r < c->nranks; r++)
            if (band_floats(c, r))
                NCCL_TRY(c, nccl().Broadcast(c->d_shard_avg + band_off(c, r), c->d_shard_avg + band_off(c, r), band_floats(c, r), ncclFloat, r,
                                             static_cast<ncclComm_t>(c->nccl), c->stream));
        NCCL_TRY(c, nccl().GroupEnd());
        c->launches += 1;
    }
    const bool dev = rc_is_device_ptr(avg);
    CU_TRY(c, cudaMemcpyAsync(avg, c->d_shard_avg, sizeof(float) * flow_floats(c), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    if (!dev) CU_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

}  // extern "C"
