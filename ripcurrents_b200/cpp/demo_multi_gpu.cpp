// Multi-GPU from a C++ host, no Python and no NCCL headers: ONE video stream split by frame pair over the GPUs of a box through
// the C ABI's own collective layer (rc_comm_* / rc_shard_*, include/ripcurrents_b200.h).
// One host thread per GPU plays the role a rank (process) would: the same calls work across processes once the 128-byte
// communicator id has been handed around (MPI_Bcast, a socket, a file).
//
//   demo_multi_gpu frames.raw W H N NGPUS B out.bin
// frames.raw: N frames of W*H u8 (the stream).  The stream is consumed in super-blocks of NGPUS*B pairs; rank r takes the
// r-th run of B pairs of each super-block (B+1 frames: one duplicated frame per block edge).  out.bin:
//   int32 npairs; float UPPER[npairs]; int64 histsum[npairs]; float acc[W*H]; uint8 mask[W*H]; float avg[W*H*2]
// (per-frame thresholds in stream order, the all-reduced accumulator, the reporting-point outmask and the window mean) --
// tests/test_gpu_sharded_nccl.py compares them bit for bit with one GPU running the stream sequentially.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/ripcurrents_b200.h"

static void die(rc_ctx* c, const char* what, int rc)
{
    std::fprintf(stderr, "demo_multi_gpu: %s failed: %s (%s)\n", what, rc_error_string(rc), c ? rc_last_error(c) : "");
    std::exit(1);
}
#define CHECK(c, call) do { int _rc = (call); if (_rc < 0) die((c), #call, _rc); } while (0)

int main(int argc, char** argv)
{
    if (argc < 8) { std::fprintf(stderr, "usage: %s frames.raw W H N NGPUS B out.bin\n", argv[0]); return 2; }
    const int W = std::atoi(argv[2]), H = std::atoi(argv[3]), N = std::atoi(argv[4]), R = std::atoi(argv[5]), B = std::atoi(argv[6]);
    const size_t n = (size_t)W * H;
    const int FC0 = 25, WIN = 5;                                  // loop counter of frame 0; sliding-window length
    std::vector<uint8_t> raw(n * N);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(raw.data(), 1, raw.size(), f) != raw.size()) { std::fprintf(stderr, "cannot read frames\n"); return 1; }
    std::fclose(f);
    const int npairs = N - 1;
    std::vector<float> upper(npairs);
    std::vector<int64_t> histsum(npairs);
    std::vector<float> acc(n), avg(2 * n);
    std::vector<uint8_t> mask(n);

    char id[RC_COMM_ID_BYTES];
    if (R > 1) CHECK(nullptr, rc_comm_unique_id(id));             // "rank 0" creates the id, every rank receives the 128 bytes

    auto rank_main = [&](int rank) {
        rc_ctx* c = nullptr;
        CHECK(nullptr, rc_create(&c, rank));
        if (R > 1) CHECK(c, rc_comm_init(c, id, rank, R));        // collective: all threads arrive here
        CHECK(c, rc_flow_configure_batch(c, W, H, 0.5, 2, 3, 2, 15, 1.2, 0, B + 1));       // ripcurrents.cpp:215 parameters
        CHECK(c, rc_shard_configure(c, WIN, 0));
        std::vector<int> ppr(R);
        std::vector<rc_frame_result> res(B);
        for (int s0 = 0; s0 < npairs; s0 += R * B) {
            for (int r = 0; r < R; r++) { int left = npairs - (s0 + r * B); ppr[r] = left < 0 ? 0 : (left < B ? left : B); }
            const int lo = s0 + rank * B, nb = ppr[rank];
            if (nb) {
                CHECK(c, rc_shard_step(c, raw.data() + (size_t)lo * n, W, n, nb + 1, FC0 + lo + 1, ppr.data(), res.data()));
                for (int i = 0; i < nb; i++) { upper[lo + i] = res[i].UPPER; histsum[lo + i] = res[i].histsum; }
            } else {
                CHECK(c, rc_shard_step(c, nullptr, 0, 0, 0, 0, ppr.data(), nullptr));
            }
        }
        // collectives: every rank takes part; rank 0 keeps the results
        std::vector<float> a(n), v(2 * n);
        std::vector<uint8_t> m(n);
        CHECK(c, rc_shard_report(c, FC0 + N - 1, m.data(), a.data(), nullptr));
        CHECK(c, rc_shard_window_get(c, v.data()));
        if (rank == 0) { acc = a; mask = m; avg = v; }
        rc_destroy(c);
    };
    std::vector<std::thread> th;
    for (int r = 0; r < R; r++) th.emplace_back(rank_main, r);
    for (auto& t : th) t.join();

    FILE* o = std::fopen(argv[7], "wb");
    if (!o) return 1;
    std::fwrite(&npairs, sizeof npairs, 1, o);
    std::fwrite(upper.data(), sizeof(float), npairs, o);
    std::fwrite(histsum.data(), sizeof(int64_t), npairs, o);
    std::fwrite(acc.data(), sizeof(float), n, o);
    std::fwrite(mask.data(), 1, n, o);
    std::fwrite(avg.data(), sizeof(float), 2 * n, o);
    std::fclose(o);
    std::printf("demo_multi_gpu ok: %d pairs on %d GPU(s)\n", npairs, R);
    return 0;
}
