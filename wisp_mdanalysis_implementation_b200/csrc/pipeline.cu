// pipeline.cu -- wisp_pipeline: coordinates -> cost matrix in ONE call, on 1..8 B200s of one box.
//
// Replaces the three stage calls of the driver (wisp_w_mdanalysis.py:86-97:
// traj_alignment_and_averaging, cartesian_covariance + pearson_correlation_analysis,
// calc_contact_map) and func_pearson_correlation (:141-144) without the node trajectory or the
// 3N x 3N covariance ever leaving the device.
//
// A context made by wisp_ctx_create_multi owns one child context (device, streams, scratch) per GPU.
// Frames are the data-parallel axis: shard r = frames [T r / R, T (r+1) / R), one host
// thread per GPU, every shard streams ITS frames over ITS PCIe link:
//   S1  chunked H2D double-buffered against K1 (translate + node COM, + the alignment landmarks);
//   S2  K7 iterative alignment (trajectory_analysis_functions.py:86-126): every shard rotates its
//       frames onto the current landmark average; the new [landmark || node] column sums are
//       combined over the shards in shard order on the host (a few KB per iteration) and every
//       shard takes the same convergence decision from the same numbers;
//   S3  mean -> centring (+ FP32 tile copy) -> K2 Gram -> K3 distance sums: partial [G], [D] per GPU;
//   S4  ONE fused kernel per GPU does the cross-GPU reduction AND the epilogue for its row slice of
//       the outputs: it reads the R partial sums of each entry straight from the peers' HBM over
//       NVLink (P2P loads, fixed shard order -> every GPU would compute identical bits), applies K4
//       (Pearson, mean distance, strict-< map, -log|C|) and the slice goes to the caller's host
//       buffers over that GPU's own PCIe link.  No all-reduce buffer, no all-gather, no NCCL.
// With one GPU the same code runs inline (no threads, no peers).
#include "common.cuh"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <chrono>
#include <thread>

namespace {

double host_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Host barrier over the shard threads.  abort() releases everybody, now and at every later wait():
// a shard that fails can never leave the others blocked inside a reduction.
class HostBarrier {
  public:
    explicit HostBarrier(int n) : n_(n) {}
    bool wait() {                        // false: somebody aborted
        if (n_ <= 1) return !aborted_;
        std::unique_lock<std::mutex> lk(m_);
        if (aborted_) return false;
        const uint64_t gen = gen_;
        if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen_ != gen || aborted_; });
        return !aborted_;
    }
    void abort() {
        std::lock_guard<std::mutex> lk(m_);
        aborted_ = true;
        cv_.notify_all();
    }
  private:
    int n_, count_ = 0;
    uint64_t gen_ = 0;
    bool aborted_ = false;
    std::mutex m_;
    std::condition_variable cv_;
};

struct PipeArgs {
    const float* coords; int64_t T; int A; const double* masses; const int32_t* node_ptr;
    const int32_t* atom_idx; int N; const int32_t* align_idx; int n_align; int atomic;
    double align_threshold; int max_align_iterations; double cutoff; int which_map;
    double *out_avg, *out_corr, *out_avgdist; int64_t* out_binary; double *out_w, *out_nodes;
    double* out_log; int* out_n_iter;
};

struct Shared {
    int R = 1;
    HostBarrier bar;
    std::atomic<int> failed{0};
    std::mutex err_m;
    std::string err;
    std::vector<std::vector<double>> slots;     // per-shard host vectors for the small reductions
    std::vector<const double*> d_gram, d_dsum;  // per-shard partial sums (device pointers)
    std::vector<int> can_p2p;                   // [R*R]
    wisp_ctx* root = nullptr;                   // stage timings of shard 0 go to root->pipe_stats
    std::vector<wisp_ctx*> ctxs;                // shard -> context
    std::vector<int> via;                       // shard -> shard whose PCIe link carries its coordinates
    double t_begin = 0.0;
    explicit Shared(int r) : R(r), bar(r), slots(r), d_gram(r, nullptr), d_dsum(r, nullptr) {}
};

// sum `buf` over the shards in shard order; every shard ends with the same numbers
int reduce_host(Shared& S, int r, double* buf, size_t n) {
    if (S.R == 1) return 0;
    S.slots[r].assign(buf, buf + n);
    WISP_CHECK(S.bar.wait(), "pipeline: another shard failed");
    for (size_t k = 0; k < n; ++k) {
        double s = S.slots[0][k];
        for (int q = 1; q < S.R; ++q) s += S.slots[q][k];
        buf[k] = s;
    }
    WISP_CHECK(S.bar.wait(), "pipeline: another shard failed");
    return 0;
}

// the shards run `step` one after the other in frame order, each starting from its predecessor's totals
int chain_host(Shared& S, int r, double* buf, size_t n, const AlignStep& step) {
    for (int q = 0; q < S.R; ++q) {
        if (q == r) {
            if (q > 0) memcpy(buf, S.slots[q - 1].data(), n * 8);
            WISP_TRY(step(buf));
            S.slots[r].assign(buf, buf + n);
        }
        WISP_CHECK(S.bar.wait(), "pipeline: another shard failed");
    }
    memcpy(buf, S.slots[S.R - 1].data(), n * 8);
    WISP_CHECK(S.bar.wait(), "pipeline: another shard failed");
    return 0;
}

int shard_body(Shared& S, int r, wisp_ctx* ctx, const PipeArgs& a) {
    const int R = S.R, N = a.N;
    const int64_t t0 = a.T * r / R, t1 = a.T * (r + 1) / R, Tl = t1 - t0;
    const int64_t ld = wisp_node_ld(N), n_pad = round_up(N, kTile);
    const int64_t na3 = 3 * (int64_t)a.n_align, nn3 = 3 * (int64_t)N;
    const size_t nn = (size_t)N * N;
    WISP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const double t_begin = host_ms();
    // small device vectors + the partial sums live in this shard's arena
    void* pw = nullptr;
    const size_t work_d = (size_t)(2 * round_up(na3, 32) + round_up(nn3, 32)) + 3 * (size_t)ld;
    WISP_TRY(wisp_scratch(ctx, SCR_MISC1, work_d * 8, &pw));
    double* d_work = (double*)pw;
    double* d_ref = d_work + (2 * round_up(na3, 32) + round_up(nn3, 32));
    double* d_mean = d_ref + 2 * ld;
    void *pg = nullptr, *pd = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_PIPE_G, (size_t)n_pad * n_pad * 8, &pg));
    WISP_TRY(wisp_scratch(ctx, SCR_PIPE_D, (size_t)n_pad * n_pad * 8, &pd));
    double* d_gram = (double*)pg;
    double* d_dsum = (double*)pd;
    S.d_gram[r] = d_gram;
    S.d_dsum[r] = d_dsum;
    // ---- S1: H2D || K1 || K3 ----
    // Pair distances do not change under the per-frame rigid motions of the alignment, so K3 -- 60 % of
    // the device time of a pass -- does not have to wait for K7: every coordinate block goes K1 -> FP32
    // tile-referenced copy (references = the block-0 positions of the tiles' middle nodes; any fixed point
    // per tile serves) -> K3, accumulated into the shard's distance sums, while the next block is still
    // crossing PCIe.  The P block is a few MB and never leaves L2.
    float* d_align = nullptr;
    if (a.max_align_iterations > 0) {
        void* pa = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_ALIGN, (size_t)Tl * na3 * 4, &pa));
        d_align = (float*)pa;
    }
    double* d_x = nullptr;
    int64_t ldx = 0;
    wisp_ctx* via = (S.via[r] != r) ? S.ctxs[S.via[r]] : nullptr;
    const double* d_k1_colsum = nullptr;      // node column sums of this shard, out of K1
    ComChunkHook hook = [&](double* dx, int64_t ldc, int64_t c0, int64_t nt) -> int {
        if (c0 == 0) {
            WISP_CUDA(cudaMemsetAsync(d_dsum, 0, (size_t)n_pad * n_pad * 8, st));
            WISP_CUDA(cudaMemcpyAsync(d_ref, dx, (size_t)ldc * 8, cudaMemcpyDeviceToDevice, st));   // frame 0 of the shard
        }
        void* pp = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_P32, (size_t)nt * n_pad * 12, &pp));
        WISP_TRY(launch_p32_convert(ctx, dx + c0 * ldc, nt, ldc, N, d_ref, (float*)pp, st));
        return launch_contact_sum(ctx, (const float*)pp, nt, N, d_ref, d_dsum, st, true);
    };
    WISP_TRY(wisp_com_from_host(ctx, a.coords + (size_t)t0 * a.A * 3, Tl, a.A, a.masses, a.node_ptr, a.atom_idx, N,
                                a.align_idx, a.n_align, a.atomic, &d_x, &ldx, d_align, via, &d_k1_colsum, &hook));
    const double t_s1 = host_ms();
    // ---- S2: K7 ----
    std::vector<double> mean(nn3);
    bool have_mean = false;
    // the alignment leaves the nodes as K1 wrote them and hands back the per-frame total rotations: the one
    // in-place rotation rides in the centring pass (unless the caller wants the aligned trajectory itself)
    const double* d_rot = nullptr;
    if (a.max_align_iterations > 0) {
        int n_iter = 0;
        AlignReduce red = [&](double* buf, size_t n, bool f32, const AlignStep& step) {
            return f32 ? chain_host(S, r, buf, n, step) : reduce_host(S, r, buf, n);
        };
        WISP_TRY(run_iterative_alignment(ctx, d_align, a.n_align, d_x, ld, N, Tl, a.align_threshold,
                                         a.max_align_iterations, d_work, mean.data(), r == 0 ? a.out_log : nullptr,
                                         &n_iter, st, a.T, R > 1 ? red : AlignReduce(), d_k1_colsum,
                                         a.out_nodes ? nullptr : &d_rot));
        if (r == 0 && a.out_n_iter) *a.out_n_iter = n_iter;
        have_mean = true;       // the node average of the last iteration IS the frame mean (:112)
    }
    const double t_s2 = host_ms();
    // ---- S3: mean, centring, K2, K3 ----
    if (!have_mean) {
        WISP_CUDA(cudaMemcpyAsync(mean.data(), d_k1_colsum, nn3 * 8, cudaMemcpyDeviceToHost, st));
        WISP_CUDA(cudaStreamSynchronize(st));
        WISP_TRY(reduce_host(S, r, mean.data(), (size_t)nn3));
        for (auto& v : mean) v /= (double)a.T;
    }
    if (r == 0 && a.out_avg) memcpy(a.out_avg, mean.data(), nn3 * 8);
    WISP_CUDA(cudaMemsetAsync(d_mean, 0, (size_t)ld * 8, st));
    WISP_CUDA(cudaMemcpyAsync(d_mean, mean.data(), nn3 * 8, cudaMemcpyHostToDevice, st));
    if (a.out_nodes)   // un-centred node trajectory, as traj_alignment_and_averaging returns it
        WISP_CUDA(cudaMemcpy2DAsync(a.out_nodes + (size_t)t0 * nn3, (size_t)nn3 * 8, d_x, (size_t)ld * 8, (size_t)nn3 * 8,
                                    (size_t)Tl, cudaMemcpyDeviceToHost, st));
    // centring (+ the deferred rotation) with the INT8 column statistics, no FP32 copy: K3 is already done
    WISP_TRY(launch_center(ctx, d_x, Tl, ld, nn3, d_mean, st, nullptr, N, d_rot, true));
    WISP_TRY(launch_gram(ctx, d_x, Tl, ld, N, 3, d_gram, st));
    WISP_CUDA(cudaStreamSynchronize(st));     // `mean` (host) was the source of an async copy; partials complete
    if (r == 0) {
        double* ps = S.root->pipe_stats;
        ps[0] = t_s1 - t_begin; ps[1] = t_s2 - t_s1; ps[2] = host_ms() - t_s2; ps[7] = (double)R;
        S.t_begin = t_begin;
    }
    return 0;
}

int shard_finish(Shared& S, int r, wisp_ctx* ctx, const PipeArgs& a) {
    const int R = S.R, N = a.N;
    const int64_t n_pad = round_up(N, kTile);
    WISP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const double t_fin = host_ms();
    // row slice of the outputs this GPU produces
    const int row0 = (int)((int64_t)N * r / R), row1 = (int)((int64_t)N * (r + 1) / R), rows = row1 - row0;
    if (rows <= 0) return 0;
    PeerSums ps;
    ps.n = R;
    for (int q = 0; q < R; ++q) {
        const double *g = S.d_gram[q], *d = S.d_dsum[q];
        if (q != r && !S.can_p2p[r * R + q]) {
            // no peer mapping between this pair: stage the partials through a peer copy
            void *sg = nullptr, *sd = nullptr;
            WISP_TRY(wisp_scratch(ctx, SCR_PIPE_STAGE0 + 2 * q, (size_t)n_pad * n_pad * 8, &sg));
            WISP_TRY(wisp_scratch(ctx, SCR_PIPE_STAGE0 + 2 * q + 1, (size_t)n_pad * n_pad * 8, &sd));
            WISP_CUDA(cudaMemcpyPeerAsync(sg, ctx->device, g, ctx->peer_devices[q], (size_t)n_pad * n_pad * 8, st));
            WISP_CUDA(cudaMemcpyPeerAsync(sd, ctx->device, d, ctx->peer_devices[q], (size_t)n_pad * n_pad * 8, st));
            g = (const double*)sg;
            d = (const double*)sd;
        }
        ps.gram[q] = g;
        ps.dsum[q] = d;
    }
    const size_t sl = (size_t)rows * N;
    void* po = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_PIPE_OUT, sl * 8 * 4 + (size_t)N * 8 + 1024, &po));
    double* d_corr = (double*)po;
    double* d_avgdist = d_corr + sl;
    double* d_w = d_avgdist + sl;
    int64_t* d_binary = (int64_t*)(d_w + sl);
    double* d_diag = (double*)(d_binary + sl);
    WISP_TRY(launch_epilogue_peers(ctx, ps, a.T, N, row0, row1, a.cutoff, a.which_map, d_diag, d_corr, d_avgdist,
                                   d_binary, d_w, st));
    const size_t off = (size_t)row0 * N;
    if (a.out_corr) WISP_CUDA(cudaMemcpyAsync(a.out_corr + off, d_corr, sl * 8, cudaMemcpyDeviceToHost, st));
    if (a.out_avgdist) WISP_CUDA(cudaMemcpyAsync(a.out_avgdist + off, d_avgdist, sl * 8, cudaMemcpyDeviceToHost, st));
    if (a.out_binary) WISP_CUDA(cudaMemcpyAsync(a.out_binary + off, d_binary, sl * 8, cudaMemcpyDeviceToHost, st));
    if (a.out_w) WISP_CUDA(cudaMemcpyAsync(a.out_w + off, d_w, sl * 8, cudaMemcpyDeviceToHost, st));
    WISP_CUDA(cudaStreamSynchronize(st));
    if (r == 0) {
        double* ps = S.root->pipe_stats;
        ps[3] = t_fin - (S.t_begin + ps[0] + ps[1] + ps[2]);      // waiting for the slowest shard
        ps[4] = host_ms() - t_fin;
        ps[5] = host_ms() - S.t_begin;
    }
    return 0;
}

void shard_thread(Shared* S, int r, wisp_ctx* ctx, const PipeArgs* a) {
    auto fail = [&]() {
        {
            std::lock_guard<std::mutex> lk(S->err_m);
            if (!S->failed.exchange(1)) S->err = std::string("shard ") + std::to_string(r) + ": " + wisp_last_error();
        }
        S->bar.abort();                  // nobody stays blocked at a barrier or inside a reduction
    };
    if (shard_body(*S, r, ctx, *a)) { fail(); return; }
    if (!S->bar.wait()) return;          // all partial sums complete
    if (shard_finish(*S, r, ctx, *a)) { fail(); return; }
    S->bar.wait();                       // nobody reuses its partials while a peer still reads them
}

}  // namespace

int wisp_pipeline_run(wisp_ctx* ctx, const float* coords, int64_t T, int A, const double* masses,
                      const int32_t* node_ptr, const int32_t* atom_idx, int N, const int32_t* align_idx,
                      int n_align, int atomic, double align_threshold, int max_align_iterations,
                      double cutoff, int which_map, double* out_avg, double* out_corr, double* out_avgdist,
                      int64_t* out_binary, double* out_w, double* out_nodes, double* out_log, int* out_n_iter) {
    PipeArgs a{coords, T, A, masses, node_ptr, atom_idx, N, align_idx, n_align, atomic, align_threshold,
               max_align_iterations, cutoff, which_map, out_avg, out_corr, out_avgdist, out_binary, out_w,
               out_nodes, out_log, out_n_iter};
    std::vector<wisp_ctx*> ctxs;
    ctxs.push_back(ctx);
    for (wisp_ctx* c : ctx->children) ctxs.push_back(c);
    int R = (int)std::min<int64_t>((int64_t)ctxs.size(), T);     // never an empty shard
    R = std::min(R, N);
    ctxs.resize(R);
    Shared S(R);
    S.root = ctx;
    S.ctxs = ctxs;
    S.via.resize(R);
    for (int r = 0; r < R; ++r) {
        const int v = (int)ctx->ingest_via.size() == (int)ctx->children.size() + 1 ? ctx->ingest_via[r] : r;
        S.via[r] = (v >= 0 && v < R) ? v : r;
    }
    S.can_p2p.assign((size_t)R * R, 1);
    for (int p = 0; p < R; ++p)
        for (int q = 0; q < R; ++q)
            if (p != q) S.can_p2p[p * R + q] = ctx->peer_ok.empty() ? 0 : ctx->peer_ok[p * (int)(ctx->children.size() + 1) + q];
    if (R == 1) {
        // inline: a failure cannot leave anybody at a barrier
        WISP_TRY(shard_body(S, 0, ctx, a));
        return shard_finish(S, 0, ctx, a);
    }
    std::vector<std::thread> th;
    for (int r = 0; r < R; ++r) th.emplace_back(shard_thread, &S, r, ctxs[r], &a);
    for (auto& t : th) t.join();
    WISP_CUDA(cudaSetDevice(ctx->device));
    if (S.failed.load()) { wisp_set_error("pipeline: %s", S.err.c_str()); return 1; }
    int64_t launches = 0;
    for (wisp_ctx* c : ctx->children) { launches += c->launches - c->launches_reported; c->launches_reported = c->launches; }
    ctx->launches += launches;
    return 0;
}

// ---------------------------------------------------------------------------------------
// multi-GPU context
// ---------------------------------------------------------------------------------------
namespace {

// Concurrent host -> device rate of every GPU in `active` (all pulling at once from one page-locked
// buffer); rates in GB/s, 0 for inactive GPUs; returns the aggregate (total bytes / slowest finish).
double probe_h2d(const std::vector<int>& devs, const std::vector<int>& active, void* h_buf, size_t bytes,
                 std::vector<void*>& d_buf, std::vector<cudaStream_t>& st, std::vector<double>* rates) {
    const int n = (int)devs.size(), reps = 3;
    std::vector<cudaEvent_t> e0(n, nullptr), e1(n, nullptr);
    for (int i = 0; i < n; ++i) {
        if (!active[i]) continue;
        cudaSetDevice(devs[i]);
        cudaEventCreate(&e0[i]); cudaEventCreate(&e1[i]);
        cudaMemcpyAsync(d_buf[i], h_buf, bytes, cudaMemcpyHostToDevice, st[i]);     // warm-up
    }
    for (int i = 0; i < n; ++i) if (active[i]) { cudaSetDevice(devs[i]); cudaStreamSynchronize(st[i]); }
    for (int i = 0; i < n; ++i) {
        if (!active[i]) continue;
        cudaSetDevice(devs[i]);
        cudaEventRecord(e0[i], st[i]);
        for (int k = 0; k < reps; ++k) cudaMemcpyAsync(d_buf[i], h_buf, bytes, cudaMemcpyHostToDevice, st[i]);
        cudaEventRecord(e1[i], st[i]);
    }
    double slowest = 0.0;
    int n_active = 0;
    rates->assign(n, 0.0);
    for (int i = 0; i < n; ++i) {
        if (!active[i]) continue;
        cudaSetDevice(devs[i]);
        cudaStreamSynchronize(st[i]);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0[i], e1[i]);
        (*rates)[i] = reps * (double)bytes / (ms * 1e-3) / 1e9;
        slowest = std::max(slowest, (double)ms);
        ++n_active;
        cudaEventDestroy(e0[i]); cudaEventDestroy(e1[i]);
    }
    return slowest > 0 ? n_active * reps * (double)bytes / (slowest * 1e-3) / 1e9 : 0.0;
}

// Which PCIe links should carry the coordinates?  On some hosts a group of GPUs shares an upstream
// limit (measured on this pool: GPUs 0-3 get 29 GB/s each when they pull together, GPUs 4-7 55 GB/s each,
// and all eight together only 186 GB/s against 220 GB/s for GPUs 4-7 alone).  When pulling through the
// fast GPUs only beats pulling through all of them, every slow GPU is fed by a fast partner: host -> the
// partner's HBM over the partner's link -> NVLink peer copy (700 GB/s, off the critical path).
void probe_ingest_topology(wisp_ctx* root, const std::vector<int>& devs, bool force_relay) {
    const int n = (int)devs.size();
    const size_t bytes = (size_t)192 << 20;
    void* h_buf = nullptr;
    if (cudaHostAlloc(&h_buf, bytes, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return; }
    memset(h_buf, 0, bytes);
    std::vector<void*> d_buf(n, nullptr);
    std::vector<cudaStream_t> st(n, nullptr);
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
        cudaSetDevice(devs[i]);
        ok &= cudaMalloc(&d_buf[i], bytes) == cudaSuccess;
        ok &= cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking) == cudaSuccess;
    }
    if (ok) {
        std::vector<int> all(n, 1), fast(n, 0);
        std::vector<double> r_all, r_fast;
        const double agg_all = probe_h2d(devs, all, h_buf, bytes, d_buf, st, &r_all);
        const double best = *std::max_element(r_all.begin(), r_all.end());
        int n_fast = 0;
        for (int i = 0; i < n; ++i) { fast[i] = r_all[i] >= 0.8 * best; n_fast += fast[i]; }
        if (force_relay && n_fast == n) {       // testing on a symmetric box: relay the second half anyway
            for (int i = (n + 1) / 2; i < n; ++i) fast[i] = 0;
            n_fast = (n + 1) / 2;
        }
        double agg_fast = 0.0;
        r_fast.assign(n, 0.0);
        if (n_fast < n && n_fast >= 1) agg_fast = probe_h2d(devs, fast, h_buf, bytes, d_buf, st, &r_fast);
        root->ingest_probe = r_all;
        root->ingest_probe.insert(root->ingest_probe.end(), r_fast.begin(), r_fast.end());
        root->ingest_probe.push_back(agg_all);
        root->ingest_probe.push_back(agg_fast);
        if (n_fast < n && n_fast >= 1 && (force_relay || agg_fast > 1.08 * agg_all)) {
            // every slow GPU gets a fast partner with peer access, dealt round-robin
            std::vector<int> fl;
            for (int i = 0; i < n; ++i) if (fast[i]) fl.push_back(i);
            int k = 0;
            for (int i = 0; i < n; ++i) {
                if (fast[i]) continue;
                for (int tries = 0; tries < (int)fl.size(); ++tries) {
                    const int p = fl[(k + tries) % fl.size()];
                    if (root->peer_ok[p * n + i]) { root->ingest_via[i] = p; k = (k + tries + 1) % (int)fl.size(); break; }
                }
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        cudaSetDevice(devs[i]);
        if (d_buf[i]) cudaFree(d_buf[i]);
        if (st[i]) cudaStreamDestroy(st[i]);
    }
    cudaFreeHost(h_buf);
    (void)cudaGetLastError();
}

}  // namespace

extern "C" int wisp_ctx_create_multi(int n_gpus, const int* devices, wisp_ctx** out) {
    WISP_CHECK(out, "ctx_create_multi: null output pointer");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        wisp_set_error("ctx_create_multi: no CUDA device available (%s) -- this library has no CPU fallback",
                       cudaGetErrorString(e));
        return 1;
    }
    if (n_gpus <= 0) n_gpus = n_dev;                       // 0 = every GPU of the box
    WISP_CHECK(devices || n_gpus <= n_dev, "ctx_create_multi: %d GPUs requested, %d visible", n_gpus, n_dev);
    WISP_CHECK(n_gpus <= kMaxPeers, "ctx_create_multi: at most %d GPUs", kMaxPeers);
    std::vector<int> devs(n_gpus);
    // A device may be listed more than once: the shards then share that GPU (separate streams and
    // scratch, plain device pointers instead of peer mappings).  This is how the sharding protocol is
    // exercised on a one-GPU box (tests); it is not a way to go faster.
    for (int i = 0; i < n_gpus; ++i) {
        devs[i] = devices ? devices[i] : i;
        WISP_CHECK(devs[i] >= 0 && devs[i] < n_dev, "ctx_create_multi: device %d out of range (0..%d)", devs[i], n_dev - 1);
    }
    wisp_ctx* root = nullptr;
    WISP_TRY(wisp_ctx_create(devs[0], &root));
    for (int i = 1; i < n_gpus; ++i) {
        wisp_ctx* c = nullptr;
        if (wisp_ctx_create(devs[i], &c)) { wisp_ctx_destroy(root); return 1; }
        root->children.push_back(c);
    }
    root->peer_devices = devs;
    for (wisp_ctx* c : root->children) c->peer_devices = devs;
    root->peer_ok.assign((size_t)n_gpus * n_gpus, 1);
    for (int p = 0; p < n_gpus; ++p) {
        cudaSetDevice(devs[p]);
        for (int q = 0; q < n_gpus; ++q) {
            if (devs[p] == devs[q]) continue;      // same device: plain pointers
            int ok = 0;
            cudaDeviceCanAccessPeer(&ok, devs[p], devs[q]);
            if (ok) {
                cudaError_t pe = cudaDeviceEnablePeerAccess(devs[q], 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
                (void)cudaGetLastError();
            }
            root->peer_ok[p * n_gpus + q] = ok;
        }
    }
    root->shard_index = 0;
    for (size_t k = 0; k < root->children.size(); ++k) root->children[k]->shard_index = (int)k + 1;
    root->ingest_via.resize(n_gpus);
    for (int i = 0; i < n_gpus; ++i) root->ingest_via[i] = i;
    bool distinct = true;
    for (int i = 0; i < n_gpus; ++i)
        for (int j = 0; j < i; ++j) distinct &= devs[i] != devs[j];
    const char* mode = getenv("WISP_INGEST");            // auto (default) | direct | relay
    const bool force = mode && !strcmp(mode, "relay");
    if (distinct && (n_gpus >= 4 || (force && n_gpus >= 2)) && !(mode && !strcmp(mode, "direct")))
        probe_ingest_topology(root, devs, force);
    cudaSetDevice(devs[0]);
    *out = root;
    return 0;
}

// ingest plan of a multi-GPU context: via [n] (shard -> relaying shard), rates [2n + 2] = concurrent H2D
// GB/s per GPU with every GPU pulling, then with only the fast set pulling (0 for the others), then the
// two aggregates.  Zeros when no probe ran.
extern "C" int wisp_ctx_ingest_info(wisp_ctx* ctx, int* via, double* rates) {
    WISP_CHECK(ctx, "ctx_ingest_info: null context");
    const int n = 1 + (int)ctx->children.size();
    for (int i = 0; i < n; ++i) if (via) via[i] = (int)ctx->ingest_via.size() == n ? ctx->ingest_via[i] : i;
    if (rates)
        for (int i = 0; i < 2 * n + 2; ++i) rates[i] = (int)ctx->ingest_probe.size() == 2 * n + 2 ? ctx->ingest_probe[i] : 0.0;
    return 0;
}

// shard 0's host-clock view of the last wisp_pipeline: [0] H2D || K1, [1] K7, [2] centring + K2 + K3,
// [3] waiting for the slowest shard, [4] fused epilogue + D2H, [5] total (ms), [6] unused, [7] shards
// ([0] includes K3, which runs block by block under the transfer; [2] is centring + K2)
extern "C" int wisp_pipeline_stats(wisp_ctx* ctx, double* out8) {
    WISP_CHECK(ctx && out8, "pipeline_stats: null argument");
    memcpy(out8, ctx->pipe_stats, sizeof(ctx->pipe_stats));
    return 0;
}

extern "C" int wisp_ctx_gpus(wisp_ctx* ctx) { return ctx ? 1 + (int)ctx->children.size() : 0; }

extern "C" int wisp_host_register(void* ptr, size_t bytes) {
    WISP_CHECK(ptr && bytes, "host_register: empty range");
    WISP_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}

extern "C" int wisp_host_unregister(void* ptr) {
    WISP_CHECK(ptr, "host_unregister: null pointer");
    WISP_CUDA(cudaHostUnregister(ptr));
    return 0;
}
