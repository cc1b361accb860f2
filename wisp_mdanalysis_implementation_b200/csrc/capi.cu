// capi.cu -- C ABI of libwisp_b200.so: context, host-buffer orchestration (H2D / D2H with
// the copies overlapped against the kernels), device-buffer entry points, device synth.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

// ---------------------------------------------------------------------------------------
// errors / context / scratch
// ---------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void wisp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* wisp_last_error(void) { return g_err; }
extern "C" int wisp_abi_version(void) { return WISP_ABI_VERSION; }
extern "C" int64_t wisp_node_ld(int n_nodes) { return 3 * round_up(n_nodes, kTile); }
extern "C" int64_t wisp_frames_pad(int64_t n_frames) { return round_up(n_frames, kFramePad); }
extern "C" int64_t wisp_gram_ld(int n_cols) { return round_up(n_cols, kTile); }

extern "C" int wisp_ctx_create(int device, wisp_ctx** out) {
    WISP_CHECK(out, "ctx_create: null output pointer");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        wisp_set_error("ctx_create: no CUDA device available (%s) -- this library has no CPU fallback",
                       cudaGetErrorString(e));
        return 1;
    }
    WISP_CHECK(device >= 0 && device < n_dev, "ctx_create: device %d out of range (0..%d)", device, n_dev - 1);
    WISP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    WISP_CUDA(cudaGetDeviceProperties(&prop, device));
    WISP_CHECK(prop.major == 10, "ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
               device, prop.major, prop.minor);
    wisp_ctx* ctx = new wisp_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    WISP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    WISP_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    *out = ctx;
    return 0;
}

extern "C" void wisp_ctx_destroy(wisp_ctx* ctx) {
    if (!ctx) return;
    for (wisp_ctx* c : ctx->children) wisp_ctx_destroy(c);
    ctx->children.clear();
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->pg && ctx->pg_free) ctx->pg_free(ctx->pg);
    for (auto& b : ctx->scratch)
        if (b.p) cudaFree(b.p);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->relay_stream) cudaStreamDestroy(ctx->relay_stream);      // (lives on the relaying GPU; the handle is process-wide)
    for (auto& e : ctx->aux_events) if (e) cudaEventDestroy(e);
    if (ctx->align_host) cudaFreeHost(ctx->align_host);
    delete ctx;
}

extern "C" int wisp_sync(wisp_ctx* ctx) {
    WISP_CHECK(ctx, "sync: null context");
    WISP_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    WISP_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int wisp_gram_mode(wisp_ctx* ctx) {
    if (ctx->gram_mode < 0) {
        const char* e = getenv("WISP_K2_MODE");
        ctx->gram_mode = !e ? WISP_GRAM_AUTO : !strcmp(e, "dmma") ? WISP_GRAM_DMMA : !strcmp(e, "i8") ? WISP_GRAM_I8 : WISP_GRAM_AUTO;
    }
    return ctx->gram_mode;
}

extern "C" int wisp_set_gram_mode(wisp_ctx* ctx, int mode) {
    WISP_CHECK(ctx && mode >= 0 && mode <= 2, "set_gram_mode: mode must be 0 (auto), 1 (dmma) or 2 (i8)");
    ctx->gram_mode = mode;
    for (wisp_ctx* c : ctx->children) c->gram_mode = mode;
    return 0;
}

extern "C" int64_t wisp_kernel_launches(wisp_ctx* ctx) { return ctx ? ctx->launches : -1; }

int wisp_scratch(wisp_ctx* ctx, int slot, size_t bytes, void** out) {
    WISP_CHECK(slot >= 0 && slot < SCR_COUNT, "scratch: bad slot %d", slot);
    DevBuf& b = ctx->scratch[slot];
    if (b.bytes < bytes) {
        if (b.p) {
            // earlier work on any stream (ours or the caller's) may still use the old buffer
            WISP_CUDA(cudaDeviceSynchronize());
            WISP_CUDA(cudaFree(b.p));
            b.p = nullptr;
            b.bytes = 0;
        }
        const size_t want = round_up((int64_t)bytes, 1 << 20);
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            wisp_set_error("scratch: cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
            b.p = nullptr;
            return 1;
        }
        b.bytes = want;
    }
    *out = b.p;
    return 0;
}

namespace {

// bump allocator over one scratch slot
struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    template <typename T>
    T* take(size_t count) {
        off = (size_t)round_up((int64_t)off, 256);
        T* p = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return p;
    }
};

struct ArenaPlan {
    size_t bytes = 0;
    void add(size_t b) { bytes = (size_t)round_up((int64_t)bytes, 256) + b; }
};

int arena_open(wisp_ctx* ctx, size_t bytes, Arena* a) {
    void* p = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_ARENA, bytes + 4096, &p));
    a->base = (char*)p;
    a->cap = bytes + 4096;
    a->off = 0;
    return 0;
}

__global__ void div_kernel(const double* __restrict__ in, double d, int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] / d;
}

int dev_div(wisp_ctx* ctx, const double* in, double d, int64_t n, double* out, cudaStream_t st) {
    div_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(in, d, n, out);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

// upload a host (T,N,3) float64 trajectory into the pitched, zero-padded device layout
int upload_nodes(wisp_ctx* ctx, const double* nodes, int64_t T, int N, double** d_x, int64_t* ld_out) {
    const int64_t ld = wisp_node_ld(N), Tpad = wisp_frames_pad(T);
    void* p = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_X, (size_t)Tpad * ld * 8, &p));
    cudaStream_t st = ctx->stream;
    ctx->i8_stats_x = nullptr;                 // new content: statistics of an earlier centring pass are void
    WISP_CUDA(cudaMemsetAsync(p, 0, (size_t)Tpad * ld * 8, st));
    WISP_CUDA(cudaMemcpy2DAsync(p, (size_t)ld * 8, nodes, (size_t)3 * N * 8, (size_t)3 * N * 8, (size_t)T,
                                cudaMemcpyHostToDevice, st));
    *d_x = (double*)p;
    *ld_out = ld;
    return 0;
}

// K1 over a host coordinate block: chunked H2D on the copy stream, double-buffered against
// the COM kernels on the compute stream.
int com_from_host(wisp_ctx* ctx, const float* coords, int64_t T, int A, const double* masses,
                  const int32_t* node_ptr, const int32_t* atom_idx, int N, const int32_t* align_idx,
                  int n_align, int atomic, double** d_x_out, int64_t* ld_out,
                  float* d_align_out = nullptr, const double** d_colsum_out = nullptr) {
    return wisp_com_from_host(ctx, coords, T, A, masses, node_ptr, atom_idx, N, align_idx, n_align, atomic,
                              d_x_out, ld_out, d_align_out, nullptr, d_colsum_out);
}

}  // namespace

// `via` (may be NULL): another GPU's context through which this shard's coordinates travel -- host ->
// via's HBM over via's PCIe link -> this GPU over NVLink (peer copy) -- used when this GPU's own host link
// is the slow one of the box (pipeline.cu: ingest topology).
int wisp_com_from_host(wisp_ctx* ctx, const float* coords, int64_t T, int A, const double* masses,
                       const int32_t* node_ptr, const int32_t* atom_idx, int N, const int32_t* align_idx,
                       int n_align, int atomic, double** d_x_out, int64_t* ld_out, float* d_align_out,
                       wisp_ctx* via, const double** d_colsum_out, const ComChunkHook* on_chunk) {
    const int64_t ld = wisp_node_ld(N), Tpad = wisp_frames_pad(T);
    wisp_com_plan* plan = nullptr;
    WISP_TRY(com_plan_create(ctx, A, masses, node_ptr, atom_idx, N, align_idx, n_align, atomic, &plan));
    void* px = nullptr;
    int rc = wisp_scratch(ctx, SCR_X, (size_t)Tpad * ld * 8, &px);
    if (rc) { com_plan_destroy(ctx, plan); return rc; }
    double* d_x = (double*)px;
    cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
    const size_t frame_bytes = (size_t)A * 12;
    int64_t chunk = std::max<int64_t>(1, ((size_t)128 << 20) / frame_bytes);
    chunk = std::min(chunk, T);
    // d_colsum_out: the node column sums come out of K1 (one row per chunk, summed in chunk order at the
    // end) -- the alignment's initial node average then needs no pass over the trajectory
    const int64_t n_chunks = (T + chunk - 1) / chunk;
    double* d_chunk_sums = nullptr;
    if (d_colsum_out) {
        void* pc = nullptr;
        rc = wisp_scratch(ctx, SCR_COM_COLSUM, (size_t)(n_chunks + 1) * ld * 8, &pc);
        if (rc) { com_plan_destroy(ctx, plan); return rc; }
        d_chunk_sums = (double*)pc;
    }
    void* bufs[2] = {nullptr, nullptr};
    void* relay[2] = {nullptr, nullptr};
    cudaEvent_t copied[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
    // every CUDA call is checked; the first failure is reported with its own message after the
    // events and the plan have been released
    cudaError_t err = cudaMemsetAsync(d_x, 0, (size_t)Tpad * ld * 8, st);
    const char* what = "cudaMemsetAsync(node trajectory)";
    auto step = [&](cudaError_t e, const char* w) { if (err == cudaSuccess && e != cudaSuccess) { err = e; what = w; } };
    rc = wisp_scratch(ctx, SCR_COORDS0, (size_t)chunk * frame_bytes + 64, &bufs[0]);
    if (!rc) rc = wisp_scratch(ctx, SCR_COORDS1, (size_t)chunk * frame_bytes + 64, &bufs[1]);
    for (int b = 0; b < 2 && !rc && err == cudaSuccess; ++b) step(cudaEventCreateWithFlags(&freed[b], cudaEventDisableTiming), "cudaEventCreate");
    cudaStream_t vs = nullptr;
    if (via && !rc && err == cudaSuccess) {
        // relay buffers, stream and the "copied" events live on the relaying GPU (one relayed shard per
        // slot pair: this shard's index picks the slots)
        step(cudaSetDevice(via->device), "cudaSetDevice(relay)");
        const int slot = SCR_RELAY0 + 2 * (ctx->shard_index % kMaxPeers);
        rc = wisp_scratch(via, slot, (size_t)chunk * frame_bytes + 64, &relay[0]);
        if (!rc) rc = wisp_scratch(via, slot + 1, (size_t)chunk * frame_bytes + 64, &relay[1]);
        if (!ctx->relay_stream) step(cudaStreamCreateWithFlags(&ctx->relay_stream, cudaStreamNonBlocking), "cudaStreamCreate(relay)");
        vs = ctx->relay_stream;
        for (int b = 0; b < 2 && !rc && err == cudaSuccess; ++b) step(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming), "cudaEventCreate");
        step(cudaSetDevice(ctx->device), "cudaSetDevice");
    } else {
        for (int b = 0; b < 2 && !rc && err == cudaSuccess; ++b) step(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming), "cudaEventCreate");
    }
    int it = 0;
    for (int64_t t0 = 0; t0 < T && rc == 0 && err == cudaSuccess; t0 += chunk, ++it) {
        const int b = it & 1;
        const int64_t nt = std::min(chunk, T - t0);
        const float* src = coords + (size_t)t0 * A * 3;
        if (via) {
            step(cudaSetDevice(via->device), "cudaSetDevice(relay)");
            if (it >= 2) step(cudaStreamWaitEvent(vs, freed[b], 0), "cudaStreamWaitEvent(relay stream)");
            step(cudaMemcpyAsync(relay[b], src, (size_t)nt * frame_bytes, cudaMemcpyHostToDevice, vs),
                 "cudaMemcpyAsync(coordinates H2D, relay)");
            step(cudaMemcpyPeerAsync(bufs[b], ctx->device, relay[b], via->device, (size_t)nt * frame_bytes, vs),
                 "cudaMemcpyPeerAsync(relay -> shard)");
            step(cudaEventRecord(copied[b], vs), "cudaEventRecord(copied)");
            step(cudaSetDevice(ctx->device), "cudaSetDevice");
        } else {
            if (it >= 2) step(cudaStreamWaitEvent(cs, freed[b], 0), "cudaStreamWaitEvent(copy stream)");
            step(cudaMemcpyAsync(bufs[b], src, (size_t)nt * frame_bytes, cudaMemcpyHostToDevice, cs),
                 "cudaMemcpyAsync(coordinates H2D)");
            step(cudaEventRecord(copied[b], cs), "cudaEventRecord(copied)");
        }
        step(cudaStreamWaitEvent(st, copied[b], 0), "cudaStreamWaitEvent(compute stream)");
        if (err != cudaSuccess) break;
        rc = launch_com_plan(ctx, plan, (const float*)bufs[b], nt, d_x, ld, t0, st, d_align_out,
                             d_chunk_sums ? d_chunk_sums + (size_t)it * ld : nullptr);
        step(cudaEventRecord(freed[b], st), "cudaEventRecord(freed)");
        // whatever the caller wants done with the node rows [t0, t0 + nt) while the next block is in flight
        if (on_chunk && *on_chunk && rc == 0 && err == cudaSuccess) rc = (*on_chunk)(d_x, ld, t0, nt);
    }
    if (d_chunk_sums && rc == 0 && err == cudaSuccess) {
        rc = launch_colsum_finish(ctx, d_chunk_sums, (int)n_chunks, ld, 3 * (int64_t)N, d_chunk_sums + (size_t)n_chunks * ld, st);
        *d_colsum_out = d_chunk_sums + (size_t)n_chunks * ld;
    }
    if (via) {
        step(cudaSetDevice(via->device), "cudaSetDevice(relay)");
        step(cudaStreamSynchronize(vs), "cudaStreamSynchronize(relay stream)");
        step(cudaSetDevice(ctx->device), "cudaSetDevice");
    } else {
        step(cudaStreamSynchronize(cs), "cudaStreamSynchronize(copy stream)");
    }
    step(cudaStreamSynchronize(st), "cudaStreamSynchronize(compute stream)");
    if (via) cudaSetDevice(via->device);
    for (int b = 0; b < 2; ++b) if (copied[b]) cudaEventDestroy(copied[b]);
    cudaSetDevice(ctx->device);
    for (int b = 0; b < 2; ++b) if (freed[b]) cudaEventDestroy(freed[b]);
    com_plan_destroy(ctx, plan);
    if (rc) return rc;
    if (err != cudaSuccess) {
        wisp_set_error("com_from_host: %s failed: %s", what, cudaGetErrorString(err));
        return 1;
    }
    *d_x_out = d_x;
    *ld_out = ld;
    return 0;
}

namespace {

int validate_node_map(int A, const int32_t* node_ptr, const int32_t* atom_idx, int N,
                      const int32_t* align_idx, int n_align) {
    WISP_CHECK(node_ptr && atom_idx && align_idx, "node map: null pointer");
    WISP_CHECK(node_ptr[0] == 0, "node map: node_ptr[0] must be 0");
    for (int n = 0; n < N; ++n)
        WISP_CHECK(node_ptr[n + 1] > node_ptr[n], "node map: node %d has no atoms", n);
    for (int e = 0; e < node_ptr[N]; ++e)
        WISP_CHECK(atom_idx[e] >= 0 && atom_idx[e] < A, "node map: atom index %d out of range", atom_idx[e]);
    for (int e = 0; e < n_align; ++e)
        WISP_CHECK(align_idx[e] >= 0 && align_idx[e] < A, "alignment atom index %d out of range", align_idx[e]);
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------------------
extern "C" int wisp_com_nodes(wisp_ctx* ctx, const float* coords, int64_t T, int A,
                              const double* masses, const int32_t* node_ptr,
                              const int32_t* atom_idx, int N, const int32_t* align_idx,
                              int n_align, int atomic, double* out_nodes, double* out_avg) {
    WISP_CHECK(ctx && coords && masses, "com_nodes: null argument");
    WISP_CHECK(T > 0 && A > 0 && N > 0 && n_align > 0, "com_nodes: empty input");
    WISP_TRY(validate_node_map(A, node_ptr, atom_idx, N, align_idx, n_align));
    WISP_CUDA(cudaSetDevice(ctx->device));
    ArenaPlan plan;
    plan.add((size_t)A * 8); plan.add((size_t)(N + 1) * 4); plan.add((size_t)node_ptr[N] * 4);
    plan.add((size_t)n_align * 4); plan.add((size_t)wisp_node_ld(N) * 8 * 2);
    Arena ar;
    WISP_TRY(arena_open(ctx, plan.bytes + 4096, &ar));
    double* d_x;
    int64_t ld;
    WISP_TRY(com_from_host(ctx, coords, T, A, masses, node_ptr, atom_idx, N, align_idx, n_align,
                           atomic, &d_x, &ld));
    cudaStream_t st = ctx->stream;
    if (out_avg) {
        double* d_sum = ar.take<double>(ld);
        double* d_avg = ar.take<double>(ld);
        WISP_TRY(launch_colsum(ctx, d_x, T, ld, 3 * (int64_t)N, d_sum, st));
        WISP_TRY(dev_div(ctx, d_sum, (double)T, 3 * (int64_t)N, d_avg, st));
        WISP_CUDA(cudaMemcpyAsync(out_avg, d_avg, (size_t)3 * N * 8, cudaMemcpyDeviceToHost, st));
    }
    if (out_nodes)
        WISP_CUDA(cudaMemcpy2DAsync(out_nodes, (size_t)3 * N * 8, d_x, (size_t)ld * 8, (size_t)3 * N * 8,
                                    (size_t)T, cudaMemcpyDeviceToHost, st));
    WISP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int wisp_traj_alignment_and_averaging(wisp_ctx* ctx, const float* coords, int64_t T, int A,
                                                 const double* masses, const int32_t* node_ptr,
                                                 const int32_t* atom_idx, int N,
                                                 const int32_t* align_idx, int n_align, int atomic,
                                                 double threshold, int max_iter, double* out_nodes,
                                                 double* out_avg, double* out_log, int* out_n_iter) {
    WISP_CHECK(ctx && coords && masses, "traj_alignment_and_averaging: null argument");
    WISP_CHECK(T > 0 && A > 0 && N > 0 && n_align > 0, "traj_alignment_and_averaging: empty input");
    WISP_CHECK(max_iter >= 0, "traj_alignment_and_averaging: negative iteration limit");
    WISP_TRY(validate_node_map(A, node_ptr, atom_idx, N, align_idx, n_align));
    WISP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t na3 = 3 * (int64_t)n_align, nn3 = 3 * (int64_t)N;
    ArenaPlan plan;
    plan.add((size_t)A * 8); plan.add((size_t)(N + 1) * 4); plan.add((size_t)node_ptr[N] * 4);
    plan.add((size_t)n_align * 4);
    plan.add((size_t)(2 * round_up(na3, 32) + round_up(nn3, 32)) * 8);
    Arena ar;
    WISP_TRY(arena_open(ctx, plan.bytes + 4096, &ar));
    void* pa = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_ALIGN, (size_t)T * na3 * 4, &pa));
    float* d_align = (float*)pa;
    double* d_x;
    int64_t ld;
    WISP_TRY(com_from_host(ctx, coords, T, A, masses, node_ptr, atom_idx, N, align_idx, n_align,
                           atomic, &d_x, &ld, d_align));
    double* d_work = ar.take<double>((size_t)(2 * round_up(na3, 32) + round_up(nn3, 32)));
    std::vector<double> avg(nn3);
    int n_iter = 0;
    WISP_TRY(run_iterative_alignment(ctx, d_align, n_align, d_x, ld, N, T, threshold, max_iter, d_work,
                                     avg.data(), out_log, &n_iter, st));
    if (out_avg) memcpy(out_avg, avg.data(), (size_t)nn3 * 8);
    if (out_n_iter) *out_n_iter = n_iter;
    if (out_nodes)
        WISP_CUDA(cudaMemcpy2DAsync(out_nodes, (size_t)nn3 * 8, d_x, (size_t)ld * 8, (size_t)nn3 * 8,
                                    (size_t)T, cudaMemcpyDeviceToHost, st));
    WISP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int wisp_cartesian_covariance(wisp_ctx* ctx, const double* nodes, int64_t T, int N,
                                         const double* avg, double* out_cov) {
    WISP_CHECK(ctx && nodes && avg && out_cov, "cartesian_covariance: null argument");
    WISP_CHECK(T > 0 && N > 0, "cartesian_covariance: empty input");
    WISP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n3 = 3 * N;
    const int64_t ldp = wisp_gram_ld(n3);
    double* d_x;
    int64_t ld;
    WISP_TRY(upload_nodes(ctx, nodes, T, N, &d_x, &ld));
    ArenaPlan plan;
    plan.add((size_t)ld * 8); plan.add((size_t)ld * 8); plan.add((size_t)ld * 8);
    plan.add((size_t)ldp * ldp * 8); plan.add((size_t)n3 * n3 * 8);
    Arena ar;
    WISP_TRY(arena_open(ctx, plan.bytes, &ar));
    double* d_avg = ar.take<double>(ld);
    double* d_sum = ar.take<double>(ld);
    double* d_delta = ar.take<double>(ld);
    double* d_p = ar.take<double>((size_t)ldp * ldp);
    double* d_cov = ar.take<double>((size_t)n3 * n3);
    WISP_CUDA(cudaMemsetAsync(d_avg, 0, (size_t)ld * 8, st));
    WISP_CUDA(cudaMemcpyAsync(d_avg, avg, (size_t)n3 * 8, cudaMemcpyHostToDevice, st));
    // centre by the SUPPLIED average (reference :169-170 subtracts a a^T with the supplied a)
    WISP_TRY(launch_center(ctx, d_x, T, ld, n3, d_avg, st));
    WISP_TRY(launch_colsum(ctx, d_x, T, ld, n3, d_sum, st));
    WISP_TRY(dev_div(ctx, d_sum, (double)T, n3, d_delta, st));       // delta = E[x] - a
    WISP_TRY(launch_gram(ctx, d_x, T, ld, n3, 1, d_p, st));
    WISP_TRY(launch_cov_finalize(ctx, d_p, ldp, n3, T, d_avg, d_delta, d_cov, st));
    WISP_CUDA(cudaMemcpyAsync(out_cov, d_cov, (size_t)n3 * n3 * 8, cudaMemcpyDeviceToHost, st));
    WISP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int wisp_pearson_from_covariance(wisp_ctx* ctx, const double* cov, int N,
                                            double* out_var, double* out_ncov, double* out_corr) {
    WISP_CHECK(ctx && cov && N > 0, "pearson_from_covariance: bad argument");
    WISP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n3 = 3 * (size_t)N, nn = (size_t)N * N;
    ArenaPlan plan;
    plan.add(n3 * n3 * 8); plan.add((size_t)N * 8); plan.add(nn * 8); plan.add(nn * 8);
    Arena ar;
    WISP_TRY(arena_open(ctx, plan.bytes, &ar));
    double* d_cov = ar.take<double>(n3 * n3);
    double* d_var = ar.take<double>(N);
    double* d_ncov = ar.take<double>(nn);
    double* d_corr = ar.take<double>(nn);
    WISP_CUDA(cudaMemcpyAsync(d_cov, cov, n3 * n3 * 8, cudaMemcpyHostToDevice, st));
    WISP_TRY(launch_pearson_from_cov(ctx, d_cov, N, d_var, d_ncov, d_corr, st));
    if (out_var) WISP_CUDA(cudaMemcpyAsync(out_var, d_var, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    if (out_ncov) WISP_CUDA(cudaMemcpyAsync(out_ncov, d_ncov, nn * 8, cudaMemcpyDeviceToHost, st));
    if (out_corr) WISP_CUDA(cudaMemcpyAsync(out_corr, d_corr, nn * 8, cudaMemcpyDeviceToHost, st));
    WISP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int wisp_contact_map(wisp_ctx* ctx, const double* nodes, int64_t T, int N, double cutoff,
                                int64_t* out_binary, double* out_avgdist) {
    WISP_CHECK(ctx && nodes, "contact_map: null argument");
    WISP_CHECK(T > 0 && N > 0, "contact_map: empty input");
    WISP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    double* d_x;
    int64_t ld;
    WISP_TRY(upload_nodes(ctx, nodes, T, N, &d_x, &ld));
    const int64_t n_pad = round_up(N, kTile);
    const size_t nn = (size_t)N * N;
    ArenaPlan plan;
    plan.add((size_t)ld * 8); plan.add((size_t)ld * 8); plan.add((size_t)n_pad * n_pad * 8);
    plan.add(nn * 8); plan.add(nn * 8);
    Arena ar;
    WISP_TRY(arena_open(ctx, plan.bytes, &ar));
    double* d_sum = ar.take<double>(ld);
    double* d_mean = ar.take<double>(ld);
    double* d_dsum = ar.take<double>((size_t)n_pad * n_pad);
    double* d_avgdist = ar.take<double>(nn);
    int64_t* d_binary = ar.take<int64_t>(nn);
    WISP_CUDA(cudaMemsetAsync(d_mean, 0, (size_t)ld * 8, st));
    WISP_TRY(launch_colsum(ctx, d_x, T, ld, 3 * (int64_t)N, d_sum, st));
    WISP_TRY(dev_div(ctx, d_sum, (double)T, 3 * (int64_t)N, d_mean, st));
    void* pp = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_P32, (size_t)T * n_pad * 12, &pp));
    WISP_TRY(launch_center(ctx, d_x, T, ld, 3 * (int64_t)N, d_mean, st, (float*)pp, N));
    WISP_TRY(launch_contact_sum(ctx, (const float*)pp, T, N, d_mean, d_dsum, st));
    WISP_TRY(launch_epilogue(ctx, nullptr, d_dsum, T, N, cutoff, 0, nullptr, nullptr, nullptr,
                             d_avgdist, d_binary, nullptr, st));
    if (out_avgdist) WISP_CUDA(cudaMemcpyAsync(out_avgdist, d_avgdist, nn * 8, cudaMemcpyDeviceToHost, st));
    if (out_binary) WISP_CUDA(cudaMemcpyAsync(out_binary, d_binary, nn * 8, cudaMemcpyDeviceToHost, st));
    WISP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int wisp_func_pearson(wisp_ctx* ctx, const double* corr, int N, const double* contact_map,
                                 double* out_w) {
    WISP_CHECK(ctx && corr && out_w && N > 0, "func_pearson: bad argument");
    WISP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t nn = (size_t)N * N;
    ArenaPlan plan;
    plan.add(nn * 8); plan.add(nn * 8); plan.add(nn * 8);
    Arena ar;
    WISP_TRY(arena_open(ctx, plan.bytes, &ar));
    double* d_c = ar.take<double>(nn);
    double* d_m = ar.take<double>(nn);
    double* d_w = ar.take<double>(nn);
    WISP_CUDA(cudaMemcpyAsync(d_c, corr, nn * 8, cudaMemcpyHostToDevice, st));
    if (contact_map) WISP_CUDA(cudaMemcpyAsync(d_m, contact_map, nn * 8, cudaMemcpyHostToDevice, st));
    WISP_TRY(launch_func_pearson(ctx, d_c, N, contact_map ? d_m : nullptr, d_w, st));
    WISP_CUDA(cudaMemcpyAsync(out_w, d_w, nn * 8, cudaMemcpyDeviceToHost, st));
    WISP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int wisp_pipeline(wisp_ctx* ctx, const float* coords, int64_t T, int A,
                             const double* masses, const int32_t* node_ptr,
                             const int32_t* atom_idx, int N, const int32_t* align_idx, int n_align,
                             int atomic, double align_threshold, int max_align_iterations,
                             double cutoff, int which_map, double* out_avg,
                             double* out_corr, double* out_avgdist, int64_t* out_binary,
                             double* out_w, double* out_nodes) {
    return wisp_pipeline_log(ctx, coords, T, A, masses, node_ptr, atom_idx, N, align_idx, n_align, atomic,
                             align_threshold, max_align_iterations, cutoff, which_map, out_avg, out_corr,
                             out_avgdist, out_binary, out_w, out_nodes, nullptr, nullptr);
}

extern "C" int wisp_pipeline_log(wisp_ctx* ctx, const float* coords, int64_t T, int A,
                                 const double* masses, const int32_t* node_ptr,
                                 const int32_t* atom_idx, int N, const int32_t* align_idx, int n_align,
                                 int atomic, double align_threshold, int max_align_iterations,
                                 double cutoff, int which_map, double* out_avg,
                                 double* out_corr, double* out_avgdist, int64_t* out_binary,
                                 double* out_w, double* out_nodes, double* out_log, int* out_n_iter) {
    WISP_CHECK(ctx && coords && masses, "pipeline: null argument");
    WISP_CHECK(max_align_iterations >= 0, "pipeline: negative alignment iteration limit");
    WISP_CHECK(T > 0 && A > 0 && N > 0 && n_align > 0, "pipeline: empty input");
    WISP_CHECK(which_map >= 0 && which_map <= 2, "pipeline: which_map must be 0, 1 or 2");
    WISP_TRY(validate_node_map(A, node_ptr, atom_idx, N, align_idx, n_align));
    return wisp_pipeline_run(ctx, coords, T, A, masses, node_ptr, atom_idx, N, align_idx, n_align, atomic,
                             align_threshold, max_align_iterations, cutoff, which_map, out_avg, out_corr,
                             out_avgdist, out_binary, out_w, out_nodes, out_log, out_n_iter);
}

extern "C" int wisp_apsp(wisp_ctx* ctx, const double* w, int N, int precision, double* out_dist,
                         int32_t* out_pred) {
    WISP_CHECK(ctx && w && out_dist && N > 0, "apsp: bad argument");
    WISP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t nn = (size_t)N * N;
    ArenaPlan plan;
    plan.add(nn * 8); plan.add(nn * 8); plan.add(nn * 4);
    Arena ar;
    WISP_TRY(arena_open(ctx, plan.bytes, &ar));
    double* d_w = ar.take<double>(nn);
    double* d_dist = ar.take<double>(nn);
    int32_t* d_pred = ar.take<int32_t>(nn);
    WISP_CUDA(cudaMemcpyAsync(d_w, w, nn * 8, cudaMemcpyHostToDevice, st));
    WISP_TRY(launch_apsp(ctx, d_w, N, precision, d_dist, d_pred, st));
    WISP_CUDA(cudaMemcpyAsync(out_dist, d_dist, nn * 8, cudaMemcpyDeviceToHost, st));
    if (out_pred) WISP_CUDA(cudaMemcpyAsync(out_pred, d_pred, nn * 4, cudaMemcpyDeviceToHost, st));
    WISP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ---------------------------------------------------------------------------------------
// device-buffer entry points
// ---------------------------------------------------------------------------------------
extern "C" int wisp_com_plan_create(wisp_ctx* ctx, int A, const double* masses, const int32_t* node_ptr,
                                    const int32_t* atom_idx, int N, const int32_t* align_idx,
                                    int n_align, int atomic, wisp_com_plan** out) {
    return com_plan_create(ctx, A, masses, node_ptr, atom_idx, N, align_idx, n_align, atomic, out);
}

extern "C" void wisp_com_plan_destroy(wisp_ctx* ctx, wisp_com_plan* plan) { com_plan_destroy(ctx, plan); }

extern "C" int wisp_com_plan_info(const wisp_com_plan* plan, int* streaming, int* n_pieces) {
    WISP_CHECK(plan, "com_plan_info: null plan");
    return com_plan_info(plan, streaming, n_pieces);
}

extern "C" int wisp_dev_com(wisp_ctx* ctx, wisp_com_plan* plan, const float* d_coords, int64_t T, int N,
                            double* d_nodes, double* d_colsum, void* stream) {
    WISP_CHECK(ctx && plan && d_coords && d_nodes, "dev_com: null argument");
    cudaStream_t st = wisp_stream(ctx, stream);
    const int64_t ld = wisp_node_ld(N);
    return launch_com_plan(ctx, plan, d_coords, T, d_nodes, ld, 0, st, nullptr, d_colsum);
}

extern "C" int wisp_dev_com_align(wisp_ctx* ctx, wisp_com_plan* plan, const float* d_coords, int64_t T, int N,
                                  double* d_nodes, double* d_colsum, float* d_align_out, void* stream) {
    WISP_CHECK(ctx && plan && d_coords && d_nodes, "dev_com_align: null argument");
    return launch_com_plan(ctx, plan, d_coords, T, d_nodes, wisp_node_ld(N), 0, wisp_stream(ctx, stream), d_align_out,
                           d_colsum);
}

extern "C" int wisp_dev_align_sums(wisp_ctx* ctx, const float* d_align, int n_align, const double* d_nodes,
                                   int64_t T, int N, int first, double* d_sums, void* stream) {
    WISP_CHECK(ctx && d_align && d_sums && T > 0, "dev_align_sums: bad argument");
    return launch_align_sums(ctx, d_align, n_align, d_nodes, wisp_node_ld(N), N, T, first != 0, d_sums,
                             wisp_stream(ctx, stream));
}

extern "C" int wisp_dev_align_rotate(wisp_ctx* ctx, float* d_align, int n_align, const double* d_avg_align,
                                     double* d_nodes, int64_t T, int N, double* d_sums, void* stream) {
    WISP_CHECK(ctx && d_align && d_avg_align && d_nodes && T > 0, "dev_align_rotate: bad argument");
    return launch_align_rotate(ctx, d_align, n_align, d_avg_align, d_nodes, wisp_node_ld(N), N, T,
                               wisp_stream(ctx, stream), d_sums);
}

extern "C" int wisp_dev_align_iter(wisp_ctx* ctx, float* d_align, int n_align, const double* d_avg_align,
                                   const double* d_nodes, int64_t T, int N, double* d_rot_total, int first,
                                   double* d_sums, const int* d_gate, void* stream) {
    WISP_CHECK(ctx && d_align && d_avg_align && d_nodes && d_rot_total && T > 0, "dev_align_iter: bad argument");
    return launch_align_iter(ctx, d_align, n_align, d_avg_align, d_nodes, wisp_node_ld(N), N, T, d_rot_total,
                             first != 0, wisp_stream(ctx, stream), d_sums, d_gate);
}

extern "C" int wisp_dev_align_update_gated(wisp_ctx* ctx, const double* d_sums, int64_t T_total, int n_align, int N,
                                           double* d_avg, double* d_res3, double threshold, int* d_gate, void* stream) {
    WISP_CHECK(ctx && d_sums && d_avg && d_res3 && d_gate && T_total > 0, "dev_align_update_gated: bad argument");
    return launch_align_update(ctx, d_sums, T_total, n_align, N, d_avg, d_res3, wisp_stream(ctx, stream), threshold, d_gate);
}

extern "C" int wisp_dev_align_nodes(wisp_ctx* ctx, double* d_nodes, int64_t T, int N, const double* d_rot_total,
                                    void* stream) {
    WISP_CHECK(ctx && d_nodes && d_rot_total && T > 0, "dev_align_nodes: bad argument");
    return launch_align_nodes(ctx, d_nodes, wisp_node_ld(N), N, T, d_rot_total, wisp_stream(ctx, stream));
}

extern "C" int wisp_dev_align_update(wisp_ctx* ctx, const double* d_sums, int64_t T_total, int n_align, int N,
                                     double* d_avg, double* d_res2, void* stream) {
    WISP_CHECK(ctx && d_sums && d_avg && d_res2 && T_total > 0, "dev_align_update: bad argument");
    return launch_align_update(ctx, d_sums, T_total, n_align, N, d_avg, d_res2, wisp_stream(ctx, stream));
}

extern "C" int wisp_dev_center(wisp_ctx* ctx, double* d_nodes, int64_t T, int N, const double* d_mean,
                               float* d_p32, void* stream) {
    WISP_CHECK(ctx && d_nodes && d_mean, "dev_center: null argument");
    return launch_center(ctx, d_nodes, T, wisp_node_ld(N), 3 * (int64_t)N, d_mean, wisp_stream(ctx, stream),
                         d_p32, N);
}

extern "C" int wisp_dev_center_rotated(wisp_ctx* ctx, double* d_nodes, int64_t T, int N, const double* d_mean,
                                       const double* d_rot_total, float* d_p32, void* stream) {
    WISP_CHECK(ctx && d_nodes && d_mean, "dev_center_rotated: null argument");
    return launch_center(ctx, d_nodes, T, wisp_node_ld(N), 3 * (int64_t)N, d_mean, wisp_stream(ctx, stream),
                         d_p32, N, d_rot_total);
}

extern "C" int wisp_dev_colsum(wisp_ctx* ctx, const double* d_nodes, int64_t T, int N, double* d_colsum,
                               void* stream) {
    WISP_CHECK(ctx && d_nodes && d_colsum, "dev_colsum: null argument");
    return launch_colsum(ctx, d_nodes, T, wisp_node_ld(N), 3 * (int64_t)N, d_colsum, wisp_stream(ctx, stream));
}

extern "C" int wisp_dev_gram(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int n_out,
                             int stride, double* d_out, void* stream) {
    WISP_CHECK(ctx && d_x && d_out, "dev_gram: null argument");
    return launch_gram(ctx, d_x, T, ld, n_out, stride, d_out, wisp_stream(ctx, stream));
}

extern "C" int wisp_dev_gram_i8(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int N,
                                double* d_out, void* stream) {
    WISP_CHECK(ctx && d_x && d_out, "dev_gram_i8: null argument");
    return launch_gram_i8(ctx, d_x, T, ld, N, d_out, wisp_stream(ctx, stream), nullptr);
}

extern "C" int wisp_dev_contact_sum(wisp_ctx* ctx, const float* d_p32, int64_t T, int N,
                                    const double* d_mean, double* d_out, void* stream) {
    WISP_CHECK(ctx && d_p32 && d_mean && d_out, "dev_contact_sum: null argument");
    return launch_contact_sum(ctx, d_p32, T, N, d_mean, d_out, wisp_stream(ctx, stream));
}

extern "C" int wisp_dev_epilogue(wisp_ctx* ctx, const double* d_gram, const double* d_dsum,
                                 int64_t T_total, int N, double cutoff, int which_map, double* d_var,
                                 double* d_ncov, double* d_corr, double* d_avgdist, int64_t* d_binary,
                                 double* d_w, const double* d_offset_sum, void* stream) {
    WISP_CHECK(ctx, "dev_epilogue: null context");
    return launch_epilogue(ctx, d_gram, d_dsum, T_total, N, cutoff, which_map, d_var, d_ncov, d_corr,
                           d_avgdist, d_binary, d_w, wisp_stream(ctx, stream), d_offset_sum);
}

extern "C" int wisp_dev_reduce_partials(wisp_ctx* ctx, const double* d_parts, int n_parts, int N,
                                        double* d_out, void* stream) {
    WISP_CHECK(ctx && d_parts && d_out && n_parts > 0, "dev_reduce_partials: bad argument");
    return launch_reduce_partials(ctx, d_parts, n_parts, round_up(N, kTile), d_out, kTile, kTile,
                                  wisp_stream(ctx, stream));
}

extern "C" int wisp_dev_apsp(wisp_ctx* ctx, const double* d_w, int N, int precision, double* d_dist,
                             int32_t* d_pred, void* stream) {
    WISP_CHECK(ctx && d_w, "dev_apsp: null argument");
    return launch_apsp(ctx, d_w, N, precision, d_dist, d_pred, wisp_stream(ctx, stream));
}

// ---------------------------------------------------------------------------------------
// device-side synthetic coordinates (synth.py model; counter-based RNG)
// ---------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void synth_kernel(float* __restrict__ coords, int64_t t0, int64_t T, int A,
                             const double* __restrict__ atom_base, const int32_t* __restrict__ node_of_atom,
                             const double* __restrict__ modes, int n_modes,
                             const double* __restrict__ coeff, double sigma, uint64_t seed) {
    extern __shared__ double cf[];   // coefficients of this frame
    const int64_t f = blockIdx.x;
    if (f >= T) return;
    for (int m = threadIdx.x; m < n_modes; m += blockDim.x) cf[m] = coeff[(t0 + f) * n_modes + m];
    __syncthreads();
    float* out = coords + f * (int64_t)A * 3;
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
        const int n = node_of_atom[a];
        double d[3] = {0, 0, 0};
        for (int m = 0; m < n_modes; ++m) {
            const double* md = modes + ((int64_t)n * n_modes + m) * 3;
            d[0] += cf[m] * md[0]; d[1] += cf[m] * md[1]; d[2] += cf[m] * md[2];
        }
        const uint64_t key = splitmix64(seed ^ splitmix64((uint64_t)(t0 + f) * 0x100000001B3ull + (uint64_t)a));
        const uint64_t k1 = splitmix64(key), k2 = splitmix64(key + 1), k3 = splitmix64(key + 2), k4 = splitmix64(key + 3);
        const double u1 = ((k1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        const double u2 = ((k2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        const double u3 = ((k3 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        const double u4 = ((k4 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        const double r1 = sqrt(-2.0 * log(u1)), r2 = sqrt(-2.0 * log(u3));
        double s1, c1, s2, c2;
        sincospi(2.0 * u2, &s1, &c1);
        sincospi(2.0 * u4, &s2, &c2);
        (void)s2;
        out[3 * (int64_t)a + 0] = (float)(atom_base[3 * (int64_t)a + 0] + d[0] + sigma * r1 * c1);
        out[3 * (int64_t)a + 1] = (float)(atom_base[3 * (int64_t)a + 1] + d[1] + sigma * r1 * s1);
        out[3 * (int64_t)a + 2] = (float)(atom_base[3 * (int64_t)a + 2] + d[2] + sigma * r2 * c2);
    }
}

}  // namespace

extern "C" int wisp_dev_synth_coords(wisp_ctx* ctx, float* d_coords, int64_t t0, int64_t T, int A,
                                     const double* d_atom_base, const int32_t* d_node_of_atom,
                                     const double* d_modes, int n_modes, const double* d_coeff,
                                     double noise_sigma, uint64_t seed, void* stream) {
    WISP_CHECK(ctx && d_coords && T > 0 && A > 0, "dev_synth_coords: bad argument");
    cudaStream_t st = wisp_stream(ctx, stream);
    synth_kernel<<<(unsigned)T, 256, n_modes * sizeof(double), st>>>(d_coords, t0, T, A, d_atom_base,
                                                                    d_node_of_atom, d_modes, n_modes,
                                                                    d_coeff, noise_sigma, seed);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}
