// K5 -- blocked Floyd-Warshall all-pairs shortest paths with predecessors, sm_100a.
//
// The reference has no APSP: get_paths calls networkx Dijkstra once per (source, sink) pair
// and twice more per pair on EVERY cutoff pass (network_analysis_functions.py:30,71-72).
// One min-plus APSP replaces all of them: rows s and t of (dist, pred) are what those calls
// return.  Graph semantics follow networkx 1.x Graph(data=W): entry != 0 <=> edge; the
// diagonal is never an edge.
//
// Classic 3-phase tiling with 64x64 tiles (round kb = pivot block):
//   phase 1  pivot tile (kb,kb) in shared memory, 64 dependent k-steps;
//   phase 2  pivot row tiles (kb,*) and pivot column tiles (*,kb): pivot tile + own tile in
//            shared memory;
//   phase 3  every other tile: min-plus product of its pivot-column tile (staged transposed
//            so a thread's 4 rows are one vector load) and pivot-row tile (+ its predecessor
//            tile), 4x4 outputs per thread held in registers, no synchronisation inside the
//            k loop.  This phase is N^3 (1 - 2/nb) of the work.
// Relaxation: cand = d_ik + d_kj; if (cand < d_ij) { d_ij = cand; pred_ij = pred_kj; }
// (strict '<': the first k that reaches the minimum wins -> deterministic predecessors).
// Templated on float / double: float64 is the parity mode (path lengths to 1e-9), float32 is
// the FP32-pipe mode the north star asks to be measured.
#include "common.cuh"
#include <limits>

namespace {

constexpr int kB = 64;            // tile edge
constexpr int kFwThreads = 256;   // 16 x 16 threads, 4 x 4 elements each

template <typename T>
__device__ __forceinline__ T inf_v() { return std::numeric_limits<T>::infinity(); }
template <> __device__ __forceinline__ float inf_v<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double inf_v<double>() { return __longlong_as_double(0x7ff0000000000000LL); }

template <typename T>
__global__ void fw_init_kernel(const double* __restrict__ w, int N, int64_t n_pad,
                               T* __restrict__ dist, int32_t* __restrict__ pred) {
    const int64_t total = n_pad * n_pad;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / n_pad, c = e - r * n_pad;
        T d = inf_v<T>();
        int32_t p = -1;
        if (r == c) {
            d = (T)0;
        } else if (r < N && c < N) {
            const double v = w[r * N + c];
            if (v != 0.0) { d = (T)v; p = (int32_t)r; }
        }
        dist[e] = d;
        pred[e] = p;
    }
}

template <typename T>
__global__ void fw_final_kernel(const T* __restrict__ dist, const int32_t* __restrict__ pred,
                                int N, int64_t n_pad, double* __restrict__ out_dist,
                                int32_t* __restrict__ out_pred) {
    const int64_t total = (int64_t)N * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / N, c = e - r * N;
        if (out_dist) out_dist[e] = (double)dist[r * n_pad + c];
        if (out_pred) out_pred[e] = pred[r * n_pad + c];
    }
}

// ---- phase 1: pivot tile ----
template <typename T>
__global__ void __launch_bounds__(kFwThreads)
fw_phase1_kernel(T* __restrict__ dist, int32_t* __restrict__ pred, int64_t ld, int kb) {
    extern __shared__ __align__(16) unsigned char fw_smem[];
    T* d = reinterpret_cast<T*>(fw_smem);                       // [64][65]
    int32_t* p = reinterpret_cast<int32_t*>(d + kB * (kB + 1)); // [64][65]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t base = (int64_t)kb * kB * ld + (int64_t)kb * kB;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            const int i = 4 * ty + a, j = 4 * tx + b;
            d[i * (kB + 1) + j] = dist[base + i * ld + j];
            p[i * (kB + 1) + j] = pred[base + i * ld + j];
        }
    __syncthreads();
    for (int k = 0; k < kB; ++k) {
        T nd[4][4];
        int32_t np[4][4];
        bool up[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int i = 4 * ty + a, j = 4 * tx + b;
                const T cand = d[i * (kB + 1) + k] + d[k * (kB + 1) + j];
                up[a][b] = cand < d[i * (kB + 1) + j];
                nd[a][b] = cand;
                np[a][b] = p[k * (kB + 1) + j];
            }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (up[a][b]) {
                    const int i = 4 * ty + a, j = 4 * tx + b;
                    d[i * (kB + 1) + j] = nd[a][b];
                    p[i * (kB + 1) + j] = np[a][b];
                }
        __syncthreads();
    }
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            const int i = 4 * ty + a, j = 4 * tx + b;
            dist[base + i * ld + j] = d[i * (kB + 1) + j];
            pred[base + i * ld + j] = p[i * (kB + 1) + j];
        }
}

// ---- phase 2: pivot row (blockIdx.y == 0) and pivot column (blockIdx.y == 1) tiles ----
template <typename T>
__global__ void __launch_bounds__(kFwThreads)
fw_phase2_kernel(T* __restrict__ dist, int32_t* __restrict__ pred, int64_t ld, int kb) {
    const int other = blockIdx.x;
    if (other == kb) return;
    const bool is_row = blockIdx.y == 0;
    extern __shared__ __align__(16) unsigned char fw_smem[];
    T* pv = reinterpret_cast<T*>(fw_smem);          // pivot dist   [64][65]
    T* ow = pv + kB * (kB + 1);                      // own dist     [64][65]
    int32_t* pp = reinterpret_cast<int32_t*>(ow + kB * (kB + 1));   // pivot pred
    int32_t* op = pp + kB * (kB + 1);                                // own pred
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t pbase = (int64_t)kb * kB * ld + (int64_t)kb * kB;
    const int64_t obase = is_row ? (int64_t)kb * kB * ld + (int64_t)other * kB
                                 : (int64_t)other * kB * ld + (int64_t)kb * kB;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            const int i = 4 * ty + a, j = 4 * tx + b;
            pv[i * (kB + 1) + j] = dist[pbase + i * ld + j];
            pp[i * (kB + 1) + j] = pred[pbase + i * ld + j];
            ow[i * (kB + 1) + j] = dist[obase + i * ld + j];
            op[i * (kB + 1) + j] = pred[obase + i * ld + j];
        }
    __syncthreads();
    for (int k = 0; k < kB; ++k) {
        T nd[4][4];
        int32_t np[4][4];
        bool up[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int i = 4 * ty + a, j = 4 * tx + b;
                T cand;
                int32_t cp;
                if (is_row) {   // own = (kb, other): d[i][j] <- pivot[i][k] + own[k][j]
                    cand = pv[i * (kB + 1) + k] + ow[k * (kB + 1) + j];
                    cp = op[k * (kB + 1) + j];
                } else {        // own = (other, kb): d[i][j] <- own[i][k] + pivot[k][j]
                    cand = ow[i * (kB + 1) + k] + pv[k * (kB + 1) + j];
                    cp = pp[k * (kB + 1) + j];
                }
                up[a][b] = cand < ow[i * (kB + 1) + j];
                nd[a][b] = cand;
                np[a][b] = cp;
            }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (up[a][b]) {
                    const int i = 4 * ty + a, j = 4 * tx + b;
                    ow[i * (kB + 1) + j] = nd[a][b];
                    op[i * (kB + 1) + j] = np[a][b];
                }
        __syncthreads();
    }
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            const int i = 4 * ty + a, j = 4 * tx + b;
            dist[obase + i * ld + j] = ow[i * (kB + 1) + j];
            pred[obase + i * ld + j] = op[i * (kB + 1) + j];
        }
}

// ---- phase 3: remaining tiles (register-tiled min-plus product) ----
template <typename T>
__global__ void __launch_bounds__(kFwThreads)
fw_phase3_kernel(T* __restrict__ dist, int32_t* __restrict__ pred, int64_t ld, int kb) {
    const int bj = blockIdx.x, bi = blockIdx.y;
    if (bi == kb || bj == kb) return;
    extern __shared__ __align__(16) unsigned char fw_smem[];
    T* ct = reinterpret_cast<T*>(fw_smem);          // column tile transposed: ct[k][i]
    T* rt = ct + kB * kB;                            // row tile: rt[k][j]
    int32_t* rp = reinterpret_cast<int32_t*>(rt + kB * kB);   // row tile preds: rp[k][j]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t cbase = (int64_t)bi * kB * ld + (int64_t)kb * kB;   // (bi, kb)
    const int64_t rbase = (int64_t)kb * kB * ld + (int64_t)bj * kB;   // (kb, bj)
    const int64_t obase = (int64_t)bi * kB * ld + (int64_t)bj * kB;
    for (int e = threadIdx.x; e < kB * kB; e += kFwThreads) {
        const int r = e >> 6, c = e & 63;
        ct[c * kB + r] = dist[cbase + r * ld + c];
        rt[r * kB + c] = dist[rbase + r * ld + c];
        rp[r * kB + c] = pred[rbase + r * ld + c];
    }
    T d[4][4];
    int32_t p[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            d[a][b] = dist[obase + (4 * ty + a) * ld + 4 * tx + b];
            p[a][b] = pred[obase + (4 * ty + a) * ld + 4 * tx + b];
        }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kB; ++k) {
        T cv[4], rv[4];
        int32_t pv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) cv[a] = ct[k * kB + 4 * ty + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) { rv[b] = rt[k * kB + 4 * tx + b]; pv[b] = rp[k * kB + 4 * tx + b]; }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const T cand = cv[a] + rv[b];
                if (cand < d[a][b]) { d[a][b] = cand; p[a][b] = pv[b]; }
            }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            dist[obase + (4 * ty + a) * ld + 4 * tx + b] = d[a][b];
            pred[obase + (4 * ty + a) * ld + 4 * tx + b] = p[a][b];
        }
}

template <typename T>
int run_apsp(wisp_ctx* ctx, const double* d_w, int N, double* d_dist, int32_t* d_pred,
             cudaStream_t st) {
    const int64_t n_pad = round_up(N, kB);
    const int nb = (int)(n_pad / kB);
    void *pd = nullptr, *pp = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_APSP_D, (size_t)n_pad * n_pad * sizeof(T), &pd));
    WISP_TRY(wisp_scratch(ctx, SCR_APSP_P, (size_t)n_pad * n_pad * sizeof(int32_t), &pp));
    T* dist = (T*)pd;
    int32_t* pred = (int32_t*)pp;
    const int g = (int)std::min<int64_t>((n_pad * n_pad + 255) / 256, (int64_t)ctx->sm_count * 16);
    fw_init_kernel<T><<<g, 256, 0, st>>>(d_w, N, n_pad, dist, pred);
    ctx->launches++;
    const int smem1 = kB * (kB + 1) * (sizeof(T) + 4);
    const int smem2 = 2 * kB * (kB + 1) * (sizeof(T) + 4);
    const int smem3 = kB * kB * (2 * sizeof(T) + 4);
    WISP_CUDA(cudaFuncSetAttribute(fw_phase1_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    WISP_CUDA(cudaFuncSetAttribute(fw_phase2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    WISP_CUDA(cudaFuncSetAttribute(fw_phase3_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    for (int kb = 0; kb < nb; ++kb) {
        fw_phase1_kernel<T><<<1, kFwThreads, smem1, st>>>(dist, pred, n_pad, kb);
        if (nb > 1) {
            fw_phase2_kernel<T><<<dim3(nb, 2), kFwThreads, smem2, st>>>(dist, pred, n_pad, kb);
            fw_phase3_kernel<T><<<dim3(nb, nb), kFwThreads, smem3, st>>>(dist, pred, n_pad, kb);
            ctx->launches += 2;
        }
        ctx->launches++;
    }
    WISP_CUDA(cudaGetLastError());
    const int g2 = (int)std::min<int64_t>(((int64_t)N * N + 255) / 256, (int64_t)ctx->sm_count * 16);
    fw_final_kernel<T><<<g2, 256, 0, st>>>(dist, pred, N, n_pad, d_dist, d_pred);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_apsp(wisp_ctx* ctx, const double* d_w, int N, int precision, double* d_dist,
                int32_t* d_pred, cudaStream_t st) {
    WISP_CHECK(N > 0, "apsp: empty graph");
    WISP_CHECK(precision == 64 || precision == 32, "apsp: precision must be 64 or 32 (got %d)", precision);
    if (precision == 64) return run_apsp<double>(ctx, d_w, N, d_dist, d_pred, st);
    return run_apsp<float>(ctx, d_w, N, d_dist, d_pred, st);
}
