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
#include <cstdlib>
#include <algorithm>

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


// =======================================================================================
// FP32 throughput mode: distance-only Floyd-Warshall with 128-wide rounds.
//   * phase 3 is a 128x128 register-tiled min-plus product, 8x8 outputs per thread, two
//     k-steps per inner iteration: candidates from packed FADD2 (two outputs per issue slot),
//     folded with the 3-input minimum FMNMX3  d = min(d, c1+r1, c2+r2)  -> one instruction per
//     relaxation instead of four (add, compare, 2 selects) in the predecessor-tracking form;
//   * predecessors are recovered afterwards from the distances:
//     pred[i][j] = argmin_u ( dist[i][u] + w[u][j] ) over the in-neighbours u of j
//     (N^2 * degree work on the CSR graph, first minimum wins).
// =======================================================================================
constexpr int kT = 128;            // tile edge / round width
constexpr int kH = 64;             // k-depth staged per pass in phase 3

typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t fw_add2(float c, f32x2_t r) {
    f32x2_t cc, o;
    asm("mov.b64 %0, {%1,%1};" : "=l"(cc) : "f"(c));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(o) : "l"(cc), "l"(r));
    return o;
}
__device__ __forceinline__ float fw_min3(float a, float b, float c) {
    float o;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(o) : "f"(a), "f"(b), "f"(c));
    return o;
}
__device__ __forceinline__ f32x2_t fw_pack(float lo, float hi) {
    f32x2_t o;
    asm("mov.b64 %0, {%1,%2};" : "=l"(o) : "f"(lo), "f"(hi));
    return o;
}
__device__ __forceinline__ void fw_unpack(f32x2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

__global__ void fw32_init_kernel(const double* __restrict__ w, int N, int64_t n_pad, float* __restrict__ dist) {
    const int64_t total = n_pad * n_pad;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / n_pad, c = e - r * n_pad;
        float d = __int_as_float(0x7f800000);
        if (r == c) d = 0.f;
        else if (r < N && c < N) { const double v = w[r * N + c]; if (v != 0.0) d = (float)v; }
        dist[e] = d;
    }
}

// phase 1: pivot tile, 1024 threads, 16 elements each, in-place (row k / column k are fixed
// points of iteration k because d[k][k] == 0 and weights are non-negative)
__global__ void __launch_bounds__(1024)
fw32_phase1_kernel(float* __restrict__ dist, int64_t ld, int kb) {
    // `dist` is the pivot ROW PANEL (128 rows x ld): the pivot tile sits at column block kb
    extern __shared__ __align__(16) unsigned char fw_smem[];
    float* d = reinterpret_cast<float*>(fw_smem);     // [128][129]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // ty 0..31 -> rows 4ty..4ty+3
    const int64_t base = (int64_t)kb * kT;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) d[(4 * ty + a) * (kT + 1) + tx + 32 * b] = dist[base + (4 * ty + a) * ld + tx + 32 * b];
    __syncthreads();
    // own 16 elements live in registers; shared memory is only the exchange medium for row k /
    // column k, written when an element improves (row k / column k are fixed points of step k)
    float own[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) own[a][b] = d[(4 * ty + a) * (kT + 1) + tx + 32 * b];
    for (int k = 0; k < kT; ++k) {
        float rk[4], ck[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) rk[b] = d[k * (kT + 1) + tx + 32 * b];
#pragma unroll
        for (int a = 0; a < 4; ++a) ck[a] = d[(4 * ty + a) * (kT + 1) + k];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float cand = ck[a] + rk[b];
                if (cand < own[a][b]) { own[a][b] = cand; d[(4 * ty + a) * (kT + 1) + tx + 32 * b] = cand; }
            }
        __syncthreads();
    }
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) dist[base + (4 * ty + a) * ld + tx + 32 * b] = own[a][b];
}

// phase 2: pivot row tiles (blockIdx.y == 0, inside the row panel) and pivot column tiles
// (blockIdx.y == 1, tile (bi, kb) of the LOCAL row block; local tile-row `skip` holds the pivot
// rows themselves and is left alone)
__global__ void __launch_bounds__(1024)
fw32_phase2_kernel(float* __restrict__ local, int64_t ld, float* __restrict__ panel, int64_t ldp,
                   int kb, int n_col_tiles, int n_local_tiles, int skip) {
    const int other = blockIdx.x;
    const bool is_row = blockIdx.y == 0;
    if (is_row) { if (other >= n_col_tiles || other == kb) return; }
    else        { if (other >= n_local_tiles || other == skip) return; }
    extern __shared__ __align__(16) unsigned char fw_smem[];
    float* pv = reinterpret_cast<float*>(fw_smem);     // pivot [128][129]
    float* ow = pv + kT * (kT + 1);                    // own   [128][129]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* pbase = panel + (int64_t)kb * kT;
    float* obase = is_row ? panel + (int64_t)other * kT : local + (int64_t)other * kT * ld + (int64_t)kb * kT;
    const int64_t ldo = is_row ? ldp : ld;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            const int i = 4 * ty + a, j = tx + 32 * b;
            pv[i * (kT + 1) + j] = pbase[i * ldp + j];
            ow[i * (kT + 1) + j] = obase[i * ldo + j];
        }
    __syncthreads();
    float own[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) own[a][b] = ow[(4 * ty + a) * (kT + 1) + tx + 32 * b];
    for (int k = 0; k < kT; ++k) {
        float rk[4], ck[4];
        if (is_row) {      // own = (kb, other): d[i][j] <- pivot[i][k] + own[k][j]
#pragma unroll
            for (int b = 0; b < 4; ++b) rk[b] = ow[k * (kT + 1) + tx + 32 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a) ck[a] = pv[(4 * ty + a) * (kT + 1) + k];
        } else {           // own = (other, kb): d[i][j] <- own[i][k] + pivot[k][j]
#pragma unroll
            for (int b = 0; b < 4; ++b) rk[b] = pv[k * (kT + 1) + tx + 32 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a) ck[a] = ow[(4 * ty + a) * (kT + 1) + k];
        }
        // (row k / column k of the own tile are fixed points of step k: pivot[k][k] == 0, so no
        //  barrier is needed between these reads and the writes below)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float cand = ck[a] + rk[b];
                if (cand < own[a][b]) { own[a][b] = cand; ow[(4 * ty + a) * (kT + 1) + tx + 32 * b] = cand; }
            }
        __syncthreads();
    }
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            const int i = 4 * ty + a, j = tx + 32 * b;
            obase[i * ldo + j] = own[a][b];
        }
}

// phase 3: all tiles outside the pivot row / column.  The 128-deep k range is streamed in four
// 32-deep chunks through a 3-stage cp.async ring (column panel kept i-major: no transpose),
// so the panel loads of chunk s+2 overlap the min-plus arithmetic of chunk s.
constexpr int kC = 32;                 // k-depth per stage
constexpr int kColLd = kC + 4;         // padded row of the column-panel stage (floats)
constexpr int kP3Stages = 3;
constexpr int kP3StageFloats = kT * kColLd + kC * kT;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(256, 2)
fw32_phase3_kernel(float* __restrict__ dist, int64_t ld, const float* __restrict__ panel, int64_t ldp,
                   int kb, int skip) {
    // dist = LOCAL row block; tile (bi, bj) <- min-plus of local column tile (bi, kb) and the
    // pivot row panel's tile (., bj); local tile-row `skip` (the pivot rows) is left alone
    const int bj = blockIdx.x, bi = blockIdx.y;
    if (bi == skip || bj == kb) return;
    extern __shared__ __align__(16) unsigned char fw_smem[];
    float* stages = reinterpret_cast<float*>(fw_smem);
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;            // cols {4tx.., 64+4tx..}, rows {4ty.., 64+4ty..}
    const int64_t obase = (int64_t)bi * kT * ld + (int64_t)bj * kT;
    const float* cg = dist + (int64_t)bi * kT * ld + (int64_t)kb * kT;    // column tile (bi, kb): [i][k]
    const float* rg = panel + (int64_t)bj * kT;                            // row tile (kb, bj):   [k][j]

    auto issue = [&](int chunk) {
        float* cs = stages + (size_t)(chunk % kP3Stages) * kP3StageFloats;     // [128][kColLd]
        float* rs = cs + kT * kColLd;                                            // [kC][128]
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + 256 * q;              // 1024 16-byte granules per panel
            const int i = e >> 3, g = e & 7;          // column panel: row i, granule g (4 floats of k)
            cp_async16(cs + i * kColLd + 4 * g, cg + i * ld + chunk * kC + 4 * g);
            const int k = e >> 5, h = e & 31;         // row panel: row k, granule h (4 floats of j)
            cp_async16(rs + k * kT + 4 * h, rg + (int64_t)(chunk * kC + k) * ldp + 4 * h);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    issue(0);
    issue(1);
    float d[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int row = (a < 4) ? 4 * ty + a : 64 + 4 * ty + (a - 4);
        const float4 v0 = *reinterpret_cast<const float4*>(dist + obase + row * ld + 4 * tx);
        const float4 v1 = *reinterpret_cast<const float4*>(dist + obase + row * ld + 64 + 4 * tx);
        d[a][0] = v0.x; d[a][1] = v0.y; d[a][2] = v0.z; d[a][3] = v0.w;
        d[a][4] = v1.x; d[a][5] = v1.y; d[a][6] = v1.z; d[a][7] = v1.w;
    }
    constexpr int kChunks = kT / kC;
    for (int chunk = 0; chunk < kChunks; ++chunk) {
        if (chunk + 1 < kChunks) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                               // chunk landed for everyone; chunk-1 fully consumed
        if (chunk + 2 < kChunks) issue(chunk + 2);
        const float* cs = stages + (size_t)(chunk % kP3Stages) * kP3StageFloats;
        const float* rs = cs + kT * kColLd;
#pragma unroll 2
        for (int k = 0; k < kC; k += 2) {
            float c1[8], c2[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int row = (a < 4) ? 4 * ty + a : 64 + 4 * ty + (a - 4);
                const float2 cc = *reinterpret_cast<const float2*>(cs + row * kColLd + k);
                c1[a] = cc.x; c2[a] = cc.y;
            }
            const float4 r1a = *reinterpret_cast<const float4*>(rs + k * kT + 4 * tx);
            const float4 r1b = *reinterpret_cast<const float4*>(rs + k * kT + 64 + 4 * tx);
            const float4 r2a = *reinterpret_cast<const float4*>(rs + (k + 1) * kT + 4 * tx);
            const float4 r2b = *reinterpret_cast<const float4*>(rs + (k + 1) * kT + 64 + 4 * tx);
            const f32x2_t r1[4] = {fw_pack(r1a.x, r1a.y), fw_pack(r1a.z, r1a.w), fw_pack(r1b.x, r1b.y), fw_pack(r1b.z, r1b.w)};
            const f32x2_t r2[4] = {fw_pack(r2a.x, r2a.y), fw_pack(r2a.z, r2a.w), fw_pack(r2b.x, r2b.y), fw_pack(r2b.z, r2b.w)};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int bp = 0; bp < 4; ++bp) {
                    float l1, h1, l2, h2;
                    fw_unpack(fw_add2(c1[a], r1[bp]), l1, h1);
                    fw_unpack(fw_add2(c2[a], r2[bp]), l2, h2);
                    d[a][2 * bp] = fw_min3(d[a][2 * bp], l1, l2);
                    d[a][2 * bp + 1] = fw_min3(d[a][2 * bp + 1], h1, h2);
                }
        }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int row = (a < 4) ? 4 * ty + a : 64 + 4 * ty + (a - 4);
        *reinterpret_cast<float4*>(dist + obase + row * ld + 4 * tx) = make_float4(d[a][0], d[a][1], d[a][2], d[a][3]);
        *reinterpret_cast<float4*>(dist + obase + row * ld + 64 + 4 * tx) = make_float4(d[a][4], d[a][5], d[a][6], d[a][7]);
    }
}

// ---- CSR of in-edges from the dense cost matrix, and the argmin predecessor pass ----
__global__ void csr_count_kernel(const double* __restrict__ w, int N, int32_t* __restrict__ count) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    int c = 0;
    for (int u = 0; u < N; ++u) c += (u != j && w[(int64_t)u * N + j] != 0.0) ? 1 : 0;
    count[j] = c;
}
__global__ void csr_fill_kernel(const double* __restrict__ w, int N, const int32_t* __restrict__ ptr,
                                int32_t* __restrict__ idx, float* __restrict__ val) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    int e = ptr[j];
    for (int u = 0; u < N; ++u) {
        const double v = w[(int64_t)u * N + j];
        if (u != j && v != 0.0) { idx[e] = u; val[e] = (float)v; ++e; }
    }
}
__global__ void fw32_final_kernel(const float* __restrict__ dist, int N, int64_t n_pad,
                                  const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                  const float* __restrict__ val, double* __restrict__ out_dist,
                                  int32_t* __restrict__ out_pred) {
    const int64_t total = (int64_t)N * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / N), j = (int)(e - (int64_t)i * N);
        const float dij = dist[(int64_t)i * n_pad + j];
        if (out_dist) out_dist[e] = (double)dij;
        if (out_pred) {
            int p = -1;
            if (i != j && dij < __int_as_float(0x7f800000)) {
                float best = __int_as_float(0x7f800000);
                const float* di = dist + (int64_t)i * n_pad;
                for (int q = ptr[j]; q < ptr[j + 1]; ++q) {
                    const int u = idx[q];
                    const float c = di[u] + val[q];
                    if (c < best) { best = c; p = u; }
                }
            }
            out_pred[e] = p;
        }
    }
}

// dense graphs: pred[i][j] = argmin_u (dist[i][u] + w[u][j]) as a tiled min-plus product with
// argmin (64x64 outputs per CTA, 4x4 per thread); used when the CSR form would be ~N^3 anyway
__global__ void __launch_bounds__(256)
fw32_pred_dense_kernel(const float* __restrict__ dist, int64_t n_pad, const double* __restrict__ w, int N,
                       int32_t* __restrict__ out_pred) {
    __shared__ float dt[64][64 + 4];    // dt[u][i]
    __shared__ float wt[64][64];        // wt[u][j]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    const float INF = __int_as_float(0x7f800000);
    float best[4][4];
    int arg[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { best[a][b] = INF; arg[a][b] = -1; }
    for (int u0 = 0; u0 < N; u0 += 64) {
        __syncthreads();
        for (int e = threadIdx.x; e < 64 * 64; e += 256) {
            const int r = e >> 6, c = e & 63;
            // D tile: row i0+r, col u0+c  -> dt[c][r]
            dt[c][r] = dist[(int64_t)(i0 + r) * n_pad + u0 + c];     // padded: always in range
            // W tile: row u0+r, col j0+c  -> wt[r][c]
            const int u = u0 + r, j = j0 + c;
            float v = INF;
            if (u < N && j < N && u != j) { const double x = w[(int64_t)u * N + j]; if (x != 0.0) v = (float)x; }
            wt[r][c] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < 64; ++k) {
            const float4 cv = *reinterpret_cast<const float4*>(&dt[k][4 * ty]);
            const float4 rv = *reinterpret_cast<const float4*>(&wt[k][4 * tx]);
            const float c4[4] = {cv.x, cv.y, cv.z, cv.w}, r4[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float cand = c4[a] + r4[b];
                    if (cand < best[a][b]) { best[a][b] = cand; arg[a][b] = u0 + k; }
                }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = i0 + 4 * ty + a, j = j0 + 4 * tx + b;
            if (i < N && j < N) out_pred[(int64_t)i * N + j] = (i == j) ? -1 : arg[a][b];
        }
}

int run_apsp_f32_fast(wisp_ctx* ctx, const double* d_w, int N, double* d_dist, int32_t* d_pred,
                      cudaStream_t st) {
    const int64_t n_pad = round_up(N, kT);
    const int nb = (int)(n_pad / kT);
    void* pd = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_APSP_D, (size_t)n_pad * n_pad * sizeof(float), &pd));
    float* dist = (float*)pd;
    const int g = (int)std::min<int64_t>((n_pad * n_pad + 255) / 256, (int64_t)ctx->sm_count * 16);
    fw32_init_kernel<<<g, 256, 0, st>>>(d_w, N, n_pad, dist);
    ctx->launches++;
    const int smem1 = kT * (kT + 1) * 4, smem2 = 2 * kT * (kT + 1) * 4;
    const int smem3 = (int)sizeof(float) * kP3Stages * kP3StageFloats;
    WISP_CUDA(cudaFuncSetAttribute(fw32_phase3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    WISP_CUDA(cudaFuncSetAttribute(fw32_phase1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    WISP_CUDA(cudaFuncSetAttribute(fw32_phase2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    for (int kb = 0; kb < nb; ++kb) {
        float* panel = dist + (int64_t)kb * kT * n_pad;     // single GPU: the panel is the pivot rows in place
        fw32_phase1_kernel<<<1, 1024, smem1, st>>>(panel, n_pad, kb);
        ctx->launches++;
        if (nb > 1) {
            fw32_phase2_kernel<<<dim3(nb, 2), 1024, smem2, st>>>(dist, n_pad, panel, n_pad, kb, nb, nb, kb);
            fw32_phase3_kernel<<<dim3(nb, nb), 256, smem3, st>>>(dist, n_pad, panel, n_pad, kb, kb);
            ctx->launches += 2;
        }
    }
    WISP_CUDA(cudaGetLastError());
    // predecessors from the distances (CSR of in-edges built from the dense matrix)
    int32_t *ptr = nullptr, *idx = nullptr;
    float* val = nullptr;
    if (d_pred) {
        void* pp = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_APSP_P, (size_t)(N + 1) * 4 + 256, &pp));
        ptr = (int32_t*)pp;
        csr_count_kernel<<<(N + 127) / 128, 128, 0, st>>>(d_w, N, ptr + 1);
        ctx->launches++;
        std::vector<int32_t> h(N + 1, 0);
        WISP_CUDA(cudaMemcpyAsync(h.data() + 1, ptr + 1, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
        WISP_CUDA(cudaStreamSynchronize(st));
        for (int j = 0; j < N; ++j) h[j + 1] += h[j];
        const size_t E = (size_t)h[N];
        if (E > (size_t)N * N / 8) {
            // dense graph: tiled argmin product instead of the CSR scan
            fw32_pred_dense_kernel<<<dim3((N + 63) / 64, (N + 63) / 64), 256, 0, st>>>(dist, n_pad, d_w, N, d_pred);
            ctx->launches++;
            const int gd = (int)std::min<int64_t>(((int64_t)N * N + 255) / 256, (int64_t)ctx->sm_count * 16);
            fw32_final_kernel<<<gd, 256, 0, st>>>(dist, N, n_pad, nullptr, nullptr, nullptr, d_dist, nullptr);
            ctx->launches++;
            WISP_CUDA(cudaGetLastError());
            return 0;
        }
        void* pe = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_MISC0, E * 8 + 256, &pe));
        idx = (int32_t*)pe;
        val = (float*)((char*)pe + round_up((int64_t)E * 4, 16));
        WISP_CUDA(cudaMemcpyAsync(ptr, h.data(), (size_t)(N + 1) * 4, cudaMemcpyHostToDevice, st));
        csr_fill_kernel<<<(N + 127) / 128, 128, 0, st>>>(d_w, N, ptr, idx, val);
        ctx->launches++;
        WISP_CUDA(cudaStreamSynchronize(st));   // h (host vector) must outlive the H2D copy
    }
    const int g2 = (int)std::min<int64_t>(((int64_t)N * N + 255) / 256, (int64_t)ctx->sm_count * 16);
    fw32_final_kernel<<<g2, 256, 0, st>>>(dist, N, n_pad, ptr, idx, val, d_dist, d_pred);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

int fw32_set_attrs() {
    const int smem1 = kT * (kT + 1) * 4, smem2 = 2 * kT * (kT + 1) * 4;
    const int smem3 = (int)sizeof(float) * kP3Stages * kP3StageFloats;
    WISP_CUDA(cudaFuncSetAttribute(fw32_phase3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    WISP_CUDA(cudaFuncSetAttribute(fw32_phase1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    WISP_CUDA(cudaFuncSetAttribute(fw32_phase2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    return 0;
}

__global__ void fw32_init_rows_kernel(const double* __restrict__ w_rows, int N, int64_t n_pad, int64_t row0,
                                      int64_t rows_pad, float* __restrict__ dist) {
    const int64_t total = rows_pad * n_pad;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t rl = e / n_pad, c = e - rl * n_pad, r = row0 + rl;
        float d = __int_as_float(0x7f800000);
        if (r == c) d = 0.f;
        else if (r < N && c < N) { const double v = w_rows[rl * N + c]; if (v != 0.0) d = (float)v; }
        dist[e] = d;
    }
}

}  // namespace

int launch_apsp(wisp_ctx* ctx, const double* d_w, int N, int precision, double* d_dist,
                int32_t* d_pred, cudaStream_t st) {
    WISP_CHECK(N > 0, "apsp: empty graph");
    WISP_CHECK(precision == 64 || precision == 32, "apsp: precision must be 64 or 32 (got %d)", precision);
    if (precision == 64) return run_apsp<double>(ctx, d_w, N, d_dist, d_pred, st);
    if (getenv("WISP_FW32_TRACKED")) return run_apsp<float>(ctx, d_w, N, d_dist, d_pred, st);
    return run_apsp_f32_fast(ctx, d_w, N, d_dist, d_pred, st);
}


// ---------------------------------------------------------------------------------------
// Row-block sharded FP32 Floyd-Warshall (BASELINE configs[3]): every rank owns a
// block of 128-row tile rows.  Round kb: the owner of pivot rows kb finishes the pivot ROW
// PANEL (phase 1 + row half of phase 2) in a separate buffer and broadcasts it (NCCL, done by
// the caller); every rank then updates its own pivot-column tiles and its remainder tiles
// from that panel.  The single-GPU path above is the same kernels with the panel aliased to
// the pivot rows in place.
// ---------------------------------------------------------------------------------------
extern "C" int wisp_dev_fw32_init_rows(wisp_ctx* ctx, const double* d_w_rows, int N, int64_t row0,
                                       int64_t rows_pad, float* d_dist_local, void* stream) {
    WISP_CHECK(ctx && d_w_rows && d_dist_local && rows_pad % kT == 0, "fw32_init_rows: bad argument");
    cudaStream_t st = wisp_stream(ctx, stream);
    const int64_t n_pad = round_up(N, kT);
    const int g = (int)std::min<int64_t>((rows_pad * n_pad + 255) / 256, (int64_t)ctx->sm_count * 16);
    fw32_init_rows_kernel<<<g, 256, 0, st>>>(d_w_rows, N, n_pad, row0, rows_pad, d_dist_local);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

// phase 1 + pivot-row tiles on a [128][n_pad] panel (owner only)
extern "C" int wisp_dev_fw32_pivot_panel(wisp_ctx* ctx, float* d_panel, int N, int kb, void* stream) {
    WISP_CHECK(ctx && d_panel, "fw32_pivot_panel: bad argument");
    cudaStream_t st = wisp_stream(ctx, stream);
    const int64_t n_pad = round_up(N, kT);
    const int nb = (int)(n_pad / kT);
    WISP_TRY(fw32_set_attrs());
    const int smem1 = kT * (kT + 1) * 4, smem2 = 2 * kT * (kT + 1) * 4;
    fw32_phase1_kernel<<<1, 1024, smem1, st>>>(d_panel, n_pad, kb);
    fw32_phase2_kernel<<<dim3(nb, 1), 1024, smem2, st>>>(d_panel, n_pad, d_panel, n_pad, kb, nb, 0, -1);
    ctx->launches += 2;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

// pivot-column tiles + remainder tiles of a local row block from a received panel;
// skip_local_tile = local tile-row index holding the pivot rows (-1 if they live elsewhere)
extern "C" int wisp_dev_fw32_update_rows(wisp_ctx* ctx, float* d_dist_local, int64_t rows_pad,
                                         const float* d_panel, int N, int kb, int skip_local_tile,
                                         void* stream) {
    WISP_CHECK(ctx && d_dist_local && d_panel && rows_pad % kT == 0, "fw32_update_rows: bad argument");
    cudaStream_t st = wisp_stream(ctx, stream);
    const int64_t n_pad = round_up(N, kT);
    const int nb = (int)(n_pad / kT), nl = (int)(rows_pad / kT);
    WISP_TRY(fw32_set_attrs());
    const int smem2 = 2 * kT * (kT + 1) * 4;
    const int smem3 = (int)sizeof(float) * kP3Stages * kP3StageFloats;
    // column tiles: launch only the y == 1 half (grid.y index 1)
    fw32_phase2_kernel<<<dim3(nl, 2), 1024, smem2, st>>>(d_dist_local, n_pad, const_cast<float*>(d_panel), n_pad,
                                                        kb, 0, nl, skip_local_tile);
    fw32_phase3_kernel<<<dim3(nb, nl), 256, smem3, st>>>(d_dist_local, n_pad, d_panel, n_pad, kb, skip_local_tile);
    ctx->launches += 2;
    WISP_CUDA(cudaGetLastError());
    return 0;
}
