// K3 -- mean pairwise node distance (contact map numerator), sm_100a.
//
// Replaces the reference's T calls of MDAnalysis distance_array on an N x N result plus an
// N^2 numpy add per frame (trajectory_analysis_functions.py:207-215):
//     D[i][j] = sum_t | x_i(t) - x_j(t) |          (divided by T and thresholded in K4)
//
// Input: the FP32 tile-referenced structure-of-arrays copy P of the node trajectory that the
// centring pass writes (k1_com.cu: center_convert_kernel):
//     P[t][tile][c][n] = float( x_{128*tile+n, c}(t) - ref_{tile, c} ),  ref_tile = mean
//     position of the node in the middle of the tile.
// FP32 differences of tile-referenced coordinates keep ~1e-7 relative accuracy at any box
// size (the reference stack itself casts ABSOLUTE coordinates to float32 inside
// distance_array); the reference offset between two tiles is added once per staged j value.
//
// One CTA = one 128 x 128 pair tile (upper triangle only) x one frame segment.  Per pair-frame:
// 3 sub + mul + 2 fma + MUFU.SQRT + add, the FP32 part issued as packed FADD2/FMUL2/FFMA2 over
// two pairs.  The MUFU pipe (16 lanes/clk/SM) is the bound, and the FP32 pipe is nearly
// saturated with it (7 of every 8 cycles), so the production kernel is the warp-specialised
// contact_sum_ws_kernel further down: compute warps run only the pair loop, a producer warp feeds
// them by TMA.  contact_sum_kernel (in-warp staging, block barriers) is kept as the measured
// baseline: 31.1 ms at C2 (512 threads x 8 x 4 pairs) / 31.5 ms (256 x 8 x 8) against 27.9 ms.
// FP32 partial sums are promoted every 128 frames into FP64 per-pair totals kept in shared
// memory (128 KB), written once at the end.
#include "common.cuh"
#include <cstdlib>

namespace {

constexpr int kFC = 8;             // frames per staged chunk
constexpr int kFlushChunks = 16;   // promote FP32 -> FP64 every 128 frames
constexpr int kTileFloats = 3 * kTile;   // 384 floats per tile-frame

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Blackwell packed FP32x2 arithmetic (SASS FADD2 / FMUL2 / FFMA2): two pairs per issue
// slot, which takes the FP32 issue pressure off the MUFU-bound loop (4.5 instead of 8 issue
// slots per pair-frame).  A scalar `a` is broadcast to both halves by ptxas (R.F32 operand).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

struct ContactSmem {
    double tot[kTile * kTile];             // [pair][thread] -> conflict-free flush
    float si[2][kFC][3][kTile];
    float sj[2][kFC][3][kTile];
};

template <int JH, int UNROLL>   // packed j pairs per thread: 2 -> 512 threads x (8x4 pairs), 4 -> 256 x (8x8)
__global__ void __launch_bounds__(1024 / JH, 1)
contact_sum_kernel(const float* __restrict__ p32, int64_t T, int N, const double* __restrict__ mean,
                   int n_tiles, int n_seg, double* __restrict__ out, int64_t out_ld) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ContactSmem& sm = *reinterpret_cast<ContactSmem*>(smem_raw);

    int bi = 0, rem = blockIdx.x;
    while (rem >= n_tiles - bi) { rem -= n_tiles - bi; ++bi; }
    const int bj = bi + rem;
    const int seg = blockIdx.y;
    const int64_t f_begin = T * seg / n_seg;
    const int64_t f_end = T * (seg + 1) / n_seg;

    constexpr int kThreadsC = 1024 / JH;
    constexpr int kPerThread = kFC * 192 / kThreadsC; // staged float4 per thread per chunk
    const int tid = threadIdx.x;
    const int ti = tid & 15;     // i nodes: 4ti..4ti+3 and 64+4ti..64+4ti+3
    const int tj = tid >> 4;     // j nodes: 4tj..4tj+3 (and 64+4tj.. when JH == 4)

    // reference offset between the two tiles: x_i - x_j = a_i - (b_j + (ref_J - ref_I))
    int ri = kTile * bi + kTile / 2, rj = kTile * bj + kTile / 2;
    if (ri > N - 1) ri = N - 1;
    if (rj > N - 1) rj = N - 1;
    float roff[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) roff[c] = (float)(mean[3 * rj + c] - mean[3 * ri + c]);

    // staging: kFC frames x (i tile + j tile) x 96 float4 per chunk
    auto load_chunk = [&](int64_t f0, float4 (&v)[kPerThread]) {
#pragma unroll
        for (int q = 0; q < kPerThread; ++q) {
            const int e = tid + kThreadsC * q;        // 0..1535
            const int fl = e / 192;
            const int r = e - fl * 192;
            const int sel = r / 96;                   // 0 = i tile, 1 = j tile
            const int w = r - sel * 96;               // float4 index inside the tile-frame
            const int64_t f = f0 + fl;
            if (f < f_end) {
                const float4* src = reinterpret_cast<const float4*>(
                    p32 + (f * n_tiles + (sel ? bj : bi)) * kTileFloats) + w;
                v[q] = __ldg(src);
            } else {
                v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    auto store_chunk = [&](int buf, const float4 (&v)[kPerThread]) {
#pragma unroll
        for (int q = 0; q < kPerThread; ++q) {
            const int e = tid + kThreadsC * q;
            const int fl = e / 192;
            const int r = e - fl * 192;
            const int sel = r / 96;
            const int w = r - sel * 96;
            const int c = w / 32;
            float4 x = v[q];
            if (sel) {
                const float o = c == 0 ? roff[0] : (c == 1 ? roff[1] : roff[2]);
                x.x += o; x.y += o; x.z += o; x.w += o;
                *reinterpret_cast<float4*>(&sm.sj[buf][fl][0][0] + 4 * w) = x;
            } else {
                *reinterpret_cast<float4*>(&sm.si[buf][fl][0][0] + 4 * w) = x;
            }
        }
    };

    f32x2 acc[8][JH];       // packed FP32 partial sums: pairs (a, 2h) and (a, 2h+1)
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int h = 0; h < JH; ++h) acc[a][h] = pack2(0.f, 0.f);
#pragma unroll
    for (int p = 0; p < 16 * JH; ++p) sm.tot[p * kThreadsC + tid] = 0.0;

    float4 stage[kPerThread];
    load_chunk(f_begin, stage);
    store_chunk(0, stage);
    __syncthreads();

    int buf = 0;
    int chunk_idx = 0;
    for (int64_t f0 = f_begin; f0 < f_end; f0 += kFC, ++chunk_idx) {
        const bool has_next = (f0 + kFC) < f_end;
        if (has_next) load_chunk(f0 + kFC, stage);
        const int nf = (int)((f_end - f0) < kFC ? (f_end - f0) : kFC);
#pragma unroll UNROLL
        for (int fl = 0; fl < nf; ++fl) {
            float ax[8], ay[8], az[8];
            f32x2 bx[JH], by[JH], bz[JH];
            {
                const float4 t0 = *reinterpret_cast<const float4*>(&sm.si[buf][fl][0][4 * ti]);
                const float4 t1 = *reinterpret_cast<const float4*>(&sm.si[buf][fl][0][64 + 4 * ti]);
                ax[0] = t0.x; ax[1] = t0.y; ax[2] = t0.z; ax[3] = t0.w;
                ax[4] = t1.x; ax[5] = t1.y; ax[6] = t1.z; ax[7] = t1.w;
                const float4 u0 = *reinterpret_cast<const float4*>(&sm.si[buf][fl][1][4 * ti]);
                const float4 u1 = *reinterpret_cast<const float4*>(&sm.si[buf][fl][1][64 + 4 * ti]);
                ay[0] = u0.x; ay[1] = u0.y; ay[2] = u0.z; ay[3] = u0.w;
                ay[4] = u1.x; ay[5] = u1.y; ay[6] = u1.z; ay[7] = u1.w;
                const float4 v0 = *reinterpret_cast<const float4*>(&sm.si[buf][fl][2][4 * ti]);
                const float4 v1 = *reinterpret_cast<const float4*>(&sm.si[buf][fl][2][64 + 4 * ti]);
                az[0] = v0.x; az[1] = v0.y; az[2] = v0.z; az[3] = v0.w;
                az[4] = v1.x; az[5] = v1.y; az[6] = v1.z; az[7] = v1.w;
                const float4 w0 = *reinterpret_cast<const float4*>(&sm.sj[buf][fl][0][4 * tj]);
                const float4 w1 = *reinterpret_cast<const float4*>(&sm.sj[buf][fl][1][4 * tj]);
                const float4 w2 = *reinterpret_cast<const float4*>(&sm.sj[buf][fl][2][4 * tj]);
                bx[0] = pack2(w0.x, w0.y); bx[1] = pack2(w0.z, w0.w);
                by[0] = pack2(w1.x, w1.y); by[1] = pack2(w1.z, w1.w);
                bz[0] = pack2(w2.x, w2.y); bz[1] = pack2(w2.z, w2.w);
                if (JH == 4) {
                    const float4 x0 = *reinterpret_cast<const float4*>(&sm.sj[buf][fl][0][64 + 4 * tj]);
                    const float4 x1 = *reinterpret_cast<const float4*>(&sm.sj[buf][fl][1][64 + 4 * tj]);
                    const float4 x2 = *reinterpret_cast<const float4*>(&sm.sj[buf][fl][2][64 + 4 * tj]);
                    bx[JH - 2] = pack2(x0.x, x0.y); bx[JH - 1] = pack2(x0.z, x0.w);
                    by[JH - 2] = pack2(x1.x, x1.y); by[JH - 1] = pack2(x1.z, x1.w);
                    bz[JH - 2] = pack2(x2.x, x2.y); bz[JH - 1] = pack2(x2.z, x2.w);
                }
            }
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const f32x2 axx = pack2(ax[a], ax[a]), ayy = pack2(ay[a], ay[a]), azz = pack2(az[a], az[a]);
#pragma unroll
                for (int h = 0; h < JH; ++h) {
                    const f32x2 dx = sub2(axx, bx[h]);
                    const f32x2 dy = sub2(ayy, by[h]);
                    const f32x2 dz = sub2(azz, bz[h]);
                    const f32x2 d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                    float lo, hi;
                    unpack2(d2, lo, hi);
                    acc[a][h] = add2(acc[a][h], pack2(fast_sqrt(lo), fast_sqrt(hi)));
                }
            }
        }
        if ((chunk_idx % kFlushChunks) == kFlushChunks - 1 || !has_next) {
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int h = 0; h < JH; ++h) {
                    float lo, hi;
                    unpack2(acc[a][h], lo, hi);
                    sm.tot[(a * 2 * JH + 2 * h) * kThreadsC + tid] += (double)lo;
                    sm.tot[(a * 2 * JH + 2 * h + 1) * kThreadsC + tid] += (double)hi;
                    acc[a][h] = pack2(0.f, 0.f);
                }
        }
        if (has_next) store_chunk(buf ^ 1, stage);
        __syncthreads();
        buf ^= 1;
    }

    double* o = out + (size_t)seg * out_ld * out_ld;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int64_t row = (int64_t)kTile * bi + ((a < 4) ? (4 * ti + a) : (64 + 4 * ti + (a - 4)));
#pragma unroll
        for (int g = 0; g < JH / 2; ++g) {       // groups of 4 consecutive j nodes
            const int64_t col0 = (int64_t)kTile * bj + 64 * g + 4 * tj;
            double* p = o + row * out_ld + col0;
            const int pb = a * 2 * JH + 4 * g;
            *reinterpret_cast<double2*>(p) =
                make_double2(sm.tot[(pb + 0) * kThreadsC + tid], sm.tot[(pb + 1) * kThreadsC + tid]);
            *reinterpret_cast<double2*>(p + 2) =
                make_double2(sm.tot[(pb + 2) * kThreadsC + tid], sm.tot[(pb + 3) * kThreadsC + tid]);
        }
    }
}

// ---- warp-specialised form --------------------------------------------------------------
// Same arithmetic, other plumbing: the 16 compute warps execute nothing but the pair loop (the
// loop needs 7 issue cycles of packed FP32 per MUFU cycle of 8, so every staging instruction a
// compute warp issues is lost MUFU time).  Warp 16 is the producer: one lane issues the TMA bulk
// copies of the next 16 frames (one 1.5 KB row per frame and tile, cp.async.bulk + mbarrier
// complete_tx) into a 2-stage ring, the warp then adds the tile-pair reference offset to the j
// rows in place and releases the stage (full barrier); compute warps hand stages back through
// an empty barrier.  No __syncthreads in the frame loop.
#ifndef WISP_WS_FC
#define WISP_WS_FC 16
#endif
#ifndef WISP_WS_STAGES
#define WISP_WS_STAGES 2
#endif
constexpr int kWsFC = WISP_WS_FC;                 // frames per stage
constexpr int kWsStages = WISP_WS_STAGES;
constexpr int kWsComputeThreads = 512;
constexpr int kWsThreads = kWsComputeThreads + 32;

struct ContactWsSmem {
    double tot[kTile * kTile];
    float ring[kWsStages][2][kWsFC][3][kTile];                // [stage][0 = i rows, 1 = j rows]
    unsigned long long landed[kWsStages], full[kWsStages], empty[kWsStages];
};

__device__ __forceinline__ uint32_t ws_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void ws_mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ws_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ws_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(ws_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ws_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(ws_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ws_smem_u32(bar)) : "memory");
}

// diagonal tiles: 32 x 32 sub-block (row, column) of compute warp w; w < 10: on or above the diagonal
__constant__ int kDiagSubRow[16] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 1, 2, 2, 3, 3, 3};
__constant__ int kDiagSubCol[16] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3, 0, 0, 1, 0, 1, 2};

template <int UNROLL>
__global__ void __launch_bounds__(kWsThreads, 1)
contact_sum_ws_kernel(const float* __restrict__ p32, int64_t T, int N, const double* __restrict__ mean,
                      int n_tiles, int n_seg, double* __restrict__ out, int64_t out_ld, int allow_idle,
                      int accumulate, int balance) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ContactWsSmem& sm = *reinterpret_cast<ContactWsSmem*>(smem_raw);

    int bi = 0, rem = blockIdx.x;
    while (rem >= n_tiles - bi) { rem -= n_tiles - bi; ++bi; }
    const int bj = bi + rem;
    const int seg = blockIdx.y;
    const int64_t f_begin = T * seg / n_seg;
    const int64_t f_end = T * (seg + 1) / n_seg;
    const int n_chunks = (int)((f_end - f_begin + kWsFC - 1) / kWsFC);
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int s = 0; s < kWsStages; ++s) {
            ws_mbar_init(&sm.landed[s], 1);
            ws_mbar_init(&sm.full[s], 1);
            ws_mbar_init(&sm.empty[s], kWsComputeThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < kWsComputeThreads) {
#pragma unroll
        for (int p = 0; p < 32; ++p) sm.tot[p * kWsComputeThreads + tid] = 0.0;
    }
    __syncthreads();

    if (tid >= kWsComputeThreads) {
        // ---------------- producer warp ----------------
        const int lane = tid & 31;
        int ri = kTile * bi + kTile / 2, rj = kTile * bj + kTile / 2;
        if (ri > N - 1) ri = N - 1;
        if (rj > N - 1) rj = N - 1;
        float roff[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) roff[c] = (float)(mean[3 * rj + c] - mean[3 * ri + c]);
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int slot = ch % kWsStages;
            const uint32_t ph = (uint32_t)((ch / kWsStages) & 1);
            const int64_t f0 = f_begin + (int64_t)ch * kWsFC;
            const int nf = (int)((f_end - f0) < kWsFC ? (f_end - f0) : kWsFC);
            ws_mbar_wait(&sm.empty[slot], ph ^ 1);
            if (lane == 0) {
                ws_mbar_expect_tx(&sm.landed[slot], (uint32_t)nf * 2u * kTileFloats * 4u);
                for (int fl = 0; fl < nf; ++fl) {
                    const float* row = p32 + (f0 + fl) * (int64_t)n_tiles * kTileFloats;
                    ws_bulk_g2s(&sm.ring[slot][0][fl][0][0], row + (int64_t)bi * kTileFloats, kTileFloats * 4, &sm.landed[slot]);
                    ws_bulk_g2s(&sm.ring[slot][1][fl][0][0], row + (int64_t)bj * kTileFloats, kTileFloats * 4, &sm.landed[slot]);
                }
            }
            ws_mbar_wait(&sm.landed[slot], ph);
            if (bi != bj) {          // x_i - x_j = a_i - (b_j + (ref_J - ref_I)): shift the j rows once
                float4* q = reinterpret_cast<float4*>(&sm.ring[slot][1][0][0][0]);
                for (int fl = 0; fl < nf; ++fl)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {              // 32 float4 per component row
                        float4 v = q[fl * 96 + c * 32 + lane];
                        v.x += roff[c]; v.y += roff[c]; v.z += roff[c]; v.w += roff[c];
                        q[fl * 96 + c * 32 + lane] = v;
                    }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // before the next TMA write here
            }
            __syncwarp();
            if (lane == 0) ws_mbar_arrive(&sm.full[slot]);
        }
        return;
    }

    // ---------------- compute warps ----------------
    // Thread -> pair block (8 i nodes x 4 j nodes).  Off-diagonal tiles: i nodes 4ti..4ti+3 and
    // 64+4ti..+3, j nodes 4tj..4tj+3 (a warp = 8 j columns x all 128 i rows); warps whose 8 columns lie
    // beyond the last real node (the padded part of the last tile) have nothing to do.  Diagonal tiles:
    // the warps form a 4 x 4 grid of 32 x 32 sub-blocks and the 6 sub-blocks strictly below the diagonal
    // (whose pairs the sub-blocks above it already cover) are idle.  Idle warps only keep the stage
    // hand-shake going; the MUFU pipe of the SM is then shared by fewer warps, so the CTA ends earlier:
    // 8.5 % less K3 time at C2 (16 diagonal tiles x 6/16 + 15 last-column tiles x 48/128).
    // balance: a tile with A < 16 active blocks leaves the four schedulers unevenly loaded (A = 10: 3, 3, 2, 2
    // warps -- the CTA takes 3/4 of a full tile's time for 10/16 of its work).  The blocks are therefore dealt
    // out again: 2A - 16 blocks keep a warp of their own, every other block is SHARED by two consecutive
    // warps (different schedulers), one taking the even and one the odd frames of every stage; each scheduler
    // then carries the same number of warps and within half a block the same work (A = 10: 2.5 blocks each).
    // The two halves meet in the per-thread FP64 totals after the frame loop.
    constexpr int JH = 2;
    int ib0, ib1, jb0;
    bool idle;
    int fstart = 0, fstep = 1;          // frames of a stage this warp takes
    bool shared_primary = false, shared_secondary = false;
    const int w = tid >> 5, lane = tid & 31;
    const int valid_i = min(kTile, N - kTile * bi), valid_j = min(kTile, N - kTile * bj);
    auto block_active = [&](int ow) -> bool {        // does the block of (original) warp ow hold real pairs?
        if (bi == bj) return ow < 10 && 32 * kDiagSubRow[ow] < valid_i && 32 * kDiagSubCol[ow] < valid_j;
        return 8 * ow < valid_j;
    };
    auto block_coords = [&](int ow, int& i0, int& i1, int& j0) {
        if (bi == bj) {
            // the 10 sub-blocks on or above the diagonal belong to (original) warps 0..9; 10..15 own the 6
            // sub-blocks below the diagonal, whose pairs the sub-blocks above it already cover
            i0 = 32 * kDiagSubRow[ow] + 8 * (lane >> 3);
            i1 = i0 + 4;
            j0 = 32 * kDiagSubCol[ow] + 4 * (lane & 7);
        } else {
            const int vt = 32 * ow + lane;
            i0 = 4 * (vt & 15);
            i1 = 64 + i0;
            j0 = 4 * (vt >> 4);
        }
    };
    int task = w;
    idle = !block_active(w) && allow_idle;
    bool balanced = false;
    if (balance && allow_idle) {
        int n_active = 0;
#pragma unroll
        for (int ow = 0; ow < 16; ++ow) n_active += block_active(ow) ? 1 : 0;
        if (n_active > 0 && n_active < 16) {
            balanced = true;
            const int n_own = n_active > 8 ? 2 * n_active - 16 : 0;
            int k;                                           // which active block (in original warp order)
            if (w < n_own) {
                k = w;
            } else {
                const int h = w - n_own;
                k = n_own + (h >> 1);
                if (k < n_active) {
                    fstart = h & 1;
                    fstep = 2;
                    shared_primary = (h & 1) == 0;
                    shared_secondary = (h & 1) == 1;
                }
            }
            idle = k >= n_active;
            if (!idle) {
                int seen = 0;
                task = 0;
#pragma unroll
                for (int ow = 0; ow < 16; ++ow) {
                    if (block_active(ow)) { if (seen == k) task = ow; ++seen; }
                }
            }
        }
    }
    block_coords(task, ib0, ib1, jb0);
    f32x2 acc[8][JH];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int h = 0; h < JH; ++h) acc[a][h] = pack2(0.f, 0.f);

    for (int ch = 0; ch < n_chunks; ++ch) {
        const int slot = ch % kWsStages;
        const uint32_t ph = (uint32_t)((ch / kWsStages) & 1);
        const int64_t f0 = f_begin + (int64_t)ch * kWsFC;
        const int nf = (int)((f_end - f0) < kWsFC ? (f_end - f0) : kWsFC);
        ws_mbar_wait(&sm.full[slot], ph);
        if (idle) {
            __syncwarp();
            if ((tid & 31) == 0) ws_mbar_arrive(&sm.empty[slot]);
            continue;
        }
        const float* si = &sm.ring[slot][0][0][0][0];
        const float* sj = &sm.ring[slot][1][0][0][0];
        auto frame_body = [&](const float* fi, const float* fj) {
            float ax[8], ay[8], az[8];
            f32x2 bx[JH], by[JH], bz[JH];
            {
                const float4 t0 = *reinterpret_cast<const float4*>(fi + ib0);
                const float4 t1 = *reinterpret_cast<const float4*>(fi + ib1);
                ax[0] = t0.x; ax[1] = t0.y; ax[2] = t0.z; ax[3] = t0.w;
                ax[4] = t1.x; ax[5] = t1.y; ax[6] = t1.z; ax[7] = t1.w;
                const float4 u0 = *reinterpret_cast<const float4*>(fi + kTile + ib0);
                const float4 u1 = *reinterpret_cast<const float4*>(fi + kTile + ib1);
                ay[0] = u0.x; ay[1] = u0.y; ay[2] = u0.z; ay[3] = u0.w;
                ay[4] = u1.x; ay[5] = u1.y; ay[6] = u1.z; ay[7] = u1.w;
                const float4 v0 = *reinterpret_cast<const float4*>(fi + 2 * kTile + ib0);
                const float4 v1 = *reinterpret_cast<const float4*>(fi + 2 * kTile + ib1);
                az[0] = v0.x; az[1] = v0.y; az[2] = v0.z; az[3] = v0.w;
                az[4] = v1.x; az[5] = v1.y; az[6] = v1.z; az[7] = v1.w;
                const float4 w0 = *reinterpret_cast<const float4*>(fj + jb0);
                const float4 w1 = *reinterpret_cast<const float4*>(fj + kTile + jb0);
                const float4 w2 = *reinterpret_cast<const float4*>(fj + 2 * kTile + jb0);
                bx[0] = pack2(w0.x, w0.y); bx[1] = pack2(w0.z, w0.w);
                by[0] = pack2(w1.x, w1.y); by[1] = pack2(w1.z, w1.w);
                bz[0] = pack2(w2.x, w2.y); bz[1] = pack2(w2.z, w2.w);
            }
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const f32x2 axx = pack2(ax[a], ax[a]), ayy = pack2(ay[a], ay[a]), azz = pack2(az[a], az[a]);
#pragma unroll
                for (int h = 0; h < JH; ++h) {
                    const f32x2 dx = sub2(axx, bx[h]);
                    const f32x2 dy = sub2(ayy, by[h]);
                    const f32x2 dz = sub2(azz, bz[h]);
                    const f32x2 d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                    float lo, hi;
                    unpack2(d2, lo, hi);
                    acc[a][h] = add2(acc[a][h], pack2(fast_sqrt(lo), fast_sqrt(hi)));
                }
            }
        };
        // two loops with compile-time strides: a run-time frame step in the unrolled loop costs integer
        // address arithmetic in a loop that has no issue slot to spare (28.7 instead of 27.7 ms at C2)
        if (fstep == 1) {
#pragma unroll UNROLL
            for (int fl = 0; fl < nf; ++fl) frame_body(si + fl * kTileFloats, sj + fl * kTileFloats);
        } else {
            const float* si2 = si + fstart * kTileFloats;
            const float* sj2 = sj + fstart * kTileFloats;
            const int nh = (nf - fstart + 1) >> 1;           // frames fstart, fstart + 2, ... < nf
#pragma unroll UNROLL
            for (int fl = 0; fl < nh; ++fl) frame_body(si2 + fl * (2 * kTileFloats), sj2 + fl * (2 * kTileFloats));
        }
        __syncwarp();
        if ((tid & 31) == 0) ws_mbar_arrive(&sm.empty[slot]);        // this warp is done with the stage
        if ((ch % (128 / kWsFC)) == (128 / kWsFC) - 1 || ch == n_chunks - 1) {      // every 128 frames
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int h = 0; h < JH; ++h) {
                    float lo, hi;
                    unpack2(acc[a][h], lo, hi);
                    sm.tot[(a * 2 * JH + 2 * h) * kWsComputeThreads + tid] += (double)lo;
                    sm.tot[(a * 2 * JH + 2 * h + 1) * kWsComputeThreads + tid] += (double)hi;
                    acc[a][h] = pack2(0.f, 0.f);
                }
        }
    }

    // Store.  Without `accumulate` the partial tile must be completely defined: blocks without real pairs get
    // zeros.  accumulate: the tile already holds the sums of earlier frame blocks (every entry belongs to
    // exactly one thread and the launches are stream-ordered: plain read-modify-write).
    double* o = out + (size_t)seg * out_ld * out_ld;
    auto store_block = [&](int i0, int i1, int j0, bool zeros, int partner) {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int64_t row = (int64_t)kTile * bi + ((a < 4) ? (i0 + a) : (i1 + (a - 4)));
            const int64_t col0 = (int64_t)kTile * bj + j0;
            double* p = o + row * out_ld + col0;
            const int pb = a * 2 * JH;
            double v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] = zeros ? 0.0 : sm.tot[(pb + k) * kWsComputeThreads + tid];
                if (partner >= 0) v[k] += sm.tot[(pb + k) * kWsComputeThreads + partner];
            }
            if (accumulate) {
                const double2 o0 = *reinterpret_cast<const double2*>(p), o1 = *reinterpret_cast<const double2*>(p + 2);
                v[0] += o0.x; v[1] += o0.y; v[2] += o1.x; v[3] += o1.y;
            }
            *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
            *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
        }
    };
    if (!balanced) {
        // every warp owns the block of its own index; idle threads store the zeros their totals were
        // initialised with
        if (accumulate && idle) return;
        store_block(ib0, ib1, jb0, false, -1);
        return;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kWsComputeThreads) : "memory");    // the halves of shared blocks are complete
    if (!idle && !shared_secondary) store_block(ib0, ib1, jb0, false, shared_primary ? tid + 32 : -1);
    if (!accumulate && !block_active(w)) {          // zeros for the block without real pairs that warp w would own
        int z0, z1, zj;
        block_coords(w, z0, z1, zj);
        store_block(z0, z1, zj, true, -1);
    }
}

}  // namespace

int launch_contact_sum(wisp_ctx* ctx, const float* d_p32, int64_t T, int N, const double* d_mean,
                       double* d_out, cudaStream_t st, bool accumulate) {
    WISP_CHECK(T > 0 && N > 0, "contact: empty input (T=%lld, N=%d)", (long long)T, N);
    const int64_t n_pad = round_up(N, kTile);
    const int n_tiles = (int)(n_pad / kTile);
    const int n_items = n_tiles * (n_tiles + 1) / 2;
    const int chunks_total = (int)((T + kFC - 1) / kFC);
    // accumulate: one segment per tile, the tile of d_out is the accumulator
    const int n_seg = accumulate ? 1 : wisp_pick_segments(n_items, chunks_total, ctx->sm_count, 16);
    double* part = d_out;
    if (n_seg > 1) {
        void* p = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_CONTACT_PART, (size_t)n_seg * n_pad * n_pad * 8, &p));
        part = (double*)p;
    }
    const int smem = (int)sizeof(ContactSmem);
    dim3 grid(n_items, n_seg);
    static const int variant_env = getenv("WISP_K3_VARIANT") ? atoi(getenv("WISP_K3_VARIANT")) : 58;
    const int variant = accumulate ? 58 : variant_env;      // only the production kernel accumulates
    const int allow_idle = getenv("WISP_K3_NOIDLE") ? 0 : 1;      // experiment switch: every warp computes its block
    // WISP_K3_BALANCE=0|1: deal the active blocks of diagonal / padded tiles out evenly over the schedulers
    static const int balance = getenv("WISP_K3_BALANCE") ? atoi(getenv("WISP_K3_BALANCE")) : 1;
#define WISP_K3_LAUNCH(JH_, UN_)                                                                          \
    do {                                                                                                  \
        WISP_CUDA(cudaFuncSetAttribute(contact_sum_kernel<JH_, UN_>,                                      \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem));               \
        contact_sum_kernel<JH_, UN_><<<grid, 1024 / JH_, smem, st>>>(d_p32, T, N, d_mean, n_tiles, n_seg, \
                                                                     part, n_pad);                        \
    } while (0)
    if (variant >= 50 && variant < 60) {
        const int smem_ws = (int)sizeof(ContactWsSmem);
#define WISP_K3_WS(UN_)                                                                                         \
    do {                                                                                                        \
        WISP_CUDA(cudaFuncSetAttribute(contact_sum_ws_kernel<UN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       smem_ws));                                                               \
        contact_sum_ws_kernel<UN_><<<grid, kWsThreads, smem_ws, st>>>(d_p32, T, N, d_mean, n_tiles, n_seg, part, \
                                                                      n_pad, allow_idle, accumulate ? 1 : 0,    \
                                                                      balance);                                 \
    } while (0)
        if (variant == 51) WISP_K3_WS(1);      // frame loop not unrolled: 30.1 ms at C2 with 8-frame stages
        else WISP_K3_WS(8);                    // 58 (default)
#undef WISP_K3_WS
    } else
    switch (variant) {
        case 41: WISP_K3_LAUNCH(4, 1); break;      // 256 threads x (8 x 8 pairs), 246 registers
        default: WISP_K3_LAUNCH(2, 1); break;      // 21: 512 threads x (8 x 4 pairs), 126 registers: 4 warps per scheduler
    }
#undef WISP_K3_LAUNCH
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    if (n_seg > 1) WISP_TRY(launch_reduce_partials(ctx, part, n_seg, n_pad, d_out, kTile, kTile, st));
    return 0;
}
