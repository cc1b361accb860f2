// K2 (tensor-core form) -- the correlation contraction on tcgen05 INT8 tensor cores with FP64-grade
// results (exact digit slicing, "Ozaki scheme"), sm_100a.
//
// tcgen05.mma has no FP64 kind, and FP32-accumulating TF32/BF16 cannot meet the 1e-6 relative bar
// on small |C|.  Instead every centred coordinate becomes a 32-bit fixed-point
// number (per-node scale, full range) written as four BALANCED base-256 digits d0..d3 in
// [-128, 127].  Then
//
//     G_ij = u_i u_j * sum_{s,t} 2^(48-8(s+t)) * ( Q_s^T Q_t )_ij ,   Q_s = digit plane s, u = 1/scale,
//
// and each  Q_s^T Q_t  is an EXACT integer GEMM on the tensor cores (kind::i8, int32 accumulation
// in TMEM).  The 10 digit pairs with s+t <= 3 accumulate into 4 TMEM accumulators (one per weight
// s+t): 4 x 128 columns = the whole tensor memory of the SM.  The dropped pairs (s+t >= 4) are
// zero-mean digit noise, 2^-33 of the leading term per row and averaging down with sqrt(rows):
// measured |dG| <= 1e-10 sqrt(G_ii G_jj) at 50,000 frames.  Because that noise is fixed in
// fixed-point units, spiky data (max|x| >> rms) loses accuracy; node_scale_kernel evaluates the
// bound on the device and raises a flag that turns every kernel of this path into a no-op and lets
// the DMMA kernel (k2_gram.cu), launched behind the same flag, produce the Gram instead.  The
// diagonal G_ii is summed in plain FP64 (the dropped d2*d2 term is NOT zero-mean there).
// K is cut into segments of <= 16384 rows so the int32 accumulators cannot overflow
// (4 pairs x 128 x 128 x 16384 = 2^30); per-segment partial
// tiles are reduced in FP64 in fixed order, like the DMMA kernel's: results are deterministic.
//
// Operand layout: digit planes are stored [plane][k = 3*frame + c][node] -- node-contiguous rows --
// which is an "MN-major" UMMA operand (valid for INT8).  One TMA box {128 nodes, 32 k-rows} with
// 128-byte swizzle lands exactly in the canonical MN-major SW128 layout ((8,1),(8,4)):((1,_),(8,64))
// in 16-byte units, so a K=32 instruction takes one descriptor per operand plane.
//
// CTA = one 128x128 output tile (upper triangle) x one K segment; 6 warps:
//   warp 0  TMA producer   : 8 box loads per stage (4 planes x 2 operands, 32 KB), 6-stage ring
//   warp 1  MMA issuer     : 10 tcgen05.mma per stage from one elected lane; owns TMEM alloc/dealloc
//   warps 2-5 epilogue     : tcgen05.ld 32 lanes x 16 columns per weight, combine the four weights in
//                            FP64, scale by u_i u_j, store the partial tile
#include "common.cuh"
#include <cuda.h>
#include <cmath>
#include <cstdlib>

namespace {

constexpr int kI8Stages = 6;
constexpr int kI8Rows = 32;                       // k-rows per stage = K of one tcgen05.mma (i8)
constexpr int kPlaneBytes = kI8Rows * 128;        // 4 KB: one plane tile of one operand
constexpr int kI8StageBytes = 8 * kPlaneBytes;    // A planes 0..3, B planes 0..3
constexpr int kSegRows = 16384;                   // K rows per CTA: 4 digit pairs x 128^2 x 16384 = 2^30, half the int32 range
constexpr int kI8Threads = 6 * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// MN-major, 128-byte swizzle, one 128-byte atom along MN, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 16;     // leading byte offset (unused: single MN atom)
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;     // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
// instruction descriptor: S32 accumulate, A/B int8 (signed or unsigned), MN-major both, M=128, N=128
__device__ __forceinline__ uint32_t idesc_i8(int a_signed, int b_signed) {
    uint32_t d = 0;
    d |= 2u << 4;                       // c_format = S32
    d |= (uint32_t)(a_signed ? 1 : 0) << 7;
    d |= (uint32_t)(b_signed ? 1 : 0) << 10;
    d |= 1u << 15;                      // a_major = MN
    d |= 1u << 16;                      // b_major = MN
    d |= (uint32_t)(128 >> 3) << 17;    // N
    d |= (uint32_t)(128 >> 4) << 24;    // M
    return d;
}

__global__ void __launch_bounds__(kI8Threads, 1)
gram_i8_kernel(const __grid_constant__ CUtensorMap tmap, int k_plane_rows, int n_tiles, int n_seg,
               int rows_total, const double* __restrict__ unit, const double* __restrict__ diag,
               double* __restrict__ out, int64_t out_ld, const int* __restrict__ unsafe) {
    extern __shared__ unsigned char smem_dyn[];
    // TMA 128-byte swizzle and the UMMA descriptors need 1024-byte aligned tiles
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ uint64_t full[kI8Stages], empty[kI8Stages], accum_full;
    __shared__ uint32_t tmem_base_smem;
    if (*unsafe) return;                 // accuracy bound not met: the DMMA kernel takes over

    int bi = 0, rem = blockIdx.x;
    while (rem >= n_tiles - bi) { rem -= n_tiles - bi; ++bi; }
    const int bj = bi + rem;
    const int seg = blockIdx.y;
    const int chunks_total = rows_total / kI8Rows;
    const int c0 = (int)((int64_t)chunks_total * seg / n_seg);
    const int c1 = (int)((int64_t)chunks_total * (seg + 1) / n_seg);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kI8Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            int it = 0;
            for (int c = c0; c < c1; ++c, ++it) {
                const int slot = it % kI8Stages;
                const uint32_t ph = (it / kI8Stages) & 1;
                mbar_wait(&empty[slot], ph ^ 1);
                mbar_expect_tx(&full[slot], kI8StageBytes);
                unsigned char* st = smem_raw + (size_t)slot * kI8StageBytes;
                for (int p = 0; p < 4; ++p) {
                    const int row = p * k_plane_rows + c * kI8Rows;
                    tma_load_2d(st + p * kPlaneBytes, &tmap, bi * kTile, row, &full[slot]);
                    tma_load_2d(st + (4 + p) * kPlaneBytes, &tmap, bj * kTile, row, &full[slot]);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t idesc = idesc_i8(1, 1);
            int it = 0;
            for (int c = c0; c < c1; ++c, ++it) {
                const int slot = it % kI8Stages;
                const uint32_t ph = (it / kI8Stages) & 1;
                mbar_wait(&full[slot], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(smem_raw + (size_t)slot * kI8StageBytes);
#pragma unroll
                for (int w = 0; w < 4; ++w)
#pragma unroll
                    for (int s = 0; s <= w; ++s) {
                        const int t = w - s;
                        const uint64_t da = smem_desc_mn_sw128(st + s * kPlaneBytes);
                        const uint64_t db = smem_desc_mn_sw128(st + (4 + t) * kPlaneBytes);
                        const uint32_t acc = (it > 0 || s > 0) ? 1u : 0u;
                        umma_i8(tmem_base + (uint32_t)(w * kTile), da, db, idesc, acc);
                    }
                umma_commit(&empty[slot]);        // frees the smem slot when these MMAs retire
            }
            umma_commit(&accum_full);             // accumulators complete
        }
    } else {
        // ---------------- epilogue warps ----------------
        mbar_wait(&accum_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int r = 32 * q + lane;                   // tile row = TMEM lane
        const int64_t grow = (int64_t)kTile * bi + r;
        const double ui = unit[grow];
        double* orow = out + (size_t)seg * out_ld * out_ld + grow * out_ld + (int64_t)kTile * bj;
        const bool any_work = c1 > c0;
#pragma unroll 1
        for (int cc = 0; cc < kTile; cc += 16) {
            double v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.0;
            if (any_work) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    uint32_t a[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(w * kTile + cc);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
                          "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const double wt = (w == 0) ? 281474976710656.0 : (w == 1) ? 1099511627776.0 : (w == 2) ? 4294967296.0 : 16777216.0;   // 2^(48-8w)
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += wt * (double)(int32_t)a[j];
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int64_t gcol = (int64_t)kTile * bj + cc + j;
                double g = (v[j] * ui) * unit[gcol];
                if (gcol == grow) g = (seg == 0) ? diag[grow] : 0.0;     // exact FP64 diagonal
                orow[cc + j] = g;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

constexpr double kFixedMax = 2130706432.0;      // 127 * 2^24: top balanced digit stays <= 127
constexpr int kStatRows = 64;                   // frame classes of the per-node statistics pass

// per node: max |x| (any component, any frame) and, per frame class, sum of squares
__global__ void node_stats_kernel(const double* __restrict__ x, int64_t T, int64_t ld, int n_pad,
                                  unsigned long long* __restrict__ absmax_bits, double* __restrict__ sumsq_part) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_pad) return;
    double m = 0.0, ss = 0.0;
    for (int64_t t = blockIdx.y; t < T; t += gridDim.y) {
        const double* p = x + t * ld + 3 * (int64_t)n;
        const double a = p[0], b = p[1], c = p[2];
        m = fmax(m, fmax(fabs(a), fmax(fabs(b), fabs(c))));
        ss += (a * a + b * b) + c * c;
    }
    sumsq_part[(int64_t)blockIdx.y * n_pad + n] = ss;
    atomicMax(absmax_bits + n, (unsigned long long)__double_as_longlong(m));    // non-negative doubles order like integers
}
// scale[n] = kFixedMax / max|x_n|, unit[n] = 1/scale[n], diag[n] = sum of squares (fixed order)
__global__ void node_scale_kernel(const unsigned long long* __restrict__ absmax_bits, const double* __restrict__ sumsq_part,
                                  int n_parts, int n_pad, double rows, double rho2_min, double* __restrict__ scale,
                                  double* __restrict__ unit, double* __restrict__ diag, int* __restrict__ unsafe) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_pad) return;
    const double m = __longlong_as_double((long long)absmax_bits[n]);
    scale[n] = m > 0.0 ? kFixedMax / m : 0.0;
    unit[n] = m > 0.0 ? m / kFixedMax : 0.0;
    double ss = 0.0;
    for (int p = 0; p < n_parts; ++p) ss += sumsq_part[(int64_t)p * n_pad + n];
    diag[n] = ss;
    // the digit products that are dropped are noise of fixed size in fixed-point units, so the
    // error relative to sqrt(G_ii G_jj) grows as max|x|^2 / mean(x^2): refuse spiky data
    if (m > 0.0 && ss < rho2_min * rows * m * m) atomicOr(unsafe, 1);
}
// the same from the per-column statistics the centring pass left behind ([parts][2][width]: max|x|, sum x^2)
__global__ void node_scale_from_columns_kernel(const double* __restrict__ stat, int n_parts, int64_t width, int n_pad,
                                               double rows, double rho2_min, double* __restrict__ scale,
                                               double* __restrict__ unit, double* __restrict__ diag, int* __restrict__ unsafe) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_pad) return;
    double m = 0.0, ss = 0.0;
    for (int p = 0; p < n_parts; ++p) {
        const double* mx = stat + ((int64_t)p * 2 + 0) * width + 3 * (int64_t)n;
        const double* sq = stat + ((int64_t)p * 2 + 1) * width + 3 * (int64_t)n;
        m = fmax(m, fmax(mx[0], fmax(mx[1], mx[2])));
        ss += (sq[0] + sq[1]) + sq[2];
    }
    scale[n] = m > 0.0 ? kFixedMax / m : 0.0;
    unit[n] = m > 0.0 ? m / kFixedMax : 0.0;
    diag[n] = ss;
    if (m > 0.0 && ss < rho2_min * rows * m * m) atomicOr(unsafe, 1);
}
// X (float64 [T][ld]) -> four planes of BALANCED base-256 digits [4][rows_total][n_pad] (int8),
// row = 3*frame + c, node contiguous:  rint(x * scale) = d0 2^24 + d1 2^16 + d2 2^8 + d3, d in [-128, 127]
constexpr int kSliceFrames = 4;     // frames per block (threadIdx.z)

// block (32, 3, kSliceFrames): lane = 4 consecutive nodes, y = component, z = frame.  The three
// component warps of a frame read the same 3 KB of the row (stride-3 picks, served by L1); every
// warp writes one full 128-byte row per plane.  No shared memory, no barriers.
__global__ void __launch_bounds__(96 * kSliceFrames)
slice_planes_kernel(const double* __restrict__ x, int64_t T, int64_t ld, int n_pad,
                    const double* __restrict__ scale, int rows_total,
                    unsigned char* __restrict__ planes, const int* __restrict__ unsafe) {
    if (*unsafe) return;
    const int tile = blockIdx.x;
    const int w = threadIdx.x, c = threadIdx.y;
    const int n0 = tile * kTile + 4 * w;
    double sc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sc[j] = scale[n0 + j];
    const double* xcol = x + 3 * (int64_t)n0 + c;
    for (int64_t t = (int64_t)blockIdx.y * kSliceFrames + threadIdx.z; t < T; t += (int64_t)gridDim.y * kSliceFrames) {
        uint32_t word[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int32_t v = __double2int_rn(xcol[t * ld + 3 * j] * sc[j]);
            const int32_t d3 = (int32_t)(int8_t)(v & 255); v = (v - d3) >> 8;
            const int32_t d2 = (int32_t)(int8_t)(v & 255); v = (v - d2) >> 8;
            const int32_t d1 = (int32_t)(int8_t)(v & 255); v = (v - d1) >> 8;      // v is now d0
            word[0] |= ((uint32_t)v & 255u) << (8 * j);
            word[1] |= ((uint32_t)d1 & 255u) << (8 * j);
            word[2] |= ((uint32_t)d2 & 255u) << (8 * j);
            word[3] |= ((uint32_t)d3 & 255u) << (8 * j);
        }
        const int64_t row = 3 * t + c;
#pragma unroll
        for (int p = 0; p < 4; ++p)
            *reinterpret_cast<uint32_t*>(planes + ((int64_t)p * rows_total + row) * n_pad + n0) = word[p];
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

int launch_gram_i8(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int N, double* d_out,
                   cudaStream_t st, const int** gate_out) {
    WISP_CHECK(T > 0 && N > 0, "gram_i8: empty input");
    const int n_tiles = (N + kTile - 1) / kTile;
    const int n_pad = n_tiles * kTile;
    WISP_CHECK(3 * (int64_t)n_pad <= ld, "gram_i8: ld too small");
    const int rows_total = (int)round_up(3 * T, kI8Rows);
    const bool force_safe = gate_out == nullptr;          // explicit wisp_dev_gram_i8 call: no fallback behind it
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        WISP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        WISP_CHECK(fn && qres == cudaDriverEntryPointSuccess, "gram_i8: cuTensorMapEncodeTiled not available");
        encode = (EncodeTiledFn)fn;
    }
    void *p_planes, *p_small;
    const int gy = (int)std::min<int64_t>(T, kStatRows);
    WISP_TRY(wisp_scratch(ctx, SCR_I8_PLANES, (size_t)4 * rows_total * n_pad + 1024, &p_planes));
    WISP_TRY(wisp_scratch(ctx, SCR_I8_SMALL, (size_t)n_pad * 8 * (5 + kStatRows), &p_small));
    // [unsafe flag | absmax | scale | unit | diag | sumsq partials]
    int* d_unsafe = (int*)p_small;
    unsigned long long* d_absmax = (unsigned long long*)p_small + n_pad;
    double* d_scale = (double*)d_absmax + n_pad;
    double* d_unit = d_scale + n_pad;
    double* d_diag = d_unit + n_pad;
    double* d_sumsq = d_diag + n_pad;
    unsigned char* planes = (unsigned char*)p_planes;
    WISP_CUDA(cudaMemsetAsync(p_small, 0, (size_t)n_pad * 16, st));       // flag + absmax
    if (rows_total > 3 * T)      // zero the padding rows of every plane
        for (int p = 0; p < 4; ++p)
            WISP_CUDA(cudaMemsetAsync(planes + ((size_t)p * rows_total + 3 * T) * n_pad, 0,
                                      (size_t)(rows_total - 3 * T) * n_pad, st));
    // the statistics left by the centring pass are trusted only if this Gram is the very next launch of
    // the library on this context, on the same block and stream (any upload / COM / rotation / plain
    // centring in between bumps the launch counter or clears the pointer); otherwise X is read once more
    const bool have_stats = ctx->i8_stats_x == (const void*)d_x && ctx->i8_stats_T == T &&
                            ctx->i8_stats_stream == st && ld == 3 * (int64_t)n_pad &&
                            ctx->i8_stats_launch == ctx->launches;
    ctx->i8_stats_x = nullptr;                                   // single use
    // max error / sqrt(G_ii G_jj) ~ 6.5e-10 / (sqrt(rows) * rho_i * rho_j), rho^2 = mean(x^2) / max(x^2);
    // hold it below 5e-10 (half the 1e-9 absolute bar on C, tests/test_gpu_parity.py)
    const double rho2_min = force_safe ? 0.0 : 1.3 / std::sqrt(3.0 * (double)T);
    if (have_stats) {
        void* p_stat = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_I8_STATS, 8, &p_stat));
        node_scale_from_columns_kernel<<<(n_pad + 127) / 128, 128, 0, st>>>(
            (const double*)p_stat, ctx->i8_stats_parts, ld, n_pad, 3.0 * (double)T, rho2_min, d_scale, d_unit,
            d_diag, d_unsafe);
        ctx->launches += 1;
    } else {
        node_stats_kernel<<<dim3((n_pad + 127) / 128, gy), 128, 0, st>>>(d_x, T, ld, n_pad, d_absmax, d_sumsq);
        node_scale_kernel<<<(n_pad + 127) / 128, 128, 0, st>>>(d_absmax, d_sumsq, gy, n_pad, 3.0 * (double)T, rho2_min,
                                                               d_scale, d_unit, d_diag, d_unsafe);
        ctx->launches += 2;
    }
    const int gy2 = (int)std::min<int64_t>((T + kSliceFrames - 1) / kSliceFrames,
                                           std::max<int64_t>(1, (int64_t)ctx->sm_count * 16 / n_tiles));
    slice_planes_kernel<<<dim3(n_tiles, gy2), dim3(32, 3, kSliceFrames), 0, st>>>(d_x, T, ld, n_pad, d_scale, rows_total, planes, d_unsafe);
    ctx->launches += 1;
    WISP_CUDA(cudaGetLastError());

    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)n_pad, (cuuint64_t)4 * rows_total};
    const cuuint64_t strides[1] = {(cuuint64_t)n_pad};
    const cuuint32_t box[2] = {128, (cuuint32_t)kI8Rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, planes, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WISP_CHECK(cr == CUDA_SUCCESS, "gram_i8: cuTensorMapEncodeTiled failed (%d)", (int)cr);

    const int n_pairs = n_tiles * (n_tiles + 1) / 2;
    // K segments: at most kSegRows rows each (int32 bound), count chosen for full waves of CTAs
    const int chunks_total = rows_total / kI8Rows;
    const int seg_min = (rows_total + kSegRows - 1) / kSegRows;
    int n_seg = seg_min;
    double best = 0.0;
    for (int sgm = seg_min; sgm <= seg_min + 12 && sgm <= chunks_total; ++sgm) {
        const int64_t ctas = (int64_t)n_pairs * sgm;
        const int64_t waves = (ctas + ctx->sm_count - 1) / ctx->sm_count;
        const double eff = (double)ctas / (double)(waves * ctx->sm_count) - 0.004 * (sgm - seg_min);
        if (eff > best + 1e-9) { best = eff; n_seg = sgm; }
    }
    double* part = d_out;
    if (n_seg > 1) {
        void* p = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_I8_PART, (size_t)n_seg * n_pad * n_pad * 8, &p));
        part = (double*)p;
    }
    const int smem = kI8Stages * kI8StageBytes + 1024;
    WISP_CUDA(cudaFuncSetAttribute(gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gram_i8_kernel<<<dim3(n_pairs, n_seg), kI8Threads, smem, st>>>(tmap, rows_total, n_tiles, n_seg, rows_total,
                                                                   d_unit, d_diag, part, n_pad, d_unsafe);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    if (n_seg > 1) WISP_TRY(launch_reduce_partials(ctx, part, n_seg, n_pad, d_out, kTile, kTile, st, d_unsafe, 0));
    if (gate_out) *gate_out = d_unsafe;
    return 0;
}
