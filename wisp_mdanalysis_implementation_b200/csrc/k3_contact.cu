// K3 -- mean pairwise node distance (contact map numerator), sm_100a.
//
// Replaces the reference's T calls of MDAnalysis distance_array on an N x N result plus an
// N^2 numpy add per frame (trajectory_analysis_functions.py:207-215):
//     D[i][j] = sum_t | x_i(t) - x_j(t) |          (divided by T and thresholded in K4)
//
// One CTA = one 64(i) x 128(j) pair tile x one frame segment, upper-triangular tiles only.
// 512 threads, each owns 4 x 4 pairs.  The centred FP64 trajectory is staged 8 frames at a
// time into shared memory as FP32 structure-of-arrays [frame][xyz][node], re-referenced to a
// point inside the i-tile (x + mean - ref: one FP64 add per staged element, not per pair) so
// FP32 differences keep ~1e-7 relative accuracy at any box size.  Inner loop per pair-frame:
// 3 FSUB + FMUL + 2 FFMA + MUFU.SQRT + FADD; FP32 partial sums are promoted to the FP64
// per-pair totals (registers) every 16 frames.  The kernel is co-limited by the MUFU pipe
// (1 sqrt per pair-frame, 16 lanes/clk/SM) and FP32 issue (8 slots per pair-frame).
#include "common.cuh"

namespace {

constexpr int kTI = 64;
constexpr int kTJ = 128;
constexpr int kFC = 8;            // frames per staged chunk
constexpr int kThreadsC = 512;
constexpr int kColsI = 3 * kTI;   // 192
constexpr int kColsJ = 3 * kTJ;   // 384
constexpr int kCols = kColsI + kColsJ;                 // 576 doubles per frame
constexpr int kPerThread = kFC * kCols / kThreadsC;    // 9 staged elements per thread
static_assert(kFC * kCols % kThreadsC == 0, "staging must divide evenly");

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct ContactSmem {
    float si[2][kFC][3][kTI];
    float sj[2][kFC][3][kTJ];
    double offs[kCols];
};

__global__ void __launch_bounds__(kThreadsC, 1)
contact_sum_kernel(const double* __restrict__ x, int64_t ld, int64_t T, int N,
                   const double* __restrict__ mean, int n_ti, int n_tj, int n_seg,
                   double* __restrict__ out, int64_t out_ld) {
    __shared__ ContactSmem sm;

    // decode (bi, bj): bj >= bi/2
    int bi = 0, rem = blockIdx.x;
    while (rem >= n_tj - bi / 2) { rem -= n_tj - bi / 2; ++bi; }
    const int bj = bi / 2 + rem;
    const int seg = blockIdx.y;
    const int64_t f_begin = T * seg / n_seg;
    const int64_t f_end = T * (seg + 1) / n_seg;

    const int tid = threadIdx.x;
    const int ti = tid & 15;     // i group: nodes 4ti..4ti+3
    const int tj = tid >> 4;     // j group: nodes 4tj..4tj+3

    // reference point: mean position of a node in the middle of the i-tile
    int nref = kTI * bi + kTI / 2;
    if (nref > N - 1) nref = N - 1;
    for (int c = tid; c < kCols; c += kThreadsC) {
        const int64_t gcol = (c < kColsI) ? (int64_t)kColsI * bi + c : (int64_t)kColsJ * bj + (c - kColsI);
        sm.offs[c] = mean[gcol] - mean[3 * nref + (c % 3)];
    }
    __syncthreads();

    auto load_chunk = [&](int64_t f0, double (&v)[kPerThread]) {
#pragma unroll
        for (int q = 0; q < kPerThread; ++q) {
            const int e = tid + kThreadsC * q;
            const int fr = e / kCols;
            const int c = e - fr * kCols;
            const int64_t f = f0 + fr;
            const int64_t gcol = (c < kColsI) ? (int64_t)kColsI * bi + c
                                              : (int64_t)kColsJ * bj + (c - kColsI);
            v[q] = (f < f_end) ? __ldg(x + f * ld + gcol) : 0.0;
        }
    };
    auto store_chunk = [&](int buf, const double (&v)[kPerThread]) {
#pragma unroll
        for (int q = 0; q < kPerThread; ++q) {
            const int e = tid + kThreadsC * q;
            const int fr = e / kCols;
            const int c = e - fr * kCols;
            const float val = (float)(v[q] + sm.offs[c]);
            if (c < kColsI) {
                sm.si[buf][fr][c % 3][c / 3] = val;
            } else {
                const int cc = c - kColsI;
                sm.sj[buf][fr][cc % 3][cc / 3] = val;
            }
        }
    };

    double tot[4][4];
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { tot[a][b] = 0.0; acc[a][b] = 0.f; }

    double stage[kPerThread];
    load_chunk(f_begin, stage);
    store_chunk(0, stage);
    __syncthreads();

    int buf = 0;
    int chunk_idx = 0;
    for (int64_t f0 = f_begin; f0 < f_end; f0 += kFC, ++chunk_idx) {
        const bool has_next = (f0 + kFC) < f_end;
        if (has_next) load_chunk(f0 + kFC, stage);

        const int nf = (int)((f_end - f0) < kFC ? (f_end - f0) : kFC);
        if (nf == kFC) {
#pragma unroll
            for (int fl = 0; fl < kFC; ++fl) {
                const float4 ax = *reinterpret_cast<const float4*>(&sm.si[buf][fl][0][4 * ti]);
                const float4 ay = *reinterpret_cast<const float4*>(&sm.si[buf][fl][1][4 * ti]);
                const float4 az = *reinterpret_cast<const float4*>(&sm.si[buf][fl][2][4 * ti]);
                const float4 bx = *reinterpret_cast<const float4*>(&sm.sj[buf][fl][0][4 * tj]);
                const float4 by = *reinterpret_cast<const float4*>(&sm.sj[buf][fl][1][4 * tj]);
                const float4 bz = *reinterpret_cast<const float4*>(&sm.sj[buf][fl][2][4 * tj]);
                const float axs[4] = {ax.x, ax.y, ax.z, ax.w}, ays[4] = {ay.x, ay.y, ay.z, ay.w},
                            azs[4] = {az.x, az.y, az.z, az.w};
                const float bxs[4] = {bx.x, bx.y, bx.z, bx.w}, bys[4] = {by.x, by.y, by.z, by.w},
                            bzs[4] = {bz.x, bz.y, bz.z, bz.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const float dx = axs[a] - bxs[b];
                        const float dy = ays[a] - bys[b];
                        const float dz = azs[a] - bzs[b];
                        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        acc[a][b] += fast_sqrt(d2);
                    }
            }
        } else {
            for (int fl = 0; fl < nf; ++fl) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const float dx = sm.si[buf][fl][0][4 * ti + a] - sm.sj[buf][fl][0][4 * tj + b];
                        const float dy = sm.si[buf][fl][1][4 * ti + a] - sm.sj[buf][fl][1][4 * tj + b];
                        const float dz = sm.si[buf][fl][2][4 * ti + a] - sm.sj[buf][fl][2][4 * tj + b];
                        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        acc[a][b] += fast_sqrt(d2);
                    }
            }
        }
        // promote FP32 partial sums to FP64 every 16 frames (and at the end)
        if ((chunk_idx & 1) || !has_next) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) { tot[a][b] += (double)acc[a][b]; acc[a][b] = 0.f; }
        }
        if (has_next) store_chunk(buf ^ 1, stage);
        __syncthreads();
        buf ^= 1;
    }

    double* o = out + (size_t)seg * out_ld * out_ld;
    const int64_t row0 = (int64_t)kTI * bi + 4 * ti;
    const int64_t col0 = (int64_t)kTJ * bj + 4 * tj;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double* p = o + (row0 + a) * out_ld + col0;
        *reinterpret_cast<double2*>(p) = make_double2(tot[a][0], tot[a][1]);
        *reinterpret_cast<double2*>(p + 2) = make_double2(tot[a][2], tot[a][3]);
    }
}

}  // namespace

int launch_contact_sum(wisp_ctx* ctx, const double* d_x, int64_t T, int N, int64_t ld,
                       const double* d_mean, double* d_out, cudaStream_t st) {
    WISP_CHECK(T > 0 && N > 0, "contact: empty input (T=%lld, N=%d)", (long long)T, N);
    const int64_t n_pad = round_up(N, kTile);
    WISP_CHECK(3 * n_pad <= ld, "contact: ld=%lld too small for N=%d", (long long)ld, N);
    const int n_ti = (int)(n_pad / kTI), n_tj = (int)(n_pad / kTJ);
    int n_items = 0;
    for (int bi = 0; bi < n_ti; ++bi) n_items += n_tj - bi / 2;
    const int chunks_total = (int)((T + kFC - 1) / kFC);
    const int n_seg = wisp_pick_segments(n_items, chunks_total, ctx->sm_count, 16);
    double* part = d_out;
    if (n_seg > 1) {
        void* p = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_CONTACT_PART, (size_t)n_seg * n_pad * n_pad * 8, &p));
        part = (double*)p;
    }
    dim3 grid(n_items, n_seg);
    contact_sum_kernel<<<grid, kThreadsC, 0, st>>>(d_x, ld, T, N, d_mean, n_ti, n_tj, n_seg, part,
                                                   n_pad);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    if (n_seg > 1) WISP_TRY(launch_reduce_partials(ctx, part, n_seg, n_pad, d_out, kTI, kTJ, st));
    return 0;
}
