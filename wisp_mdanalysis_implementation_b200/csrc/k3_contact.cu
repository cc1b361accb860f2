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
// One CTA = one 128 x 128 pair tile (upper triangle only) x one frame segment; 512 threads,
// each owns 8 x 4 pairs.  Frames are staged 8 at a time (float4 loads, double-buffered).
// Per pair-frame: 3 FADD + FMUL + 2 FFMA + MUFU.SQRT + FADD -- the kernel is co-limited by
// the MUFU pipe (16 lanes/clk/SM) and FP32 issue.  FP32 partial sums are promoted every 64
// frames into FP64 per-pair totals kept in shared memory (128 KB), written once at the end.
#include "common.cuh"

namespace {

constexpr int kFC = 8;             // frames per staged chunk
constexpr int kThreadsC = 512;
constexpr int kFlushChunks = 8;    // promote FP32 -> FP64 every 64 frames
constexpr int kTileFloats = 3 * kTile;   // 384 floats per tile-frame

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct ContactSmem {
    double tot[32 * kThreadsC];            // [pair][thread] -> conflict-free flush
    float si[2][kFC][3][kTile];
    float sj[2][kFC][3][kTile];
};

__global__ void __launch_bounds__(kThreadsC, 1)
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

    const int tid = threadIdx.x;
    const int ti = tid & 15;     // i nodes: 4ti..4ti+3 and 64+4ti..64+4ti+3
    const int tj = tid >> 4;     // j nodes: 4tj..4tj+3

    // reference offset between the two tiles: x_i - x_j = a_i - (b_j + (ref_J - ref_I))
    int ri = kTile * bi + kTile / 2, rj = kTile * bj + kTile / 2;
    if (ri > N - 1) ri = N - 1;
    if (rj > N - 1) rj = N - 1;
    float roff[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) roff[c] = (float)(mean[3 * rj + c] - mean[3 * ri + c]);

    // staging: 8 frames x (i tile + j tile) x 96 float4 = 1536 float4 per chunk, 3 per thread
    auto load_chunk = [&](int64_t f0, float4 (&v)[3]) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
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
    auto store_chunk = [&](int buf, const float4 (&v)[3]) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int e = tid + kThreadsC * q;
            const int fl = e / 192;
            const int r = e - fl * 192;
            const int sel = r / 96;
            const int w = r - sel * 96;
            const int c = w / 32;
            float4 x = v[q];
            if (sel) {
                const float o = roff[c];
                x.x += o; x.y += o; x.z += o; x.w += o;
                *reinterpret_cast<float4*>(&sm.sj[buf][fl][0][0] + 4 * w) = x;
            } else {
                *reinterpret_cast<float4*>(&sm.si[buf][fl][0][0] + 4 * w) = x;
            }
        }
    };

    float acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll
    for (int p = 0; p < 32; ++p) sm.tot[p * kThreadsC + tid] = 0.0;

    float4 stage[3];
    load_chunk(f_begin, stage);
    store_chunk(0, stage);
    __syncthreads();

    int buf = 0;
    int chunk_idx = 0;
    for (int64_t f0 = f_begin; f0 < f_end; f0 += kFC, ++chunk_idx) {
        const bool has_next = (f0 + kFC) < f_end;
        if (has_next) load_chunk(f0 + kFC, stage);
        const int nf = (int)((f_end - f0) < kFC ? (f_end - f0) : kFC);
#pragma unroll 2
        for (int fl = 0; fl < nf; ++fl) {
            float ax[8], ay[8], az[8], bx[4], by[4], bz[4];
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
                bx[0] = w0.x; bx[1] = w0.y; bx[2] = w0.z; bx[3] = w0.w;
                by[0] = w1.x; by[1] = w1.y; by[2] = w1.z; by[3] = w1.w;
                bz[0] = w2.x; bz[1] = w2.y; bz[2] = w2.z; bz[3] = w2.w;
            }
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float dx = ax[a] - bx[b];
                    const float dy = ay[a] - by[b];
                    const float dz = az[a] - bz[b];
                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    acc[a][b] += fast_sqrt(d2);
                }
        }
        if ((chunk_idx % kFlushChunks) == kFlushChunks - 1 || !has_next) {
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    sm.tot[(a * 4 + b) * kThreadsC + tid] += (double)acc[a][b];
                    acc[a][b] = 0.f;
                }
        }
        if (has_next) store_chunk(buf ^ 1, stage);
        __syncthreads();
        buf ^= 1;
    }

    double* o = out + (size_t)seg * out_ld * out_ld;
    const int64_t col0 = (int64_t)kTile * bj + 4 * tj;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int64_t row = (int64_t)kTile * bi + ((a < 4) ? (4 * ti + a) : (64 + 4 * ti + (a - 4)));
        double* p = o + row * out_ld + col0;
        *reinterpret_cast<double2*>(p) =
            make_double2(sm.tot[(a * 4 + 0) * kThreadsC + tid], sm.tot[(a * 4 + 1) * kThreadsC + tid]);
        *reinterpret_cast<double2*>(p + 2) =
            make_double2(sm.tot[(a * 4 + 2) * kThreadsC + tid], sm.tot[(a * 4 + 3) * kThreadsC + tid]);
    }
}

}  // namespace

int launch_contact_sum(wisp_ctx* ctx, const float* d_p32, int64_t T, int N, const double* d_mean,
                       double* d_out, cudaStream_t st) {
    WISP_CHECK(T > 0 && N > 0, "contact: empty input (T=%lld, N=%d)", (long long)T, N);
    const int64_t n_pad = round_up(N, kTile);
    const int n_tiles = (int)(n_pad / kTile);
    const int n_items = n_tiles * (n_tiles + 1) / 2;
    const int chunks_total = (int)((T + kFC - 1) / kFC);
    const int n_seg = wisp_pick_segments(n_items, chunks_total, ctx->sm_count, 16);
    double* part = d_out;
    if (n_seg > 1) {
        void* p = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_CONTACT_PART, (size_t)n_seg * n_pad * n_pad * 8, &p));
        part = (double*)p;
    }
    const int smem = (int)sizeof(ContactSmem);
    WISP_CUDA(cudaFuncSetAttribute(contact_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid(n_items, n_seg);
    contact_sum_kernel<<<grid, kThreadsC, smem, st>>>(d_p32, T, N, d_mean, n_tiles, n_seg, part, n_pad);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    if (n_seg > 1) WISP_TRY(launch_reduce_partials(ctx, part, n_seg, n_pad, d_out, kTile, kTile, st));
    return 0;
}
