// K1 -- per-frame translate + per-node centre of mass, plus column sums / centring.
//
// Replaces trajectory_analysis_functions.py:70-79: N+1 Python->MDAnalysis center_of_mass()
// calls per frame.  Per frame (one CTA per frame, grid-stride over frames):
//   phase 1: c = sum_a m_a x_a / sum_a m_a over the alignment atoms (float64)      (:71)
//   phase 2: every atom position becomes  float32( float64(x) - c )  -- MDAnalysis'
//            in-place translate on float32 positions -- and each node's mass-weighted mean
//            of those float32 values is taken in float64                           (:76)
//            (ATOMIC nodes: the translated float32 position itself, :78).
// The node map is a general CSR gather (node_ptr / atom_idx): nucleic-acid nodes of the
// reference span residues (make_node_selections.py:210), so "residue ranges" are not enough.
// Eight lanes cooperate on a node (coalesced 96-byte runs for contiguous residues),
// shuffle-reduce, and lane 0..2 write x,y,z into the pitched device trajectory
// X[row0 + frame][3*node + c].
//
// HBM-bound: algorithmic bytes per frame = 12*A (coords read once) + 24*N (COMs written).
// The alignment atoms are re-read in phase 2 from L1/L2, not from DRAM.
#include "common.cuh"

namespace {

constexpr int kComThreads = 256;
constexpr int kLanesPerNode = 8;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kComThreads)
com_kernel(const float* __restrict__ coords, int64_t T, int A, const double* __restrict__ masses,
           const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ atom_idx, int N,
           const int32_t* __restrict__ align_idx, int n_align, int atomic,
           double* __restrict__ nodes, int64_t ld, int64_t row0, float* __restrict__ align_out) {
    __shared__ double red[4][kComThreads / 32];
    __shared__ double shift[3];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    for (int64_t f = blockIdx.x; f < T; f += gridDim.x) {
        const float* fr = coords + f * (int64_t)A * 3;
        // ---- phase 1: alignment-selection centre of mass ----
        double sx = 0, sy = 0, sz = 0, sm = 0;
        for (int a = tid; a < n_align; a += kComThreads) {
            const int idx = __ldg(align_idx + a);
            const double m = __ldg(masses + idx);
            sx += m * (double)__ldg(fr + 3 * (int64_t)idx + 0);
            sy += m * (double)__ldg(fr + 3 * (int64_t)idx + 1);
            sz += m * (double)__ldg(fr + 3 * (int64_t)idx + 2);
            sm += m;
        }
        sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); sm = warp_sum(sm);
        if (lane == 0) { red[0][warp] = sx; red[1][warp] = sy; red[2][warp] = sz; red[3][warp] = sm; }
        __syncthreads();
        if (tid == 0) {
            double tx = 0, ty = 0, tz = 0, tm = 0;
            for (int w = 0; w < kComThreads / 32; ++w) {
                tx += red[0][w]; ty += red[1][w]; tz += red[2][w]; tm += red[3][w];
            }
            shift[0] = -(tx / tm); shift[1] = -(ty / tm); shift[2] = -(tz / tm);
        }
        __syncthreads();
        const double hx = shift[0], hy = shift[1], hz = shift[2];

        if (align_out) {   // all_pos_Align.append(u_alignment.positions)  (:72)
            float* ao = align_out + (row0 + f) * (int64_t)n_align * 3;
            for (int a = tid; a < n_align; a += kComThreads) {
                const float* p = fr + 3 * (int64_t)__ldg(align_idx + a);
                ao[3 * a + 0] = __double2float_rn((double)__ldg(p + 0) + hx);
                ao[3 * a + 1] = __double2float_rn((double)__ldg(p + 1) + hy);
                ao[3 * a + 2] = __double2float_rn((double)__ldg(p + 2) + hz);
            }
        }
        // ---- phase 2: node centres of mass ----
        const int group = tid / kLanesPerNode;
        const int gl = tid % kLanesPerNode;
        double* out = nodes + (row0 + f) * ld;
        // trip count is uniform across the block: the shuffles below need every lane present
        for (int n0 = 0; n0 < N; n0 += kComThreads / kLanesPerNode) {
            const int n = n0 + group;
            const bool valid = n < N;
            const int e0 = valid ? __ldg(node_ptr + n) : 0, e1 = valid ? __ldg(node_ptr + n + 1) : 0;
            double ax = 0, ay = 0, az = 0, am = 0;
            for (int e = e0 + gl; e < e1; e += kLanesPerNode) {
                const int idx = __ldg(atom_idx + e);
                const float* p = fr + 3 * (int64_t)idx;
                // translate(): float32 position += float64 vector, rounded back to float32
                const double px = (double)__double2float_rn((double)__ldg(p + 0) + hx);
                const double py = (double)__double2float_rn((double)__ldg(p + 1) + hy);
                const double pz = (double)__double2float_rn((double)__ldg(p + 2) + hz);
                if (atomic) {
                    if (e == e0) { ax = px; ay = py; az = pz; am = 1.0; }
                } else {
                    const double m = __ldg(masses + idx);
                    ax += m * px; ay += m * py; az += m * pz; am += m;
                }
            }
#pragma unroll
            for (int o = kLanesPerNode / 2; o > 0; o >>= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, o);
                ay += __shfl_xor_sync(0xffffffffu, ay, o);
                az += __shfl_xor_sync(0xffffffffu, az, o);
                am += __shfl_xor_sync(0xffffffffu, am, o);
            }
            // one division per warp instruction: lanes 0..2 of the group pick x, y, z
            const double num = (gl == 0) ? ax : ((gl == 1) ? ay : az);
            if (valid && gl < 3) out[3 * (int64_t)n + gl] = num / am;
        }
        __syncthreads();   // shift[] / red[] reuse
    }
}

// ---- column sums: part[seg][col] then fixed-order reduction ----
constexpr int kCsCols = 32, kCsRows = 8;

template <typename TIn>
__global__ void colsum_part_kernel(const TIn* __restrict__ x, int64_t T, int64_t ld,
                                   int64_t ncols, int n_seg, double* __restrict__ part) {
    __shared__ double sm[kCsRows][kCsCols + 1];
    const int cx = threadIdx.x, ry = threadIdx.y;
    const int64_t col = (int64_t)blockIdx.x * kCsCols + cx;
    const int seg = blockIdx.y;
    const int64_t t0 = T * seg / n_seg, t1 = T * (seg + 1) / n_seg;
    double s = 0.0;
    if (col < ncols)
        for (int64_t t = t0 + ry; t < t1; t += kCsRows) s += (double)x[t * ld + col];
    sm[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && col < ncols) {
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < kCsRows; ++k) tot += sm[k][cx];
        part[(int64_t)seg * ncols + col] = tot;
    }
}

__global__ void colsum_final_kernel(const double* __restrict__ part, int n_seg, int64_t row_stride,
                                    int64_t ncols, double* __restrict__ out) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    double s = 0.0;
    for (int k = 0; k < n_seg; ++k) s += part[(int64_t)k * row_stride + col];
    out[col] = s;
}

__global__ void center_kernel(double* __restrict__ x, int64_t T, int64_t ld, int64_t ncols,
                              const double* __restrict__ mean) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    const double m = mean[col];
    for (int64_t t = blockIdx.y; t < T; t += gridDim.y) x[t * ld + col] -= m;
}

// Centring fused with the FP32 tile-referenced copy K3 consumes:
//   X[t][col] -= mean[col]                                   (in place, float64)
//   P[t][tile][c][n] = float( x - ref_tile,c )               (x = un-centred value)
// One CTA of 384 threads per tile, kCcFrames frames per step: thread = column, so the 3 KB rows
// of X are read and written coalesced; the (node, component) -> (component, node) transpose of P
// goes through shared memory, 1.5 KB coalesced writes.  Each thread also keeps max|x| and sum x^2
// of its centred column: the statistics the INT8 K2 engine scales by.
// ROT: x is first rotated by the frame's matrix rot[t] (the node rotation the K7 loop deferred:
// trajectory_analysis_functions.py:107 applied once with the composed rotation) -- the thread of column
// 3n + c picks up the node's other two components from shared memory and forms row c of R x.
constexpr int kCcFrames = 4;

// CENTER == false: X is left alone and only P is written, against the references in `mean` (any fixed
// point per tile will do for K3, which takes the tile-pair offset from the same vector): the conversion the
// streamed pipeline runs on every coordinate block as it arrives (pipeline.cu), long before the mean is known.
// WRITE_P == false: centring (+ rotation) + statistics only -- P was already consumed.
template <bool ROT, bool CENTER, bool WRITE_P>
__global__ void __launch_bounds__(384)
center_convert_kernel(double* __restrict__ x, int64_t T, int64_t ld, int N,
                      const double* __restrict__ mean, float* __restrict__ p32, int n_tiles,
                      double* __restrict__ stat, const double* __restrict__ rot) {
    __shared__ float s[WRITE_P ? kCcFrames : 1][3][kTile + 1];
    __shared__ double raw[ROT ? kCcFrames : 1][ROT ? 3 * kTile : 1];
    const int tile = blockIdx.x;
    const int k = threadIdx.x;
    const int64_t col = (int64_t)3 * kTile * tile + k;
    const int c = k % 3, n = k / 3;
    int nref = kTile * tile + kTile / 2;
    if (nref > N - 1) nref = N - 1;
    const double m = CENTER ? mean[col] : 0.0;
    const double ref = mean[3 * nref + c];
    double amax = 0.0, ssq = 0.0;
    for (int64_t t0 = (int64_t)blockIdx.y * kCcFrames; t0 < T; t0 += (int64_t)gridDim.y * kCcFrames) {
        double v[kCcFrames];
#pragma unroll
        for (int f = 0; f < kCcFrames; ++f) v[f] = (t0 + f < T) ? x[(t0 + f) * ld + col] : (ROT ? 0.0 : m);
        if (ROT) {
#pragma unroll
            for (int f = 0; f < kCcFrames; ++f) raw[f][k] = v[f];
            __syncthreads();
#pragma unroll
            for (int f = 0; f < kCcFrames; ++f) {
                if (t0 + f < T) {
                    const double* R = rot + (t0 + f) * 9 + 3 * c;
                    v[f] = raw[f][3 * n] * __ldg(R) + raw[f][3 * n + 1] * __ldg(R + 1) + raw[f][3 * n + 2] * __ldg(R + 2);
                } else {
                    v[f] = m;
                }
            }
        }
#pragma unroll
        for (int f = 0; f < kCcFrames; ++f) {
            if (CENTER) {
                const double xc = v[f] - m;
                if (t0 + f < T) x[(t0 + f) * ld + col] = xc;
                amax = fmax(amax, fabs(xc));
                ssq = fma(xc, xc, ssq);
            }
            if (WRITE_P) s[f][c][n] = (float)(v[f] - ref);
        }
        if (WRITE_P || ROT) __syncthreads();
        if (WRITE_P) {
#pragma unroll
            for (int f = 0; f < kCcFrames; ++f)
                if (t0 + f < T)
                    p32[((t0 + f) * n_tiles + tile) * (int64_t)(3 * kTile) + k] = s[f][k / kTile][k % kTile];
            __syncthreads();
        }
    }
    if (CENTER && stat) {              // [gridDim.y][2][3 * 128 * n_tiles]
        const int64_t w = (int64_t)3 * kTile * n_tiles;
        stat[((int64_t)blockIdx.y * 2 + 0) * w + col] = amax;
        stat[((int64_t)blockIdx.y * 2 + 1) * w + col] = ssq;
    }
}

}  // namespace

int launch_com(wisp_ctx* ctx, const float* d_coords, int64_t T, int A, const double* d_masses,
               const int32_t* d_node_ptr, const int32_t* d_atom_idx, int N,
               const int32_t* d_align_idx, int n_align, int atomic, double* d_nodes, int64_t ld,
               int64_t row0, cudaStream_t st, float* d_align_out) {
    WISP_CHECK(T > 0 && A > 0 && N > 0 && n_align > 0, "com: empty input (T=%lld A=%d N=%d nAlign=%d)",
               (long long)T, A, N, n_align);
    ctx->i8_stats_x = nullptr;        // the node trajectory is being rewritten
    const int blocks = (int)std::min<int64_t>(T, (int64_t)ctx->sm_count * 8);
    com_kernel<<<blocks, kComThreads, 0, st>>>(d_coords, T, A, d_masses, d_node_ptr, d_atom_idx, N,
                                              d_align_idx, n_align, atomic, d_nodes, ld, row0, d_align_out);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

template <typename TIn>
static int launch_colsum_t(wisp_ctx* ctx, const TIn* d_x, int64_t T, int64_t ld, int64_t ncols,
                           double* d_colsum, cudaStream_t st) {
    const int col_blocks = (int)((ncols + kCsCols - 1) / kCsCols);
    int n_seg = (int)std::max<int64_t>(1, std::min<int64_t>(T / 64, (ctx->sm_count * 8 + col_blocks - 1) / col_blocks));
    n_seg = std::min(n_seg, 256);
    void* p = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_COLSUM_PART, (size_t)n_seg * ncols * 8, &p));
    colsum_part_kernel<TIn><<<dim3(col_blocks, n_seg), dim3(kCsCols, kCsRows), 0, st>>>(
        d_x, T, ld, ncols, n_seg, (double*)p);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    colsum_final_kernel<<<(int)((ncols + 255) / 256), 256, 0, st>>>((const double*)p, n_seg, ncols, ncols,
                                                                    d_colsum);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

int launch_colsum_finish(wisp_ctx* ctx, const double* d_part, int n_rows, int64_t row_stride, int64_t ncols,
                         double* d_colsum, cudaStream_t st) {
    colsum_final_kernel<<<(int)((ncols + 255) / 256), 256, 0, st>>>(d_part, n_rows, row_stride, ncols, d_colsum);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

int launch_colsum(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int64_t ncols,
                  double* d_colsum, cudaStream_t st) {
    return launch_colsum_t<double>(ctx, d_x, T, ld, ncols, d_colsum, st);
}

int launch_colsum_f32(wisp_ctx* ctx, const float* d_x, int64_t T, int64_t ld, int64_t ncols,
                      double* d_colsum, cudaStream_t st) {
    return launch_colsum_t<float>(ctx, d_x, T, ld, ncols, d_colsum, st);
}

// FP32 tile-referenced copy of the rows [0, T) of d_x against the references in d_ref (layout of a mean
// vector); X is not modified.
int launch_p32_convert(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int N, const double* d_ref,
                       float* d_p32, cudaStream_t st) {
    const int n_tiles = (int)(round_up(N, kTile) / kTile);
    const int gy = (int)std::min<int64_t>((T + kCcFrames - 1) / kCcFrames,
                                          std::max<int64_t>(1, (int64_t)ctx->sm_count * 8 / n_tiles));
    center_convert_kernel<false, false, true><<<dim3(n_tiles, gy), 384, 0, st>>>(
        const_cast<double*>(d_x), T, ld, N, d_ref, d_p32, n_tiles, nullptr, nullptr);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

int launch_center(wisp_ctx* ctx, double* d_x, int64_t T, int64_t ld, int64_t ncols,
                  const double* d_mean, cudaStream_t st, float* d_p32, int N, const double* d_rot,
                  bool stats_without_p32) {
    if (d_rot && !d_p32 && !stats_without_p32) {
        // no fused form without the FP32 copy: rotate in place, then centre
        WISP_CHECK(N > 0, "center: the node count is needed to rotate");
        WISP_TRY(launch_align_nodes(ctx, d_x, ld, N, T, d_rot, st));
    }
    if (d_p32 || stats_without_p32) {
        WISP_CHECK(N > 0, "center: the node count is needed for the fused centring pass");
        const int n_tiles = (int)(round_up(N, kTile) / kTile);
        const int gy = (int)std::min<int64_t>((T + kCcFrames - 1) / kCcFrames,
                                              std::max<int64_t>(1, (int64_t)ctx->sm_count * 8 / n_tiles));
        const int parts = gy;
        // per-column statistics of the centred block ride along: the INT8 K2 engine takes them from
        // here instead of reading X once more, if the next wisp_dev_gram is on this very block
        void* p_stat = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_I8_STATS, (size_t)parts * 2 * 3 * kTile * n_tiles * 8, &p_stat));
        const dim3 grid(n_tiles, gy);
        double* ps = (double*)p_stat;
        if (d_p32) {
            if (d_rot) center_convert_kernel<true, true, true><<<grid, 384, 0, st>>>(d_x, T, ld, N, d_mean, d_p32, n_tiles, ps, d_rot);
            else center_convert_kernel<false, true, true><<<grid, 384, 0, st>>>(d_x, T, ld, N, d_mean, d_p32, n_tiles, ps, nullptr);
        } else {
            if (d_rot) center_convert_kernel<true, true, false><<<grid, 384, 0, st>>>(d_x, T, ld, N, d_mean, nullptr, n_tiles, ps, d_rot);
            else center_convert_kernel<false, true, false><<<grid, 384, 0, st>>>(d_x, T, ld, N, d_mean, nullptr, n_tiles, ps, nullptr);
        }
        ctx->launches++;
        WISP_CUDA(cudaGetLastError());
        ctx->i8_stats_x = d_x;
        ctx->i8_stats_T = T;
        ctx->i8_stats_parts = parts;
        ctx->i8_stats_stream = st;
        ctx->i8_stats_launch = ctx->launches;      // valid only while no other launch of this library follows
        return 0;
    }
    ctx->i8_stats_x = nullptr;                     // plain centring: no statistics for this content
    const int gx = (int)((ncols + 255) / 256);
    const int gy = (int)std::min<int64_t>(T, std::max<int64_t>(1, (int64_t)ctx->sm_count * 16 / gx));
    center_kernel<<<dim3(gx, gy), 256, 0, st>>>(d_x, T, ld, ncols, d_mean);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}
