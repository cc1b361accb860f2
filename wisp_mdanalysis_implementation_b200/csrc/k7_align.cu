// K7 -- iterative alignment to the running average, sm_100a.
//
// Replaces the second half of traj_alignment_and_averaging
// (trajectory_analysis_functions.py:86-126): up to 100 passes over all frames, each frame
// needing MDAnalysis' rotation_matrix (QCP) and two numpy dots from Python.
// One iteration = two kernels:
//   align_solve_kernel   one WARP per frame: S = sum_i a_i b_i^T over the alignment landmarks (a = this
//                        frame, b = current average), 9 FP64 sums reduced by shuffles; lane 0 builds
//                        Horn's 4x4 quaternion matrix and diagonalises it with cyclic Jacobi rotations
//                        in FP64 -- the eigenvector of the largest eigenvalue is the optimal proper
//                        rotation (same minimiser as QCP / Kabsch with the reflection fix) -- and
//                        stores R[frame] (72 bytes).  64 frames per SM are in flight, so the serial
//                        eigen-solve never leaves the memory system idle;
//   align_apply_kernel   HBM-bound streaming pass: thread = one node (float64, :107) or one landmark
//                        (float32 like the reference's all_pos_Align array, :106), walks a segment of
//                        frames, rotates in place and keeps the running column sums of what it stores in
//                        registers -- the new averages (:110-111) come out of the same pass (per-segment
//                        partial rows, fixed-order finish: deterministic) instead of re-reading both arrays.
// Round 1 ran one CTA per frame with a block reduction and thread 0 solving while 255 threads waited,
// then two separate column-sum passes: 5.6 ms per iteration at C2; this form moves 7.2 GB per iteration.
// The host loop checks the RMSD between successive landmark averages against the threshold (:118).
#include "common.cuh"
#include <cmath>

int launch_colsum_f32(wisp_ctx* ctx, const float* d_x, int64_t T, int64_t ld, int64_t ncols,
                      double* d_colsum, cudaStream_t st);

namespace {

constexpr int kAlThreads = 256;

__device__ void jacobi4(double A[4][4], double V[4][4]) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 32; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
        double diag = 0.0;
        for (int p = 0; p < 4; ++p) diag += A[p][p] * A[p][p];
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
}

__global__ void __launch_bounds__(kAlThreads)
align_solve_kernel(const float* __restrict__ align, int n_align, const double* __restrict__ avg_align, int64_t T,
                   double* __restrict__ rot) {
    const int lane = threadIdx.x & 31;
    const int64_t f = (int64_t)blockIdx.x * (kAlThreads / 32) + (threadIdx.x >> 5);
    if (f >= T) return;
    const float* fa = align + f * (int64_t)n_align * 3;
    double s[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = 0.0;
    for (int i = lane; i < n_align; i += 32) {
        const double ax = fa[3 * i], ay = fa[3 * i + 1], az = fa[3 * i + 2];
        const double bx = __ldg(avg_align + 3 * i), by = __ldg(avg_align + 3 * i + 1), bz = __ldg(avg_align + 3 * i + 2);
        s[0] += ax * bx; s[1] += ax * by; s[2] += ax * bz;
        s[3] += ay * bx; s[4] += ay * by; s[5] += ay * bz;
        s[6] += az * bx; s[7] += az * by; s[8] += az * bz;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if (lane != 0) return;
    const double Sxx = s[0], Sxy = s[1], Sxz = s[2], Syx = s[3], Syy = s[4], Syz = s[5],
                 Szx = s[6], Szy = s[7], Szz = s[8];
    double A[4][4] = {
        {Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
        {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
        {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
        {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
    double V[4][4];
    jacobi4(A, V);
    int best = 0;
    for (int k = 1; k < 4; ++k)
        if (A[k][k] > A[best][best]) best = k;
    double q0 = V[0][best], qx = V[1][best], qy = V[2][best], qz = V[3][best];
    const double nq = sqrt(q0 * q0 + qx * qx + qy * qy + qz * qz);
    q0 /= nq; qx /= nq; qy /= nq; qz /= nq;
    double* R = rot + f * 9;
    R[0] = q0 * q0 + qx * qx - qy * qy - qz * qz; R[1] = 2 * (qx * qy - q0 * qz); R[2] = 2 * (qx * qz + q0 * qy);
    R[3] = 2 * (qy * qx + q0 * qz); R[4] = q0 * q0 - qx * qx + qy * qy - qz * qz; R[5] = 2 * (qy * qz - q0 * qx);
    R[6] = 2 * (qz * qx - q0 * qy); R[7] = 2 * (qz * qy + q0 * qx); R[8] = q0 * q0 - qx * qx - qy * qy + qz * qz;
}

// x <- x . R^T (row vector times R transposed == R x) for every frame of segment blockIdx.y.
// A CTA owns 128 consecutive landmarks (float32; blocks [0, gl)) or nodes (float64; blocks [gl, ..)),
// i.e. 384 consecutive columns of a frame row.  Global traffic is fully coalesced: thread t moves columns
// t, t+128, t+256 of the span (four frames per step, all loads in flight together), the (x, y, z) triples
// are regrouped through shared memory, and the running sums of the columns a thread STORES stay in its
// registers.  (A thread-per-node version with three stride-24-byte accesses per frame ran at 2.5 TB/s:
// partial-sector stores.)  part [n_seg][3 n_align + 3 N]: column sums of the values stored by a segment.
constexpr int kApplyThreads = 128;
constexpr int kApplyFrames = 4;

template <typename T>
__device__ __forceinline__ void align_apply_block(T* __restrict__ base, int64_t row_stride, int n_items, int item0,
                                                  int64_t f0, int64_t f1, const double* __restrict__ rot,
                                                  double* __restrict__ part_cols, T (*sm)[3 * kApplyThreads]) {
    const int t = threadIdx.x;
    const int ncol = 3 * min(kApplyThreads, n_items - item0);
    const int64_t col0 = 3 * (int64_t)item0;
    double s[3] = {0.0, 0.0, 0.0};
    for (int64_t f = f0; f < f1; f += kApplyFrames) {
        const int nf = (int)min((int64_t)kApplyFrames, f1 - f);
        T v[kApplyFrames][3];
#pragma unroll
        for (int q = 0; q < kApplyFrames; ++q)
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int c = t + kApplyThreads * u;
                v[q][u] = (q < nf && c < ncol) ? base[(f + q) * row_stride + col0 + c] : (T)0;
            }
#pragma unroll
        for (int q = 0; q < kApplyFrames; ++q)
#pragma unroll
            for (int u = 0; u < 3; ++u) sm[q][t + kApplyThreads * u] = v[q][u];
        __syncthreads();
        if (3 * t < ncol) {
#pragma unroll
            for (int q = 0; q < kApplyFrames; ++q)
                if (q < nf) {
                    const double* R = rot + (f + q) * 9;
                    const double x = (double)sm[q][3 * t], y = (double)sm[q][3 * t + 1], z = (double)sm[q][3 * t + 2];
                    sm[q][3 * t + 0] = (T)(x * __ldg(R + 0) + y * __ldg(R + 1) + z * __ldg(R + 2));
                    sm[q][3 * t + 1] = (T)(x * __ldg(R + 3) + y * __ldg(R + 4) + z * __ldg(R + 5));
                    sm[q][3 * t + 2] = (T)(x * __ldg(R + 6) + y * __ldg(R + 7) + z * __ldg(R + 8));
                }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kApplyFrames; ++q)
            if (q < nf) {
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int c = t + kApplyThreads * u;
                    if (c < ncol) {
                        const T o = sm[q][c];
                        base[(f + q) * row_stride + col0 + c] = o;
                        s[u] += (double)o;
                    }
                }
            }
        // (the next step's shared-memory writes of this thread go to the very slots it has just read)
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int c = t + kApplyThreads * u;
        if (c < ncol) part_cols[col0 + c] = s[u];
    }
}

__global__ void __launch_bounds__(kApplyThreads)
align_apply_kernel(float* __restrict__ align, int n_align, double* __restrict__ nodes, int64_t ld, int N,
                   int64_t T, const double* __restrict__ rot, int64_t frames_per_seg, double* __restrict__ part) {
    __shared__ __align__(16) double smd[kApplyFrames][3 * kApplyThreads];
    const int gl = (n_align + kApplyThreads - 1) / kApplyThreads;      // landmark blocks come first
    const int64_t f0 = (int64_t)blockIdx.y * frames_per_seg, f1 = min(T, f0 + frames_per_seg);
    double* prow = part + (int64_t)blockIdx.y * (3 * (int64_t)(n_align + N));
    if ((int)blockIdx.x < gl)
        align_apply_block<float>(align, 3 * (int64_t)n_align, n_align, blockIdx.x * kApplyThreads, f0, f1, rot, prow,
                                 reinterpret_cast<float (*)[3 * kApplyThreads]>(smd));
    else
        align_apply_block<double>(nodes, ld, N, (blockIdx.x - gl) * kApplyThreads, f0, f1, rot,
                                  prow + 3 * (int64_t)n_align, smd);
}

// np.sum(float32 array, axis=0) of the reference (:86): sequential float32 adds down the frames.
// `inout` holds the running float32 totals of all EARLIER frames on entry (zeros for the first frame
// block) and the totals including this block on exit (doubles holding floats): frame shards chain
// their blocks in frame order and reproduce the single sequential sum bit for bit.
// The add chain of a column is inherently serial, the memory traffic is not: a CTA owns 32 columns,
// all 256 threads stream [256 frames][32 columns] tiles (one 128-byte row per warp load, the next
// tile's loads in flight while the current one is summed) into shared memory, and warp 0 walks the
// tile in frame order.  (One thread per column reading global memory directly took 6 ms at C2 for
// 1.2 GB -- latency bound; this form is HBM bound.)
constexpr int kSeqCols = 32, kSeqFrames = 256, kSeqThreads = 256;
__global__ void __launch_bounds__(kSeqThreads)
sum32_sequential_kernel(const float* __restrict__ a, int64_t T, int64_t ncols, double* __restrict__ inout) {
    __shared__ float tile[kSeqFrames][kSeqCols];
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int64_t c = (int64_t)blockIdx.x * kSeqCols + lane;
    const bool ok = c < ncols;
    float s = (wrp == 0 && ok) ? (float)inout[c] : 0.f;
    constexpr int PER = kSeqFrames / (kSeqThreads / 32);      // rows of a tile each warp loads
    float r[PER];
    auto load = [&](int64_t f0) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int64_t f = f0 + wrp + (int64_t)k * (kSeqThreads / 32);
            r[k] = (ok && f < T) ? a[f * ncols + c] : 0.f;
        }
    };
    load(0);
    for (int64_t f0 = 0; f0 < T; f0 += kSeqFrames) {
        __syncthreads();                                   // the previous tile has been summed
#pragma unroll
        for (int k = 0; k < PER; ++k) tile[wrp + k * (kSeqThreads / 32)][lane] = r[k];
        __syncthreads();
        if (f0 + kSeqFrames < T) load(f0 + kSeqFrames);     // in flight while warp 0 sums this tile
        if (wrp == 0) {
            const int nf = (int)min((int64_t)kSeqFrames, T - f0);
#pragma unroll 8
            for (int f = 0; f < nf; ++f) s = __fadd_rn(s, tile[f][lane]);
        }
    }
    if (wrp == 0 && ok) inout[c] = (double)s;
}

// new averages and the two RMSDs' squared sums in ONE launch (a frame-sharded caller has just all-reduced
// `sums`): avg <- sums / T_total in place; res2[0] = sum over landmark columns of (old - new)^2,
// res2[1] = the same over the node columns.  Single CTA, fixed-order tree -> deterministic.
__global__ void __launch_bounds__(1024)
align_update_kernel(const double* __restrict__ sums, double T_total, int64_t na3, int64_t nn3,
                    double* __restrict__ avg, double* __restrict__ res2) {
    __shared__ double red[2][32];
    double ra = 0.0, rn = 0.0;
    for (int64_t c = threadIdx.x; c < na3 + nn3; c += blockDim.x) {
        const double v = sums[c] / T_total, d = avg[c] - v;
        if (c < na3) ra += d * d; else rn += d * d;
        avg[c] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ra += __shfl_xor_sync(0xffffffffu, ra, o); rn += __shfl_xor_sync(0xffffffffu, rn, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ra; red[1][threadIdx.x >> 5] = rn; }
    __syncthreads();
    if (threadIdx.x < 32) {
        ra = threadIdx.x < (blockDim.x >> 5) ? red[0][threadIdx.x] : 0.0;
        rn = threadIdx.x < (blockDim.x >> 5) ? red[1][threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { ra += __shfl_xor_sync(0xffffffffu, ra, o); rn += __shfl_xor_sync(0xffffffffu, rn, o); }
        if (threadIdx.x == 0) { res2[0] = ra; res2[1] = rn; }
    }
}

}  // namespace

int launch_align_update(wisp_ctx* ctx, const double* d_sums, int64_t T_total, int n_align, int N, double* d_avg,
                        double* d_res2, cudaStream_t st) {
    align_update_kernel<<<1, 1024, 0, st>>>(d_sums, (double)T_total, 3 * (int64_t)n_align, 3 * (int64_t)N, d_avg, d_res2);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

// ---- building blocks (also the device-level C ABI: wisp_dev_align_sums / wisp_dev_align_iter) ----
// d_sums [na3 + nn3]: first the landmark column sums, then the node column sums of this frame block.
// first == true: landmark sums as the reference's sequential float32 sums (initial average, :86):
// d_sums[0, na3) must hold the running totals of the earlier frame blocks on entry (zeros for the
// first block) and holds the totals including this block on exit.
int launch_align_sums(wisp_ctx* ctx, const float* d_align, int n_align, const double* d_nodes, int64_t ld,
                      int N, int64_t T, bool first, double* d_sums, cudaStream_t st) {
    const int64_t na3 = 3 * (int64_t)n_align, nn3 = 3 * (int64_t)N;
    if (first) {
        sum32_sequential_kernel<<<(int)((na3 + kSeqCols - 1) / kSeqCols), kSeqThreads, 0, st>>>(d_align, T, na3, d_sums);
        ctx->launches++;
        WISP_CUDA(cudaGetLastError());
    } else {
        WISP_TRY(launch_colsum_f32(ctx, d_align, T, na3, na3, d_sums, st));
    }
    WISP_TRY(launch_colsum(ctx, d_nodes, T, ld, nn3, d_sums + na3, st));
    return 0;
}

// one alignment pass over this frame block: rotate every frame onto d_avg_align (in place); d_sums
// [3 n_align + 3 N] (may be NULL) receives the column sums of the rotated landmarks, then of the rotated
// nodes -- what launch_align_sums(first = false) would compute, without reading either array again
int launch_align_rotate(wisp_ctx* ctx, float* d_align, int n_align, const double* d_avg_align, double* d_nodes,
                        int64_t ld, int N, int64_t T, cudaStream_t st, double* d_sums) {
    ctx->i8_stats_x = nullptr;               // the node trajectory is rotated in place
    void *pr = nullptr, *pp = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_ALIGN_ROT, (size_t)T * 9 * 8, &pr));
    double* d_rot = (double*)pr;
    align_solve_kernel<<<(unsigned)((T + kAlThreads / 32 - 1) / (kAlThreads / 32)), kAlThreads, 0, st>>>(
        d_align, n_align, d_avg_align, T, d_rot);
    const int items = n_align + N;
    const int gx = (n_align + kApplyThreads - 1) / kApplyThreads + (N + kApplyThreads - 1) / kApplyThreads;
    int n_seg = (int)std::max<int64_t>(1, std::min<int64_t>(T, (int64_t)ctx->sm_count * 6 / gx));
    const int64_t per = (T + n_seg - 1) / n_seg;
    n_seg = (int)((T + per - 1) / per);
    const int64_t ncols = 3 * (int64_t)items;
    WISP_TRY(wisp_scratch(ctx, SCR_ALIGN_PART, (size_t)n_seg * ncols * 8, &pp));
    align_apply_kernel<<<dim3(gx, n_seg), kApplyThreads, 0, st>>>(d_align, n_align, d_nodes, ld, N, T, d_rot, per,
                                                                 (double*)pp);
    ctx->launches += 2;
    WISP_CUDA(cudaGetLastError());
    if (d_sums) WISP_TRY(launch_colsum_finish(ctx, (const double*)pp, n_seg, ncols, ncols, d_sums, st));
    return 0;
}

// d_align: float32 [T][n_align][3] (in/out); d_nodes: pitched float64 trajectory (in/out), T = the
// frames of THIS shard, T_total = all frames.  `reduce` (may be empty) sums a host vector over the
// frame shards in shard order (or, for the float32 chain, runs the shards' steps one after the other in
// frame order); every shard then takes the same decisions from the same numbers.  d_work: [2 * round_up(na3,32) + round_up(nn3,32)].
// Returns iterations run; log gets (iteration, residual, node RMSD) triples.
int run_iterative_alignment(wisp_ctx* ctx, float* d_align, int n_align, double* d_nodes, int64_t ld,
                            int N, int64_t T, double threshold, int max_iter, double* d_work,
                            double* h_avg_nodes, double* log, int* n_iter_out, cudaStream_t st,
                            int64_t T_total, const AlignReduce& reduce) {
    const int64_t na3 = 3 * (int64_t)n_align, nn3 = 3 * (int64_t)N;
    if (T_total <= 0) T_total = T;
    double* d_avgA = d_work;                 // [na3]
    double* d_sums = d_avgA + round_up(na3, 32);   // [na3 + nn3]
    std::vector<double> avgA(na3), avgN(nn3), sums(na3 + nn3);
    // initial landmark totals: float32, chained over the shards in frame order (`reduce` with
    // f32_chain = true hands this shard the totals of the earlier shards, runs `step`, and returns the
    // grand totals); node sums: plain float64 sums combined afterwards
    auto step = [&](double* h_inout) -> int {
        WISP_CUDA(cudaMemcpyAsync(d_sums, h_inout, na3 * 8, cudaMemcpyHostToDevice, st));
        WISP_TRY(launch_align_sums(ctx, d_align, n_align, d_nodes, ld, N, T, true, d_sums, st));
        WISP_CUDA(cudaMemcpyAsync(sums.data(), d_sums, (na3 + nn3) * 8, cudaMemcpyDeviceToHost, st));
        WISP_CUDA(cudaStreamSynchronize(st));
        memcpy(h_inout, sums.data(), na3 * 8);
        return 0;
    };
    if (reduce) {
        std::vector<double> chain(na3, 0.0);
        WISP_TRY(reduce(chain.data(), (size_t)na3, true, step));
        memcpy(sums.data(), chain.data(), na3 * 8);
        WISP_TRY(reduce(sums.data() + na3, (size_t)nn3, false, AlignStep()));
    } else {
        std::vector<double> chain(na3, 0.0);
        WISP_TRY(step(chain.data()));
    }
    for (int64_t k = 0; k < na3; ++k) avgA[k] = (double)((float)sums[k] / (float)T_total);   // float32 mean (:86)
    for (int64_t k = 0; k < nn3; ++k) avgN[k] = sums[na3 + k] / (double)T_total;
    WISP_CUDA(cudaMemcpyAsync(d_avgA, avgA.data(), na3 * 8, cudaMemcpyHostToDevice, st));
    int iteration = 0;
    double residual = threshold + 9999.0;
    while (residual > threshold && iteration < max_iter) {
        WISP_TRY(launch_align_rotate(ctx, d_align, n_align, d_avgA, d_nodes, ld, N, T, st, d_sums));
        WISP_CUDA(cudaMemcpyAsync(sums.data(), d_sums, (na3 + nn3) * 8, cudaMemcpyDeviceToHost, st));
        WISP_CUDA(cudaStreamSynchronize(st));
        if (reduce) WISP_TRY(reduce(sums.data(), (size_t)(na3 + nn3), false, AlignStep()));
        double ra = 0.0, rn = 0.0;
        for (int64_t k = 0; k < na3; ++k) { const double v = sums[k] / (double)T_total; const double d = avgA[k] - v; ra += d * d; avgA[k] = v; }
        for (int64_t k = 0; k < nn3; ++k) { const double v = sums[na3 + k] / (double)T_total; const double d = avgN[k] - v; rn += d * d; avgN[k] = v; }
        residual = std::sqrt(ra / n_align);
        const double analysis = std::sqrt(rn / N);
        ++iteration;
        WISP_CUDA(cudaMemcpyAsync(d_avgA, avgA.data(), na3 * 8, cudaMemcpyHostToDevice, st));
        if (log) { log[3 * (iteration - 1)] = iteration; log[3 * (iteration - 1) + 1] = residual; log[3 * (iteration - 1) + 2] = analysis; }
    }
    WISP_CUDA(cudaStreamSynchronize(st));
    if (h_avg_nodes) memcpy(h_avg_nodes, avgN.data(), nn3 * 8);
    if (n_iter_out) *n_iter_out = iteration;
    return 0;
}
