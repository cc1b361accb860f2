// K7 -- iterative alignment to the running average, sm_100a.
//
// Replaces the second half of traj_alignment_and_averaging
// (trajectory_analysis_functions.py:86-126): up to 100 passes over all frames, each frame
// needing MDAnalysis' rotation_matrix (QCP) and two numpy dots from Python.
// One CTA per frame (grid-stride):
//   1. S = sum_i a_i b_i^T over the alignment landmarks (a = this frame, b = current average)
//      -- block reduction of 9 FP64 sums;
//   2. thread 0 builds Horn's 4x4 quaternion matrix and diagonalises it with cyclic Jacobi
//      rotations in FP64; the eigenvector of the largest eigenvalue is the optimal proper
//      rotation (same minimiser as QCP / Kabsch with the reflection fix);
//   3. all threads rotate the landmarks (stored back as float32, like the reference's float32
//      all_pos_Align array, :106) and the node positions (float64, :107) in place.
// New averages are column sums (k1_com.cu) divided by T; the host loop checks the RMSD
// between successive landmark averages against the threshold (:118).
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
align_rotate_kernel(float* __restrict__ align, int n_align, const double* __restrict__ avg_align,
                    double* __restrict__ nodes, int64_t ld, int N, int64_t T) {
    __shared__ double red[9][kAlThreads / 32];
    __shared__ double Rm[9];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t f = blockIdx.x; f < T; f += gridDim.x) {
        float* fa = align + f * (int64_t)n_align * 3;
        double s[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) s[k] = 0.0;
        for (int i = tid; i < n_align; i += kAlThreads) {
            const double ax = fa[3 * i], ay = fa[3 * i + 1], az = fa[3 * i + 2];
            const double bx = avg_align[3 * i], by = avg_align[3 * i + 1], bz = avg_align[3 * i + 2];
            s[0] += ax * bx; s[1] += ax * by; s[2] += ax * bz;
            s[3] += ay * bx; s[4] += ay * by; s[5] += ay * bz;
            s[6] += az * bx; s[7] += az * by; s[8] += az * bz;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            double v = s[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[k][warp] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double S[9];
            for (int k = 0; k < 9; ++k) {
                double v = 0.0;
                for (int w = 0; w < kAlThreads / 32; ++w) v += red[k][w];
                S[k] = v;
            }
            const double Sxx = S[0], Sxy = S[1], Sxz = S[2], Syx = S[3], Syy = S[4], Syz = S[5],
                         Szx = S[6], Szy = S[7], Szz = S[8];
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
            Rm[0] = q0 * q0 + qx * qx - qy * qy - qz * qz; Rm[1] = 2 * (qx * qy - q0 * qz); Rm[2] = 2 * (qx * qz + q0 * qy);
            Rm[3] = 2 * (qy * qx + q0 * qz); Rm[4] = q0 * q0 - qx * qx + qy * qy - qz * qz; Rm[5] = 2 * (qy * qz - q0 * qx);
            Rm[6] = 2 * (qz * qx - q0 * qy); Rm[7] = 2 * (qz * qy + q0 * qx); Rm[8] = q0 * q0 - qx * qx - qy * qy + qz * qz;
        }
        __syncthreads();
        const double r0 = Rm[0], r1 = Rm[1], r2 = Rm[2], r3 = Rm[3], r4 = Rm[4], r5 = Rm[5],
                     r6 = Rm[6], r7 = Rm[7], r8 = Rm[8];
        // x <- x . R^T  (row vector times R transposed == R x)
        for (int i = tid; i < n_align; i += kAlThreads) {
            const double x = fa[3 * i], y = fa[3 * i + 1], z = fa[3 * i + 2];
            fa[3 * i + 0] = (float)(x * r0 + y * r1 + z * r2);
            fa[3 * i + 1] = (float)(x * r3 + y * r4 + z * r5);
            fa[3 * i + 2] = (float)(x * r6 + y * r7 + z * r8);
        }
        double* fn = nodes + f * ld;
        for (int i = tid; i < N; i += kAlThreads) {
            const double x = fn[3 * i], y = fn[3 * i + 1], z = fn[3 * i + 2];
            fn[3 * i + 0] = x * r0 + y * r1 + z * r2;
            fn[3 * i + 1] = x * r3 + y * r4 + z * r5;
            fn[3 * i + 2] = x * r6 + y * r7 + z * r8;
        }
        __syncthreads();
    }
}

// np.sum(float32 array, axis=0) of the reference (:86): sequential float32 adds down the frames.
// `inout` holds the running float32 totals of all EARLIER frames on entry (zeros for the first frame
// block) and the totals including this block on exit (doubles holding floats): frame shards chain
// their blocks in frame order and reproduce the single sequential sum bit for bit.
__global__ void sum32_sequential_kernel(const float* __restrict__ a, int64_t T, int64_t ncols,
                                        double* __restrict__ inout) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    float s = (float)inout[c];
    for (int64_t t = 0; t < T; ++t) s = __fadd_rn(s, a[t * ncols + c]);
    inout[c] = (double)s;
}

}  // namespace

// ---- building blocks (also the device-level C ABI: wisp_dev_align_sums / wisp_dev_align_iter) ----
// d_sums [na3 + nn3]: first the landmark column sums, then the node column sums of this frame block.
// first == true: landmark sums as the reference's sequential float32 sums (initial average, :86):
// d_sums[0, na3) must hold the running totals of the earlier frame blocks on entry (zeros for the
// first block) and holds the totals including this block on exit.
int launch_align_sums(wisp_ctx* ctx, const float* d_align, int n_align, const double* d_nodes, int64_t ld,
                      int N, int64_t T, bool first, double* d_sums, cudaStream_t st) {
    const int64_t na3 = 3 * (int64_t)n_align, nn3 = 3 * (int64_t)N;
    if (first) {
        sum32_sequential_kernel<<<(int)((na3 + 127) / 128), 128, 0, st>>>(d_align, T, na3, d_sums);
        ctx->launches++;
        WISP_CUDA(cudaGetLastError());
    } else {
        WISP_TRY(launch_colsum_f32(ctx, d_align, T, na3, na3, d_sums, st));
    }
    WISP_TRY(launch_colsum(ctx, d_nodes, T, ld, nn3, d_sums + na3, st));
    return 0;
}

// one alignment pass over this frame block: rotate every frame onto d_avg_align (in place)
int launch_align_rotate(wisp_ctx* ctx, float* d_align, int n_align, const double* d_avg_align, double* d_nodes,
                        int64_t ld, int N, int64_t T, cudaStream_t st) {
    ctx->i8_stats_x = nullptr;               // the node trajectory is rotated in place
    const int blocks = (int)std::min<int64_t>(T, (int64_t)ctx->sm_count * 8);
    align_rotate_kernel<<<blocks, kAlThreads, 0, st>>>(d_align, n_align, d_avg_align, d_nodes, ld, N, T);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
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
        WISP_TRY(launch_align_rotate(ctx, d_align, n_align, d_avgA, d_nodes, ld, N, T, st));
        WISP_TRY(launch_align_sums(ctx, d_align, n_align, d_nodes, ld, N, T, false, d_sums, st));
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
