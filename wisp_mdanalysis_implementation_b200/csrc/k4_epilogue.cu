// K4 -- fused N^2 epilogues (HBM-bound, one pass each).
//
//  epilogue_kernel        Gram sums + distance sums -> node covariance, variance, Pearson
//                         correlation (adjacency_matrix_functions.py:34-46), mean distance +
//                         strict '<' binary contact map (trajectory_analysis_functions.py:
//                         215-225) and the cost matrix -log|C| (.) map
//                         (functionalize_adj_matrix_functions.py:17-20), all in one pass.
//  cov_finalize_kernel    3N x 3N covariance  E[xx^T] - a a^T  from the centred product
//                         (trajectory_analysis_functions.py:168-170, SURVEY App. D rank-2 term)
//  pearson_from_cov       the literal trace / normalise loop on a given 3N x 3N matrix
//  func_pearson           -log|C| (.) map on given matrices
//
// IEEE notes that parity depends on: sqrt and division are correctly rounded (nvcc default
// -prec-sqrt/-prec-div), so C_ii = v/sqrt(v*v) == 1.0 exactly and W_ii == -0.0, never an
// edge -- same as numpy.
#include "common.cuh"

namespace {

// offs (may be NULL): column sums s = sum_t (x - r) of a trajectory that was centred with a
// provisional reference r instead of its mean; the Gram sums are then re-centred exactly with
// the parallel-axis term  G_ij -= (s_i . s_j) / T.
__device__ __forceinline__ double gram_entry(const double* __restrict__ gram, const double* __restrict__ offs,
                                             int64_t n_pad, int i, int j, double Td) {
    double g = gram[(int64_t)i * n_pad + j];
    if (offs) {
        const double dot = (offs[3 * i] * offs[3 * j] + offs[3 * i + 1] * offs[3 * j + 1]) + offs[3 * i + 2] * offs[3 * j + 2];
        g -= dot / Td;
    }
    return g / Td;
}

__global__ void epilogue_kernel(const double* __restrict__ gram, const double* __restrict__ dsum,
                                const double* __restrict__ offs,
                                int64_t n_pad, int64_t T, int N, double cutoff, int which_map,
                                double* __restrict__ var, double* __restrict__ ncov,
                                double* __restrict__ corr, double* __restrict__ avgdist,
                                int64_t* __restrict__ binary, double* __restrict__ w) {
    const int64_t total = (int64_t)N * N;
    const double Td = (double)T;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / N), c = (int)(e - (int64_t)r * N);
        const int i = r < c ? r : c, j = r < c ? c : r;
        double cij = 0.0;
        if (gram) {
            const double g = gram_entry(gram, offs, n_pad, i, j, Td);
            const double vi = gram_entry(gram, offs, n_pad, i, i, Td);
            const double vj = gram_entry(gram, offs, n_pad, j, j, Td);
            cij = g / sqrt(vi * vj);
            if (ncov) ncov[e] = g;
            if (corr) corr[e] = cij;
            if (var && r == c) var[r] = vi;
        }
        double d = 0.0;
        int64_t b = 0;
        if (dsum) {
            d = dsum[(int64_t)i * n_pad + j] / Td;
            b = (r != c && d < cutoff) ? 1 : 0;
            if (avgdist) avgdist[e] = d;
            if (binary) binary[e] = b;
        }
        if (w && gram) {
            double v = -log(fabs(cij));
            if (which_map == 1) v *= (double)b;
            else if (which_map == 2) v *= d;
            w[e] = v;
        }
    }
}

// ---- cross-GPU reduction fused into the epilogue (pipeline.cu, stage S4) ----
__device__ __forceinline__ double peer_sum(const double* const* part, int n, int64_t idx) {
    double s = part[0][idx];
    for (int q = 1; q < n; ++q) s += part[q][idx];       // fixed shard order: deterministic
    return s;
}

__global__ void diag_peers_kernel(PeerSums ps, int64_t n_pad, int64_t T, int N, double* __restrict__ diag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) diag[i] = peer_sum(ps.gram, ps.n, (int64_t)i * n_pad + i) / (double)T;
}

__global__ void epilogue_peers_kernel(PeerSums ps, const double* __restrict__ diag, int64_t n_pad, int64_t T, int N,
                                      int row0, int row1, double cutoff, int which_map,
                                      double* __restrict__ corr, double* __restrict__ avgdist,
                                      int64_t* __restrict__ binary, double* __restrict__ w) {
    const int64_t total = (int64_t)(row1 - row0) * N;
    const double Td = (double)T;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int r = row0 + (int)(e / N), c = (int)(e % N);
        const int i = r < c ? r : c, j = r < c ? c : r;
        const int64_t idx = (int64_t)i * n_pad + j;
        const double g = peer_sum(ps.gram, ps.n, idx) / Td;
        const double cij = g / sqrt(diag[i] * diag[j]);
        const double d = peer_sum(ps.dsum, ps.n, idx) / Td;
        const int64_t b = (r != c && d < cutoff) ? 1 : 0;
        corr[e] = cij;
        avgdist[e] = d;
        binary[e] = b;
        double v = -log(fabs(cij));
        if (which_map == 1) v *= (double)b;
        else if (which_map == 2) v *= d;
        w[e] = v;
    }
}

__global__ void cov_finalize_kernel(const double* __restrict__ p, int64_t ldp, int n3, int64_t T,
                                    const double* __restrict__ avg, const double* __restrict__ delta,
                                    double* __restrict__ cov) {
    const int64_t total = (int64_t)n3 * n3;
    const double Td = (double)T;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / n3), c = (int)(e - (int64_t)r * n3);
        const int i = r < c ? r : c, j = r < c ? c : r;
        // E[(x-a)(x-a)^T] + delta a^T + a delta^T  ==  E[xx^T] - a a^T   (delta = E[x] - a)
        cov[e] = p[(int64_t)i * ldp + j] / Td + (delta[i] * avg[j] + avg[i] * delta[j]);
    }
}

__global__ void pearson_from_cov_kernel(const double* __restrict__ cov, int N,
                                        double* __restrict__ var, double* __restrict__ ncov,
                                        double* __restrict__ corr) {
    const int64_t total = (int64_t)N * N;
    const int64_t n3 = 3 * (int64_t)N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / N), c = (int)(e - (int64_t)r * N);
        const int i = r < c ? r : c, j = r < c ? c : r;     // value computed for i<=j, mirrored
        const double* bi = cov + (3 * (int64_t)i) * n3;
        const double tr = __dadd_rn(__dadd_rn(bi[3 * j], bi[n3 + 3 * j + 1]), bi[2 * n3 + 3 * j + 2]);
        const double* di = cov + (3 * (int64_t)i) * n3 + 3 * i;
        const double* dj = cov + (3 * (int64_t)j) * n3 + 3 * j;
        const double vi = __dadd_rn(__dadd_rn(di[0], di[n3 + 1]), di[2 * n3 + 2]);
        const double vj = __dadd_rn(__dadd_rn(dj[0], dj[n3 + 1]), dj[2 * n3 + 2]);
        if (ncov) ncov[e] = tr;
        if (corr) corr[e] = tr / sqrt(__dmul_rn(vi, vj));
        if (var && r == c) var[r] = vi;
    }
}

__global__ void func_pearson_kernel(const double* __restrict__ corr, int64_t total,
                                    const double* __restrict__ map, double* __restrict__ w) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        double v = -log(fabs(corr[e]));
        if (map) v *= map[e];
        w[e] = v;
    }
}

inline int grid_for(wisp_ctx* ctx, int64_t total) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 16));
}

}  // namespace

int launch_epilogue(wisp_ctx* ctx, const double* d_gram, const double* d_dsum, int64_t T_total,
                    int N, double cutoff, int which_map, double* d_var, double* d_ncov,
                    double* d_corr, double* d_avgdist, int64_t* d_binary, double* d_w,
                    cudaStream_t st, const double* d_offs) {
    WISP_CHECK(N > 0 && T_total > 0, "epilogue: empty input");
    WISP_CHECK(which_map == 0 || d_dsum, "epilogue: contact-map weighting needs distance sums");
    const int64_t n_pad = round_up(N, kTile);
    epilogue_kernel<<<grid_for(ctx, (int64_t)N * N), 256, 0, st>>>(
        d_gram, d_dsum, d_offs, n_pad, T_total, N, cutoff, which_map, d_var, d_ncov, d_corr, d_avgdist,
        d_binary, d_w);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

int launch_epilogue_peers(wisp_ctx* ctx, const PeerSums& ps, int64_t T_total, int N, int row0, int row1,
                          double cutoff, int which_map, double* d_diag, double* d_corr, double* d_avgdist,
                          int64_t* d_binary, double* d_w, cudaStream_t st) {
    WISP_CHECK(N > 0 && T_total > 0 && ps.n >= 1 && ps.n <= kMaxPeers && row1 > row0, "epilogue_peers: bad argument");
    const int64_t n_pad = round_up(N, kTile);
    diag_peers_kernel<<<(N + 127) / 128, 128, 0, st>>>(ps, n_pad, T_total, N, d_diag);
    epilogue_peers_kernel<<<grid_for(ctx, (int64_t)(row1 - row0) * N), 256, 0, st>>>(
        ps, d_diag, n_pad, T_total, N, row0, row1, cutoff, which_map, d_corr, d_avgdist, d_binary, d_w);
    ctx->launches += 2;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

int launch_cov_finalize(wisp_ctx* ctx, const double* d_p, int64_t ldp, int n3, int64_t T,
                        const double* d_avg, const double* d_delta, double* d_cov,
                        cudaStream_t st) {
    cov_finalize_kernel<<<grid_for(ctx, (int64_t)n3 * n3), 256, 0, st>>>(d_p, ldp, n3, T, d_avg,
                                                                         d_delta, d_cov);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

int launch_pearson_from_cov(wisp_ctx* ctx, const double* d_cov, int N, double* d_var,
                            double* d_ncov, double* d_corr, cudaStream_t st) {
    pearson_from_cov_kernel<<<grid_for(ctx, (int64_t)N * N), 256, 0, st>>>(d_cov, N, d_var, d_ncov,
                                                                           d_corr);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

int launch_func_pearson(wisp_ctx* ctx, const double* d_corr, int N, const double* d_map,
                        double* d_w, cudaStream_t st) {
    func_pearson_kernel<<<grid_for(ctx, (int64_t)N * N), 256, 0, st>>>(d_corr, (int64_t)N * N,
                                                                       d_map, d_w);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}
