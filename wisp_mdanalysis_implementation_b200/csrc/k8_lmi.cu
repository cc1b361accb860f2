// K8 -- linear mutual information adjacency and its functionalisation.
//
//  lmi_kernel      adjacency_matrix_functions.py:62-103: for every node pair the reference
//                  calls np.linalg.det three times from Python (two 3x3, one 6x6 gathered with
//                  np.ix_).  One thread per pair: closed-form 3x3 determinants, the 6x6 one by
//                  LU with partial pivoting in registers (the same factorisation LAPACK's
//                  getrf performs), MI = 0.5 ln(det_i det_j / det_ij), +inf on the diagonal.
//  gencorr_kernel  functionalize_adj_matrix_functions.py:27-46: r = sqrt(1 - exp(-2 MI/3)),
//                  optional exp(-d/Lambda) damping, cost = -ln r.
#include "common.cuh"

namespace {

__device__ __forceinline__ double det3(const double* __restrict__ c, int64_t ld) {
    // LU with partial pivoting on a 3x3 (matches getrf's pivot order)
    double a[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) a[r][q] = c[r * ld + q];
    double det = 1.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        double best = fabs(a[k][k]);
#pragma unroll
        for (int r = k + 1; r < 3; ++r)
            if (fabs(a[r][k]) > best) { best = fabs(a[r][k]); piv = r; }
        if (piv != k) {
#pragma unroll
            for (int q = 0; q < 3; ++q) { const double t = a[k][q]; a[k][q] = a[piv][q]; a[piv][q] = t; }
            det = -det;
        }
        det *= a[k][k];
        if (a[k][k] == 0.0) return 0.0;
#pragma unroll
        for (int r = k + 1; r < 3; ++r) {
            const double f = a[r][k] / a[k][k];
#pragma unroll
            for (int q = k + 1; q < 3; ++q) a[r][q] -= f * a[k][q];
        }
    }
    return det;
}

__device__ double det6(double a[6][6]) {
    double det = 1.0;
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        double best = fabs(a[k][k]);
        for (int r = k + 1; r < 6; ++r)
            if (fabs(a[r][k]) > best) { best = fabs(a[r][k]); piv = r; }
        if (piv != k) {
            for (int q = 0; q < 6; ++q) { const double t = a[k][q]; a[k][q] = a[piv][q]; a[piv][q] = t; }
            det = -det;
        }
        det *= a[k][k];
        if (a[k][k] == 0.0) return 0.0;
        for (int r = k + 1; r < 6; ++r) {
            const double f = a[r][k] / a[k][k];
            for (int q = k + 1; q < 6; ++q) a[r][q] -= f * a[k][q];
        }
    }
    return det;
}

__global__ void lmi_kernel(const double* __restrict__ cov, int N, double* __restrict__ mi) {
    const int64_t n3 = 3 * (int64_t)N;
    const int64_t total = (int64_t)N * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / N), c = (int)(e - (int64_t)r * N);
        if (r == c) { mi[e] = __longlong_as_double(0x7ff0000000000000LL); continue; }   // 0 + inf (:89)
        if (r > c) continue;                                      // computed for i<j, mirrored
        const int i = r, j = c;
        const double di = det3(cov + (3 * (int64_t)i) * n3 + 3 * i, n3);
        const double dj = det3(cov + (3 * (int64_t)j) * n3 + 3 * j, n3);
        double a[6][6];
        for (int p = 0; p < 6; ++p) {
            const int64_t row = (p < 3) ? 3 * (int64_t)i + p : 3 * (int64_t)j + (p - 3);
            for (int q = 0; q < 6; ++q) {
                const int64_t col = (q < 3) ? 3 * (int64_t)i + q : 3 * (int64_t)j + (q - 3);
                a[p][q] = cov[row * n3 + col];
            }
        }
        const double dij = det6(a);
        const double v = 0.5 * log((di * dj) / dij);
        mi[(int64_t)i * N + j] = v;
        mi[(int64_t)j * N + i] = v;
    }
}

__global__ void gencorr_kernel(const double* __restrict__ mi, int64_t total, const double* __restrict__ map,
                               double lambda, int has_lambda, double* __restrict__ rmi,
                               double* __restrict__ func) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        double r = sqrt(1.0 - exp(-2.0 * mi[e] / 3));
        if (map && has_lambda) r *= exp(-map[e] / lambda);
        if (rmi) rmi[e] = r;
        if (func) func[e] = -log(r);
    }
}

}  // namespace

extern "C" int wisp_lmi_from_covariance(wisp_ctx* ctx, const double* cov, int N, double* out_mi) {
    WISP_CHECK(ctx && cov && out_mi && N > 0, "lmi_from_covariance: bad argument");
    WISP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n3 = 3 * (size_t)N, nn = (size_t)N * N;
    void *d_cov, *d_mi;
    WISP_TRY(wisp_scratch(ctx, SCR_ARENA, n3 * n3 * 8 + nn * 8 + 512, &d_cov));
    d_mi = (char*)d_cov + round_up((int64_t)(n3 * n3 * 8), 256);
    WISP_CUDA(cudaMemcpyAsync(d_cov, cov, n3 * n3 * 8, cudaMemcpyHostToDevice, st));
    const int g = (int)std::max<int64_t>(1, std::min<int64_t>((nn + 127) / 128, (int64_t)ctx->sm_count * 16));
    lmi_kernel<<<g, 128, 0, st>>>((const double*)d_cov, N, (double*)d_mi);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    WISP_CUDA(cudaMemcpyAsync(out_mi, d_mi, nn * 8, cudaMemcpyDeviceToHost, st));
    WISP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int wisp_func_generalized_correlation(wisp_ctx* ctx, const double* mi, int N,
                                                 const double* contact_map, double lambda, int has_lambda,
                                                 double* out_rmi, double* out_func) {
    WISP_CHECK(ctx && mi && N > 0 && (out_rmi || out_func), "func_generalized_correlation: bad argument");
    WISP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t nn = (size_t)N * N;
    void* base;
    WISP_TRY(wisp_scratch(ctx, SCR_ARENA, 4 * (nn * 8 + 256), &base));
    double* d_mi = (double*)base;
    double* d_map = (double*)((char*)base + round_up((int64_t)(nn * 8), 256));
    double* d_r = (double*)((char*)d_map + round_up((int64_t)(nn * 8), 256));
    double* d_f = (double*)((char*)d_r + round_up((int64_t)(nn * 8), 256));
    WISP_CUDA(cudaMemcpyAsync(d_mi, mi, nn * 8, cudaMemcpyHostToDevice, st));
    if (contact_map) WISP_CUDA(cudaMemcpyAsync(d_map, contact_map, nn * 8, cudaMemcpyHostToDevice, st));
    const int g = (int)std::max<int64_t>(1, std::min<int64_t>((nn + 255) / 256, (int64_t)ctx->sm_count * 16));
    gencorr_kernel<<<g, 256, 0, st>>>(d_mi, (int64_t)nn, contact_map ? d_map : nullptr, lambda, has_lambda, d_r, d_f);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    if (out_rmi) WISP_CUDA(cudaMemcpyAsync(out_rmi, d_r, nn * 8, cudaMemcpyDeviceToHost, st));
    if (out_func) WISP_CUDA(cudaMemcpyAsync(out_func, d_f, nn * 8, cudaMemcpyDeviceToHost, st));
    WISP_CUDA(cudaStreamSynchronize(st));
    return 0;
}
