// K7 -- iterative alignment to the running average, sm_100a.
//
// Replaces the second half of traj_alignment_and_averaging
// (trajectory_analysis_functions.py:86-126): up to 100 passes over all frames, each frame
// needing MDAnalysis' rotation_matrix (QCP) and two numpy dots from Python.
// One iteration = three kernels:
//   align_corr_kernel + align_quat_kernel
//                        one WARP per frame: S = sum_i a_i b_i^T over the alignment landmarks (a = this
//                        frame, b = current average), 9 FP64 sums reduced by shuffles; then one THREAD per
//                        frame finds the optimal proper rotation (align_solve.cuh: Newton on the quartic of
//                        Horn's quaternion matrix + adjugate column + Rayleigh polish; Jacobi only for
//                        degenerate spectra), stores R[frame] and composes it into the frame's TOTAL rotation
//                        R_tot[frame] <- R R_tot[frame];
//   align_apply_kernel   HBM-bound streaming pass, warp-autonomous (no block barrier anywhere): a warp
//                        owns 32 landmarks (float32 like the reference's all_pos_Align array, :106) or 32
//                        nodes (float64, :107) and a segment of frames.  Landmarks are rotated in place
//                        by R; the NODE trajectory is only READ -- the column sums of R_tot x_original
//                        are what the iteration needs of it (:111, the printed node RMSD and, after the
//                        last iteration, the frame mean), so the 2.4 GB array is not rewritten every
//                        iteration.  The one in-place rotation of the nodes is fused into the centring
//                        pass that follows the loop (launch_center with d_rot), or done by
//                        launch_align_nodes for callers that want the aligned trajectory itself.
//                        Global traffic is fully coalesced (lane l moves columns l, l+32, l+64 of the
//                        warp's 96-column span, four frames in flight), the (x, y, z) triples are regrouped
//                        through a warp-private shared-memory patch, and the running sums of the columns
//                        stay in registers (per-segment partial rows, fixed-order finish: deterministic).
// Rounding: the reference rounds the nodes to float64 after every iteration's rotation; composing the
// rotations first differs from that by ~1e-16 relative (the landmarks, whose float32 rounding per
// iteration IS visible in the result, are rotated every iteration exactly as before).
// Bytes per iteration and frame: 36 n_align (landmarks: read by the solve, read + written by the apply
// pass) + 24 N (nodes read) -- round 2's first cut moved 36 n_align + 48 N and round 1 twice that.
// The host loop checks the RMSD between successive landmark averages against the threshold (:118).
#include "common.cuh"
#include "align_solve.cuh"
#include <cmath>
#include <cstdlib>
#include <cuda.h>

int launch_colsum_f32(wisp_ctx* ctx, const float* d_x, int64_t T, int64_t ld, int64_t ncols,
                      double* d_colsum, cudaStream_t st);

namespace {

constexpr int kAlThreads = 256;

// Two kernels per solve: the correlation sums are a pure stream of the landmark array and want many light
// warps in flight; the eigen-solve is ~2 us of dependent FP64 arithmetic per frame on ~100 registers.  In
// one kernel (one warp per frame, lane 0 solving) the solver's registers cut the occupancy to 16 warps per
// SM and the stream ran at 3.0 TB/s (0.40 ms per iteration at C2); apart, the stream kernel keeps 40+
// warps per SM busy and the solve kernel runs one THREAD per frame over all frames at once.
//   align_corr_kernel   rot[f][0..9) <- S = sum_i a_i b_i^T (row-major), one warp per frame;
//   align_quat_kernel   rot[f] <- R(S) in place; rot_total[f] (may be NULL) <- compose ? R rot_total[f] : R.
// gate (may be NULL): device flag set by align_update_kernel once the residual is below the threshold.
// SMEM_B: the reference landmarks live in shared memory, transposed to [12][n_align / 4] so that the lanes
// of a warp (which own consecutive groups of four landmarks) read consecutive doubles.  Read through L1
// instead (lane stride 96 bytes) every one of the 12 loads per group touches ~24 cache lines and the L1
// wavefront rate, not HBM, sets the pace (0.37 ms per iteration at C2 against 0.2 ms for the stream).
// Persistent CTAs: the 24 n_align bytes are staged once per CTA.
template <bool SMEM_B>
__global__ void __launch_bounds__(kAlThreads)
align_corr_kernel(const float* __restrict__ align, int n_align, const double* __restrict__ avg_align, int64_t T,
                  double* __restrict__ rot, const int* __restrict__ gate) {
    extern __shared__ __align__(16) double corr_b[];                 // [12][groups] when SMEM_B
    if (gate && *gate) return;          // the loop has converged: a speculatively launched iteration is a no-op
    const int lane = threadIdx.x & 31;
    const int groups = n_align >> 2;
    if (SMEM_B) {
        for (int e = threadIdx.x; e < 3 * n_align; e += kAlThreads) {
            const int g = e / 12, j = e - 12 * g;
            corr_b[(size_t)j * groups + g] = avg_align[e];
        }
        __syncthreads();
    }
    const int64_t f_stride = (int64_t)gridDim.x * (kAlThreads / 32);
    for (int64_t f = (int64_t)blockIdx.x * (kAlThreads / 32) + (threadIdx.x >> 5); f < T; f += f_stride) {
        const float* fa = align + f * (int64_t)n_align * 3;
        double s[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) s[k] = 0.0;
        auto acc3 = [&](double ax, double ay, double az, double bx, double by, double bz) {
            s[0] += ax * bx; s[1] += ax * by; s[2] += ax * bz;
            s[3] += ay * bx; s[4] += ay * by; s[5] += ay * bz;
            s[6] += az * bx; s[7] += az * by; s[8] += az * bz;
        };
        if (SMEM_B) {
            // every frame starts on a 16-byte boundary: a lane takes four consecutive landmarks (48 bytes) as
            // three 16-byte loads, two such groups in flight -- the pass is a pure stream of the landmark array
            const float4* fv = reinterpret_cast<const float4*>(fa);
#pragma unroll 2
            for (int gI = lane; gI < groups; gI += 32) {
                const float4 p = __ldg(fv + 3 * gI), q = __ldg(fv + 3 * gI + 1), r = __ldg(fv + 3 * gI + 2);
                const double* b = corr_b + gI;
                acc3(p.x, p.y, p.z, b[0], b[groups], b[2 * (size_t)groups]);
                acc3(p.w, q.x, q.y, b[3 * (size_t)groups], b[4 * (size_t)groups], b[5 * (size_t)groups]);
                acc3(q.z, q.w, r.x, b[6 * (size_t)groups], b[7 * (size_t)groups], b[8 * (size_t)groups]);
                acc3(r.y, r.z, r.w, b[9 * (size_t)groups], b[10 * (size_t)groups], b[11 * (size_t)groups]);
            }
        } else {
#pragma unroll 4
            for (int i = lane; i < n_align; i += 32)
                acc3(fa[3 * i], fa[3 * i + 1], fa[3 * i + 2], __ldg(avg_align + 3 * i), __ldg(avg_align + 3 * i + 1),
                     __ldg(avg_align + 3 * i + 2));
        }
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if (lane < 9) {
            double v = s[0];
#pragma unroll
            for (int k = 1; k < 9; ++k) v = (lane == k) ? s[k] : v;
            rot[f * 9 + lane] = v;
        }
    }
}

__global__ void __launch_bounds__(128)
align_quat_kernel(int64_t T, double* __restrict__ rot, double* __restrict__ rot_total, int compose,
                  const int* __restrict__ gate) {
    if (gate && *gate) return;
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= T) return;
    double* out = rot + f * 9;
    double s[9], R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = out[k];
    horn_rotation(s, R);
#pragma unroll
    for (int k = 0; k < 9; ++k) out[k] = R[k];
    if (rot_total) {
        double* tot = rot_total + f * 9;
        if (compose) {
            double P[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) P[k] = tot[k];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    tot[3 * i + j] = R[3 * i] * P[j] + R[3 * i + 1] * P[3 + j] + R[3 * i + 2] * P[6 + j];
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) tot[k] = R[k];
        }
    }
}

// x <- x . R^T (row vector times R transposed == R x) for the frames [f0, f1) and the 32 items
// [item0, item0 + 32) of a [frame][item][3] array with row stride `row_stride`.
// WRITE: store the rotated values (rounded to T) in place; SUMS: part_cols[3 item + c] = column sums of the
// rotated values (of the STORED values when WRITE: the reference sums its float32 array, :110).
constexpr int kApplyWarps = 8;
constexpr int kApplyFrames = 4;

template <typename T, bool WRITE, bool SUMS>
__device__ __forceinline__ void align_apply_warp(T* __restrict__ base, int64_t row_stride, int n_items, int item0,
                                                 int64_t f0, int64_t f1, const double* __restrict__ rot,
                                                 double* __restrict__ part_cols, T (*sm)[96]) {
    const int lane = threadIdx.x & 31;
    const int ncol = 3 * min(32, n_items - item0);
    T* col = base + 3 * (int64_t)item0;
    double s[3] = {0.0, 0.0, 0.0};
    // (issuing the loads of step k + 1 before step k is processed -- a register double buffer -- costs 16
    // registers, drops the occupancy from 4 to 3 CTAs per SM and was slower: 1.26 instead of 1.11 ms at C2)
    for (int64_t f = f0; f < f1; f += kApplyFrames) {
        const int nf = (int)min((int64_t)kApplyFrames, f1 - f);
        T v[kApplyFrames][3];
#pragma unroll
        for (int q = 0; q < kApplyFrames; ++q)
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int c = lane + 32 * u;
                v[q][u] = (q < nf && c < ncol) ? col[(f + q) * row_stride + c] : (T)0;
            }
#pragma unroll
        for (int q = 0; q < kApplyFrames; ++q)
#pragma unroll
            for (int u = 0; u < 3; ++u) sm[q][lane + 32 * u] = v[q][u];
        __syncwarp();
        if (3 * lane < ncol) {
#pragma unroll
            for (int q = 0; q < kApplyFrames; ++q)
                if (q < nf) {
                    const double* R = rot + (f + q) * 9;
                    const double x = (double)sm[q][3 * lane], y = (double)sm[q][3 * lane + 1], z = (double)sm[q][3 * lane + 2];
                    const double rx = x * __ldg(R + 0) + y * __ldg(R + 1) + z * __ldg(R + 2);
                    const double ry = x * __ldg(R + 3) + y * __ldg(R + 4) + z * __ldg(R + 5);
                    const double rz = x * __ldg(R + 6) + y * __ldg(R + 7) + z * __ldg(R + 8);
                    if (WRITE) {
                        sm[q][3 * lane + 0] = (T)rx;
                        sm[q][3 * lane + 1] = (T)ry;
                        sm[q][3 * lane + 2] = (T)rz;
                    } else if (SUMS) {
                        s[0] += rx; s[1] += ry; s[2] += rz;
                    }
                }
        }
        if (WRITE) {
            __syncwarp();
#pragma unroll
            for (int q = 0; q < kApplyFrames; ++q)
                if (q < nf) {
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int c = lane + 32 * u;
                        if (c < ncol) {
                            const T o = sm[q][c];
                            col[(f + q) * row_stride + c] = o;
                            if (SUMS) s[u] += (double)o;
                        }
                    }
                }
            // (a lane's next shared-memory writes go to the very slots it has just read)
        } else {
            __syncwarp();       // every lane has read its triple before the patch is overwritten
        }
    }
    if (SUMS) {
        if (WRITE) {            // lane owns columns lane, lane + 32, lane + 64 of the span
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int c = lane + 32 * u;
                if (c < ncol) part_cols[3 * (int64_t)item0 + c] = s[u];
            }
        } else if (3 * lane < ncol) {   // lane owns item item0 + lane
#pragma unroll
            for (int u = 0; u < 3; ++u) part_cols[3 * (int64_t)(item0 + lane) + u] = s[u];
        }
    }
}

// grid (ceil(groups / 8), n_seg); a warp = one group of 32 items; landmark groups come first.
// NODE_WRITE: the nodes are rotated in place by rot_nodes (the old per-iteration behaviour and the final
// in-place rotation); otherwise they are only read.  align == NULL: nodes only.  part == NULL: no sums.
template <bool NODE_WRITE, bool SUMS>
__global__ void __launch_bounds__(kApplyWarps * 32)
align_apply_kernel(float* __restrict__ align, int n_align, double* __restrict__ nodes, int64_t ld, int N,
                   int64_t T, const double* __restrict__ rot, const double* __restrict__ rot_nodes,
                   int64_t frames_per_seg, double* __restrict__ part, const int* __restrict__ gate) {
    __shared__ __align__(16) double smd[kApplyWarps][kApplyFrames][96];
    if (gate && *gate) return;
    const int w = threadIdx.x >> 5;
    const int gl = align ? (n_align + 31) / 32 : 0;
    const int gn = (N + 31) / 32;
    const int g = blockIdx.x * kApplyWarps + w;
    if (g >= gl + gn) return;
    const int64_t f0 = (int64_t)blockIdx.y * frames_per_seg, f1 = min(T, f0 + frames_per_seg);
    double* prow = SUMS ? part + (int64_t)blockIdx.y * (3 * (int64_t)((align ? n_align : 0) + N)) : nullptr;
    if (g < gl)
        align_apply_warp<float, true, SUMS>(align, 3 * (int64_t)n_align, n_align, g * 32, f0, f1, rot, prow,
                                            reinterpret_cast<float (*)[96]>(smd[w]));
    else
        align_apply_warp<double, NODE_WRITE, SUMS>(nodes, ld, N, (g - gl) * 32, f0, f1, rot_nodes,
                                                   SUMS ? prow + 3 * (int64_t)(align ? n_align : 0) : nullptr, smd[w]);
}

// np.sum(float32 array, axis=0) of the reference (:86): sequential float32 adds down the frames.
// `inout` holds the running float32 totals of all EARLIER frames on entry (zeros for the first frame
// block) and the totals including this block on exit (doubles holding floats): frame shards chain
// their blocks in frame order and reproduce the single sequential sum bit for bit.
// The add chain of a column is inherently serial (50,000 dependent FADDs = 0.1 ms), the memory traffic is
// not -- but register-staged loads never got it flowing: [256 frames][32 columns] tiles with 32 loads per
// thread in flight 0.92 ms at C2 for 1.2 GB, 16 columns 0.83 ms, 64 columns x 64 loads 1.03 ms (one SM does
// not keep that many scalar loads outstanding), and one cp.async.bulk per 176-byte row segment 1.40 ms (the
// TMA unit starts a copy every ~50 cycles, whatever its size).  sum32_tma_kernel therefore describes the
// array to the TMA unit as a 2-D tensor and moves a whole [128 rows][48 columns] tile (24 KB) with ONE
// cp.async.bulk.tensor instruction into a 4-stage ring (completion on mbarriers, no registers held); 48
// consumer threads walk each tile in frame order; 125 CTAs at C2.  The tensor map needs a 16-byte multiple
// as row pitch (3 n_align a multiple of 4); sum32_sequential_kernel is the fallback for any other shape.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int kSqStages = 4, kSqRows = 128, kSqCols = 48;

__device__ __forceinline__ uint32_t sq_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void sq_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sq_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sq_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sq_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sq_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sq_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sq_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(sq_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void sq_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(sq_smem_u32(dst)), "l"(map), "r"(sq_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// block = 32 (producer warp) + 64 (consumers); dynamic smem = kSqStages * kSqRows * kSqCols floats.
// tmap: the landmark array as a 2-D tensor {3 n_align columns, T rows}, box {kSqCols, kSqRows}; rows and
// columns beyond the array are zero-filled by the TMA unit (and never added: the consumers stop at the last
// real row, threads of columns beyond the array idle), and every box completes the full box byte count.
__global__ void __launch_bounds__(96)
sum32_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t T, int64_t ncols, double* __restrict__ inout) {
    extern __shared__ __align__(128) float sq_ring[];
    __shared__ uint64_t full[kSqStages], empty[kSqStages];
    const int tid = threadIdx.x;
    const int64_t c0 = (int64_t)blockIdx.x * kSqCols;
    const int nc = (int)min((int64_t)kSqCols, ncols - c0);
    const int64_t n_tiles = (T + kSqRows - 1) / kSqRows;
    if (tid == 0) {
        for (int s = 0; s < kSqStages; ++s) { sq_mbar_init(&full[s], 1); sq_mbar_init(&empty[s], 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) {
        // ---------------- producer: one lane, one tensor copy per tile ----------------
        if (tid == 0) {
            for (int64_t k = 0; k < n_tiles; ++k) {
                const int slot = (int)(k % kSqStages);
                const uint32_t ph = (uint32_t)((k / kSqStages) & 1);
                sq_mbar_wait(&empty[slot], ph ^ 1);
                sq_mbar_expect_tx(&full[slot], (uint32_t)(kSqRows * kSqCols * sizeof(float)));
                sq_tma_load_2d(sq_ring + (size_t)slot * kSqRows * kSqCols, &tmap, (int)c0, (int)(k * kSqRows), &full[slot]);
            }
        }
        return;
    }
    // ---------------- consumers: thread = column ----------------
    const int cl = tid - 32;
    const bool ok = cl < nc;
    float s = ok ? (float)inout[c0 + cl] : 0.f;
    for (int64_t k = 0; k < n_tiles; ++k) {
        const int slot = (int)(k % kSqStages);
        const uint32_t ph = (uint32_t)((k / kSqStages) & 1);
        const int nr = (int)min((int64_t)kSqRows, T - k * kSqRows);
        sq_mbar_wait(&full[slot], ph);
        if (ok) {
            const float* tile = sq_ring + (size_t)slot * kSqRows * kSqCols + cl;
#pragma unroll 8
            for (int r = 0; r < nr; ++r) s = __fadd_rn(s, tile[(size_t)r * kSqCols]);
        }
        __syncwarp();
        if ((tid & 31) == 0) sq_mbar_arrive(&empty[slot]);
    }
    if (ok) inout[c0 + cl] = (double)s;
}

// fallback for row segments that are not 16-byte aligned: register-staged [512 frames][16 columns] tiles
constexpr int kSeqCols = 16, kSeqFrames = 512, kSeqThreads = 256;
__global__ void __launch_bounds__(kSeqThreads)
sum32_sequential_kernel(const float* __restrict__ a, int64_t T, int64_t ncols, double* __restrict__ inout) {
    __shared__ float tile[kSeqFrames][kSeqCols];
    const int tid = threadIdx.x, cl = tid % kSeqCols, rg = tid / kSeqCols;   // column, row group
    constexpr int RG = kSeqThreads / kSeqCols;           // rows loaded per step by the CTA
    constexpr int PER = kSeqFrames / RG;                 // rows of a tile each thread loads
    const int64_t c = (int64_t)blockIdx.x * kSeqCols + cl;
    const bool ok = c < ncols;
    float s = (rg == 0 && ok) ? (float)inout[c] : 0.f;
    float r[PER];
    auto load = [&](int64_t f0) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int64_t f = f0 + rg + (int64_t)k * RG;
            r[k] = (ok && f < T) ? a[f * ncols + c] : 0.f;
        }
    };
    load(0);
    for (int64_t f0 = 0; f0 < T; f0 += kSeqFrames) {
        __syncthreads();                                   // the previous tile has been summed
#pragma unroll
        for (int k = 0; k < PER; ++k) tile[rg + k * RG][cl] = r[k];
        __syncthreads();
        if (f0 + kSeqFrames < T) load(f0 + kSeqFrames);     // in flight while row group 0 sums this tile
        if (rg == 0) {
            const int nf = (int)min((int64_t)kSeqFrames, T - f0);
#pragma unroll 8
            for (int f = 0; f < nf; ++f) s = __fadd_rn(s, tile[f][cl]);
        }
    }
    if (rg == 0 && ok) inout[c] = (double)s;
}

// new averages and the two RMSDs' squared sums in ONE launch (a frame-sharded caller has just all-reduced
// `sums`): avg <- sums / T_total in place; res2[0] = sum over landmark columns of (old - new)^2,
// res2[1] = the same over the node columns.  Single CTA, fixed-order tree -> deterministic.
// gate (may be NULL): when already set the kernel does nothing; otherwise res2[2] counts the iterations
// run and the flag is raised once sqrt(res2[0] / n_align) <= threshold (:96).
__global__ void __launch_bounds__(1024)
align_update_kernel(const double* __restrict__ sums, double T_total, int64_t na3, int64_t nn3,
                    double* __restrict__ avg, double* __restrict__ res2, double threshold, int* __restrict__ gate) {
    __shared__ double red[2][32];
    if (gate && *gate) return;
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
        if (threadIdx.x == 0) {
            res2[0] = ra; res2[1] = rn;
            if (gate) {
                res2[2] += 1.0;
                if (sqrt(ra / (double)(na3 / 3)) <= threshold) *gate = 1;
            }
        }
    }
}

// initial averages from the initial sums (:86-87): landmarks float32(sum) / float32(T) -- the reference's
// array is float32 --, nodes sum / T; avg [na3 + nn3]
__global__ void align_init_avg_kernel(const double* __restrict__ sums, double T_total, int64_t na3, int64_t nn3,
                                      double* __restrict__ avg) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= na3 + nn3) return;
    avg[c] = c < na3 ? (double)__fdiv_rn((float)sums[c], (float)T_total) : sums[c] / T_total;
}

}  // namespace

int launch_align_update(wisp_ctx* ctx, const double* d_sums, int64_t T_total, int n_align, int N, double* d_avg,
                        double* d_res2, cudaStream_t st, double threshold, int* d_gate) {
    align_update_kernel<<<1, 1024, 0, st>>>(d_sums, (double)T_total, 3 * (int64_t)n_align, 3 * (int64_t)N, d_avg, d_res2,
                                            threshold, d_gate);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

// ---- building blocks (also the device-level C ABI: wisp_dev_align_sums / _iter / _nodes / _rotate) ----
// d_sums [na3 + nn3]: first the landmark column sums, then the node column sums of this frame block.
// first == true: landmark sums as the reference's sequential float32 sums (initial average, :86):
// d_sums[0, na3) must hold the running totals of the earlier frame blocks on entry (zeros for the
// first block) and holds the totals including this block on exit.
// d_nodes == NULL: the caller already has the node column sums (K1 produces them, launch_com_plan's
// d_colsum) and d_sums[na3, ..) is left alone.
int launch_align_sums(wisp_ctx* ctx, const float* d_align, int n_align, const double* d_nodes, int64_t ld,
                      int N, int64_t T, bool first, double* d_sums, cudaStream_t st) {
    const int64_t na3 = 3 * (int64_t)n_align, nn3 = 3 * (int64_t)N;
    if (first) {
        static EncodeTiledFn encode = nullptr;
        static bool encode_tried = false;
        if (!encode_tried) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
                qres == cudaDriverEntryPointSuccess) encode = (EncodeTiledFn)fn;
            (void)cudaGetLastError();
            encode_tried = true;
        }
        CUtensorMap tmap;
        bool use_tma = encode && (na3 & 3) == 0 && (reinterpret_cast<size_t>(d_align) & 15) == 0 && T < (1ll << 31) &&
                       getenv("WISP_K7_SUM_SEQ") == nullptr;
        if (use_tma) {
            const cuuint64_t dims[2] = {(cuuint64_t)na3, (cuuint64_t)T};
            const cuuint64_t strides[1] = {(cuuint64_t)na3 * 4};
            const cuuint32_t box[2] = {(cuuint32_t)kSqCols, (cuuint32_t)kSqRows};
            const cuuint32_t estr[2] = {1, 1};
            use_tma = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(d_align), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        if (use_tma) {
            const int smem = kSqStages * kSqRows * kSqCols * (int)sizeof(float);
            WISP_CUDA(cudaFuncSetAttribute(sum32_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            sum32_tma_kernel<<<(int)((na3 + kSqCols - 1) / kSqCols), 96, smem, st>>>(tmap, T, na3, d_sums);
        } else {
            sum32_sequential_kernel<<<(int)((na3 + kSeqCols - 1) / kSeqCols), kSeqThreads, 0, st>>>(d_align, T, na3, d_sums);
        }
        ctx->launches++;
        WISP_CUDA(cudaGetLastError());
    } else {
        WISP_TRY(launch_colsum_f32(ctx, d_align, T, na3, na3, d_sums, st));
    }
    if (d_nodes) WISP_TRY(launch_colsum(ctx, d_nodes, T, ld, nn3, d_sums + na3, st));
    return 0;
}

namespace {

// Segments of frames so that the apply grid is exactly ONE wave of resident CTAs: every CTA moves the
// same number of bytes per frame (a landmark warp reads and writes 384 bytes, a node warp reads 768), so a
// grid of 1.5 waves -- what a fixed "six CTAs per SM" gave at C2: 880 CTAs on 592 slots -- leaves half the
// machine idle for the second half of the kernel (1.33 ms instead of 1.0 ms per iteration).
template <typename Kernel>
int apply_segments(wisp_ctx* ctx, Kernel kernel, int groups, int64_t T, int64_t* per_out) {
    const int gx = (groups + kApplyWarps - 1) / kApplyWarps;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kApplyWarps * 32, 0) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 4;
    }
    int n_seg = (int)std::max<int64_t>(1, std::min<int64_t>((T + 15) / 16, (int64_t)ctx->sm_count * per_sm / gx));
    const int64_t per = round_up((T + n_seg - 1) / n_seg, kApplyFrames);
    *per_out = per;
    return (int)((T + per - 1) / per);
}

int launch_solve(wisp_ctx* ctx, const float* d_align, int n_align, const double* d_avg_align, int64_t T,
                 double* d_rot, double* d_rot_total, bool compose, cudaStream_t st, const int* d_gate = nullptr) {
    const int64_t ctas_all = (T + kAlThreads / 32 - 1) / (kAlThreads / 32);
    const size_t b_bytes = (size_t)n_align * 24;
    if ((n_align & 3) == 0 && (reinterpret_cast<size_t>(d_align) & 15) == 0 && b_bytes <= 96 * 1024) {
        WISP_CUDA(cudaFuncSetAttribute(align_corr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b_bytes));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, align_corr_kernel<true>, kAlThreads, b_bytes) != cudaSuccess || per_sm < 1) {
            (void)cudaGetLastError();
            per_sm = 1;
        }
        const unsigned grid = (unsigned)std::min<int64_t>(ctas_all, (int64_t)ctx->sm_count * per_sm);
        align_corr_kernel<true><<<grid, kAlThreads, b_bytes, st>>>(d_align, n_align, d_avg_align, T, d_rot, d_gate);
    } else {
        align_corr_kernel<false><<<(unsigned)std::min<int64_t>(ctas_all, (int64_t)ctx->sm_count * 16), kAlThreads, 0, st>>>(
            d_align, n_align, d_avg_align, T, d_rot, d_gate);
    }
    align_quat_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(T, d_rot, d_rot_total, compose ? 1 : 0, d_gate);
    ctx->launches += 2;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

// One alignment iteration over this frame block with the node rotation DEFERRED: the landmarks are
// rotated onto d_avg_align in place, d_rot_total [T][9] <- R (first) or R d_rot_total, and d_sums
// [3 n_align + 3 N] (may be NULL) receives the column sums of the rotated landmarks, then of
// d_rot_total applied to the (unchanged) nodes.
// d_gate (may be NULL): device flag; when set, every kernel of the iteration is a no-op (a caller that
// launches iteration k + 1 before it has seen the residual of iteration k, see FramePipeline.align).
int launch_align_iter(wisp_ctx* ctx, float* d_align, int n_align, const double* d_avg_align, const double* d_nodes,
                      int64_t ld, int N, int64_t T, double* d_rot_total, bool first, cudaStream_t st, double* d_sums,
                      const int* d_gate) {
    void *pr = nullptr, *pp = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_ALIGN_ROT, (size_t)T * 9 * 8, &pr));
    double* d_rot = (double*)pr;
    WISP_TRY(launch_solve(ctx, d_align, n_align, d_avg_align, T, d_rot, d_rot_total, !first, st, d_gate));
    const int groups = (n_align + 31) / 32 + (N + 31) / 32;
    int64_t per = 0;
    const int n_seg = d_sums ? apply_segments(ctx, align_apply_kernel<false, true>, groups, T, &per)
                             : apply_segments(ctx, align_apply_kernel<false, false>, (n_align + 31) / 32, T, &per);
    const int64_t ncols = 3 * (int64_t)(n_align + N);
    const dim3 grid((groups + kApplyWarps - 1) / kApplyWarps, n_seg);
    if (d_sums) {
        WISP_TRY(wisp_scratch(ctx, SCR_ALIGN_PART, (size_t)n_seg * ncols * 8, &pp));
        align_apply_kernel<false, true><<<grid, kApplyWarps * 32, 0, st>>>(
            d_align, n_align, const_cast<double*>(d_nodes), ld, N, T, d_rot, d_rot_total, per, (double*)pp, d_gate);
    } else {
        // nothing to sum: only the landmarks move
        const dim3 grid_l(((n_align + 31) / 32 + kApplyWarps - 1) / kApplyWarps, n_seg);
        align_apply_kernel<false, false><<<grid_l, kApplyWarps * 32, 0, st>>>(
            d_align, n_align, nullptr, ld, 0, T, d_rot, d_rot_total, per, nullptr, d_gate);
    }
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    if (d_sums) WISP_TRY(launch_colsum_finish(ctx, (const double*)pp, n_seg, ncols, ncols, d_sums, st));
    return 0;
}

// the deferred node rotation, in place: x[f] <- d_rot_total[f] x[f]
int launch_align_nodes(wisp_ctx* ctx, double* d_nodes, int64_t ld, int N, int64_t T, const double* d_rot_total,
                       cudaStream_t st) {
    ctx->i8_stats_x = nullptr;               // the node trajectory is rotated in place
    const int groups = (N + 31) / 32;
    int64_t per = 0;
    const int n_seg = apply_segments(ctx, align_apply_kernel<true, false>, groups, T, &per);
    align_apply_kernel<true, false><<<dim3((groups + kApplyWarps - 1) / kApplyWarps, n_seg), kApplyWarps * 32, 0, st>>>(
        nullptr, 0, d_nodes, ld, N, T, nullptr, d_rot_total, per, nullptr, nullptr);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

// one alignment pass over this frame block with BOTH arrays rotated in place (the reference's literal
// data flow, :106-107): rotate every frame onto d_avg_align; d_sums [3 n_align + 3 N] (may be NULL) receives
// the column sums of the rotated landmarks, then of the rotated nodes
int launch_align_rotate(wisp_ctx* ctx, float* d_align, int n_align, const double* d_avg_align, double* d_nodes,
                        int64_t ld, int N, int64_t T, cudaStream_t st, double* d_sums) {
    ctx->i8_stats_x = nullptr;               // the node trajectory is rotated in place
    void *pr = nullptr, *pp = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_ALIGN_ROT, (size_t)T * 9 * 8, &pr));
    double* d_rot = (double*)pr;
    WISP_TRY(launch_solve(ctx, d_align, n_align, d_avg_align, T, d_rot, nullptr, false, st));
    const int groups = (n_align + 31) / 32 + (N + 31) / 32;
    int64_t per = 0;
    const int n_seg = d_sums ? apply_segments(ctx, align_apply_kernel<true, true>, groups, T, &per)
                             : apply_segments(ctx, align_apply_kernel<true, false>, groups, T, &per);
    const int64_t ncols = 3 * (int64_t)(n_align + N);
    const dim3 grid((groups + kApplyWarps - 1) / kApplyWarps, n_seg);
    if (d_sums) {
        WISP_TRY(wisp_scratch(ctx, SCR_ALIGN_PART, (size_t)n_seg * ncols * 8, &pp));
        align_apply_kernel<true, true><<<grid, kApplyWarps * 32, 0, st>>>(d_align, n_align, d_nodes, ld, N, T, d_rot,
                                                                        d_rot, per, (double*)pp, nullptr);
    } else {
        align_apply_kernel<true, false><<<grid, kApplyWarps * 32, 0, st>>>(d_align, n_align, d_nodes, ld, N, T, d_rot,
                                                                         d_rot, per, nullptr, nullptr);
    }
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    if (d_sums) WISP_TRY(launch_colsum_finish(ctx, (const double*)pp, n_seg, ncols, ncols, d_sums, st));
    return 0;
}

// d_align: float32 [T][n_align][3] (in/out); d_nodes: pitched float64 trajectory, T = the frames of THIS
// shard, T_total = all frames.  `reduce` (may be empty) sums a host vector over the frame shards in shard
// order (or, for the float32 chain, runs the shards' steps one after the other in frame order); every shard
// then takes the same decisions from the same numbers.  d_work: [2 * round_up(na3,32) + round_up(nn3,32)].
// d_node_colsum (may be NULL): the node column sums of this shard, already on the device (from K1).
// d_rot_out == NULL: the nodes are rotated in place before returning.  Otherwise they are left as they
// came and *d_rot_out = the per-frame total rotations [T][9] (NULL when no iteration ran) for the caller
// to apply (launch_center's d_rot, launch_align_nodes).
// Returns iterations run; log gets (iteration, residual, node RMSD) triples.
int run_iterative_alignment(wisp_ctx* ctx, float* d_align, int n_align, double* d_nodes, int64_t ld,
                            int N, int64_t T, double threshold, int max_iter, double* d_work,
                            double* h_avg_nodes, double* log, int* n_iter_out, cudaStream_t st,
                            int64_t T_total, const AlignReduce& reduce, const double* d_node_colsum,
                            const double** d_rot_out) {
    const int64_t na3 = 3 * (int64_t)n_align, nn3 = 3 * (int64_t)N;
    if (T_total <= 0) T_total = T;
    double* d_avgA = d_work;                 // [na3]
    double* d_sums = d_avgA + round_up(na3, 32);   // [na3 + nn3]
    void* prt = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_ALIGN_ROTTOT, (size_t)T * 9 * 8, &prt));
    double* d_rot_total = (double*)prt;
    if (!reduce) {
        // ---- one shard: the whole loop stays on the device ----
        // averages [landmarks || nodes], the two squared residual sums and the convergence flag live in
        // HBM; per iteration 24 bytes come back for the log, and iteration k + 1 is launched BEFORE the
        // residual of iteration k has been read (its kernels are no-ops once the flag is up), so the GPU
        // never waits for the host between iterations.
        void* pav = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_ALIGN_AVG, (size_t)(na3 + nn3 + 16) * 8, &pav));
        double* d_avg = (double*)pav;
        double* d_res3 = d_avg + na3 + nn3;
        int* d_gate = (int*)(d_res3 + 4);
        WISP_CUDA(cudaMemsetAsync(d_sums, 0, na3 * 8, st));
        WISP_CUDA(cudaMemsetAsync(d_res3, 0, 8 * 8, st));
        WISP_TRY(launch_align_sums(ctx, d_align, n_align, d_node_colsum ? nullptr : d_nodes, ld, N, T, true, d_sums, st));
        if (d_node_colsum)
            WISP_CUDA(cudaMemcpyAsync(d_sums + na3, d_node_colsum, nn3 * 8, cudaMemcpyDeviceToDevice, st));
        align_init_avg_kernel<<<(unsigned)((na3 + nn3 + 255) / 256), 256, 0, st>>>(d_sums, (double)T_total, na3, nn3, d_avg);
        ctx->launches++;
        WISP_CUDA(cudaGetLastError());
        if (ctx->align_host_cap < max_iter + 1) {
            if (ctx->align_host) cudaFreeHost(ctx->align_host);
            ctx->align_host = nullptr;
            WISP_CUDA(cudaHostAlloc((void**)&ctx->align_host, (size_t)(max_iter + 1) * 3 * 8, cudaHostAllocDefault));
            ctx->align_host_cap = max_iter + 1;
        }
        double* h_res = ctx->align_host;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        WISP_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
        WISP_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        auto launch = [&](int k) -> int {
            WISP_TRY(launch_align_iter(ctx, d_align, n_align, d_avg, d_nodes, ld, N, T, d_rot_total, k == 0, st, d_sums, d_gate));
            WISP_TRY(launch_align_update(ctx, d_sums, T_total, n_align, N, d_avg, d_res3, st, threshold, d_gate));
            WISP_CUDA(cudaMemcpyAsync(h_res + 3 * k, d_res3, 24, cudaMemcpyDeviceToHost, st));
            WISP_CUDA(cudaEventRecord(ev[k & 1], st));
            return 0;
        };
        int iteration = 0, rc = 0;
        if (max_iter > 0) rc = launch(0);
        while (!rc && iteration < max_iter) {
            if (iteration + 1 < max_iter) rc = launch(iteration + 1);      // speculative
            if (rc) break;
            if (cudaEventSynchronize(ev[iteration & 1]) != cudaSuccess) { wisp_set_error("alignment: device error"); rc = 1; break; }
            const double residual = std::sqrt(h_res[3 * iteration] / n_align);
            const double analysis = std::sqrt(h_res[3 * iteration + 1] / N);
            ++iteration;
            if (log) { log[3 * (iteration - 1)] = iteration; log[3 * (iteration - 1) + 1] = residual; log[3 * (iteration - 1) + 2] = analysis; }
            if (!(residual > threshold)) break;
        }
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
        if (rc) return rc;
        if (d_rot_out) *d_rot_out = iteration > 0 ? d_rot_total : nullptr;
        else if (iteration > 0) WISP_TRY(launch_align_nodes(ctx, d_nodes, ld, N, T, d_rot_total, st));
        if (h_avg_nodes) WISP_CUDA(cudaMemcpyAsync(h_avg_nodes, d_avg + na3, nn3 * 8, cudaMemcpyDeviceToHost, st));
        WISP_CUDA(cudaStreamSynchronize(st));
        if (n_iter_out) *n_iter_out = iteration;
        return 0;
    }
    std::vector<double> avgA(na3), avgN(nn3), sums(na3 + nn3);
    // initial landmark totals: float32, chained over the shards in frame order (`reduce` with
    // f32_chain = true hands this shard the totals of the earlier shards, runs `step`, and returns the
    // grand totals); node sums: plain float64 sums combined afterwards
    auto step = [&](double* h_inout) -> int {
        WISP_CUDA(cudaMemcpyAsync(d_sums, h_inout, na3 * 8, cudaMemcpyHostToDevice, st));
        WISP_TRY(launch_align_sums(ctx, d_align, n_align, d_node_colsum ? nullptr : d_nodes, ld, N, T, true, d_sums, st));
        if (d_node_colsum)
            WISP_CUDA(cudaMemcpyAsync(d_sums + na3, d_node_colsum, nn3 * 8, cudaMemcpyDeviceToDevice, st));
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
        WISP_TRY(launch_align_iter(ctx, d_align, n_align, d_avgA, d_nodes, ld, N, T, d_rot_total, iteration == 0, st, d_sums, nullptr));
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
    if (d_rot_out) *d_rot_out = iteration > 0 ? d_rot_total : nullptr;
    else if (iteration > 0) WISP_TRY(launch_align_nodes(ctx, d_nodes, ld, N, T, d_rot_total, st));
    WISP_CUDA(cudaStreamSynchronize(st));
    if (h_avg_nodes) memcpy(h_avg_nodes, avgN.data(), nn3 * 8);
    if (n_iter_out) *n_iter_out = iteration;
    return 0;
}
