// K5 -- blocked Floyd-Warshall all-pairs shortest paths with predecessors, sm_100a.
//
// The reference has no APSP: get_paths calls networkx Dijkstra once per (source, sink) pair
// and twice more per pair on EVERY cutoff pass (network_analysis_functions.py:30,71-72).
// One min-plus APSP replaces all of them: rows s and t of (dist, pred) are what those calls
// return.  Graph semantics follow networkx 1.x Graph(data=W): entry != 0 <=> edge; the
// diagonal is never an edge.
//
// Round kb of the blocked algorithm (tile edge B = 128 in FP32, 64 in FP64):
//   P1   pivot tile (kb,kb): B dependent k-steps in one CTA, own elements in registers, one
//        barrier per step;
//   P2   pivot-row tiles (kb,*) and pivot-column tiles (*,kb).  WARP-AUTONOMOUS: a warp owns 8
//        whole columns (rows) of its tile -- lane (q, v) holds B/4 consecutive elements of vector
//        v in registers -- so the value of row k of its own tile is one register shuffle away and
//        the pivot tile sits read-only in shared memory: no barrier in the k loop at all;
//   P3   every other tile: min-plus product of its pivot-column tile and pivot-row tile,
//        register-tiled, accumulated into fresh minima m and merged into the tile at the end.
//        N^3 (1 - 2/nb) of the work.
// LOOK-AHEAD.  Only the tiles of row kb+1 and column kb+1 are needed by the next round's P1/P2,
// so P3 is split: a small "lead" launch for those tiles, then P1/P2 of round kb+1 run on a
// second, high-priority stream WHILE the bulk of round kb's P3 is still running on the main
// stream.  The serial pivot phases disappear from the critical path once N is more than a few
// tiles (round 1 spent 62 % of an N = 2,000 run and 25 % of an N = 10,000 run in them).
// PREDECESSORS WITHOUT A SINGLE EXTRA INSTRUCTION IN THE RELAXATION LOOP.  The min-plus loop is
// dispatch bound -- tools/peaks_minplus.cu: FADD2+FMNMX3, scalar FADD+FMNMX and the DPX VIADDMNMX
// all top out at 16-18 T relaxations/s because every half-rate or packed instruction holds the
// dispatch port two cycles -- so carrying a predecessor select (round 1: x3.7 slower) is the
// wrong trade.  Instead every kernel records, per element, the ROUND in which it last improved
// (one compare when the tile is merged; uint16 matrix).  After the last round
//   mid[i][j]  = argmin over the <= B vertices k of that round of dist[i][k] + dist[k][j]
//                (N^2 * B work = B/N of the main loop; the improving k lies in that round, and
//                final distances reproduce it up to rounding),
//   pred[i][j] = follow mid along column j until the remaining hop is a direct edge
// (the classic midpoint form of Floyd-Warshall path reconstruction).  On generic weights the
// shortest path is unique and the result equals the Dijkstra predecessor tree.
#include "common.cuh"
#include <limits>
#include <cstdlib>
#include <algorithm>

namespace {

constexpr uint16_t kNoRound = 0xFFFFu;

template <typename T> __device__ __forceinline__ T inf_v();
template <> __device__ __forceinline__ float inf_v<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double inf_v<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ float fw_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double fw_min(double a, double b) { return b < a ? b : a; }     // fmin() lowers badly

template <typename T>
__global__ void fw_init_kernel(const double* __restrict__ w, int N, int64_t n_pad, int64_t row0, int64_t rows,
                               T* __restrict__ dist, uint16_t* __restrict__ rounds) {
    // rows [row0, row0 + rows) of the padded matrix; w holds the same rows of the N x N cost matrix
    const int64_t total = rows * n_pad;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t rl = e / n_pad, c = e - rl * n_pad, r = row0 + rl;
        T d = inf_v<T>();
        if (r == c) d = (T)0;
        else if (r < N && c < N) { const double v = w[rl * N + c]; if (v != 0.0) d = (T)v; }
        dist[e] = d;
        if (rounds) rounds[e] = kNoRound;
    }
}

// ---------------------------------------------------------------------------------------
// P1: pivot tile.  (B/4)^2 threads, 4 x 4 elements each (rows 4ty+a, columns tx + (B/4) b), own
// elements in registers; shared memory is only the exchange medium for row k / column k, written
// when an element improves.  Row k and column k are fixed points of step k (d[k][k] == 0), so one
// barrier per step is enough.  `panel` = the pivot ROW PANEL (B rows x ld): the pivot tile sits at
// column block kb.
// ---------------------------------------------------------------------------------------
template <typename T, int B>
__global__ void __launch_bounds__((B / 4) * (B / 4))
fw_p1_kernel(T* __restrict__ panel, int64_t ld, int kb, uint16_t* __restrict__ rounds_panel) {
    constexpr int TX = B / 4;
    extern __shared__ __align__(16) unsigned char fw_smem[];
    T* d = reinterpret_cast<T*>(fw_smem);     // [B][B+1]
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int64_t base = (int64_t)kb * B;
    T own[4][4], init[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            own[a][b] = init[a][b] = panel[base + (4 * ty + a) * ld + tx + TX * b];
            d[(4 * ty + a) * (B + 1) + tx + TX * b] = own[a][b];
        }
    __syncthreads();
    for (int k = 0; k < B; ++k) {
        T rk[4], ck[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) rk[b] = d[k * (B + 1) + tx + TX * b];
#pragma unroll
        for (int a = 0; a < 4; ++a) ck[a] = d[(4 * ty + a) * (B + 1) + k];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const T cand = ck[a] + rk[b];
                if (cand < own[a][b]) { own[a][b] = cand; d[(4 * ty + a) * (B + 1) + tx + TX * b] = cand; }
            }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t o = base + (4 * ty + a) * ld + tx + TX * b;
            if (own[a][b] < init[a][b]) {
                panel[o] = own[a][b];
                if (rounds_panel) rounds_panel[o] = (uint16_t)kb;
            }
        }
}

// ---------------------------------------------------------------------------------------
// P2: warp-autonomous pivot-row / pivot-column tiles.
//   row tile (kb, other):  own[i][j] <- min(own[i][j], piv[i][k] + own[k][j]),  vector = column j
//   col tile (other, kb):  own[i][j] <- min(own[i][j], own[i][k] + piv[k][j]),  vector = row i
// In both cases, with s the position along the vector:  o_v[s] <- min(o_v[s], M[k][s] + o_v[k]),
// M[k][s] = piv[s][k] (row tiles: pivot transposed) or piv[k][s] (column tiles).  Lane (q, v) of a
// warp holds o_v[q*E .. q*E+E), E = B/4, so o_v[k] lives in lane (k / E, v), register k % E: one
// shuffle.  CTA = 4 warps = 32 vectors; B/32 CTAs per tile; grid (tiles * B/32, 2): y = 0 row
// tiles (inside the row panel), y = 1 column tiles (tile-rows of the LOCAL row block, `skip` = the
// local tile-row that holds the pivot rows themselves).
// ---------------------------------------------------------------------------------------
template <typename T, int B>
__global__ void __launch_bounds__(128)
fw_p2_kernel(T* __restrict__ local, int64_t ld, T* __restrict__ panel, int64_t ldp, int kb,
             int n_col_tiles, int n_local_tiles, int skip, int skip2, uint16_t* __restrict__ rounds_local,
             uint16_t* __restrict__ rounds_panel, int only_tile) {
    constexpr int E = B / 4;                       // elements per lane
    constexpr int PAD = 16 / sizeof(T);            // keeps rows 16-byte aligned
    constexpr int LDM = B + PAD;
    constexpr int SUB = B / 32;                    // CTAs per tile
    const bool is_row = blockIdx.y == 0;
    const int other = only_tile >= 0 ? only_tile : blockIdx.x / SUB;
    const int sub = blockIdx.x % SUB;
    if (only_tile >= 0 && (int)(blockIdx.x / SUB) != 0) return;
    if (is_row) { if (other >= n_col_tiles || other == kb) return; }
    else        { if (other >= n_local_tiles || other == skip || other == skip2) return; }
    extern __shared__ __align__(16) unsigned char fw_smem[];
    T* M = reinterpret_cast<T*>(fw_smem);          // [B][LDM]
    const T* piv = panel + (int64_t)kb * B;
    {   // pivot tile -> shared memory with 16-byte loads, a batch of a thread's loads in flight at once
        constexpr int V = 16 / sizeof(T), PER = B * B / V / 128, BATCH = 16;
#pragma unroll 1
        for (int u0 = 0; u0 < PER; u0 += BATCH) {
            int4 buf[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const int e = (threadIdx.x + 128 * (u0 + u)) * V, r = e / B, c = e % B;
                buf[u] = *reinterpret_cast<const int4*>(piv + (int64_t)r * ldp + c);
            }
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const int e = (threadIdx.x + 128 * (u0 + u)) * V, r = e / B, c = e % B;
                if (is_row) {
                    const T* vals = reinterpret_cast<const T*>(&buf[u]);
#pragma unroll
                    for (int w = 0; w < V; ++w) M[(c + w) * LDM + r] = vals[w];
                } else {
                    *reinterpret_cast<int4*>(M + r * LDM + c) = buf[u];
                }
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane >> 3, vv = lane & 7;
    const int v = sub * 32 + warp * 8 + vv;        // vector index inside the tile
    T* own = is_row ? panel + (int64_t)other * B : local + (int64_t)other * B * ld + (int64_t)kb * B;
    uint16_t* rnd = is_row ? (rounds_panel ? rounds_panel + (int64_t)other * B : nullptr)
                           : (rounds_local ? rounds_local + (int64_t)other * B * ld + (int64_t)kb * B : nullptr);
    const int64_t ldo = is_row ? ldp : ld;
    // element s of vector v: row tile -> own[s][v]; column tile -> own[v][s]
    const int64_t e0 = is_row ? (int64_t)(q * E) * ldo + v : (int64_t)v * ldo + q * E;
    const int64_t es = is_row ? ldo : 1;
    T o[E], init[E];
#pragma unroll
    for (int r = 0; r < E; ++r) o[r] = init[r] = own[e0 + r * es];
    __syncthreads();
    for (int kq = 0; kq < 4; ++kq) {
        const int src = kq * 8 + vv;
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const T xk = __shfl_sync(0xffffffffu, o[r], src);      // o_v[k], k = kq*E + r
            const T* mrow = M + (kq * E + r) * LDM + q * E;
#pragma unroll
            for (int r2 = 0; r2 < E; ++r2) o[r2] = fw_min(o[r2], mrow[r2] + xk);
        }
    }
#pragma unroll
    for (int r = 0; r < E; ++r)
        if (o[r] < init[r]) {
            own[e0 + r * es] = o[r];
            if (rnd) rnd[e0 + r * es] = (uint16_t)kb;
        }
}

// ---------------------------------------------------------------------------------------
// P3 tile selection shared by both precisions.  part 0: every tile outside the pivot row / column;
// part 1 ("lead"): only the tiles of column `next_col` and of local tile-row `next_row`;
// part 2 ("rest"): every tile outside those (and outside the pivot row / column).
// Lead grids are 1-D: [0, nl) = (bi, next_col), [nl, nl + nb) = (next_row, bj).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool fw_p3_tile(int part, int nb, int nl, int kb, int skip, int next_col, int next_row,
                                           int& bi, int& bj) {
    if (part == 1) {
        const int x = blockIdx.x;
        if (x < nl) { if (next_col < 0) return false; bi = x; bj = next_col; }
        else { if (next_row < 0) return false; bi = next_row; bj = x - nl; if (bj == next_col) return false; }
    } else {
        bj = blockIdx.x; bi = blockIdx.y;
        if (part == 2 && (bj == next_col || bi == next_row)) return false;
    }
    return !(bi == skip || bj == kb || bi >= nl || bj >= nb);
}

// ---- FP32 P3: 128 x 128 tile, 8 x 8 outputs per thread, 32-deep k chunks through a 3-stage cp.async
// ring; candidates from packed FADD2, folded with the 3-input minimum FMNMX3 (m = min(m, c1+r1, c2+r2)):
// one dispatch cycle pair per relaxation, which is the machine's limit for a 32-bit min-plus.
constexpr int kT = 128;
constexpr int kC = 32;                 // k-depth per stage
constexpr int kColLd = kC + 4;         // padded row of the column-panel stage (floats)
constexpr int kP3Stages = 3;
constexpr int kP3StageFloats = kT * kColLd + kC * kT;

typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t fw_add2(float c, f32x2_t r) {
    f32x2_t cc, o;
    asm("mov.b64 %0, {%1,%1};" : "=l"(cc) : "f"(c));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(o) : "l"(cc), "l"(r));
    return o;
}
__device__ __forceinline__ float fw_min3(float a, float b, float c) {
    float o;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(o) : "f"(a), "f"(b), "f"(c));
    return o;
}
__device__ __forceinline__ f32x2_t fw_pack(float lo, float hi) {
    f32x2_t o;
    asm("mov.b64 %0, {%1,%2};" : "=l"(o) : "f"(lo), "f"(hi));
    return o;
}
__device__ __forceinline__ void fw_unpack(f32x2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

template <bool TRACK>     // TRACK: record the round of every improvement (predecessor reconstruction)
__global__ void __launch_bounds__(256, 2)
fw32_p3_kernel(float* __restrict__ dist, int64_t ld, const float* __restrict__ panel, int64_t ldp,
               int kb, int skip, int nb, int nl, int part, int next_col, int next_row,
               uint16_t* __restrict__ rounds) {
    // dist = LOCAL row block; tile (bi, bj) <- min-plus of local column tile (bi, kb) and the
    // pivot row panel's tile (., bj); local tile-row `skip` (the pivot rows) is left alone
    int bi, bj;
    if (!fw_p3_tile(part, nb, nl, kb, skip, next_col, next_row, bi, bj)) return;
    extern __shared__ __align__(16) unsigned char fw_smem[];
    float* stages = reinterpret_cast<float*>(fw_smem);
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;            // cols {4tx.., 64+4tx..}, rows {4ty.., 64+4ty..}
    const int64_t obase = (int64_t)bi * kT * ld + (int64_t)bj * kT;
    const float* cg = dist + (int64_t)bi * kT * ld + (int64_t)kb * kT;    // column tile (bi, kb): [i][k]
    const float* rg = panel + (int64_t)bj * kT;                            // row tile (kb, bj):   [k][j]

    auto issue = [&](int chunk) {
        float* cs = stages + (size_t)(chunk % kP3Stages) * kP3StageFloats;     // [128][kColLd]
        float* rs = cs + kT * kColLd;                                            // [kC][128]
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + 256 * q;              // 1024 16-byte granules per panel
            const int i = e >> 3, g = e & 7;          // column panel: row i, granule g (4 floats of k)
            cp_async16(cs + i * kColLd + 4 * g, cg + i * ld + chunk * kC + 4 * g);
            const int k = e >> 5, h = e & 31;         // row panel: row k, granule h (4 floats of j)
            cp_async16(rs + k * kT + 4 * h, rg + (int64_t)(chunk * kC + k) * ldp + 4 * h);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    issue(0);
    issue(1);
    // TRACK: fresh minima, merged into the tile at the end (the comparison tells which entries improved);
    // distance only: the tile itself is the accumulator, loaded under the first cp.async stages
    float m[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        if (TRACK) {
#pragma unroll
            for (int b = 0; b < 8; ++b) m[a][b] = __int_as_float(0x7f800000);
        } else {
            const int row = (a < 4) ? 4 * ty + a : 64 + 4 * ty + (a - 4);
            const float4 v0 = *reinterpret_cast<const float4*>(dist + obase + row * ld + 4 * tx);
            const float4 v1 = *reinterpret_cast<const float4*>(dist + obase + row * ld + 64 + 4 * tx);
            m[a][0] = v0.x; m[a][1] = v0.y; m[a][2] = v0.z; m[a][3] = v0.w;
            m[a][4] = v1.x; m[a][5] = v1.y; m[a][6] = v1.z; m[a][7] = v1.w;
        }
    }
    constexpr int kChunks = kT / kC;
    static_assert(kChunks >= 3 && 2 * kP3StageFloats >= kT * kT, "the tile prefetch reuses two idle stages");
    // the tile itself is not needed before the merge: it is prefetched (each thread its own 16 granules,
    // laid out [granule][thread] -> conflict-free) into the two stages that are idle during the last chunk
    float4* dpre = reinterpret_cast<float4*>(stages + (size_t)((kChunks - 1 + 1) % kP3Stages) * kP3StageFloats);
    static_assert((kChunks - 1) % kP3Stages == 0, "stages 1 and 2 must be the idle, adjacent ones");
    for (int chunk = 0; chunk < kChunks; ++chunk) {
        if (chunk + 1 < kChunks) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                               // chunk landed for everyone; chunk-1 fully consumed
        if (chunk + 2 < kChunks) issue(chunk + 2);
        if (TRACK && chunk == kChunks - 1) {
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int row = (a < 4) ? 4 * ty + a : 64 + 4 * ty + (a - 4);
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    cp_async16(dpre + (2 * a + h) * 256 + tid, dist + obase + row * ld + 64 * h + 4 * tx);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const float* cs = stages + (size_t)(chunk % kP3Stages) * kP3StageFloats;
        const float* rs = cs + kT * kColLd;
#pragma unroll 2
        for (int k = 0; k < kC; k += 2) {
            float c1[8], c2[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int row = (a < 4) ? 4 * ty + a : 64 + 4 * ty + (a - 4);
                const float2 cc = *reinterpret_cast<const float2*>(cs + row * kColLd + k);
                c1[a] = cc.x; c2[a] = cc.y;
            }
            const float4 r1a = *reinterpret_cast<const float4*>(rs + k * kT + 4 * tx);
            const float4 r1b = *reinterpret_cast<const float4*>(rs + k * kT + 64 + 4 * tx);
            const float4 r2a = *reinterpret_cast<const float4*>(rs + (k + 1) * kT + 4 * tx);
            const float4 r2b = *reinterpret_cast<const float4*>(rs + (k + 1) * kT + 64 + 4 * tx);
            const f32x2_t r1[4] = {fw_pack(r1a.x, r1a.y), fw_pack(r1a.z, r1a.w), fw_pack(r1b.x, r1b.y), fw_pack(r1b.z, r1b.w)};
            const f32x2_t r2[4] = {fw_pack(r2a.x, r2a.y), fw_pack(r2a.z, r2a.w), fw_pack(r2b.x, r2b.y), fw_pack(r2b.z, r2b.w)};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int bp = 0; bp < 4; ++bp) {
                    float l1, h1, l2, h2;
                    fw_unpack(fw_add2(c1[a], r1[bp]), l1, h1);
                    fw_unpack(fw_add2(c2[a], r2[bp]), l2, h2);
                    m[a][2 * bp] = fw_min3(m[a][2 * bp], l1, l2);
                    m[a][2 * bp + 1] = fw_min3(m[a][2 * bp + 1], h1, h2);
                }
        }
    }
    if (!TRACK) {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int row = (a < 4) ? 4 * ty + a : 64 + 4 * ty + (a - 4);
            *reinterpret_cast<float4*>(dist + obase + row * ld + 4 * tx) = make_float4(m[a][0], m[a][1], m[a][2], m[a][3]);
            *reinterpret_cast<float4*>(dist + obase + row * ld + 64 + 4 * tx) = make_float4(m[a][4], m[a][5], m[a][6], m[a][7]);
        }
        return;
    }
    // merge: the prefetched tile (own granules only: no barrier needed) is compared with the new minima,
    // written back only where it improved, and the round of the improvement is recorded
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    uint16_t* rt = rounds + obase;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int row = (a < 4) ? 4 * ty + a : 64 + 4 * ty + (a - 4);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 d4 = dpre[(2 * a + h) * 256 + tid];
            float* dv = reinterpret_cast<float*>(&d4);
            bool any = false;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (m[a][4 * h + b] < dv[b]) {
                    dv[b] = m[a][4 * h + b];
                    any = true;
                    rt[row * ld + 64 * h + 4 * tx + b] = (uint16_t)kb;
                }
            if (any) *reinterpret_cast<float4*>(dist + obase + row * ld + 64 * h + 4 * tx) = d4;
        }
    }
}

// ---- FP64 P3: 128 x 128 tile per CTA (two 64-wide pivot blocks fit in one tile edge), 8 x 8 outputs per
// thread, the 64-deep round staged by cp.async in two halves (the second half lands while the first is
// being used), both panels kept in their natural row-major layout and read as (k, k+1) / (j, j+1) double
// pairs: 16 LDS.128 per 128 relaxations instead of the 32 of a 4 x 4 tile (round 1: shared-memory bound).
// DADD + DSETP/select per relaxation.  The pivot rows / columns inside a tile are relaxed again -- a no-op,
// P2 already closed them under this very relaxation, and nothing is written where nothing improves.
constexpr int kB64 = 64;               // round width (pivot block) in FP64
constexpr int kT64 = 128;              // P3 tile edge
constexpr int kC64Ld = kB64 + 2;       // padded row of the column panel (doubles)
constexpr int kP3Smem64 = (kT64 * kC64Ld + kB64 * kT64) * (int)sizeof(double);

__global__ void __launch_bounds__(256, 1)
fw64_p3_kernel(double* __restrict__ dist, int64_t ld, const double* __restrict__ panel, int64_t ldp,
               int kb, int skip, int nb, int nl, int part, int next_col, int next_row,
               uint16_t* __restrict__ rounds) {
    // grid in 128-tiles: nb128 x nl128; kb / skip / next_* are in 64-blocks
    const int nb2 = (nb + 1) / 2, nl2 = (nl + 1) / 2;
    int bi, bj;
    {
        const int nc = next_col >= 0 ? next_col / 2 : -1, nr = next_row >= 0 ? next_row / 2 : -1;
        if (part == 1) {
            const int x = blockIdx.x;
            if (x < nl2) { if (nc < 0) return; bi = x; bj = nc; }
            else { if (nr < 0) return; bi = nr; bj = x - nl2; if (bj == nc) return; }
        } else {
            bj = blockIdx.x; bi = blockIdx.y;
            if (part == 2 && (bj == nc || bi == nr)) return;
        }
        if (bi >= nl2 || bj >= nb2) return;
    }
    extern __shared__ __align__(16) unsigned char fw_smem[];
    double* cs = reinterpret_cast<double*>(fw_smem);          // [128][kC64Ld]: column panel (i, k)
    double* rs = cs + kT64 * kC64Ld;                           // [64][128]:     row panel (k, j)
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const double* cg = dist + (int64_t)bi * kT64 * ld + (int64_t)kb * kB64;
    const double* rg = panel + (int64_t)bj * kT64;
    const int64_t obase = (int64_t)bi * kT64 * ld + (int64_t)bj * kT64;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = tid + 256 * q;                       // 2048 granules per half and panel
            const int i = e >> 4, g = (e & 15) + 16 * half;    // column panel: row i, granule g (2 doubles of k)
            cp_async16(cs + i * kC64Ld + 2 * g, cg + (int64_t)i * ld + 2 * g);
            const int k = (e >> 6) + 32 * half, h = e & 63;    // row panel: row k, granule h (2 doubles of j)
            cp_async16(rs + k * kT64 + 2 * h, rg + (int64_t)k * ldp + 2 * h);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    double m[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) m[a][b] = inf_v<double>();
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        if (half == 0) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll 2
        for (int k = 32 * half; k < 32 * half + 32; k += 2) {
            double c1[8], c2[8], r1[8], r2[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int row = (a < 4) ? 4 * ty + a : 64 + 4 * ty + (a - 4);
                const double2 cc = *reinterpret_cast<const double2*>(cs + row * kC64Ld + k);
                c1[a] = cc.x; c2[a] = cc.y;
            }
#pragma unroll
            for (int bp = 0; bp < 4; ++bp) {
                const int col = (bp < 2) ? 4 * tx + 2 * bp : 64 + 4 * tx + 2 * (bp - 2);
                const double2 ra = *reinterpret_cast<const double2*>(rs + k * kT64 + col);
                const double2 rb = *reinterpret_cast<const double2*>(rs + (k + 1) * kT64 + col);
                r1[2 * bp] = ra.x; r1[2 * bp + 1] = ra.y;
                r2[2 * bp] = rb.x; r2[2 * bp + 1] = rb.y;
            }
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    // compare + select (DSETP + 2 FSEL); fmin() lowers to DSETP.MIN + SEL/FSEL + five moves
                    const double u = c1[a] + r1[b], v = c2[a] + r2[b];
                    double w = m[a][b];
                    w = u < w ? u : w;
                    m[a][b] = v < w ? v : w;
                }
        }
    }
    // Pivot rows / columns inside this tile belong to P1 / P2 and other CTAs are reading them as operands:
    // they are never written here (in exact arithmetic nothing would improve; skipping them keeps the
    // result independent of rounding and of the order the CTAs run in).
    uint16_t* rr = rounds ? rounds + obase : nullptr;
    const int piv_row_half = (2 * bi == kb) ? 0 : (2 * bi + 1 == kb) ? 1 : -1;     // which 64-row half is pivot
    const int piv_col_half = (2 * bj == kb) ? 0 : (2 * bj + 1 == kb) ? 1 : -1;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int row = (a < 4) ? 4 * ty + a : 64 + 4 * ty + (a - 4);
        if ((a < 4 ? 0 : 1) == piv_row_half) continue;
#pragma unroll
        for (int bp = 0; bp < 4; ++bp) {
            const int col = (bp < 2) ? 4 * tx + 2 * bp : 64 + 4 * tx + 2 * (bp - 2);
            if ((bp < 2 ? 0 : 1) == piv_col_half) continue;
            double* p = dist + obase + (int64_t)row * ld + col;
            double2 d2 = *reinterpret_cast<const double2*>(p);
            bool any = false;
            if (m[a][2 * bp] < d2.x) { d2.x = m[a][2 * bp]; any = true; if (rr) rr[(int64_t)row * ld + col] = (uint16_t)kb; }
            if (m[a][2 * bp + 1] < d2.y) { d2.y = m[a][2 * bp + 1]; any = true; if (rr) rr[(int64_t)row * ld + col + 1] = (uint16_t)kb; }
            if (any) *reinterpret_cast<double2*>(p) = d2;
        }
    }
}

// ---------------------------------------------------------------------------------------
// predecessors from (dist, rounds)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void fw_transpose_kernel(const T* __restrict__ in, int64_t n_pad, T* __restrict__ out) {
    __shared__ T tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) tile[r][threadIdx.x] = in[(int64_t)(by + r) * n_pad + bx + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) out[(int64_t)(bx + r) * n_pad + by + threadIdx.x] = tile[threadIdx.x][r];
}

// midT[j][i] = the vertex k of round rounds[i][j] that minimises dist[i][k] + dist[k][j] (k != i, j;
// first minimum wins), or -1 when (i, j) never improved (direct edge, or no path).  distT = dist
// transposed, so both operands are contiguous row segments.  Eight lanes per element; one CTA
// walks the elements of one j (row j of distT stays in L1).
template <typename T, int B>
__global__ void __launch_bounds__(256)
fw_resolve_kernel(const T* __restrict__ dist, const T* __restrict__ distT, const uint16_t* __restrict__ rounds,
                  int N, int64_t n_pad, int32_t* __restrict__ midT) {
    constexpr int V = 16 / sizeof(T);                 // elements per 16-byte load
    const int j = blockIdx.x;
    const int g = threadIdx.x >> 3, l = threadIdx.x & 7;       // 32 groups of 8 lanes
    const T* dj = distT + (int64_t)j * n_pad;
    for (int i0 = blockIdx.y * 32; i0 < N; i0 += gridDim.y * 32) {
        const int i = i0 + g;
        int best_k = -1;
        T best = inf_v<T>();
        uint16_t r = kNoRound;
        if (i < N && i != j) r = rounds[(int64_t)i * n_pad + j];
        // The scan of a group stops as soon as one of its lanes holds a candidate that reproduces
        // dist[i][j] exactly (the usual case: the improving sum is recomputed from unchanged operands):
        // any such k is a valid midpoint, and the loads of the remaining sub-blocks are skipped.
        const T dij = (r != kNoRound) ? dj[i] : inf_v<T>();           // distT[j][i] == dist[i][j]
        bool done = (r == kNoRound);
        const unsigned gmask = 0xFFu << (threadIdx.x & 24);            // the 8 lanes of this group
        for (int kk = l * V; kk < B; kk += 8 * V) {
            if (!done) {
                const T* di = dist + (int64_t)i * n_pad;
                const int k0 = (int)r * B;
                T a[V], b[V];
                *reinterpret_cast<int4*>(a) = *reinterpret_cast<const int4*>(di + k0 + kk);
                *reinterpret_cast<int4*>(b) = *reinterpret_cast<const int4*>(dj + k0 + kk);
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    const int k = k0 + kk + u;
                    const T c = a[u] + b[u];
                    if (c < best && k != i && k != j) { best = c; best_k = k; }
                }
            }
            // every lane of the warp takes part in the vote, finished groups included (`done || ballot`
            // would let them skip it by short-circuit evaluation and leave the others waiting forever)
            const unsigned hit = __ballot_sync(0xffffffffu, best_k >= 0 && best <= dij);
            done = done || (hit & gmask) != 0;
        }
        // argmin over the 8 lanes of the group (ties: smaller k); executed by every lane of the warp
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const T ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
            if (ok >= 0 && (best_k < 0 || ob < best || (ob == best && ok < best_k))) { best = ob; best_k = ok; }
        }
        if (l == 0 && i < N) midT[(int64_t)j * n_pad + i] = best_k;
    }
}

// predT[j][i] = last vertex before j on the path i -> j: follow midT[j][.] (the path i -> j splits
// at mid into i -> mid and mid -> j; the predecessor of j lies on the second part) until the
// remaining hop is a direct edge.
template <typename T>
__global__ void fw_chase_kernel(const int32_t* __restrict__ midT, const T* __restrict__ distT, int N,
                                int64_t n_pad, int32_t* __restrict__ predT) {
    const int j = blockIdx.x;
    const int32_t* mj = midT + (int64_t)j * n_pad;
    const T* dj = distT + (int64_t)j * n_pad;          // dist[i][j] for all i
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        int p = -1;
        if (i != j && dj[i] < inf_v<T>()) {
            int cur = i;
            for (int hop = 0; hop < N; ++hop) {
                const int mm = mj[cur];
                if (mm < 0) break;
                cur = mm;
            }
            p = cur;
        }
        predT[(int64_t)j * n_pad + i] = p;
    }
}

// out_dist[i][j] = (double) dist[i][j]; out_pred[i][j] = predT[j][i] (transposed through shared memory)
template <typename T>
__global__ void fw_final_kernel(const T* __restrict__ dist, const int32_t* __restrict__ predT, int N,
                                int64_t n_pad, double* __restrict__ out_dist, int32_t* __restrict__ out_pred) {
    __shared__ int32_t tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;      // output block: rows by.., cols bx..
    if (out_pred) {
        for (int r = threadIdx.y; r < 32; r += blockDim.y)       // read predT rows bx.., cols by..
            tile[r][threadIdx.x] = predT[(int64_t)(bx + r) * n_pad + by + threadIdx.x];
        __syncthreads();
    }
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = by + r, j = bx + threadIdx.x;
        if (i < N && j < N) {
            if (out_dist) out_dist[(int64_t)i * N + j] = (double)dist[(int64_t)i * n_pad + j];
            if (out_pred) out_pred[(int64_t)i * N + j] = tile[threadIdx.x][r];
        }
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <typename T> struct FwCfg;
template <> struct FwCfg<float> {
    static constexpr int B = kT;
    static int p3_smem() { return (int)sizeof(float) * kP3Stages * kP3StageFloats; }
    static constexpr int p3_threads = 256;
};
template <> struct FwCfg<double> {
    static constexpr int B = kB64;
    static int p3_smem() { return kP3Smem64; }
    static constexpr int p3_threads = 256;
};

template <typename T> int fw_p1_smem() { return FwCfg<T>::B * (FwCfg<T>::B + 1) * (int)sizeof(T); }
template <typename T> int fw_p2_smem() { return FwCfg<T>::B * (FwCfg<T>::B + 16 / (int)sizeof(T)) * (int)sizeof(T); }

template <typename T>
int fw_set_attrs() {
    constexpr int B = FwCfg<T>::B;
    WISP_CUDA(cudaFuncSetAttribute(fw_p1_kernel<T, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, fw_p1_smem<T>()));
    WISP_CUDA(cudaFuncSetAttribute(fw_p2_kernel<T, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, fw_p2_smem<T>()));
    if (sizeof(T) == 4)
    {
        WISP_CUDA(cudaFuncSetAttribute(fw32_p3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwCfg<float>::p3_smem()));
        WISP_CUDA(cudaFuncSetAttribute(fw32_p3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwCfg<float>::p3_smem()));
    }
    else
        WISP_CUDA(cudaFuncSetAttribute(fw64_p3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwCfg<double>::p3_smem()));
    return 0;
}

inline void launch_p3(float* dist, int64_t ld, const float* panel, int64_t ldp, int kb, int skip, int nb, int nl,
                      int part, int next_col, int next_row, uint16_t* rounds, cudaStream_t st) {
    const dim3 grid = part == 1 ? dim3(nl + nb) : dim3(nb, nl);
    if (rounds)
        fw32_p3_kernel<true><<<grid, 256, FwCfg<float>::p3_smem(), st>>>(dist, ld, panel, ldp, kb, skip, nb, nl, part,
                                                                        next_col, next_row, rounds);
    else
        fw32_p3_kernel<false><<<grid, 256, FwCfg<float>::p3_smem(), st>>>(dist, ld, panel, ldp, kb, skip, nb, nl, part,
                                                                         next_col, next_row, nullptr);
}
inline void launch_p3(double* dist, int64_t ld, const double* panel, int64_t ldp, int kb, int skip, int nb, int nl,
                      int part, int next_col, int next_row, uint16_t* rounds, cudaStream_t st) {
    const int nb2 = (nb + 1) / 2, nl2 = (nl + 1) / 2;         // the FP64 P3 works on 128-tiles
    const dim3 grid = part == 1 ? dim3(nl2 + nb2) : dim3(nb2, nl2);
    fw64_p3_kernel<<<grid, 256, FwCfg<double>::p3_smem(), st>>>(dist, ld, panel, ldp, kb, skip, nb, nl, part,
                                                               next_col, next_row, rounds);
}

int fw_aux(wisp_ctx* ctx) {
    if (!ctx->aux_stream) {
        int lo = 0, hi = 0;
        WISP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        WISP_CUDA(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, hi));
        for (auto& e : ctx->aux_events) WISP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return 0;
}

template <typename T>
int run_apsp(wisp_ctx* ctx, const double* d_w, int N, double* d_dist, int32_t* d_pred, cudaStream_t st) {
    constexpr int B = FwCfg<T>::B;
    const int64_t n_pad = round_up(N, 128);          // both tile sizes divide it; rows stay 16-byte aligned
    const int nb = (int)(n_pad / B);
    WISP_CHECK(nb < (int)kNoRound, "apsp: too many rounds (%d)", nb);
    void *pd = nullptr, *pr = nullptr;
    WISP_TRY(wisp_scratch(ctx, SCR_APSP_D, (size_t)n_pad * n_pad * sizeof(T), &pd));
    T* dist = (T*)pd;
    uint16_t* rounds = nullptr;
    if (d_pred) {
        WISP_TRY(wisp_scratch(ctx, SCR_APSP_R, (size_t)n_pad * n_pad * sizeof(uint16_t), &pr));
        rounds = (uint16_t*)pr;
    }
    WISP_TRY(fw_set_attrs<T>());
    WISP_TRY(fw_aux(ctx));
    cudaStream_t side = ctx->aux_stream;
    cudaEvent_t ev_lead = ctx->aux_events[0], ev_piv = ctx->aux_events[1];
    const int g = (int)std::min<int64_t>((n_pad * n_pad + 255) / 256, (int64_t)ctx->sm_count * 16);
    fw_init_kernel<T><<<g, 256, 0, st>>>(d_w, N, n_pad, 0, n_pad, dist, rounds);
    ctx->launches++;
    const int smem1 = fw_p1_smem<T>(), smem2 = fw_p2_smem<T>();
    constexpr int P1T = (B / 4) * (B / 4), SUB = B / 32;
    auto pivot_phases = [&](int kb, cudaStream_t s) {
        T* panel = dist + (int64_t)kb * B * n_pad;     // single GPU: the panel is the pivot rows in place
        uint16_t* rpanel = rounds ? rounds + (int64_t)kb * B * n_pad : nullptr;
        fw_p1_kernel<T, B><<<1, P1T, smem1, s>>>(panel, n_pad, kb, rpanel);
        ctx->launches++;
        if (nb > 1) {
            fw_p2_kernel<T, B><<<dim3(nb * SUB, 2), 128, smem2, s>>>(dist, n_pad, panel, n_pad, kb, nb, nb, kb, -1,
                                                                     rounds, rpanel, -1);
            ctx->launches++;
        }
    };
    pivot_phases(0, st);
    for (int kb = 0; kb < nb && nb > 1; ++kb) {
        const T* panel = dist + (int64_t)kb * B * n_pad;
        const bool more = kb + 1 < nb;
        if (more) {
            // lead: the tiles of row / column kb+1, then the next round's pivot phases on the side
            // stream while the rest of this round runs here
            launch_p3(dist, n_pad, panel, n_pad, kb, kb, nb, nb, 1, kb + 1, kb + 1, rounds, st);
            ctx->launches++;
            WISP_CUDA(cudaEventRecord(ev_lead, st));
            WISP_CUDA(cudaStreamWaitEvent(side, ev_lead, 0));
            pivot_phases(kb + 1, side);
            WISP_CUDA(cudaEventRecord(ev_piv, side));
        }
        launch_p3(dist, n_pad, panel, n_pad, kb, kb, nb, nb, more ? 2 : 0, more ? kb + 1 : -1, more ? kb + 1 : -1,
                  rounds, st);
        ctx->launches++;
        if (more) WISP_CUDA(cudaStreamWaitEvent(st, ev_piv, 0));
    }
    WISP_CUDA(cudaGetLastError());
    // ---- predecessors: transpose, resolve the midpoints, chase, transpose back ----
    const dim3 tb(32, 8), tg((unsigned)(n_pad / 32), (unsigned)(n_pad / 32));
    int32_t* predT = nullptr;
    if (d_pred) {
        void *pt = nullptr, *pm = nullptr, *pp = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_APSP_T, (size_t)n_pad * n_pad * sizeof(T), &pt));
        WISP_TRY(wisp_scratch(ctx, SCR_APSP_M, (size_t)n_pad * n_pad * sizeof(int32_t), &pm));
        WISP_TRY(wisp_scratch(ctx, SCR_APSP_P, (size_t)n_pad * n_pad * sizeof(int32_t), &pp));
        T* distT = (T*)pt;
        int32_t* midT = (int32_t*)pm;
        predT = (int32_t*)pp;
        fw_transpose_kernel<T><<<tg, tb, 0, st>>>(dist, n_pad, distT);
        const int gy = (int)std::max<int64_t>(1, std::min<int64_t>((N + 31) / 32, (int64_t)ctx->sm_count * 8 / std::max(1, N)));
        fw_resolve_kernel<T, B><<<dim3(N, gy), 256, 0, st>>>(dist, distT, rounds, N, n_pad, midT);
        fw_chase_kernel<T><<<N, 256, 0, st>>>(midT, distT, N, n_pad, predT);
        ctx->launches += 3;
    }
    fw_final_kernel<T><<<tg, tb, 0, st>>>(dist, predT, N, n_pad, d_dist, d_pred);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_apsp(wisp_ctx* ctx, const double* d_w, int N, int precision, double* d_dist,
                int32_t* d_pred, cudaStream_t st) {
    WISP_CHECK(N > 0, "apsp: empty graph");
    WISP_CHECK(precision == 64 || precision == 32, "apsp: precision must be 64 or 32 (got %d)", precision);
    if (precision == 64) return run_apsp<double>(ctx, d_w, N, d_dist, d_pred, st);
    return run_apsp<float>(ctx, d_w, N, d_dist, d_pred, st);
}

// ---------------------------------------------------------------------------------------
// Row-block sharded FP32 Floyd-Warshall (BASELINE configs[3]): every rank owns a
// block of 128-row tile rows.  Round kb: the owner of pivot rows kb finishes the pivot ROW
// PANEL (P1 + the row half of P2) in a separate buffer and broadcasts it (NCCL, done by the
// caller); every rank then updates its own pivot-column tiles and its remainder tiles from that
// panel.  The single-GPU path above is the same kernels with the panel aliased to the pivot rows
// in place.  wisp_dev_fw32_update_rows_part splits the remainder update for look-ahead: the owner
// of the NEXT pivot rows updates that tile-row first (part 1), builds and broadcasts the next
// panel on a second stream, and runs the rest (part 2) meanwhile.
// ---------------------------------------------------------------------------------------
extern "C" int wisp_dev_fw32_init_rows(wisp_ctx* ctx, const double* d_w_rows, int N, int64_t row0,
                                       int64_t rows_pad, float* d_dist_local, void* stream) {
    WISP_CHECK(ctx && d_w_rows && d_dist_local && rows_pad % kT == 0, "fw32_init_rows: bad argument");
    cudaStream_t st = wisp_stream(ctx, stream);
    const int64_t n_pad = round_up(N, kT);
    const int g = (int)std::min<int64_t>((rows_pad * n_pad + 255) / 256, (int64_t)ctx->sm_count * 16);
    fw_init_kernel<float><<<g, 256, 0, st>>>(d_w_rows, N, n_pad, row0, rows_pad, d_dist_local, nullptr);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

// P1 + pivot-row tiles on a [128][n_pad] panel (owner only)
extern "C" int wisp_dev_fw32_pivot_panel(wisp_ctx* ctx, float* d_panel, int N, int kb, void* stream) {
    WISP_CHECK(ctx && d_panel, "fw32_pivot_panel: bad argument");
    cudaStream_t st = wisp_stream(ctx, stream);
    const int64_t n_pad = round_up(N, kT);
    const int nb = (int)(n_pad / kT);
    WISP_TRY(fw_set_attrs<float>());
    fw_p1_kernel<float, kT><<<1, 1024, fw_p1_smem<float>(), st>>>(d_panel, n_pad, kb, nullptr);
    fw_p2_kernel<float, kT><<<dim3(nb * (kT / 32), 1), 128, fw_p2_smem<float>(), st>>>(
        d_panel, n_pad, d_panel, n_pad, kb, nb, 0, -1, -1, nullptr, nullptr, -1);
    ctx->launches += 2;
    WISP_CUDA(cudaGetLastError());
    return 0;
}

// part 0: pivot-column tiles + every remainder tile of the local row block;
// part 1: pivot-column tile and remainder tiles of local tile-row `next_local_tile` only;
// part 2: everything part 1 left out.  skip_local_tile = local tile-row holding the pivot rows
// of THIS round (-1 if they live elsewhere).
extern "C" int wisp_dev_fw32_update_rows_part(wisp_ctx* ctx, float* d_dist_local, int64_t rows_pad,
                                              const float* d_panel, int N, int kb, int skip_local_tile,
                                              int next_local_tile, int part, void* stream) {
    WISP_CHECK(ctx && d_dist_local && d_panel && rows_pad % kT == 0, "fw32_update_rows: bad argument");
    WISP_CHECK(part >= 0 && part <= 2, "fw32_update_rows: part must be 0, 1 or 2");
    cudaStream_t st = wisp_stream(ctx, stream);
    const int64_t n_pad = round_up(N, kT);
    const int nb = (int)(n_pad / kT), nl = (int)(rows_pad / kT);
    WISP_TRY(fw_set_attrs<float>());
    float* panel = const_cast<float*>(d_panel);
    if (part == 1) {
        if (next_local_tile < 0 || next_local_tile >= nl) return 0;
        // column tile (next, kb) first, then the tile-row
        fw_p2_kernel<float, kT><<<dim3(kT / 32, 2), 128, fw_p2_smem<float>(), st>>>(
            d_dist_local, n_pad, panel, n_pad, kb, 0, nl, skip_local_tile, -1, nullptr, nullptr, next_local_tile);
        launch_p3(d_dist_local, n_pad, d_panel, n_pad, kb, skip_local_tile, nb, nl, 1, -1, next_local_tile, nullptr, st);
        ctx->launches += 2;
    } else {
        // column tiles (in part 2 the tile of `next_local_tile` was already done by part 1)
        fw_p2_kernel<float, kT><<<dim3(nl * (kT / 32), 2), 128, fw_p2_smem<float>(), st>>>(
            d_dist_local, n_pad, panel, n_pad, kb, 0, nl, skip_local_tile, part == 2 ? next_local_tile : -1,
            nullptr, nullptr, -1);
        launch_p3(d_dist_local, n_pad, d_panel, n_pad, kb, skip_local_tile, nb, nl, part == 2 ? 2 : 0, -1,
                  part == 2 ? next_local_tile : -1, nullptr, st);
        ctx->launches += 2;
    }
    WISP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int wisp_dev_fw32_update_rows(wisp_ctx* ctx, float* d_dist_local, int64_t rows_pad,
                                         const float* d_panel, int N, int kb, int skip_local_tile,
                                         void* stream) {
    return wisp_dev_fw32_update_rows_part(ctx, d_dist_local, rows_pad, d_panel, N, kb, skip_local_tile, -1, 0, stream);
}
