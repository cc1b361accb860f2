// K1 (streaming form) -- per-frame translate + per-node centre of mass with TMA-staged
// coordinate streams, plus the plan object that owns the device-side node map.
//
// Same arithmetic as com_kernel (k1_com.cu, the general CSR gather), restructured for HBM
// bandwidth when every node is a contiguous, ascending run of atoms (residue COM nodes,
// make_node_selections.py:39-58):
//   align_shift_kernel  one warp per frame gathers the alignment landmarks (a few percent of
//                       the frame) and writes shift[f] = -COM_align(f)           (:71);
//   com_stream_kernel   persistent CTAs; the frame is cut into "pieces" (<= 128 consecutive nodes,
//                       <= 2048 atoms); each CTA owns one piece index and walks the
//                       frames with a 3-stage ring of cp.async.bulk (TMA) copies completed
//                       on mbarriers, so a whole piece (24 KB) is in flight per stage with
//                       no registers held; one lane PAIR per node -- each lane walks every
//                       second atom of the node in shared memory for all three components
//                       (three independent FP64 chains, 27 instructions per atom instead of
//                       the 72 of a lane-per-component layout), applies the float32 translate
//                       rounding (Veltkamp split on the FP64 pipe), accumulates m*x in FP64;
//                       one shuffle step joins the pair and the COM (sum times the
//                       precomputed 1/M_node) is stored (:76).
// HBM traffic: 12*A (bulk copies) + 24*N (stores) per frame, plus the 32-byte sectors of
// the landmark gather (~17 % of a frame when there is one landmark per 16-atom node).
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int kPieceAtomsMax = 2048;
constexpr int kStStages = 3;
constexpr int kStThreads = 256;                            // compute threads (one lane pair per node)
constexpr int kStLaunchThreads = kStThreads + 32;          // + the producer warp
constexpr int kPieceNodesMax = kStThreads / 2;             // one lane pair per node, one pass per piece
constexpr int kStageBytes = kPieceAtomsMax * 12 + 32;     // + slack for the 16-byte alignment shift

struct Piece {
    int32_t atom0, n_atoms, node0, n_nodes;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float ldg_sparse(const float* p) {      // isolated 12-byte reads: smallest L2 fetch
    float v;
    asm volatile("ld.global.nc.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// ---- shift[f] = -(sum_a m_a x_a / sum_a m_a) over the alignment landmarks; one warp per frame ----
__global__ void __launch_bounds__(256)
align_shift_kernel(const float* __restrict__ coords, int64_t T, int A,
                   const int32_t* __restrict__ align_idx, const double* __restrict__ align_mass,
                   int n_align, double* __restrict__ shift,
                   float* __restrict__ align_out, int64_t row0) {
    const int lane = threadIdx.x & 31;
    const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= T) return;
    const float* fr = coords + f * (int64_t)A * 3;
    double sx = 0, sy = 0, sz = 0, sm = 0;
#pragma unroll 4
    for (int a = lane; a < n_align; a += 32) {
        const float* p = fr + 3 * (int64_t)__ldg(align_idx + a);
        const double m = __ldg(align_mass + a);
        sx += m * (double)ldg_sparse(p + 0);
        sy += m * (double)ldg_sparse(p + 1);
        sz += m * (double)ldg_sparse(p + 2);
        sm += m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
    }
    const double hx = -(sx / sm), hy = -(sy / sm), hz = -(sz / sm);
    if (lane == 0) { shift[3 * f + 0] = hx; shift[3 * f + 1] = hy; shift[3 * f + 2] = hz; }
    if (align_out) {   // all_pos_Align.append(u_alignment.positions)  (:72)
        float* ao = align_out + (row0 + f) * (int64_t)n_align * 3;
        for (int a = lane; a < n_align; a += 32) {
            const float* p = fr + 3 * (int64_t)__ldg(align_idx + a);
            ao[3 * a + 0] = __double2float_rn((double)__ldg(p + 0) + hx);
            ao[3 * a + 1] = __double2float_rn((double)__ldg(p + 1) + hy);
            ao[3 * a + 2] = __double2float_rn((double)__ldg(p + 2) + hz);
        }
    }
}

// ---- streaming node COM ----
// round a double to float precision (round-to-nearest-even) without leaving the FP64 pipe:
// Veltkamp split with C = 2^29 + 1 -- bit-identical to (double)(float)t for |t| in float's
// normal range (checked against numpy on random values, exact ties and near-ties); the
// intrinsics keep the compiler from contracting the sequence into FMAs.  This replaces two
// F2F conversions per coordinate on the quarter-rate XU pipe.
__device__ __forceinline__ double round_to_f32(double t) {
    const double g = __dmul_rn(t, 536870913.0);
    const double d = __dsub_rn(g, t);
    return __dsub_rn(g, d);
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

__global__ void __launch_bounds__(kStLaunchThreads, 2)
com_stream_kernel(const float* __restrict__ coords, int64_t T, int A, const Piece* __restrict__ pieces,
                  int P, const double* __restrict__ shift, const int4* __restrict__ node_info,
                  const double* __restrict__ ent_mass, const double* __restrict__ inv_node_mass,
                  double* __restrict__ nodes, int64_t ld, int64_t row0, int64_t total_bytes,
                  double* __restrict__ colpart, const int32_t* __restrict__ lm_ptr,
                  const int2* __restrict__ lm_ent, float* __restrict__ align_out, int n_align) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[kStStages], empty[kStStages];
    double* smass = reinterpret_cast<double*>(smem_raw + (size_t)kStStages * kStageBytes);
    const int tid = threadIdx.x;
    const int p = blockIdx.x % P;
    const int64_t f_start = blockIdx.x / P;
    const int64_t f_stride = gridDim.x / P;
    const Piece pc = pieces[p];
    const int64_t n_items = (f_start < T) ? (T - f_start + f_stride - 1) / f_stride : 0;
    const uint32_t piece_bytes = (uint32_t)pc.n_atoms * 12u;

    if (tid == 0) {
        for (int s = 0; s < kStStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kStThreads / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // this CTA always works on the same piece: its per-atom masses live in shared memory and
    // its per-lane node descriptor in registers for the whole kernel
    for (int i = tid; i < pc.n_atoms; i += kStLaunchThreads) smass[i] = 0.0;
    __syncthreads();
    for (int nl = tid; nl < pc.n_nodes; nl += kStLaunchThreads) {
        const int4 info = __ldg(node_info + pc.node0 + nl);
        for (int a = 0; a < info.y; ++a) smass[info.z + a] = __ldg(ent_mass + info.x + a);
    }

    // lane pair = one node: lane h of the pair takes the atoms a = h (mod 2), all three components
    // (three independent FP64 chains), the pair is combined with one shuffle step.  Atom order is
    // rotated per node so that the 32 lanes of a warp read 32 different banks when nodes are
    // equally sized (16 atoms: float address 48 j + 6 u + 3 h, u rotated by j / 2).
    const int nl = tid >> 1, h = tid & 1;
    const bool valid = tid < kStThreads && nl < pc.n_nodes;
    int cnt_h = 0, rot = 0;
    uint32_t q_rel = 0, m_addr = 0;
    double inv_m = 0.0;
    int64_t out_col = 0;
    if (valid) {
        const int4 info = __ldg(node_info + pc.node0 + nl);
        cnt_h = (info.y + 1 - h) >> 1;                       // atoms h, h+2, ... of this node
        q_rel = 12u * (uint32_t)(info.z + h);                // byte offset of atom h inside the piece
        m_addr = smem_u32(smass) + 8u * (uint32_t)(info.z + h);
        rot = cnt_h > 0 ? (nl >> 1) % cnt_h : 0;
        inv_m = __ldg(inv_node_mass + pc.node0 + nl);
        out_col = 3 * (int64_t)(pc.node0 + nl);
    }
    int cnt_max = cnt_h;                                     // warp-uniform trip count
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt_max = max(cnt_max, __shfl_xor_sync(0xffffffffu, cnt_max, o));
    __syncthreads();

    if (tid >= kStThreads) {
        // ---------------- producer warp: stage item k into slot k % kStStages ----------------
        // TMA bulk copy from the 16-byte-aligned address at or below the piece start; the last bytes
        // of the whole buffer are copied by hand so the bulk copy never reads past the allocation.
        const int lane = tid & 31;
        for (int64_t k = 0; k < n_items; ++k) {
            const int slot = (int)(k % kStStages);
            const uint32_t ph = (uint32_t)((k / kStStages) & 1);
            mbar_wait(&empty[slot], ph ^ 1);
            const int64_t f = f_start + k * f_stride;
            const int64_t src = (f * (int64_t)A + pc.atom0) * 12;
            const int64_t src_al = src & ~(int64_t)15;
            const uint32_t delta = (uint32_t)(src - src_al);
            const uint32_t bytes = (delta + piece_bytes + 15u) & ~15u;
            unsigned char* dst = smem_raw + (size_t)slot * kStageBytes;
            if (src_al + bytes <= total_bytes) {
                if (lane == 0) {
                    mbar_expect_tx(&full[slot], bytes);
                    bulk_g2s(dst, reinterpret_cast<const unsigned char*>(coords) + src_al, bytes, &full[slot]);
                }
            } else {
                const float* sp = coords + (f * (int64_t)A + pc.atom0) * 3;
                float* d = reinterpret_cast<float*>(dst + delta);
                for (int i = lane; i < pc.n_atoms * 3; i += 32) d[i] = __ldg(sp + i);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_expect_tx(&full[slot], 0);   // plain arrival completes the phase
            }
        }
        return;
    }

    // ---------------- compute warps ----------------
    const uint32_t smem_base = smem_u32(smem_raw);
    const int lm_lo = align_out ? __ldg(lm_ptr + p) : 0, lm_hi = align_out ? __ldg(lm_ptr + p + 1) : 0;
    double cs0 = 0.0, cs1 = 0.0;      // column sums of what this lane stores, over this CTA's frames
    for (int64_t k = 0; k < n_items; ++k) {
        const int slot = (int)(k % kStStages);
        const uint32_t ph = (uint32_t)((k / kStStages) & 1);
        mbar_wait(&full[slot], ph);
        const int64_t f = f_start + k * f_stride;
        const int64_t src = (f * (int64_t)A + pc.atom0) * 12;
        const uint32_t q_addr = smem_base + (uint32_t)slot * kStageBytes + (uint32_t)(src & 15) + q_rel;
        const double hx = __ldg(shift + 3 * f), hy = __ldg(shift + 3 * f + 1), hz = __ldg(shift + 3 * f + 2);
        double ax = 0.0, ay = 0.0, az = 0.0;
        int u = rot;
#pragma unroll 2
        for (int it = 0; it < cnt_max; ++it) {
            if (it < cnt_h) {
                // translate(): float32 position += float64 vector, rounded back to float32
                const uint32_t qa = q_addr + 24u * (uint32_t)u;
                const double m = lds_f64(m_addr + 16u * (uint32_t)u);
                const double vx = round_to_f32((double)lds_f32(qa) + hx);
                const double vy = round_to_f32((double)lds_f32(qa + 4) + hy);
                const double vz = round_to_f32((double)lds_f32(qa + 8) + hz);
                ax = fma(m, vx, ax);
                ay = fma(m, vy, ay);
                az = fma(m, vz, az);
                u = (u + 1 == cnt_h) ? 0 : u + 1;
            }
        }
        if (align_out) {
            // the translated alignment landmarks of this piece (all_pos_Align, :72) leave from the staged
            // copy as well: entry = (atom offset inside the piece, landmark index); no second gather
            const uint32_t p_addr = smem_base + (uint32_t)slot * kStageBytes + (uint32_t)(src & 15);
            float* ao = align_out + (row0 + f) * (int64_t)n_align * 3;
            for (int e = lm_lo + tid; e < lm_hi; e += kStThreads) {
                const int2 ent = __ldg(lm_ent + e);
                const uint32_t qa = p_addr + 12u * (uint32_t)ent.x;
                ao[3 * ent.y + 0] = __double2float_rn((double)lds_f32(qa) + hx);
                ao[3 * ent.y + 1] = __double2float_rn((double)lds_f32(qa + 4) + hy);
                ao[3 * ent.y + 2] = __double2float_rn((double)lds_f32(qa + 8) + hz);
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[slot]);      // this warp has read the stage
        ax += __shfl_xor_sync(0xffffffffu, ax, 1);
        ay += __shfl_xor_sync(0xffffffffu, ay, 1);
        az += __shfl_xor_sync(0xffffffffu, az, 1);
        if (valid) {
            double* o = nodes + (row0 + f) * ld + out_col;
            if (h == 0) {
                const double vx = ax * inv_m, vy = ay * inv_m;
                o[0] = vx; o[1] = vy;
                cs0 += vx; cs1 += vy;
            } else {
                const double vz = az * inv_m;
                o[2] = vz;
                cs0 += vz;
            }
        }
    }
    if (colpart && valid) {           // row = this CTA's frame class; summed over rows in fixed order later
        double* o = colpart + f_start * ld + out_col;
        if (h == 0) { o[0] = cs0; o[1] = cs1; }
        else o[2] = cs0;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------
struct wisp_com_plan {
    int A = 0, N = 0, n_align = 0, atomic = 0, n_entries = 0;
    bool streaming = false;
    int P = 0;
    // device arrays
    double* d_masses = nullptr;
    int32_t *d_node_ptr = nullptr, *d_atom_idx = nullptr, *d_align_idx = nullptr;
    double *d_align_mass = nullptr, *d_ent_mass = nullptr, *d_inv_node_mass = nullptr;
    int4* d_node_info = nullptr;
    Piece* d_pieces = nullptr;
    // alignment landmarks grouped by the piece that stages their atom: lm_ptr [P+1], lm_ent = (atom
    // offset inside the piece, landmark index); lm_fused: every landmark atom lies inside some piece
    int32_t* d_lm_ptr = nullptr;
    int2* d_lm_ent = nullptr;
    bool lm_fused = false;
    double* d_shift = nullptr;
    int64_t shift_frames = 0;
};

namespace {
template <typename T>
int upload(const std::vector<T>& h, T** d) {
    WISP_CUDA(cudaMalloc((void**)d, std::max<size_t>(h.size(), 1) * sizeof(T)));
    if (!h.empty()) WISP_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}
}  // namespace

int com_plan_create(wisp_ctx* ctx, int A, const double* masses, const int32_t* node_ptr,
                    const int32_t* atom_idx, int N, const int32_t* align_idx, int n_align, int atomic,
                    wisp_com_plan** out) {
    WISP_CHECK(ctx && masses && node_ptr && atom_idx && align_idx && out, "com_plan: null argument");
    WISP_CHECK(A > 0 && N > 0 && n_align > 0, "com_plan: empty input (A=%d N=%d nAlign=%d)", A, N, n_align);
    WISP_CHECK(node_ptr[0] == 0, "com_plan: node_ptr[0] must be 0");
    for (int n = 0; n < N; ++n) WISP_CHECK(node_ptr[n + 1] > node_ptr[n], "com_plan: node %d has no atoms", n);
    const int E = node_ptr[N];
    for (int e = 0; e < E; ++e) WISP_CHECK(atom_idx[e] >= 0 && atom_idx[e] < A, "com_plan: atom index %d out of range", atom_idx[e]);
    for (int e = 0; e < n_align; ++e) WISP_CHECK(align_idx[e] >= 0 && align_idx[e] < A, "com_plan: alignment atom index %d out of range", align_idx[e]);
    WISP_CUDA(cudaSetDevice(ctx->device));
    wisp_com_plan* pl = new wisp_com_plan();
    pl->A = A; pl->N = N; pl->n_align = n_align; pl->atomic = atomic; pl->n_entries = E;
    std::vector<double> h_masses(masses, masses + A), h_align_mass(n_align), h_ent_mass(E), h_node_mass(N);
    std::vector<int32_t> h_np(node_ptr, node_ptr + N + 1), h_ai(atom_idx, atom_idx + E),
        h_al(align_idx, align_idx + n_align), h_first(N);
    for (int a = 0; a < n_align; ++a) h_align_mass[a] = masses[align_idx[a]];
    bool contiguous = !atomic;
    for (int n = 0; n < N; ++n) {
        double m = 0.0;
        for (int e = node_ptr[n]; e < node_ptr[n + 1]; ++e) {
            h_ent_mass[e] = masses[atom_idx[e]];
            m += masses[atom_idx[e]];
            if (e > node_ptr[n] && atom_idx[e] != atom_idx[e - 1] + 1) contiguous = false;
        }
        h_node_mass[n] = m;
        h_first[n] = atom_idx[node_ptr[n]];
        if (node_ptr[n + 1] - node_ptr[n] > kPieceAtomsMax) contiguous = false;
        if (n > 0 && h_first[n] < h_first[n - 1] + (node_ptr[n] - node_ptr[n - 1])) contiguous = false;
    }
    // pieces: runs of consecutive nodes whose atoms span <= kPieceAtomsMax atoms; a gap of
    // more than 64 unused atoms starts a new piece
    std::vector<Piece> h_pieces;
    if (contiguous) {
        int n = 0;
        while (n < N) {
            Piece pc;
            pc.node0 = n;
            pc.atom0 = h_first[n];
            int end = h_first[n] + (node_ptr[n + 1] - node_ptr[n]);
            int m = n + 1;
            while (m < N) {
                const int s = h_first[m], e = s + (node_ptr[m + 1] - node_ptr[m]);
                if (s - end > 64 || e - pc.atom0 > kPieceAtomsMax || m - n >= kPieceNodesMax) break;
                end = e;
                ++m;
            }
            pc.n_nodes = m - n;
            pc.n_atoms = end - pc.atom0;
            h_pieces.push_back(pc);
            n = m;
        }
    }
    pl->streaming = contiguous && !h_pieces.empty() && getenv("WISP_K1_GATHER") == nullptr;
    pl->P = (int)h_pieces.size();
    std::vector<int32_t> h_lm_ptr(h_pieces.size() + 1, 0);
    std::vector<int2> h_lm_ent;
    if (pl->streaming) {
        // piece of every landmark atom (pieces are sorted by atom0 and do not overlap)
        std::vector<std::vector<int2>> per_piece(h_pieces.size());
        bool all_inside = true;
        for (int a = 0; a < n_align && all_inside; ++a) {
            const int atom = align_idx[a];
            int lo = 0, hi = (int)h_pieces.size() - 1, found = -1;
            while (lo <= hi) {
                const int mid = (lo + hi) / 2;
                if (atom < h_pieces[mid].atom0) hi = mid - 1;
                else if (atom >= h_pieces[mid].atom0 + h_pieces[mid].n_atoms) lo = mid + 1;
                else { found = mid; break; }
            }
            if (found < 0) all_inside = false;
            else per_piece[found].push_back(make_int2(atom - h_pieces[found].atom0, a));
        }
        pl->lm_fused = all_inside && getenv("WISP_K1_LANDMARK_GATHER") == nullptr;
        if (pl->lm_fused)
            for (size_t q = 0; q < h_pieces.size(); ++q) {
                for (const int2& e : per_piece[q]) h_lm_ent.push_back(e);
                h_lm_ptr[q + 1] = (int32_t)h_lm_ent.size();
            }
    }
    std::vector<int4> h_info(N);
    std::vector<double> h_inv(N);
    for (int n = 0; n < N; ++n) { h_info[n] = make_int4(node_ptr[n], node_ptr[n + 1] - node_ptr[n], 0, 0); h_inv[n] = 1.0 / h_node_mass[n]; }
    for (const Piece& pc : h_pieces)
        for (int n = pc.node0; n < pc.node0 + pc.n_nodes; ++n) h_info[n].z = h_first[n] - pc.atom0;
    int rc = upload(h_masses, &pl->d_masses) || upload(h_np, &pl->d_node_ptr) || upload(h_ai, &pl->d_atom_idx) ||
             upload(h_al, &pl->d_align_idx) || upload(h_align_mass, &pl->d_align_mass) ||
             upload(h_ent_mass, &pl->d_ent_mass) || upload(h_inv, &pl->d_inv_node_mass) ||
             upload(h_info, &pl->d_node_info) || upload(h_pieces, &pl->d_pieces) ||
             upload(h_lm_ptr, &pl->d_lm_ptr) || upload(h_lm_ent, &pl->d_lm_ent);
    if (rc) { com_plan_destroy(ctx, pl); return 1; }
    *out = pl;
    return 0;
}

void com_plan_destroy(wisp_ctx* ctx, wisp_com_plan* pl) {
    if (!pl) return;
    if (ctx) cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(pl->d_masses); cudaFree(pl->d_node_ptr); cudaFree(pl->d_atom_idx); cudaFree(pl->d_align_idx);
    cudaFree(pl->d_align_mass); cudaFree(pl->d_ent_mass); cudaFree(pl->d_inv_node_mass);
    cudaFree(pl->d_node_info); cudaFree(pl->d_pieces); cudaFree(pl->d_shift);
    cudaFree(pl->d_lm_ptr); cudaFree(pl->d_lm_ent);
    delete pl;
}

int com_plan_info(const wisp_com_plan* pl, int* streaming, int* n_pieces) {
    if (streaming) *streaming = pl->streaming ? 1 : 0;
    if (n_pieces) *n_pieces = pl->P;
    return 0;
}

int launch_com_plan(wisp_ctx* ctx, wisp_com_plan* pl, const float* d_coords, int64_t T, double* d_nodes,
                    int64_t ld, int64_t row0, cudaStream_t st, float* d_align_out, double* d_colsum) {
    WISP_CHECK(pl && d_coords && d_nodes && T > 0, "com: bad argument");
    ctx->i8_stats_x = nullptr;        // the node trajectory is being rewritten
    if (!pl->streaming) {
        WISP_TRY(launch_com(ctx, d_coords, T, pl->A, pl->d_masses, pl->d_node_ptr, pl->d_atom_idx, pl->N,
                            pl->d_align_idx, pl->n_align, pl->atomic, d_nodes, ld, row0, st, d_align_out));
        if (d_colsum) WISP_TRY(launch_colsum(ctx, d_nodes + row0 * ld, T, ld, 3 * (int64_t)pl->N, d_colsum, st));
        return 0;
    }
    if (pl->shift_frames < T) {
        if (pl->d_shift) { WISP_CUDA(cudaDeviceSynchronize()); WISP_CUDA(cudaFree(pl->d_shift)); pl->d_shift = nullptr; }
        WISP_CUDA(cudaMalloc((void**)&pl->d_shift, (size_t)T * 3 * 8));
        pl->shift_frames = T;
    }
    // the landmark positions leave from the streaming kernel's staged copy when every landmark atom is
    // inside a piece (the usual case: alignment atoms are node atoms); otherwise the gather kernel
    // writes them (a second, sector-granular pass over the landmarks)
    const bool lm_fused = d_align_out && pl->lm_fused;
    align_shift_kernel<<<(unsigned)((T + 7) / 8), 256, 0, st>>>(d_coords, T, pl->A, pl->d_align_idx,
                                                               pl->d_align_mass, pl->n_align, pl->d_shift,
                                                               lm_fused ? nullptr : d_align_out, row0);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    const int smem = kStStages * kStageBytes + kPieceAtomsMax * 8;
    WISP_CUDA(cudaFuncSetAttribute(com_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int P = pl->P;
    int per_piece = std::max(1, (2 * ctx->sm_count) / P);
    per_piece = (int)std::min<int64_t>(per_piece, T);
    double* d_colpart = nullptr;      // [per_piece][ld] column sums per frame class, fused into the COM pass
    if (d_colsum) {
        void* p = nullptr;
        WISP_TRY(wisp_scratch(ctx, SCR_COLSUM_PART, (size_t)per_piece * ld * 8, &p));
        d_colpart = (double*)p;
    }
    com_stream_kernel<<<P * per_piece, kStLaunchThreads, smem, st>>>(
        d_coords, T, pl->A, pl->d_pieces, P, pl->d_shift, pl->d_node_info, pl->d_ent_mass,
        pl->d_inv_node_mass, d_nodes, ld, row0, (int64_t)T * pl->A * 12, d_colpart, pl->d_lm_ptr, pl->d_lm_ent,
        lm_fused ? d_align_out : nullptr, pl->n_align);
    ctx->launches++;
    WISP_CUDA(cudaGetLastError());
    if (d_colsum) WISP_TRY(launch_colsum_finish(ctx, d_colpart, per_piece, ld, 3 * (int64_t)pl->N, d_colsum, st));
    return 0;
}
