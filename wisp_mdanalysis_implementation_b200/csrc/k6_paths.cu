// K6 -- suboptimal source->sink path enumeration on the GPU + the get_paths contract.
//
// Replaces network_analysis_functions.py:12-214 (get_paths).  Verified behaviour
// (SURVEY.md Appendix A): seed = first strict minimum over (source, sink) pairs of the
// left-to-right float64 length of a shortest path; cutoff starts there and grows by repeated
// "+= 0.1"; each pass collects, per pair, ALL simple paths with LR length <= cutoff (sink
// only as last node; paths whose 2nd node is 0 are never produced when source != 0 -- the
// reference's `node in [length, ...]` membership test), sorted by (length, nodes); a pass
// list is de-duplicated with a path equal to its reverse; the loop stops after the first
// pass with >= K entries; the union over passes plus the seed is de-duplicated, sorted, and
// truncated to K only when the last pass did not count exactly K (K+1 results possible).
//
// The reference expands every partial path below the cutoff in Python lists with no
// distance-to-sink bound.  Here each pass is two kernels over the APSP result of K5:
//   expand_level_kernel  level-synchronous expansion of the first few hops (one thread per
//                        partial path) until there are enough independent prefixes;
//   dfs_warp_kernel      one WARP per prefix runs an explicit-stack DFS (32 edges examined per
//                        step, ballot-compacted pushes) pruned by the admissible bound  len + w(u,v) + dist(v, sink) <= cutoff (1 + 1e-9),
//                        with the exact float64 LR length for the final `<= cutoff` test.
// Paths are emitted through atomically reserved slots; the host sorts, so the result does
// not depend on emission order.
#include "common.cuh"
#include <algorithm>
#include <chrono>
#include <set>

namespace {

constexpr int kMaxLen0 = 96;         // initial capacity of a path record (nodes); grows x4 on demand, up to N
constexpr int kTargetPrefixes = 4096;
constexpr int kMaxExpandLevels = 8;

struct EnumParams {
    const int32_t* adj_ptr;
    const int32_t* adj_idx;
    const double* adj_w;
    const double* dist;      // APSP distances [N][N] (row t = distances to sink t)
    const int32_t* pairs;    // [P][2]
    // A pass serves one or more QUERY GROUPS at once: the pairs of one get_paths call form one group
    // (one cutoff, one shared count); the batched API (wisp_paths_query_pairs) puts every independent
    // (source, sink) query in its own group, each at its own point of its own cutoff ratchet.
    const int32_t* pair_group;   // [P] group of every pair
    const double* g_cutoff;      // [G] LR-length cutoff of the pass
    const double* g_bound;       // [G] admissible pruning bound (cutoff + slack)
    const int32_t* g_count_only; // [G] 1: count paths, do not store them
    const int32_t* g_cap;        // [G] count-only early exit: stop once more than cap paths were seen
    int32_t* g_count;            // [G] paths found per group (both modes)
    int N;
    int quirk;
    int max_len;             // capacity of a path record / a warp's path buffer (nodes)
    // outputs
    int32_t* out_count;      // [0] stored paths, [1] n node entries, [2] overflow flags
    unsigned long long* expansions;
    double* out_len;
    int32_t* out_pair;
    int32_t* out_off;        // node offset of path
    int32_t* out_n;          // nodes in path
    int32_t* out_nodes;
    int cap_paths;
    int cap_nodes;
};

__device__ __forceinline__ void emit_path(const EnumParams& P, int pair, const int32_t* nodes,
                                          int depth, int last, double len) {
    const int g = P.pair_group[pair];
    atomicAdd(&P.g_count[g], 1);
    if (P.g_count_only[g]) return;
    const int slot = atomicAdd(&P.out_count[0], 1);
    const int off = atomicAdd(&P.out_count[1], depth + 1);
    if (slot >= P.cap_paths || off + depth + 1 > P.cap_nodes) {
        atomicOr(&P.out_count[2], 1);
        return;
    }
    P.out_len[slot] = len;
    P.out_pair[slot] = pair;
    P.out_off[slot] = off;
    P.out_n[slot] = depth + 1;
    for (int k = 0; k < depth; ++k) P.out_nodes[off + k] = nodes[k];
    P.out_nodes[off + depth] = last;
}

// frontier record layout: len[cap], meta[cap][2] = (pair, depth), nodes[cap][max_len]
__global__ void expand_level_kernel(EnumParams P, const double* __restrict__ fin_len,
                                    const int32_t* __restrict__ fin_meta,
                                    const int32_t* __restrict__ fin_nodes, int n_in,
                                    double* __restrict__ fout_len, int32_t* __restrict__ fout_meta,
                                    int32_t* __restrict__ fout_nodes, int32_t* fout_count,
                                    int cap_out) {
    const int rec = blockIdx.x * blockDim.x + threadIdx.x;
    if (rec >= n_in) return;
    const int pair = fin_meta[2 * rec], depth = fin_meta[2 * rec + 1];
    const int s = P.pairs[2 * pair], t = P.pairs[2 * pair + 1];
    const int grp = P.pair_group[pair];
    const double cutoff = P.g_cutoff[grp], bound = P.g_bound[grp];
    const int32_t* nodes = fin_nodes + (size_t)rec * P.max_len;
    const double len = fin_len[rec];
    const int u = nodes[depth - 1];
    const double* dt = P.dist + (size_t)t * P.N;
    unsigned long long nexp = 0;
    for (int e = P.adj_ptr[u]; e < P.adj_ptr[u + 1]; ++e) {
        const int v = P.adj_idx[e];
        bool on = false;
        for (int k = 0; k < depth; ++k) on |= (nodes[k] == v);
        if (on) continue;
        if (P.quirk && depth == 1 && v == 0 && s != 0) continue;
        const double nl = len + P.adj_w[e];
        ++nexp;
        if (!(nl + dt[v] <= bound)) continue;
        if (v == t) {
            if (nl <= cutoff) emit_path(P, pair, nodes, depth, v, nl);
            continue;
        }
        if (nl > cutoff) continue;
        if (depth + 1 >= P.max_len) { atomicOr(&P.out_count[2], 2); continue; }
        const int slot = atomicAdd(fout_count, 1);
        if (slot >= cap_out) { atomicOr(&P.out_count[2], 4); continue; }
        fout_len[slot] = nl;
        fout_meta[2 * slot] = pair;
        fout_meta[2 * slot + 1] = depth + 1;
        int32_t* dst = fout_nodes + (size_t)slot * P.max_len;
        for (int k = 0; k < depth; ++k) dst[k] = nodes[k];
        dst[depth] = v;
    }
    atomicAdd(P.expansions, nexp);
}

// One WARP per prefix: depth-first search with an explicit stack of pending children in global
// memory (16 bytes per entry, coalesced pushes).  Popping an entry makes it the path tip at
// its depth (the prefix below it is still valid: deeper entries are popped first); its
// adjacency list is then examined 32 edges at a time -- coalesced loads of (neighbour, weight),
// a gather of dist(v, sink), the on-path test against the warp's path in shared memory, the
// admissible bound -- survivors are pushed with one ballot-compacted store, sink hits are
// emitted with warp-aggregated slot reservation.
struct StackEntry {
    int32_t node, depth;
    double len;
};

constexpr int kDfsWarps = 4;      // warps per block

__global__ void __launch_bounds__(kDfsWarps * 32)
dfs_warp_kernel(EnumParams P, const double* __restrict__ f_len, const int32_t* __restrict__ f_meta,
                const int32_t* __restrict__ f_nodes, int n_in, StackEntry* __restrict__ stacks,
                int cap_stack) {
    extern __shared__ int32_t spath[];      // [kDfsWarps][max_len]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int rec = blockIdx.x * kDfsWarps + wib;
    if (rec >= n_in) return;
    const unsigned FULL = 0xffffffffu;
    const int pair = f_meta[2 * rec], d0 = f_meta[2 * rec + 1];
    const int s = P.pairs[2 * pair], t = P.pairs[2 * pair + 1];
    const int grp = P.pair_group[pair];
    const double cutoff = P.g_cutoff[grp], bound = P.g_bound[grp];
    const int count_only = P.g_count_only[grp], cap = P.g_cap[grp];
    const double* dt = P.dist + (size_t)t * P.N;
    int32_t* path = spath + (size_t)wib * P.max_len;
    for (int k = lane; k < d0; k += 32) path[k] = f_nodes[(size_t)rec * P.max_len + k];
    StackEntry* stack = stacks + (size_t)rec * cap_stack;
    int sp = 0;
    if (lane == 0) { stack[0].node = f_nodes[(size_t)rec * P.max_len + d0 - 1]; stack[0].depth = d0 - 1; stack[0].len = f_len[rec]; }
    sp = 1;
    __syncwarp();
    unsigned long long nexp = 0;
    unsigned it = 0;
    while (sp > 0) {
        if (count_only && (++it & 15u) == 0) {
            int c = 0;
            if (lane == 0) c = *reinterpret_cast<volatile int32_t*>(P.g_count + grp);
            c = __shfl_sync(FULL, c, 0);
            if (c > cap) break;
        }
        --sp;
        const StackEntry top = stack[sp];          // same address for all lanes: broadcast
        const int u = top.node, depth = top.depth + 1;   // path has `depth` nodes once u is placed
        if (lane == 0) path[depth - 1] = u;
        __syncwarp();
        const int e_end = P.adj_ptr[u + 1];
        for (int e0 = P.adj_ptr[u]; e0 < e_end; e0 += 32) {
            const int e = e0 + lane;
            bool ok = e < e_end;
            int v = -1;
            double nl = 0.0;
            if (ok) {
                v = P.adj_idx[e];
                nl = top.len + P.adj_w[e];
            }
            if (ok) {
                for (int k = 0; k < depth; ++k) ok &= (path[k] != v);
            }
            if (ok && P.quirk && depth == 1 && v == 0 && s != 0) ok = false;
            nexp += ok ? 1 : 0;
            if (ok) ok = (nl + dt[v] <= bound);
            const bool hit = ok && v == t && nl <= cutoff;
            const bool cont = ok && v != t && nl <= cutoff;
            // ---- sink hits ----
            const unsigned hmask = __ballot_sync(FULL, hit);
            if (hmask) {
                const int nh = __popc(hmask);
                int slot0 = 0, off0 = 0;
                if (lane == 0) {
                    atomicAdd(&P.g_count[grp], nh);
                    if (!count_only) {
                        slot0 = atomicAdd(&P.out_count[0], nh);
                        off0 = atomicAdd(&P.out_count[1], nh * (depth + 1));
                    }
                }
                slot0 = __shfl_sync(FULL, slot0, 0);
                off0 = __shfl_sync(FULL, off0, 0);
                if (hit && !count_only) {
                    const int r = __popc(hmask & ((1u << lane) - 1));
                    const int slot = slot0 + r, off = off0 + r * (depth + 1);
                    if (slot >= P.cap_paths || off + depth + 1 > P.cap_nodes) {
                        atomicOr(&P.out_count[2], 1);
                    } else {
                        P.out_len[slot] = nl;
                        P.out_pair[slot] = pair;
                        P.out_off[slot] = off;
                        P.out_n[slot] = depth + 1;
                        for (int k = 0; k < depth; ++k) P.out_nodes[off + k] = path[k];
                        P.out_nodes[off + depth] = v;
                    }
                }
            }
            // ---- push survivors ----
            const unsigned cmask = __ballot_sync(FULL, cont);
            if (cmask) {
                const int nc = __popc(cmask);
                if (depth + 1 >= P.max_len) { if (lane == 0) atomicOr(&P.out_count[2], 2); }
                else if (sp + nc > cap_stack) { if (lane == 0) atomicOr(&P.out_count[2], 8); }
                else {
                    if (cont) {
                        const int r = __popc(cmask & ((1u << lane) - 1));
                        StackEntry en;
                        en.node = v; en.depth = depth; en.len = nl;
                        stack[sp + r] = en;
                    }
                    sp += nc;
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0) atomicAdd(P.expansions, nexp);
    // lanes other than 0 counted their own edges too
    nexp = (lane == 0) ? 0 : nexp;
    for (int o = 16; o > 0; o >>= 1) nexp += __shfl_xor_sync(FULL, nexp, o);
    if (lane == 0) atomicAdd(P.expansions, nexp);
}

struct HostPath {
    double len;
    std::vector<int32_t> nodes;
};

inline bool path_less(const HostPath& a, const HostPath& b) {
    // Python list comparison of [len, n0, n1, ...]
    if (a.len != b.len) return a.len < b.len;
    return a.nodes < b.nodes;
}

// remove_redundant_paths (:157-174): same number of nodes and same node sequence after
// orienting each so that first >= last; first occurrence wins.
void dedupe(std::vector<HostPath>& v) {
    if (v.size() == 1) return;
    std::set<std::vector<int32_t>> seen;
    std::vector<HostPath> out;
    out.reserve(v.size());
    for (auto& p : v) {
        std::vector<int32_t> key = p.nodes;
        if (key.front() < key.back()) std::reverse(key.begin(), key.end());
        if (seen.insert(key).second) out.push_back(std::move(p));
    }
    v.swap(out);
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// graph + APSP state kept between wisp_paths_prepare and any number of wisp_paths_query calls
struct PathGraphState {
    int N = 0, precision = 64;
    std::vector<double> w;                       // host copy (LR path lengths are summed from it)
    std::vector<int32_t> adj_ptr, adj_idx;
    std::vector<double> adj_w;
    double max_simple_len = 0.0, apsp_ms = 0.0;
    void *d_w = nullptr, *d_dist = nullptr, *d_pred = nullptr, *d_adj_ptr = nullptr, *d_adj_idx = nullptr,
         *d_adj_w = nullptr;
};
static void free_path_graph(PathGraphState* g) { delete g; }

extern "C" int wisp_paths_query(wisp_ctx* ctx, const int32_t* sources, int n_sources,
                                const int32_t* sinks, int n_sinks, int number_of_paths, int flags);

extern "C" int wisp_paths_prepare(wisp_ctx* ctx, const double* w, int N, int flags) {
    WISP_CHECK(ctx && w && N > 1, "paths_prepare: bad arguments");
    WISP_CUDA(cudaSetDevice(ctx->device));
    const int precision = (flags & 2) ? 32 : 64;
    cudaStream_t st = ctx->stream;
    if (ctx->pg) { free_path_graph(ctx->pg); ctx->pg = nullptr; }
    PathGraphState* g = new PathGraphState();
    ctx->pg = g;
    ctx->pg_free = free_path_graph;
    g->N = N;
    g->precision = precision;
    g->w.assign(w, w + (size_t)N * N);
    // ---- graph: non-zero entry <=> edge (networkx 1.x Graph(data=W), :20) ----
    std::vector<int32_t>& adj_ptr = g->adj_ptr;
    std::vector<int32_t>& adj_idx = g->adj_idx;
    std::vector<double>& adj_w = g->adj_w;
    adj_ptr.assign(N + 1, 0);
    for (int u = 0; u < N; ++u) {
        for (int v = 0; v < N; ++v) {
            const double x = w[(size_t)u * N + v];
            WISP_CHECK(x == x, "get_paths: NaN weight at (%d,%d)", u, v);
            if (x != 0.0 && u != v) {
                WISP_CHECK(x > -1e-12, "get_paths: negative weight %g at (%d,%d)", x, u, v);
                adj_idx.push_back(v);
                adj_w.push_back(x);
            }
        }
        adj_ptr[u + 1] = (int32_t)adj_idx.size();
    }
    const size_t E = adj_idx.size();
    // upper bound on the length of any finite simple path: every node contributes at most
    // its heaviest finite edge.  Past it, no new path can ever appear.
    for (int u = 0; u < N; ++u) {
        double m = 0.0;
        for (int e = adj_ptr[u]; e < adj_ptr[u + 1]; ++e)
            if (adj_w[e] < 1e300 && adj_w[e] > m) m = adj_w[e];
        g->max_simple_len += m;
    }
    // ---- APSP on the device (K5) ----
    double t0 = now_ms();
    WISP_TRY(wisp_scratch(ctx, SCR_MISC2, (size_t)N * N * 8, &g->d_w));
    WISP_TRY(wisp_scratch(ctx, SCR_MISC3, (size_t)N * N * 8, &g->d_dist));
    WISP_TRY(wisp_scratch(ctx, SCR_MISC4, (size_t)N * N * 4, &g->d_pred));
    WISP_CUDA(cudaMemcpyAsync(g->d_w, w, (size_t)N * N * 8, cudaMemcpyHostToDevice, st));
    WISP_TRY(launch_apsp(ctx, (const double*)g->d_w, N, precision, (double*)g->d_dist, (int32_t*)g->d_pred, st));
    WISP_TRY(wisp_scratch(ctx, SCR_MISC5, round_up((size_t)(N + 1) * 4, 16) + E * 4 + 64, &g->d_adj_ptr));
    g->d_adj_idx = (char*)g->d_adj_ptr + round_up((size_t)(N + 1) * 4, 16);
    WISP_TRY(wisp_scratch(ctx, SCR_MISC6, E * 8 + 16, &g->d_adj_w));
    WISP_CUDA(cudaMemcpyAsync(g->d_adj_ptr, adj_ptr.data(), (size_t)(N + 1) * 4, cudaMemcpyHostToDevice, st));
    if (E) {
        WISP_CUDA(cudaMemcpyAsync(g->d_adj_idx, adj_idx.data(), E * 4, cudaMemcpyHostToDevice, st));
        WISP_CUDA(cudaMemcpyAsync(g->d_adj_w, adj_w.data(), E * 8, cudaMemcpyHostToDevice, st));
    }
    WISP_CUDA(cudaStreamSynchronize(st));
    g->apsp_ms = now_ms() - t0;
    return 0;
}

extern "C" int wisp_get_paths(wisp_ctx* ctx, const double* w, int N, const int32_t* sources,
                              int n_sources, const int32_t* sinks, int n_sinks,
                              int number_of_paths, int flags) {
    WISP_TRY(wisp_paths_prepare(ctx, w, N, flags));
    return wisp_paths_query(ctx, sources, n_sources, sinks, n_sinks, number_of_paths, flags);
}

// ---------------------------------------------------------------------------------------
// Enumeration engine: one PASS answers, for every query group, "all simple paths of the group's pairs
// with LR length <= the group's cutoff" -- counted (with early exit beyond the group's cap) or stored.
// All groups of a pass share the level expansion and ONE dfs launch, so a batch of independent
// queries (configs[4]: 200 pairs) keeps the whole GPU busy instead of 2 % of it per query.
// ---------------------------------------------------------------------------------------
namespace {

struct GroupRequest {
    double cutoff = 0.0;
    bool count_only = true;
    int cap = 0;
};

// fetched paths as they come off the device: path k = nodes[off[k] .. off[k] + n[k]) of pair `pair[k]`
struct RawPaths {
    std::vector<double> len;
    std::vector<int32_t> pair, off, n, nodes;
};

struct EnumEngine {
    wisp_ctx* ctx;
    PathGraphState* g;
    int quirk;
    int cap_paths = 4096, cap_front = 16 * kTargetPrefixes, cap_stack = 1024, target_prefixes = kTargetPrefixes;
    int max_len = kMaxLen0;
    double enum_ms = 0.0, host_ms = 0.0;
    unsigned long long expansions_total = 0;

    // pairs [P][2], pair_group [P], req [G].  counts [G] out; paths_per_pair [P] out (fetch groups only),
    // every pair's list sorted by (length, nodes) (:145).
    int pass(const std::vector<int32_t>& pairs, const std::vector<int32_t>& pair_group,
             const std::vector<GroupRequest>& req, std::vector<int64_t>* counts,
             std::vector<std::vector<HostPath>>* paths_per_pair, RawPaths* raw = nullptr) {
        const int n_pairs = (int)pair_group.size(), n_groups = (int)req.size();
        const int N = g->N;
        cudaStream_t st = ctx->stream;
        bool any_fetch = false;
        for (const auto& r : req) any_fetch |= !r.count_only;
        for (;;) {   // retry loop on capacity overflow
            double te = now_ms();
            const int cap_nodes = cap_paths * 24;
            void *d_out, *d_front, *d_small;
            const size_t out_bytes = (size_t)cap_paths * (8 + 4 + 4 + 4) + (size_t)cap_nodes * 4 + 256;
            WISP_TRY(wisp_scratch(ctx, SCR_MISC8, out_bytes, &d_out));
            const size_t front_rec = 8 + 8 + (size_t)max_len * 4;
            WISP_TRY(wisp_scratch(ctx, SCR_MISC9, 2 * (size_t)cap_front * front_rec + 256, &d_front));
            // small per-pass arrays in one block: pairs, pair_group, group params, counters
            const size_t o_pairs = 0, o_pg = round_up((int64_t)n_pairs * 8, 64), o_cut = o_pg + round_up((int64_t)n_pairs * 4, 64),
                         o_bnd = o_cut + round_up((int64_t)n_groups * 8, 64), o_co = o_bnd + round_up((int64_t)n_groups * 8, 64),
                         o_cap = o_co + round_up((int64_t)n_groups * 4, 64), o_gc = o_cap + round_up((int64_t)n_groups * 4, 64),
                         o_cnt = o_gc + round_up((int64_t)n_groups * 4, 64), small_bytes = o_cnt + 128;
            WISP_TRY(wisp_scratch(ctx, SCR_MISC7, small_bytes, &d_small));
            std::vector<char> h_small(small_bytes, 0);
            memcpy(h_small.data() + o_pairs, pairs.data(), (size_t)n_pairs * 8);
            memcpy(h_small.data() + o_pg, pair_group.data(), (size_t)n_pairs * 4);
            for (int q = 0; q < n_groups; ++q) {
                const double c = req[q].cutoff;
                // admissible bound: APSP distances are only bounds (summation order / FP32 mode), so leave
                // slack; the exact float64 left-to-right length decides membership
                const double bnd = c + (g->precision == 32 ? 2e-5 : 1e-9) * std::max(1.0, std::fabs(c));
                ((double*)(h_small.data() + o_cut))[q] = c;
                ((double*)(h_small.data() + o_bnd))[q] = bnd;
                ((int32_t*)(h_small.data() + o_co))[q] = req[q].count_only ? 1 : 0;
                ((int32_t*)(h_small.data() + o_cap))[q] = req[q].cap;
            }
            WISP_CUDA(cudaMemcpyAsync(d_small, h_small.data(), small_bytes, cudaMemcpyHostToDevice, st));
            char* ds = (char*)d_small;
            EnumParams P;
            P.adj_ptr = (const int32_t*)g->d_adj_ptr; P.adj_idx = (const int32_t*)g->d_adj_idx;
            P.adj_w = (const double*)g->d_adj_w; P.dist = (const double*)g->d_dist;
            P.pairs = (const int32_t*)(ds + o_pairs); P.pair_group = (const int32_t*)(ds + o_pg);
            P.g_cutoff = (const double*)(ds + o_cut); P.g_bound = (const double*)(ds + o_bnd);
            P.g_count_only = (const int32_t*)(ds + o_co); P.g_cap = (const int32_t*)(ds + o_cap);
            P.g_count = (int32_t*)(ds + o_gc);
            P.N = N; P.quirk = quirk; P.max_len = max_len;
            int32_t* cnt = (int32_t*)(ds + o_cnt);           // [0..2] out counters, [4] frontier count, [8..9] expansions
            P.out_count = cnt;
            P.expansions = (unsigned long long*)(cnt + 8);
            P.out_len = (double*)d_out;
            P.out_pair = (int32_t*)(P.out_len + cap_paths);
            P.out_off = P.out_pair + cap_paths;
            P.out_n = P.out_off + cap_paths;
            P.out_nodes = P.out_n + cap_paths;
            P.cap_paths = cap_paths; P.cap_nodes = cap_nodes;
            char* fb = (char*)d_front;
            double* f_len[2]; int32_t* f_meta[2]; int32_t* f_nodes[2];
            for (int bq = 0; bq < 2; ++bq) {
                f_len[bq] = (double*)fb; fb += (size_t)cap_front * 8;
                f_meta[bq] = (int32_t*)fb; fb += (size_t)cap_front * 8;
                f_nodes[bq] = (int32_t*)fb; fb += (size_t)cap_front * max_len * 4;
            }
            if (n_pairs > cap_front) { cap_front = 2 * n_pairs; continue; }
            {   // seed frontier: one record [s] per pair
                std::vector<double> hl(n_pairs, 0.0);
                std::vector<int32_t> hm(2 * (size_t)n_pairs), hn((size_t)n_pairs * max_len, 0);
                for (int p = 0; p < n_pairs; ++p) { hm[2 * p] = p; hm[2 * p + 1] = 1; hn[(size_t)p * max_len] = pairs[2 * p]; }
                WISP_CUDA(cudaMemcpyAsync(f_len[0], hl.data(), (size_t)n_pairs * 8, cudaMemcpyHostToDevice, st));
                WISP_CUDA(cudaMemcpyAsync(f_meta[0], hm.data(), (size_t)n_pairs * 8, cudaMemcpyHostToDevice, st));
                WISP_CUDA(cudaMemcpyAsync(f_nodes[0], hn.data(), (size_t)n_pairs * max_len * 4, cudaMemcpyHostToDevice, st));
                WISP_CUDA(cudaStreamSynchronize(st));     // the host vectors above and h_small go out of use
            }
            int n_front = n_pairs, cur = 0;
            bool overflow = false;
            int32_t hc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int level = 0; level < kMaxExpandLevels && n_front > 0 && n_front < target_prefixes; ++level) {
                WISP_CUDA(cudaMemsetAsync(cnt + 4, 0, 4, st));
                expand_level_kernel<<<(n_front + 127) / 128, 128, 0, st>>>(
                    P, f_len[cur], f_meta[cur], f_nodes[cur], n_front, f_len[cur ^ 1], f_meta[cur ^ 1],
                    f_nodes[cur ^ 1], cnt + 4, cap_front);
                ctx->launches++;
                WISP_CUDA(cudaMemcpyAsync(hc, cnt, 32, cudaMemcpyDeviceToHost, st));
                WISP_CUDA(cudaStreamSynchronize(st));
                if (hc[2] & 4) { overflow = true; cap_front *= 4; break; }
                if (hc[2] & 1) { overflow = true; cap_paths *= 4; break; }
                if (hc[2] & 2) { overflow = true; break; }
                n_front = hc[4];
                cur ^= 1;
            }
            if (!overflow && n_front > 0) {
                void* d_stack = nullptr;
                WISP_TRY(wisp_scratch(ctx, SCR_MISC0, (size_t)n_front * cap_stack * sizeof(StackEntry) + 256, &d_stack));
                const int dfs_smem = kDfsWarps * max_len * (int)sizeof(int32_t);
                if (dfs_smem > 48 * 1024)
                    WISP_CUDA(cudaFuncSetAttribute(dfs_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dfs_smem));
                dfs_warp_kernel<<<(n_front + kDfsWarps - 1) / kDfsWarps, kDfsWarps * 32, dfs_smem, st>>>(
                    P, f_len[cur], f_meta[cur], f_nodes[cur], n_front, (StackEntry*)d_stack, cap_stack);
                ctx->launches++;
            }
            unsigned long long hexp = 0;
            std::vector<int32_t> hg(n_groups, 0);
            if (!overflow) {
                WISP_CUDA(cudaMemcpyAsync(hc, cnt, 32, cudaMemcpyDeviceToHost, st));
                WISP_CUDA(cudaMemcpyAsync(&hexp, cnt + 8, 8, cudaMemcpyDeviceToHost, st));
                WISP_CUDA(cudaMemcpyAsync(hg.data(), P.g_count, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, st));
                WISP_CUDA(cudaStreamSynchronize(st));
                WISP_CUDA(cudaGetLastError());
                if (hc[2] & 2) overflow = true;
                if (hc[2] & 1) { overflow = true; cap_paths = std::max(cap_paths * 4, hc[0] + 1024); }
                if (hc[2] & 8) { overflow = true; cap_stack *= 4; }
            }
            if (overflow && (hc[2] & 2)) {
                // a path outgrew the record: a simple path has at most N nodes, so the capacity grows x4 up
                // to N + 1 and the pass is repeated
                WISP_CHECK(max_len <= N, "get_paths: a simple path cannot have more than %d nodes", N);
                max_len = std::min(4 * max_len, N + 1);
                WISP_CHECK((size_t)kDfsWarps * max_len * 4 <= 200 * 1024,
                           "get_paths: paths of more than %d nodes are not supported", 200 * 1024 / (4 * kDfsWarps));
            }
            if (overflow) continue;
            expansions_total += hexp;
            counts->assign(hg.begin(), hg.end());
            if (any_fetch && raw) {
                // batched API: keep the paths packed (no per-path host objects)
                const int np = hc[0], nn = hc[1];
                raw->len.resize(np); raw->pair.resize(np); raw->off.resize(np); raw->n.resize(np); raw->nodes.resize(nn);
                if (np) {
                    WISP_CUDA(cudaMemcpyAsync(raw->len.data(), P.out_len, (size_t)np * 8, cudaMemcpyDeviceToHost, st));
                    WISP_CUDA(cudaMemcpyAsync(raw->pair.data(), P.out_pair, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
                    WISP_CUDA(cudaMemcpyAsync(raw->off.data(), P.out_off, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
                    WISP_CUDA(cudaMemcpyAsync(raw->n.data(), P.out_n, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
                    WISP_CUDA(cudaMemcpyAsync(raw->nodes.data(), P.out_nodes, (size_t)nn * 4, cudaMemcpyDeviceToHost, st));
                    WISP_CUDA(cudaStreamSynchronize(st));
                }
                enum_ms += now_ms() - te;
            } else if (any_fetch && paths_per_pair) {
                const int np = hc[0], nn = hc[1];
                std::vector<double> hl(np);
                std::vector<int32_t> hp(np), ho(np), hn(np), hnodes(nn);
                if (np) {
                    WISP_CUDA(cudaMemcpyAsync(hl.data(), P.out_len, (size_t)np * 8, cudaMemcpyDeviceToHost, st));
                    WISP_CUDA(cudaMemcpyAsync(hp.data(), P.out_pair, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
                    WISP_CUDA(cudaMemcpyAsync(ho.data(), P.out_off, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
                    WISP_CUDA(cudaMemcpyAsync(hn.data(), P.out_n, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
                    WISP_CUDA(cudaMemcpyAsync(hnodes.data(), P.out_nodes, (size_t)nn * 4, cudaMemcpyDeviceToHost, st));
                    WISP_CUDA(cudaStreamSynchronize(st));
                }
                enum_ms += now_ms() - te;
                double th = now_ms();
                paths_per_pair->assign(n_pairs, std::vector<HostPath>());
                for (int k = 0; k < np; ++k) {
                    HostPath hpth;
                    hpth.len = hl[k];
                    hpth.nodes.assign(hnodes.begin() + ho[k], hnodes.begin() + ho[k] + hn[k]);
                    (*paths_per_pair)[hp[k]].push_back(std::move(hpth));
                }
                for (int p = 0; p < n_pairs; ++p) std::sort((*paths_per_pair)[p].begin(), (*paths_per_pair)[p].end(), path_less);
                host_ms += now_ms() - th;
            } else {
                enum_ms += now_ms() - te;
            }
            return 0;
        }
    }
};

// seed of one (source, sink) pair: a shortest path from the APSP predecessors, LR length from the matrix (:30-33)
int seed_path(const PathGraphState* g, const int32_t* pred_row, const double* dist_row, int s, int t,
              std::vector<int32_t>* sp, double* length) {
    const int N = g->N;
    WISP_CHECK(dist_row[t] < 1e300 && pred_row[t] >= 0, "get_paths: no path between node %d and node %d", s, t);
    sp->clear();
    sp->push_back(t);
    while (sp->back() != s) {
        const int pr = pred_row[sp->back()];
        WISP_CHECK(pr >= 0 && (int)sp->size() <= N, "get_paths: broken predecessor chain %d -> %d", s, t);
        sp->push_back(pr);
    }
    std::reverse(sp->begin(), sp->end());
    double len = 0.0;
    for (size_t k = 0; k + 1 < sp->size(); ++k) len += g->w[(size_t)(*sp)[k] * N + (*sp)[k + 1]];
    *length = len;
    return 0;
}

// rows `rows` of (pred, dist) in ONE round trip: [n][N] each
int fetch_rows(wisp_ctx* ctx, const PathGraphState* g, const std::vector<int>& rows, std::vector<int32_t>* pred,
               std::vector<double>* dist) {
    const int N = g->N;
    cudaStream_t st = ctx->stream;
    pred->resize(rows.size() * (size_t)N);
    dist->resize(rows.size() * (size_t)N);
    for (size_t k = 0; k < rows.size(); ++k) {
        WISP_CUDA(cudaMemcpyAsync(pred->data() + k * N, (int32_t*)g->d_pred + (size_t)rows[k] * N, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
        WISP_CUDA(cudaMemcpyAsync(dist->data() + k * N, (double*)g->d_dist + (size_t)rows[k] * N, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    }
    WISP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

void store_result(wisp_ctx* ctx, std::vector<HostPath>& paths) {
    for (auto& p : paths) {
        ctx->path_len.push_back(p.len);
        for (int32_t n : p.nodes) ctx->path_nodes.push_back(n);
        ctx->path_off.push_back((int64_t)ctx->path_nodes.size());
    }
}

}  // namespace

extern "C" int wisp_paths_query(wisp_ctx* ctx, const int32_t* sources, int n_sources,
                                const int32_t* sinks, int n_sinks, int number_of_paths, int flags) {
    WISP_CHECK(ctx && ctx->pg, "paths_query: call wisp_paths_prepare first");
    WISP_CHECK(sources && sinks && n_sources > 0 && n_sinks > 0, "get_paths: empty source or sink list");
    WISP_CHECK(number_of_paths >= 1, "get_paths: number_of_paths must be >= 1");
    WISP_CUDA(cudaSetDevice(ctx->device));
    PathGraphState* g = ctx->pg;
    const int N = g->N;
    for (int i = 0; i < n_sources; ++i) WISP_CHECK(sources[i] >= 0 && sources[i] < N, "get_paths: source %d out of range", sources[i]);
    for (int i = 0; i < n_sinks; ++i) WISP_CHECK(sinks[i] >= 0 && sinks[i] < N, "get_paths: sink %d out of range", sinks[i]);
    const double max_simple_len = g->max_simple_len;

    // ---- pairs in the reference's loop order (:27-29), seed path (:25-36) ----
    std::vector<int32_t> pairs;
    for (int a = 0; a < n_sources; ++a)
        for (int b = 0; b < n_sinks; ++b)
            if (sources[a] != sinks[b]) { pairs.push_back(sources[a]); pairs.push_back(sinks[b]); }
    const int n_pairs = (int)pairs.size() / 2;
    WISP_CHECK(n_pairs > 0, "get_paths: no (source, sink) pair with source != sink");

    std::vector<int> rows;
    std::vector<int> row_of(N, -1);
    for (int p = 0; p < n_pairs; ++p)
        if (row_of[pairs[2 * p]] < 0) { row_of[pairs[2 * p]] = (int)rows.size(); rows.push_back(pairs[2 * p]); }
    std::vector<int32_t> pred_rows;
    std::vector<double> dist_rows;
    WISP_TRY(fetch_rows(ctx, g, rows, &pred_rows, &dist_rows));
    double shortest_length = 99999999.999;
    std::vector<int32_t> shortest_path;
    for (int p = 0; p < n_pairs; ++p) {
        const int s = pairs[2 * p], t = pairs[2 * p + 1];
        std::vector<int32_t> sp;
        double length = 0.0;
        WISP_TRY(seed_path(g, pred_rows.data() + (size_t)row_of[s] * N, dist_rows.data() + (size_t)row_of[s] * N, s, t, &sp, &length));
        if (length < shortest_length) { shortest_length = length; shortest_path = sp; }
    }

    std::vector<HostPath> paths;
    paths.push_back(HostPath{shortest_length, shortest_path});
    int64_t num_paths = 1;
    double cutoff = shortest_length;

    EnumEngine eng{ctx, g, (flags & 1) ? 1 : 0};
    eng.cap_paths = std::max(4096, 8 * number_of_paths * std::max(1, n_pairs));
    eng.cap_front = std::max(16 * kTargetPrefixes, 4 * n_pairs);
    const std::vector<int32_t> pair_group(n_pairs, 0);
    int passes = 0;
    auto run_pass = [&](double pass_cutoff, bool count_only, int cap, int64_t* count_out,
                        std::vector<HostPath>* out_paths) -> int {
        std::vector<GroupRequest> req(1);
        req[0].cutoff = pass_cutoff; req[0].count_only = count_only; req[0].cap = cap;
        std::vector<int64_t> counts;
        std::vector<std::vector<HostPath>> per_pair;
        WISP_TRY(eng.pass(pairs, pair_group, req, &counts, count_only ? nullptr : &per_pair));
        *count_out = counts[0];
        if (!count_only && out_paths)
            for (auto& lst : per_pair)
                for (auto& q : lst) out_paths->push_back(std::move(q));
        return 0;
    };

    // A path and its reverse can both be enumerated only when (s,t) and (t,s) are both pairs, and the
    // same path twice only when a pair occurs twice (a repeated index in sources or sinks); otherwise
    // no two enumerated paths are "redundant" and pass counts need no dedupe.
    bool needs_dedupe = false;
    for (int p = 0; p < n_pairs && !needs_dedupe; ++p)
        for (int q = 0; q < n_pairs; ++q) {
            if (pairs[2 * p] == pairs[2 * q + 1] && pairs[2 * p + 1] == pairs[2 * q]) { needs_dedupe = true; break; }
            if (q != p && pairs[2 * p] == pairs[2 * q] && pairs[2 * p + 1] == pairs[2 * q + 1]) { needs_dedupe = true; break; }
        }

    const int64_t K = number_of_paths;
    const int cap_count = (int)std::min<int64_t>(4 * K + 64, 1 << 30);
    double prev_cutoff = -1.0;     // largest cutoff known to hold fewer than K paths
    while (num_paths < number_of_paths) {                                    // :56
        std::vector<HostPath> temp_paths;
        if (needs_dedupe) {
            // literal: full pass, per-pass dedupe decides the count (:157-179)
            int64_t n = 0;
            WISP_TRY(run_pass(cutoff, false, 0, &n, &temp_paths));
            if (temp_paths.size() != 1) dedupe(temp_paths);
            num_paths = (int64_t)temp_paths.size();
        } else {
            // Every pass's list is a subset of the next one's, and only the K smallest of
            // the last one survive the final sort + truncate.  So: count first (early exit
            // beyond 4K); fetch paths only from the last pass, and when that pass holds far
            // more than K, only those below a length tau (found by bisection on the count)
            // that still contains at least K of them.
            int64_t n = 0;
            WISP_TRY(run_pass(cutoff, true, cap_count, &n, nullptr));
            if (n < K) {
                num_paths = n;
                prev_cutoff = cutoff;
            } else if (n <= cap_count) {
                WISP_TRY(run_pass(cutoff, false, 0, &n, &temp_paths));
                num_paths = (int64_t)temp_paths.size();
            } else {
                double lo = prev_cutoff, hi = cutoff, tau = cutoff;
                bool found = false;
                for (int itb = 0; itb < 60; ++itb) {
                    const double mid = 0.5 * (lo + hi);
                    if (!(mid > lo && mid < hi)) break;
                    int64_t c = 0;
                    WISP_TRY(run_pass(mid, true, cap_count, &c, nullptr));
                    if (c < K) lo = mid;
                    else if (c > cap_count) hi = mid;
                    else { tau = mid; found = true; break; }
                }
                if (!found) tau = hi;      // massive ties: take everything up to hi
                int64_t c = 0;
                WISP_TRY(run_pass(tau, false, 0, &c, &temp_paths));
                num_paths = n;             // > 4K: certainly != K, the final list is truncated
            }
        }
        double th = now_ms();
        for (auto& q : temp_paths) paths.push_back(q);
        cutoff += 0.1;                                                       // :187
        ++passes;
        eng.host_ms += now_ms() - th;
        // the reference never terminates when fewer than K simple paths exist (App. A.6):
        // past the longest possible simple path nothing new can appear -- fail loudly instead.
        WISP_CHECK(cutoff <= max_simple_len + 0.2 || num_paths >= number_of_paths,
                   "get_paths: only %lld simple paths exist, %d requested", (long long)num_paths, number_of_paths);
    }
    double th = now_ms();
    if (paths.size() != 1) dedupe(paths);                                    // :189
    std::sort(paths.begin(), paths.end(), path_less);                        // :209
    if (num_paths != number_of_paths && (int64_t)paths.size() > number_of_paths)  // :211
        paths.resize(number_of_paths);
    eng.host_ms += now_ms() - th;

    ctx->path_len.clear(); ctx->path_off.clear(); ctx->path_nodes.clear(); ctx->path_query_off.clear();
    ctx->path_off.push_back(0);
    store_result(ctx, paths);
    ctx->path_stats[0] = passes; ctx->path_stats[1] = (double)eng.expansions_total;
    ctx->path_stats[2] = cutoff - 0.1; ctx->path_stats[3] = g->apsp_ms;
    ctx->path_stats[4] = eng.enum_ms; ctx->path_stats[5] = eng.host_ms;
    return 0;
}

// Many INDEPENDENT queries on the prepared graph -- query q = get_paths([s_q], [t_q], K) -- answered
// together (configs[4]: 200 pairs x 1,000 paths).  Every query runs the reference's own cutoff ratchet
// (count passes until >= K, bisection when a pass overshoots, one fetch); the passes of all unfinished
// queries are served by the same kernels.  Results: wisp_paths_size / wisp_paths_fetch give the
// concatenation, wisp_paths_query_offsets the first path of every query.
extern "C" int wisp_paths_query_pairs(wisp_ctx* ctx, const int32_t* pairs_in, int n_queries, int number_of_paths,
                                      int flags) {
    WISP_CHECK(ctx && ctx->pg, "paths_query_pairs: call wisp_paths_prepare first");
    WISP_CHECK(pairs_in && n_queries > 0, "paths_query_pairs: empty pair list");
    WISP_CHECK(number_of_paths >= 1, "get_paths: number_of_paths must be >= 1");
    WISP_CUDA(cudaSetDevice(ctx->device));
    PathGraphState* g = ctx->pg;
    const int N = g->N;
    const int64_t K = number_of_paths;
    const int cap_count = (int)std::min<int64_t>(4 * K + 64, 1 << 30);
    for (int q = 0; q < n_queries; ++q) {
        const int s = pairs_in[2 * q], t = pairs_in[2 * q + 1];
        WISP_CHECK(s >= 0 && s < N && t >= 0 && t < N, "get_paths: pair (%d, %d) out of range", s, t);
        WISP_CHECK(s != t, "get_paths: no (source, sink) pair with source != sink");
    }
    // ---- seeds: the needed predecessor / distance rows in one round trip ----
    std::vector<int> rows, row_of(N, -1);
    for (int q = 0; q < n_queries; ++q)
        if (row_of[pairs_in[2 * q]] < 0) { row_of[pairs_in[2 * q]] = (int)rows.size(); rows.push_back(pairs_in[2 * q]); }
    std::vector<int32_t> pred_rows;
    std::vector<double> dist_rows;
    WISP_TRY(fetch_rows(ctx, g, rows, &pred_rows, &dist_rows));
    enum Phase { COUNT, BISECT, FETCH, DONE };
    // A query is one (s, t) pair with s != t, so a pass never holds redundant paths (no reverse pair, no
    // repeated pair) and the reference's ratchet reduces to: count passes, one fetch, then the final
    // dedupe (which can only remove the seed's twin), sort and K / K+1 truncation.  The fetched paths
    // stay packed: sorted index lists instead of per-path host objects.
    struct Query {
        Phase phase = COUNT;
        std::vector<int32_t> seed;
        double seed_len = 0.0;
        int64_t num_paths = 1, n_over = 0;
        double cutoff = 0.0, prev_cutoff = -1.0, lo = 0.0, hi = 0.0, req_cutoff = 0.0;
        int bisect_it = 0, passes = 0;
        // fetched paths (sorted by (length, nodes)): lengths, node offsets, nodes
        std::vector<double> len;
        std::vector<int64_t> off;
        std::vector<int32_t> nodes;
    };
    std::vector<Query> Q(n_queries);
    for (int q = 0; q < n_queries; ++q) {
        const int s = pairs_in[2 * q], t = pairs_in[2 * q + 1];
        double length = 0.0;
        WISP_TRY(seed_path(g, pred_rows.data() + (size_t)row_of[s] * N, dist_rows.data() + (size_t)row_of[s] * N, s, t,
                           &Q[q].seed, &length));
        Q[q].seed_len = length;
        Q[q].cutoff = length;
        Q[q].req_cutoff = length;
        Q[q].off.push_back(0);
        if (Q[q].num_paths >= K) Q[q].phase = DONE;         // K == 1: the seed alone (:56 is false at once)
    }
    EnumEngine eng{ctx, g, (flags & 1) ? 1 : 0};
    eng.target_prefixes = std::min(65536, std::max(kTargetPrefixes, 512 * n_queries));
    eng.cap_front = std::max(16 * eng.target_prefixes, 4 * n_queries);
    eng.cap_paths = std::max(4096, (int)std::min<int64_t>((int64_t)8 * K * n_queries, 1 << 26));
    int rounds = 0;
    for (;;) {
        // one request per unfinished query
        std::vector<int32_t> pairs, pair_group;
        std::vector<GroupRequest> req;
        std::vector<int> who;
        bool any_fetch = false;
        for (int q = 0; q < n_queries; ++q) {
            if (Q[q].phase == DONE) continue;
            GroupRequest r;
            r.cutoff = Q[q].req_cutoff;
            r.count_only = Q[q].phase != FETCH;
            any_fetch |= !r.count_only;
            r.cap = cap_count;
            pairs.push_back(pairs_in[2 * q]); pairs.push_back(pairs_in[2 * q + 1]);
            pair_group.push_back((int32_t)req.size());
            req.push_back(r);
            who.push_back(q);
        }
        if (who.empty()) break;
        std::vector<int64_t> counts;
        RawPaths raw;
        WISP_TRY(eng.pass(pairs, pair_group, req, &counts, nullptr, &raw));
        ++rounds;
        double th = now_ms();
        // fetched paths of this round, grouped by request, each group sorted by (length, nodes) (:145)
        std::vector<std::vector<int32_t>> idx_of(any_fetch ? who.size() : 0);
        if (any_fetch)
            for (int32_t k = 0; k < (int32_t)raw.len.size(); ++k) idx_of[raw.pair[k]].push_back(k);
        for (size_t k = 0; k < who.size(); ++k) {
            Query& y = Q[who[k]];
            const int64_t n = counts[k];
            auto finish_pass = [&]() -> int {       // the tail of the reference's while body (:187) + the guard
                y.cutoff += 0.1;
                ++y.passes;
                WISP_CHECK(y.cutoff <= g->max_simple_len + 0.2 || y.num_paths >= K,
                           "get_paths: only %lld simple paths exist, %d requested", (long long)y.num_paths, number_of_paths);
                if (y.num_paths >= K) y.phase = DONE;
                else { y.phase = COUNT; y.req_cutoff = y.cutoff; }
                return 0;
            };
            if (y.phase == COUNT) {
                if (n < K) {
                    y.num_paths = n;
                    y.prev_cutoff = y.cutoff;
                    WISP_TRY(finish_pass());
                } else if (n <= cap_count) {
                    y.phase = FETCH;
                    y.req_cutoff = y.cutoff;
                } else {
                    y.n_over = n;
                    y.lo = y.prev_cutoff; y.hi = y.cutoff; y.bisect_it = 0;
                    const double mid = 0.5 * (y.lo + y.hi);
                    if (mid > y.lo && mid < y.hi) { y.phase = BISECT; y.req_cutoff = mid; }
                    else { y.phase = FETCH; y.req_cutoff = y.hi; }
                }
            } else if (y.phase == BISECT) {
                const double mid = y.req_cutoff;
                bool found = false;
                if (n < K) y.lo = mid;
                else if (n > cap_count) y.hi = mid;
                else found = true;
                ++y.bisect_it;
                if (found) { y.phase = FETCH; y.req_cutoff = mid; }
                else {
                    const double m2 = 0.5 * (y.lo + y.hi);
                    if (y.bisect_it < 60 && m2 > y.lo && m2 < y.hi) y.req_cutoff = m2;
                    else { y.phase = FETCH; y.req_cutoff = y.hi; }      // massive ties: take everything up to hi
                }
            } else {   // FETCH
                std::vector<int32_t>& idx = idx_of[k];
                std::sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) {
                    if (raw.len[a] != raw.len[b]) return raw.len[a] < raw.len[b];
                    return std::lexicographical_compare(raw.nodes.begin() + raw.off[a], raw.nodes.begin() + raw.off[a] + raw.n[a],
                                                        raw.nodes.begin() + raw.off[b], raw.nodes.begin() + raw.off[b] + raw.n[b]);
                });
                for (int32_t a : idx) {
                    y.len.push_back(raw.len[a]);
                    y.nodes.insert(y.nodes.end(), raw.nodes.begin() + raw.off[a], raw.nodes.begin() + raw.off[a] + raw.n[a]);
                    y.off.push_back((int64_t)y.nodes.size());
                }
                y.num_paths = y.n_over > 0 ? y.n_over : (int64_t)idx.size();
                y.n_over = 0;
                WISP_TRY(finish_pass());
            }
        }
        eng.host_ms += now_ms() - th;
    }
    double th = now_ms();
    ctx->path_len.clear(); ctx->path_off.clear(); ctx->path_nodes.clear(); ctx->path_query_off.clear();
    ctx->path_off.push_back(0);
    int passes = 0;
    for (int q = 0; q < n_queries; ++q) {
        Query& y = Q[q];
        // paths = [seed] + fetched (:186); dedupe (:189) keeps the first of two identical paths -- only the
        // seed can have a twin; sort (:209): the seed takes its place by (length, nodes); truncate (:211)
        const int64_t nf = (int64_t)y.len.size();
        int64_t twin = -1;
        for (int64_t k = 0; k < nf && twin < 0; ++k)
            if (y.off[k + 1] - y.off[k] == (int64_t)y.seed.size() &&
                std::equal(y.seed.begin(), y.seed.end(), y.nodes.begin() + y.off[k])) twin = k;
        auto seed_before = [&](int64_t k) {     // Python list order of [len, nodes...]
            if (y.seed_len != y.len[k]) return y.seed_len < y.len[k];
            return std::lexicographical_compare(y.seed.begin(), y.seed.end(), y.nodes.begin() + y.off[k],
                                                y.nodes.begin() + y.off[k + 1]);
        };
        int64_t total = nf + (twin < 0 ? 1 : 0);
        int64_t keep = total;
        if (y.num_paths != K && total > K) keep = K;
        ctx->path_query_off.push_back((int64_t)ctx->path_len.size());
        bool seed_done = twin >= 0 && false;
        int64_t emitted = 0;
        auto emit = [&](double len, const int32_t* a, const int32_t* b) {
            ctx->path_len.push_back(len);
            ctx->path_nodes.insert(ctx->path_nodes.end(), a, b);
            ctx->path_off.push_back((int64_t)ctx->path_nodes.size());
            ++emitted;
        };
        for (int64_t k = 0; k < nf && emitted < keep; ++k) {
            if (k == twin) {
                // the seed's twin: the reference keeps the seed copy (first occurrence) -- same nodes; its
                // length is the seed's own LR sum
                emit(y.seed_len, y.seed.data(), y.seed.data() + y.seed.size());
                continue;
            }
            if (twin < 0 && !seed_done && seed_before(k)) {
                emit(y.seed_len, y.seed.data(), y.seed.data() + y.seed.size());
                seed_done = true;
                if (emitted >= keep) break;
            }
            emit(y.len[k], y.nodes.data() + y.off[k], y.nodes.data() + y.off[k + 1]);
        }
        if (twin < 0 && !seed_done && emitted < keep) emit(y.seed_len, y.seed.data(), y.seed.data() + y.seed.size());
        passes += y.passes;
    }
    ctx->path_query_off.push_back((int64_t)ctx->path_len.size());
    eng.host_ms += now_ms() - th;
    ctx->path_stats[0] = passes; ctx->path_stats[1] = (double)eng.expansions_total;
    ctx->path_stats[2] = rounds; ctx->path_stats[3] = g->apsp_ms;
    ctx->path_stats[4] = eng.enum_ms; ctx->path_stats[5] = eng.host_ms;
    return 0;
}

extern "C" int wisp_paths_query_offsets(wisp_ctx* ctx, int64_t* n_queries, int64_t* offsets) {
    WISP_CHECK(ctx, "paths_query_offsets: null context");
    const int64_t nq = ctx->path_query_off.empty() ? 0 : (int64_t)ctx->path_query_off.size() - 1;
    if (n_queries) *n_queries = nq;
    if (offsets && nq > 0) memcpy(offsets, ctx->path_query_off.data(), (size_t)(nq + 1) * 8);
    return 0;
}

extern "C" int wisp_paths_size(wisp_ctx* ctx, int64_t* n_paths, int64_t* n_node_entries) {
    WISP_CHECK(ctx, "paths_size: null context");
    if (n_paths) *n_paths = (int64_t)ctx->path_len.size();
    if (n_node_entries) *n_node_entries = (int64_t)ctx->path_nodes.size();
    return 0;
}

extern "C" int wisp_paths_fetch(wisp_ctx* ctx, double* lengths, int64_t* offsets, int32_t* nodes) {
    WISP_CHECK(ctx, "paths_fetch: null context");
    if (lengths) memcpy(lengths, ctx->path_len.data(), ctx->path_len.size() * 8);
    if (offsets) memcpy(offsets, ctx->path_off.data(), ctx->path_off.size() * 8);
    if (nodes) memcpy(nodes, ctx->path_nodes.data(), ctx->path_nodes.size() * 4);
    return 0;
}

extern "C" int wisp_paths_stats(wisp_ctx* ctx, double* out6) {
    WISP_CHECK(ctx && out6, "paths_stats: null argument");
    memcpy(out6, ctx->path_stats, sizeof(ctx->path_stats));
    return 0;
}
