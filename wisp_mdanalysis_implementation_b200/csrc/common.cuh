// common.cuh -- shared declarations for libwisp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <vector>
#include <string>
#include <functional>

#include "../../include/wisp_b200.h"

#define WISP_ABI_VERSION 1

// ---- layout constants --------------------------------------------------------------------
// Node tile shared by K2 (Gram) / K3 (contact) / K4 (epilogue): 128 nodes.
constexpr int kTile = 128;
// K2 pipeline: one stage holds 24 k-values for both operands (8 frames x 3 components in
// node-trace mode, 24 frames in plain X^T X mode).  Frames are padded to a multiple of 24.
constexpr int kFramePad = 24;

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---- error handling ----------------------------------------------------------------------
void wisp_set_error(const char* fmt, ...);

#define WISP_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            wisp_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,            \
                           cudaGetErrorString(e__));                                     \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

#define WISP_CHECK(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            wisp_set_error(__VA_ARGS__);                                                 \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

#define WISP_TRY(call)                                                                   \
    do {                                                                                 \
        int r__ = (call);                                                                \
        if (r__ != 0) return r__;                                                        \
    } while (0)

// ---- context -----------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct PathGraphState;

struct wisp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;   // compute
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;   // high-priority side stream (K5 look-ahead), created on first use
    cudaEvent_t aux_events[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t launches = 0;
    // multi-GPU (wisp_ctx_create_multi): this context is shard 0; one child context per further GPU
    std::vector<wisp_ctx*> children;
    std::vector<int> peer_devices;           // device id of every shard, in shard order
    std::vector<int> peer_ok;                // [R][R] peer access enabled from shard p to shard q
    int64_t launches_reported = 0;           // (children) launches already added to the root's counter
    int shard_index = 0;                     // position of this context among the shards
    std::vector<int> ingest_via;             // (root) shard -> shard whose PCIe link carries its coordinates
    std::vector<double> ingest_probe;        // (root) what the topology probe measured, for reporting
    cudaStream_t relay_stream = nullptr;     // stream on the relaying GPU that feeds this shard
    double* align_host = nullptr;            // page-locked [align_host_cap][3]: per-iteration K7 residuals
    int align_host_cap = 0;
    int gram_mode = -1;              // WISP_GRAM_*; -1 = read $WISP_K2_MODE on first use
    // column statistics left by the last centring pass (launch_center) for the INT8 K2 engine:
    // valid for ONE following launch_gram on the same block and stream
    const void* i8_stats_x = nullptr;
    int64_t i8_stats_T = 0;
    int i8_stats_parts = 0;
    cudaStream_t i8_stats_stream = nullptr;
    int64_t i8_stats_launch = -1;
    // grow-only scratch buffers, keyed by slot
    DevBuf scratch[96];
    // get_paths results
    std::vector<double> path_len;
    std::vector<int64_t> path_off;
    std::vector<int32_t> path_nodes;
    std::vector<int64_t> path_query_off;    // wisp_paths_query_pairs: first path of every query (+ end)
    double path_stats[6] = {0, 0, 0, 0, 0, 0};
    double pipe_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    PathGraphState* pg = nullptr;            // graph + APSP state of wisp_paths_prepare
    void (*pg_free)(PathGraphState*) = nullptr;
};

int wisp_scratch(wisp_ctx* ctx, int slot, size_t bytes, void** out);   // grow-only device buffer
static inline cudaStream_t wisp_stream(wisp_ctx* ctx, void* s) {
    return s ? (cudaStream_t)s : ctx->stream;
}

enum ScratchSlot {
    SCR_GRAM_PART = 0, SCR_CONTACT_PART, SCR_COLSUM_PART, SCR_X, SCR_COORDS0, SCR_COORDS1,
    SCR_APSP_D, SCR_APSP_P, SCR_ARENA, SCR_ALIGN, SCR_P32,
    SCR_MISC0, SCR_MISC1, SCR_MISC2, SCR_MISC3, SCR_MISC4, SCR_MISC5, SCR_MISC6, SCR_MISC7,
    SCR_MISC8, SCR_MISC9, SCR_I8_PLANES, SCR_I8_SMALL, SCR_I8_PART, SCR_I8_STATS, SCR_APSP_R, SCR_APSP_T, SCR_APSP_M,
    SCR_ALIGN_ROT, SCR_ALIGN_ROTTOT, SCR_ALIGN_PART, SCR_ALIGN_AVG, SCR_COM_COLSUM, SCR_PIPE_G, SCR_PIPE_D, SCR_PIPE_OUT, SCR_PIPE_STAGE0, SCR_PIPE_STAGE_LAST = SCR_PIPE_STAGE0 + 15,
    SCR_RELAY0, SCR_RELAY_LAST = SCR_RELAY0 + 15, SCR_COUNT
};
static_assert(SCR_COUNT <= 96, "wisp_ctx::scratch is too small for the slot list");

// ---- kernel launchers (one per .cu) -------------------------------------------------------
int launch_com(wisp_ctx* ctx, const float* d_coords, int64_t T, int A, const double* d_masses,
               const int32_t* d_node_ptr, const int32_t* d_atom_idx, int N,
               const int32_t* d_align_idx, int n_align, int atomic, double* d_nodes, int64_t ld,
               int64_t row0, cudaStream_t st, float* d_align_out = nullptr);
struct wisp_com_plan;
int com_plan_create(wisp_ctx* ctx, int A, const double* masses, const int32_t* node_ptr,
                    const int32_t* atom_idx, int N, const int32_t* align_idx, int n_align, int atomic,
                    wisp_com_plan** out);
void com_plan_destroy(wisp_ctx* ctx, wisp_com_plan* pl);
int com_plan_info(const wisp_com_plan* pl, int* streaming, int* n_pieces);
// d_colsum != nullptr: also the column sums of the T rows just written (fused into the streaming
// kernel; the gather fallback runs the separate column-sum pass)
int launch_com_plan(wisp_ctx* ctx, wisp_com_plan* pl, const float* d_coords, int64_t T, double* d_nodes,
                    int64_t ld, int64_t row0, cudaStream_t st, float* d_align_out = nullptr,
                    double* d_colsum = nullptr);
// out[col] = sum over n_rows rows of part[row * row_stride + col], fixed order
int launch_colsum_finish(wisp_ctx* ctx, const double* d_part, int n_rows, int64_t row_stride, int64_t ncols,
                         double* d_colsum, cudaStream_t st);
int launch_colsum_f32(wisp_ctx* ctx, const float* d_x, int64_t T, int64_t ld, int64_t ncols,
                      double* d_colsum, cudaStream_t st);
// Cross-shard combination of small host vectors during the alignment.  f32_chain == false: buf <- sum
// of buf over the frame shards, in shard order.  f32_chain == true: the shards run `step` one after the
// other in frame order, each on the running totals left by its predecessor (zeros for the first), and
// buf ends as the grand totals on every shard.
typedef std::function<int(double* h_inout)> AlignStep;
typedef std::function<int(double* buf, size_t n, bool f32_chain, const AlignStep& step)> AlignReduce;
int run_iterative_alignment(wisp_ctx* ctx, float* d_align, int n_align, double* d_nodes, int64_t ld,
                            int N, int64_t T, double threshold, int max_iter, double* d_work,
                            double* h_avg_nodes, double* log, int* n_iter_out, cudaStream_t st,
                            int64_t T_total = 0, const AlignReduce& reduce = AlignReduce(),
                            const double* d_node_colsum = nullptr, const double** d_rot_out = nullptr);
int launch_align_iter(wisp_ctx* ctx, float* d_align, int n_align, const double* d_avg_align, const double* d_nodes,
                      int64_t ld, int N, int64_t T, double* d_rot_total, bool first, cudaStream_t st,
                      double* d_sums = nullptr, const int* d_gate = nullptr);
int launch_align_nodes(wisp_ctx* ctx, double* d_nodes, int64_t ld, int N, int64_t T, const double* d_rot_total,
                       cudaStream_t st);
int launch_align_sums(wisp_ctx* ctx, const float* d_align, int n_align, const double* d_nodes, int64_t ld,
                      int N, int64_t T, bool first, double* d_sums, cudaStream_t st);
int launch_align_update(wisp_ctx* ctx, const double* d_sums, int64_t T_total, int n_align, int N, double* d_avg,
                        double* d_res2, cudaStream_t st, double threshold = -1.0, int* d_gate = nullptr);
int launch_align_rotate(wisp_ctx* ctx, float* d_align, int n_align, const double* d_avg_align, double* d_nodes,
                        int64_t ld, int N, int64_t T, cudaStream_t st, double* d_sums = nullptr);
int launch_colsum(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int64_t ncols,
                  double* d_colsum, cudaStream_t st);
// d_p32 != nullptr: also write the FP32 tile-referenced copy [T][n_tiles][3][128] for K3.
// d_rot != nullptr ([T][9], needs N): every frame is first rotated by its matrix (the node rotation the
// alignment loop deferred, k7_align.cu) -- x <- R[t] x - mean in the same pass.
int launch_center(wisp_ctx* ctx, double* d_x, int64_t T, int64_t ld, int64_t ncols,
                  const double* d_mean, cudaStream_t st, float* d_p32 = nullptr, int N = 0,
                  const double* d_rot = nullptr, bool stats_without_p32 = false);
int launch_p32_convert(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int N, const double* d_ref,
                       float* d_p32, cudaStream_t st);
int launch_gram(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int n_out, int stride,
                double* d_out, cudaStream_t st);
int launch_gram_dmma(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int n_out, int stride,
                     double* d_out, cudaStream_t st, const int* gate, int gate_want);
// *gate_out (device int, valid until the next call): 0 = the INT8 result stands, 1 = accuracy
// bound not met for this data and nothing was written.  gate_out == NULL: no bound check.
int launch_gram_i8(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int N, double* d_out,
                   cudaStream_t st, const int** gate_out);
enum { WISP_GRAM_AUTO = 0, WISP_GRAM_DMMA = 1, WISP_GRAM_I8 = 2 };
constexpr int64_t kGramI8AutoRows = 12288;      // auto mode: INT8 path from 4096 frames up
int wisp_gram_mode(wisp_ctx* ctx);
int launch_reduce_partials(wisp_ctx* ctx, const double* part, int n_seg, int64_t n_pad,
                           double* out, int tile_r, int tile_c, cudaStream_t st,
                           const int* gate = nullptr, int gate_want = 0);
int wisp_pick_segments(int items_per_seg, int stages_total, int sm_count, int min_stages);
// accumulate: d_out += the sums of this frame block (one segment per tile, no partial buffer) instead of
// overwriting it -- the streamed pipeline calls K3 once per coordinate block
int launch_contact_sum(wisp_ctx* ctx, const float* d_p32, int64_t T, int N, const double* d_mean,
                       double* d_out, cudaStream_t st, bool accumulate = false);
int launch_epilogue(wisp_ctx* ctx, const double* d_gram, const double* d_dsum, int64_t T_total,
                    int N, double cutoff, int which_map, double* d_var, double* d_ncov,
                    double* d_corr, double* d_avgdist, int64_t* d_binary, double* d_w,
                    cudaStream_t st, const double* d_offs = nullptr);
// cross-GPU reduction fused into the epilogue: entry (i, j) of the Gram / distance sums is the sum,
// in shard order, of the R partial matrices addressed by `ps` (own HBM or peers' HBM over NVLink)
constexpr int kMaxPeers = 8;
struct PeerSums {
    const double* gram[kMaxPeers];
    const double* dsum[kMaxPeers];
    int n;
};
// rows [row0, row1) of the dense outputs, each [row1 - row0][N]; d_diag [N] scratch
int launch_epilogue_peers(wisp_ctx* ctx, const PeerSums& ps, int64_t T_total, int N, int row0, int row1,
                          double cutoff, int which_map, double* d_diag, double* d_corr, double* d_avgdist,
                          int64_t* d_binary, double* d_w, cudaStream_t st);
// called on the compute stream's behalf after K1 of every coordinate block: node rows [t0, t0 + nt) of d_x
typedef std::function<int(double* d_x, int64_t ld, int64_t t0, int64_t nt)> ComChunkHook;
int wisp_com_from_host(wisp_ctx* ctx, const float* coords, int64_t T, int A, const double* masses,
                       const int32_t* node_ptr, const int32_t* atom_idx, int N, const int32_t* align_idx,
                       int n_align, int atomic, double** d_x_out, int64_t* ld_out, float* d_align_out,
                       wisp_ctx* via = nullptr, const double** d_colsum_out = nullptr,
                       const ComChunkHook* on_chunk = nullptr);
int wisp_pipeline_run(wisp_ctx* ctx, const float* coords, int64_t T, int A, const double* masses,
                      const int32_t* node_ptr, const int32_t* atom_idx, int N, const int32_t* align_idx,
                      int n_align, int atomic, double align_threshold, int max_align_iterations,
                      double cutoff, int which_map, double* out_avg, double* out_corr, double* out_avgdist,
                      int64_t* out_binary, double* out_w, double* out_nodes, double* out_log, int* out_n_iter);
int launch_cov_finalize(wisp_ctx* ctx, const double* d_p, int64_t ldp, int n3, int64_t T,
                        const double* d_avg, const double* d_delta, double* d_cov,
                        cudaStream_t st);
int launch_pearson_from_cov(wisp_ctx* ctx, const double* d_cov, int N, double* d_var,
                            double* d_ncov, double* d_corr, cudaStream_t st);
int launch_func_pearson(wisp_ctx* ctx, const double* d_corr, int N, const double* d_map,
                        double* d_w, cudaStream_t st);
int launch_apsp(wisp_ctx* ctx, const double* d_w, int N, int precision, double* d_dist,
                int32_t* d_pred, cudaStream_t st);
