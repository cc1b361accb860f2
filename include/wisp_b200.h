/* wisp_b200.h -- C ABI of the B200-native WISP hot path (libwisp_b200.so).
 *
 * Plain pointers and sizes only; no torch / C++ types cross this boundary.  The reference
 * (rbdavid/WISP_mdanalysis_implementation) is pure Python and has no FFI of its own, so
 * each entry point cites the reference *function* it replaces (file:line); the Python
 * stage modules in wisp_mdanalysis_implementation_b200/ bind these with ctypes and keep
 * the reference's function names and signatures (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; wisp_last_error() gives
 *     the message of the last failure on the calling thread;
 *   - "host" pointers may be pageable or pinned (pinned makes the copies asynchronous);
 *     "dev" pointers are CUDA device pointers owned by the caller (e.g. torch tensors);
 *   - matrices are row-major; N = nodes, T = frames, A = atoms;
 *   - the node trajectory lives on the device as double [Tpad][ld] with
 *     ld = wisp_node_ld(N) columns (3*N rounded up to whole 128-node tiles, zero padded)
 *     and Tpad = wisp_frames_pad(T) rows (zero padded);
 *   - `stream` is a cudaStream_t passed as void* (NULL = the context's own stream).
 */
#ifndef WISP_B200_H
#define WISP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wisp_ctx wisp_ctx;

/* ---- context ------------------------------------------------------------------------ */
int         wisp_ctx_create(int device, wisp_ctx** out);
/* One context over n_gpus GPUs of this box (devices[] or, when NULL, 0..n_gpus-1; n_gpus = 0: every
 * visible GPU).  wisp_pipeline on such a context shards the frames over the GPUs (one host thread,
 * stream and PCIe link per GPU) and combines the partial sums over NVLink peer memory inside its
 * epilogue kernel -- the caller makes the same single call as on one GPU.  Every other entry point
 * uses the first device. */
int         wisp_ctx_create_multi(int n_gpus, const int* devices, wisp_ctx** out);
int         wisp_ctx_gpus(wisp_ctx* ctx);
/* How the coordinates reach the GPUs of a multi-GPU context.  wisp_ctx_create_multi measures the
 * concurrent host->device rate of every GPU (4+ distinct GPUs; $WISP_INGEST = auto | direct | relay); where a
 * group of GPUs shares a slow upstream link and pulling through the fast GPUs alone moves more bytes per
 * second, each slow GPU is fed by a fast partner: host -> partner HBM -> NVLink peer copy.
 * via [n_gpus]: relaying GPU (shard index) of every shard (itself = direct); rates [2 n_gpus + 2]: GB/s per
 * GPU with all GPUs pulling, with only the fast set pulling, then the two aggregates (zeros: no probe). */
int         wisp_ctx_ingest_info(wisp_ctx* ctx, int* via, double* rates);
/* Page-lock a caller-owned host range (e.g. the MemoryReader coordinate array) so that the chunked
 * H2D copies of wisp_pipeline / wisp_com_nodes run asynchronously at full PCIe rate from every GPU. */
int         wisp_host_register(void* ptr, size_t bytes);
int         wisp_host_unregister(void* ptr);
void        wisp_ctx_destroy(wisp_ctx* ctx);
const char* wisp_last_error(void);
int         wisp_abi_version(void);
int         wisp_sync(wisp_ctx* ctx);
/* layout helpers */
int64_t     wisp_node_ld(int n_nodes);        /* columns of the device node trajectory   */
int64_t     wisp_frames_pad(int64_t n_frames);/* rows of the device node trajectory       */
int64_t     wisp_gram_ld(int n_cols);         /* leading dimension of Gram / sum buffers  */
/* number of launches of this library's kernels since the context was created */
int64_t     wisp_kernel_launches(wisp_ctx* ctx);
/* K2 engine for the node Gram: 0 = auto (INT8 tensor cores from 4096 frames up, guarded by an
 * on-device accuracy bound with DMMA fallback), 1 = always the FP64 DMMA kernel, 2 = INT8
 * whenever the bound holds.  Initial value: $WISP_K2_MODE = auto | dmma | i8. */
int         wisp_set_gram_mode(wisp_ctx* ctx, int mode);

/* ---- host-buffer entry points (what the reference-facing Python modules call) ------- */

/* K1: per-frame translate by -COM(alignment atoms) and per-node mass-weighted COM.
 * Replaces trajectory_analysis_functions.py:70-79 (inside traj_alignment_and_averaging).
 * coords [T][A][3] float32; masses [A]; CSR node map node_ptr[N+1]/atom_idx[];
 * align_idx[n_align].  atomic != 0 selects node_definition 'ATOMIC' (:78).
 * out_nodes [T][N][3] float64, out_avg [N][3] = (sum_t nodes)/T (:87). */
int wisp_com_nodes(wisp_ctx* ctx, const float* coords, int64_t T, int A,
                   const double* masses, const int32_t* node_ptr, const int32_t* atom_idx,
                   int N, const int32_t* align_idx, int n_align, int atomic,
                   double* out_nodes, double* out_avg);

/* a1 + alignment: traj_alignment_and_averaging(...) -- trajectory_analysis_functions.py:39-138.
 * K1 as above, then the iterative alignment of every frame onto the running average of the
 * alignment landmarks (:86-126) until the landmark-average RMSD between two
 * iterations is <= threshold or max_iter iterations ran (max_iter = 0: COM extraction only).
 * out_nodes [T][N][3] aligned node trajectory, out_avg [N][3] its frame mean,
 * out_log [3*max_iter] = (iteration, landmark RMSD, node RMSD) per iteration (may be NULL). */
int wisp_traj_alignment_and_averaging(wisp_ctx* ctx, const float* coords, int64_t T, int A,
                                      const double* masses, const int32_t* node_ptr,
                                      const int32_t* atom_idx, int N, const int32_t* align_idx,
                                      int n_align, int atomic, double threshold, int max_iter,
                                      double* out_nodes, double* out_avg, double* out_log,
                                      int* out_n_iter);

/* a2: cartesian_covariance(node_trajectory, avg_node_positions, ...) --
 * trajectory_analysis_functions.py:141-184.  nodes [T][N][3], avg [N][3];
 * out_cov [3N][3N] = (1/T) sum_t x x^T - avg avg^T. */
int wisp_cartesian_covariance(wisp_ctx* ctx, const double* nodes, int64_t T, int N,
                              const double* avg, double* out_cov);

/* a3: pearson_correlation_analysis(cov, ...) -- adjacency_matrix_functions.py:8-59.
 * cov [3N][3N]; outputs node_variance [N], node_covariance [N][N], node_correlation [N][N]
 * (any output pointer may be NULL). */
int wisp_pearson_from_covariance(wisp_ctx* ctx, const double* cov, int N, double* out_var,
                                 double* out_ncov, double* out_corr);

/* a4: calc_contact_map(trajectory_data, ..., distance_cutoff) --
 * trajectory_analysis_functions.py:187-236.  out_binary [N][N] int64 (numpy int),
 * out_avgdist [N][N] float64. */
int wisp_contact_map(wisp_ctx* ctx, const double* nodes, int64_t T, int N, double cutoff,
                     int64_t* out_binary, double* out_avgdist);

/* a5: func_pearson_correlation(adjacency, ..., contact_map) --
 * functionalize_adj_matrix_functions.py:10-24.  map may be NULL; out_w [N][N]. */
int wisp_func_pearson(wisp_ctx* ctx, const double* corr, int N, const double* contact_map,
                      double* out_w);

/* 8f row 4: linear_mutual_information_analysis(cov, ...) -- adjacency_matrix_functions.py:62-103.
 * cov [3N][3N] -> out_mi [N][N] (+inf on the diagonal). */
int wisp_lmi_from_covariance(wisp_ctx* ctx, const double* cov, int N, double* out_mi);
/* generalized_correlation_coefficient_calc(MI, ..., contact_map, Lambda) --
 * functionalize_adj_matrix_functions.py:27-46.  contact_map may be NULL; the damping is applied
 * only when both contact_map and has_lambda are given.  out_rmi / out_func [N][N] (either may be NULL). */
int wisp_func_generalized_correlation(wisp_ctx* ctx, const double* mi, int N, const double* contact_map,
                                      double lambda, int has_lambda, double* out_rmi, double* out_func);

/* Fused a1..a5: coordinates -> node COMs -> (iterative alignment, max_align_iterations > 0) ->
 * centred Gram + mean distances -> correlation, contact maps, cost matrix, with the node
 * trajectory never leaving the device and the 3N x 3N covariance never materialised.
 * which_map: 0 none, 1 binary, 2 average-distance (wisp_w_mdanalysis.py:98-101,141-144).
 * Any output pointer may be NULL.  out_nodes [T][N][3] is optional (large). */
int wisp_pipeline(wisp_ctx* ctx, const float* coords, int64_t T, int A, const double* masses,
                  const int32_t* node_ptr, const int32_t* atom_idx, int N,
                  const int32_t* align_idx, int n_align, int atomic, double align_threshold,
                  int max_align_iterations, double cutoff, int which_map, double* out_avg,
                  double* out_corr, double* out_avgdist, int64_t* out_binary, double* out_w,
                  double* out_nodes);

/* The same call, also returning the alignment log: out_log [3 * max_align_iterations] =
 * (iteration, landmark RMSD, node RMSD) per iteration, *out_n_iter = iterations run (may be NULL). */
int wisp_pipeline_log(wisp_ctx* ctx, const float* coords, int64_t T, int A, const double* masses,
                      const int32_t* node_ptr, const int32_t* atom_idx, int N,
                      const int32_t* align_idx, int n_align, int atomic, double align_threshold,
                      int max_align_iterations, double cutoff, int which_map, double* out_avg,
                      double* out_corr, double* out_avgdist, int64_t* out_binary, double* out_w,
                      double* out_nodes, double* out_log, int* out_n_iter);

/* Host-clock stage times of the last wisp_pipeline on this context, as seen by the first GPU's thread:
 * [0] H2D || K1 || K3 (pair distances are invariant under the alignment, so K3 runs block by block under the
 * transfer), [1] K7 alignment, [2] centring + K2, [3] waiting for the slowest GPU,
 * [4] fused cross-GPU epilogue + D2H, [5] total (ms); [7] GPUs used. */
int wisp_pipeline_stats(wisp_ctx* ctx, double* out8);

/* a6 (north-star APSP; the reference calls Dijkstra, network_analysis_functions.py:30,71-72).
 * w [N][N] cost matrix (non-zero = edge).  precision 64 or 32 (distances are returned as
 * float64 either way).  out_dist [N][N], out_pred [N][N] (-1 = none); pred may be NULL. */
int wisp_apsp(wisp_ctx* ctx, const double* w, int N, int precision, double* out_dist,
              int32_t* out_pred);

/* a6-a9: get_paths(adjacency_matrix, sources, sinks, number_of_paths) --
 * network_analysis_functions.py:12-214 (Appendix A of SURVEY.md).  flags bit0: keep the
 * reference's node-0 quirk (default on), bit1: APSP in float32.  Results are kept in the
 * context; fetch them with wisp_paths_size / wisp_paths_fetch. */
int wisp_get_paths(wisp_ctx* ctx, const double* w, int N, const int32_t* sources, int n_sources,
                   const int32_t* sinks, int n_sinks, int number_of_paths, int flags);
/* The same in two steps, for many queries on one graph (configs[4]: 200 pairs x 1,000 paths):
 * prepare = upload + APSP + adjacency, query = seed + cutoff passes for one source/sink set. */
int wisp_paths_prepare(wisp_ctx* ctx, const double* w, int N, int flags);
int wisp_paths_query(wisp_ctx* ctx, const int32_t* sources, int n_sources, const int32_t* sinks,
                     int n_sinks, int number_of_paths, int flags);
/* Many independent queries on the prepared graph, answered together: query q =
 * get_paths([pairs[q][0]], [pairs[q][1]], number_of_paths).  The results are concatenated in query
 * order (wisp_paths_size / wisp_paths_fetch); wisp_paths_query_offsets gives the index of the first
 * path of every query: offsets [n_queries + 1] (pass NULL to get only the count). */
int wisp_paths_query_pairs(wisp_ctx* ctx, const int32_t* pairs /* [n_queries][2] */, int n_queries,
                           int number_of_paths, int flags);
int wisp_paths_query_offsets(wisp_ctx* ctx, int64_t* n_queries, int64_t* offsets);
int wisp_paths_size(wisp_ctx* ctx, int64_t* n_paths, int64_t* n_node_entries);
int wisp_paths_fetch(wisp_ctx* ctx, double* lengths, int64_t* offsets /* n_paths+1 */,
                     int32_t* nodes);
/* statistics of the last wisp_get_paths: [0] passes, [1] expansions, [2] final cutoff,
 * [3] apsp ms, [4] enumeration ms, [5] host finalize ms */
int wisp_paths_stats(wisp_ctx* ctx, double* out6);

/* ---- stage-boundary text files (host only) ---------------------------- */
/* Byte-identical to np.savetxt(path, a, fmt='%.18e' / '%i', header=...), formatted on all
 * host threads.  header may be NULL. */
int wisp_savetxt_f64(const char* path, const double* a, int64_t rows, int64_t cols, const char* header);
int wisp_savetxt_i64(const char* path, const int64_t* a, int64_t rows, int64_t cols, const char* header);
/* np.loadtxt for the same files: out == NULL returns only the shape. */
int wisp_loadtxt_f64(const char* path, double* out, int64_t capacity, int64_t* rows, int64_t* cols);

/* ---- device-buffer entry points (bench, multi-GPU orchestration) --------------------- */

/* Node map ("plan") for K1: uploads masses / CSR node map / alignment landmarks once and
 * decides between the TMA-streamed kernel (every node a contiguous ascending atom run) and
 * the general CSR gather kernel.  All arrays are HOST pointers. */
typedef struct wisp_com_plan wisp_com_plan;
int  wisp_com_plan_create(wisp_ctx* ctx, int A, const double* masses, const int32_t* node_ptr,
                          const int32_t* atom_idx, int N, const int32_t* align_idx, int n_align,
                          int atomic, wisp_com_plan** out);
void wisp_com_plan_destroy(wisp_ctx* ctx, wisp_com_plan* plan);
int  wisp_com_plan_info(const wisp_com_plan* plan, int* streaming, int* n_pieces);

/* K1 on device buffers.  d_coords float32 [T][A][3]; d_nodes: double [wisp_frames_pad(T)][ld]
 * (pad rows / columns zeroed by the caller); d_colsum [ld] (may be NULL) receives the sum over
 * frames of every column. */
int wisp_dev_com(wisp_ctx* ctx, wisp_com_plan* plan, const float* d_coords, int64_t T, int N,
                 double* d_nodes, double* d_colsum, void* stream);
/* The same, also writing the translated alignment landmarks (trajectory_analysis_functions.py:72):
 * d_align_out float32 [T][n_align][3] (may be NULL). */
int wisp_dev_com_align(wisp_ctx* ctx, wisp_com_plan* plan, const float* d_coords, int64_t T, int N,
                       double* d_nodes, double* d_colsum, float* d_align_out, void* stream);
/* K7 building blocks for a frame-sharded alignment loop (trajectory_analysis_functions.py:86-126).
 * wisp_dev_align_sums: d_sums [3 n_align + 3 N] = column sums of this rank's landmarks, then of its
 * node rows; first != 0: the landmark sums are the reference's sequential float32 sums (:86).
 * wisp_dev_align_rotate: rotate every frame of this rank onto d_avg_align [3 n_align] in place; d_sums
 * (may be NULL) receives the column sums of the rotated block in the same layout, from the same pass.
 * The caller all-reduces the sums, divides by the total frame count and tests the RMSD (:118). */
int wisp_dev_align_sums(wisp_ctx* ctx, const float* d_align, int n_align, const double* d_nodes,
                        int64_t T, int N, int first, double* d_sums, void* stream);
int wisp_dev_align_rotate(wisp_ctx* ctx, float* d_align, int n_align, const double* d_avg_align,
                          double* d_nodes, int64_t T, int N, double* d_sums, void* stream);
/* One alignment iteration with the node rotation DEFERRED (what the library's own loop uses): the
 * landmarks are rotated onto d_avg_align in place (:106), the frame's rotation is composed into
 * d_rot_total [T][9] (first != 0: d_rot_total <- R, else R d_rot_total), and the node trajectory is only
 * READ: d_sums [3 n_align + 3 N] (may be NULL) = column sums of the rotated landmarks, then of d_rot_total
 * applied to the unchanged nodes (:110-111).  After the loop the caller applies d_rot_total once, either
 * with wisp_dev_align_nodes (in place, :107) or fused into the centring pass (wisp_dev_center_rotated).
 * d_nodes == NULL in wisp_dev_align_sums: only the landmark sums are formed (K1 already gave the node sums).
 * d_gate (may be NULL): a device flag; while it is non-zero every kernel of the iteration is a no-op, so a
 * caller may launch iteration k + 1 before it has read the residual of iteration k (no idle GPU between
 * iterations); wisp_dev_align_update_gated raises it. */
int wisp_dev_align_iter(wisp_ctx* ctx, float* d_align, int n_align, const double* d_avg_align,
                        const double* d_nodes, int64_t T, int N, double* d_rot_total, int first,
                        double* d_sums, const int* d_gate, void* stream);
int wisp_dev_align_nodes(wisp_ctx* ctx, double* d_nodes, int64_t T, int N, const double* d_rot_total,
                         void* stream);
/* After the caller has all-reduced d_sums: d_avg [3 n_align + 3 N] <- d_sums / T_total in place (the new
 * landmark / node averages, :112-113) and d_res2 [2] = sum of squared differences old - new over the
 * landmark columns and over the node columns (the RMSDs of :118-119 are sqrt(res2 / n)). */
int wisp_dev_align_update(wisp_ctx* ctx, const double* d_sums, int64_t T_total, int n_align, int N,
                          double* d_avg, double* d_res2, void* stream);
/* wisp_dev_align_update with the loop decision on the device: nothing happens while *d_gate != 0; otherwise
 * d_res3 [3] = the two squared sums as above and a running count of the iterations done, and *d_gate is set
 * when sqrt(d_res3[0] / n_align) <= threshold (the loop condition of :96). */
int wisp_dev_align_update_gated(wisp_ctx* ctx, const double* d_sums, int64_t T_total, int n_align, int N,
                                double* d_avg, double* d_res3, double threshold, int* d_gate, void* stream);
/* X <- X - mean (in place), mean [ld]; pad columns / rows are left at zero.  When d_p32 is
 * not NULL it also receives the FP32 tile-referenced copy K3 reads:
 * float [T][wisp_gram_ld(N)/128][3][128], value = float(x - mean of the tile's middle node). */
int wisp_dev_center(wisp_ctx* ctx, double* d_nodes, int64_t T, int N, const double* d_mean,
                    float* d_p32, void* stream);
/* The same with every frame first rotated by d_rot_total[t] (row-major 3x3; NULL = wisp_dev_center):
 * X <- R[t] X - mean in one pass (trajectory_analysis_functions.py:107 with the composed rotation). */
int wisp_dev_center_rotated(wisp_ctx* ctx, double* d_nodes, int64_t T, int N, const double* d_mean,
                            const double* d_rot_total, float* d_p32, void* stream);
/* column sums of the (possibly centred) trajectory: d_colsum [ld] */
int wisp_dev_colsum(wisp_ctx* ctx, const double* d_nodes, int64_t T, int N, double* d_colsum,
                    void* stream);
/* K2: node Gram  G[i][j] = sum_t sum_c X[t][3i+c] X[t][3j+c]  (stride 3) or plain
 * X^T X over columns (stride 1), upper-triangular 128-tiles only.  d_out: double
 * [n_out_pad][n_out_pad] with n_out_pad = wisp_gram_ld(n_out); overwritten. */
int wisp_dev_gram(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int n_out,
                  int stride, double* d_out, void* stream);
/* K2 on the INT8 tensor cores (tcgen05, exact byte-plane slicing, FP64-grade result; see
 * csrc/k2_gram_i8.cu): same node Gram as wisp_dev_gram(stride 3).  Same output layout. */
int wisp_dev_gram_i8(wisp_ctx* ctx, const double* d_x, int64_t T, int64_t ld, int N,
                     double* d_out, void* stream);
/* K3: D[i][j] = sum_t |x_i(t) - x_j(t)| for i<=j tiles; d_p32 is the FP32 copy written by
 * wisp_dev_center and d_mean [ld] the mean used there.  d_out [Npad][Npad] overwritten. */
int wisp_dev_contact_sum(wisp_ctx* ctx, const float* d_p32, int64_t T, int N,
                         const double* d_mean, double* d_out, void* stream);
/* K4: fused epilogue on (all-reduced) sums.  d_gram / d_dsum are [Npad][Npad] upper-tile
 * sums over T_total frames.  Outputs (each may be NULL) are dense [N][N]:
 * corr, ncov (=G/T), avgdist, binary (int64), w (= -log|corr| (.) map). d_var [N].
 * d_offset_sum (may be NULL): column sums [3N] of a trajectory centred with a provisional
 * reference instead of its mean; G is then re-centred exactly, G_ij -= (s_i . s_j)/T. */
int wisp_dev_epilogue(wisp_ctx* ctx, const double* d_gram, const double* d_dsum,
                      int64_t T_total, int N, double cutoff, int which_map, double* d_var,
                      double* d_ncov, double* d_corr, double* d_avgdist, int64_t* d_binary,
                      double* d_w, const double* d_offset_sum, void* stream);
/* out = sum over n_parts partial [Npad][Npad] matrices (upper 128-tiles, fixed order). */
int wisp_dev_reduce_partials(wisp_ctx* ctx, const double* d_parts, int n_parts, int N,
                             double* d_out, void* stream);
/* K5 on device: d_w [N][N] cost matrix -> d_dist (float64 [N][N]) and d_pred (int32). */
int wisp_dev_apsp(wisp_ctx* ctx, const double* d_w, int N, int precision, double* d_dist,
                  int32_t* d_pred, void* stream);
/* Row-block sharded FP32 Floyd-Warshall (BASELINE configs[3]): each rank owns whole 128-row tile
 * rows of the distance matrix, float [rows_pad][wisp_gram_ld(N)].  Per round kb the owner of the
 * pivot rows copies them into a [128][wisp_gram_ld(N)] panel, finishes it with
 * wisp_dev_fw32_pivot_panel, the caller broadcasts the panel (NCCL), and every rank applies
 * wisp_dev_fw32_update_rows to its own rows (skip_local_tile = local tile-row index of the pivot
 * rows on the owner, -1 elsewhere; the owner copies the panel back into its rows). */
int wisp_dev_fw32_init_rows(wisp_ctx* ctx, const double* d_w_rows, int N, int64_t row0, int64_t rows_pad,
                            float* d_dist_local, void* stream);
int wisp_dev_fw32_pivot_panel(wisp_ctx* ctx, float* d_panel, int N, int kb, void* stream);
int wisp_dev_fw32_update_rows(wisp_ctx* ctx, float* d_dist_local, int64_t rows_pad, const float* d_panel,
                              int N, int kb, int skip_local_tile, void* stream);
/* The same update split for look-ahead: part 1 = only local tile-row next_local_tile (the NEXT
 * pivot rows, on their owner), part 2 = everything part 1 left out, part 0 = everything.  The owner
 * of the next pivot rows runs part 1, builds and broadcasts the next panel on a second stream, and
 * runs part 2 meanwhile. */
int wisp_dev_fw32_update_rows_part(wisp_ctx* ctx, float* d_dist_local, int64_t rows_pad, const float* d_panel,
                                   int N, int kb, int skip_local_tile, int next_local_tile, int part,
                                   void* stream);

/* synthetic coordinates generated on the device (bench-scale inputs; synth.py model):
 * atom_base [A][3] f64, node_of_atom [A], modes [N][M][3] f64, coeff [T][M] f64. */
int wisp_dev_synth_coords(wisp_ctx* ctx, float* d_coords, int64_t t0, int64_t T, int A,
                          const double* d_atom_base, const int32_t* d_node_of_atom,
                          const double* d_modes, int n_modes, const double* d_coeff,
                          double noise_sigma, uint64_t seed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WISP_B200_H */
