"""CPU restatement of the WISP hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Citations are ``file:line`` into rbdavid/WISP_mdanalysis_implementation (the read-only
reference).  "literal" functions keep the reference's loop structure and float64 summation
order; "fast" functions are result-equivalent vectorised forms used for large parity cases
and as the honest (multi-threaded BLAS) CPU baseline.  Reference defects that stop the code
from running at all are repaired in the minimal way listed in SURVEY.md Appendix B (B6, B7,
B9, B10, B11); behavioural quirks (node-0 rule, K+1 return) are KEPT.

Third-party arithmetic restated here because it is not vendored in the reference:
  * MDAnalysis ``AtomGroup.center_of_mass``  -> sum(m_i x_i)/sum(m_i) in float64
  * MDAnalysis ``AtomGroup.translate(v)``    -> float32 positions += float64 vector
                                                (computed in float64, rounded to float32)
  * MDAnalysis ``distance_array(a, b)``      -> plain Euclidean distance, no box
  * MDAnalysis ``rotation_matrix(a, b)``     -> least-squares rotation taking a onto b
  * networkx 1.x ``Graph(data=A)``           -> every NON-ZERO entry is an edge
  * networkx Dijkstra                        -> any exact shortest path
"""
from __future__ import annotations

import heapq
import numpy as np

__all__ = [
    "center_of_mass", "com_nodes", "com_nodes_fast", "rotation_matrix", "iterative_alignment",
    "traj_alignment_and_averaging",
    "cartesian_covariance", "cartesian_covariance_fast",
    "pearson_correlation_analysis", "pearson_correlation_fast", "node_gram_truth",
    "calc_contact_map", "calc_contact_map_fast", "func_pearson_correlation",
    "graph_edges", "dijkstra", "dijkstra_path", "apsp_floyd_warshall", "apsp_dijkstra",
    "get_paths", "get_paths_pruned", "enumerate_paths_literal", "enumerate_paths_pruned",
    "lr_length", "dedupe_paths",
    "linear_mutual_information_analysis", "generalized_correlation_coefficient_calc",
]


# ----------------------------------------------------------------------------------------
# a1: COM extraction (trajectory_analysis_functions.py:70-79)
# ----------------------------------------------------------------------------------------
def center_of_mass(positions, masses):
    """MDAnalysis AtomGroup.center_of_mass(): float32 positions, float64 masses,
    float64 result (trajectory_analysis_functions.py:71,76 call sites)."""
    p = np.asarray(positions, dtype=np.float64)
    m = np.asarray(masses, dtype=np.float64)
    return np.sum(p * m[:, None], axis=0) / m.sum()


def com_nodes(coords, masses, node_ptr, atom_idx, align_idx, node_definition="COM"):
    """Per-frame translate + per-node COM (trajectory_analysis_functions.py:70-79).

    coords   (T, A, 3) float32 -- the MemoryReader block (frame, atom, xyz)
    masses   (A,) float64
    node_ptr (N+1,), atom_idx (node_ptr[N],) -- CSR form of ``selection_list``
    align_idx (nAlign,) -- atoms of ``alignment_selection``
    Returns (all_pos_Nodes (T,N,3) float64, all_pos_Align (T,nAlign,3) float32).
    """
    coords = np.asarray(coords, dtype=np.float32)
    masses = np.asarray(masses, dtype=np.float64)
    T = coords.shape[0]
    N = len(node_ptr) - 1
    all_pos_Nodes = np.empty((T, N, 3), dtype=np.float64)
    all_pos_Align = np.empty((T, len(align_idx), 3), dtype=np.float32)
    atomic = node_definition.upper() == "ATOMIC"
    for ts in range(T):                                                 # :70
        pos = coords[ts]
        shift = -center_of_mass(pos[align_idx], masses[align_idx])      # :71
        # u_all.translate(shift): float32 array += float64 vector -> computed in float64,
        # stored back as float32
        pos = (pos.astype(np.float64) + shift).astype(np.float32)
        all_pos_Align[ts] = pos[align_idx]                              # :72
        for i in range(N):                                              # :74
            idx = atom_idx[node_ptr[i]:node_ptr[i + 1]]
            if atomic:
                all_pos_Nodes[ts, i] = pos[idx[0]]                      # :78
            else:
                all_pos_Nodes[ts, i] = center_of_mass(pos[idx], masses[idx])   # :76
    return all_pos_Nodes, all_pos_Align


def com_nodes_fast(coords, masses, node_ptr, atom_idx, align_idx, node_definition="COM", block=64):
    """Frame-blocked vectorised form of :func:`com_nodes` (trajectory_analysis_functions.py:70-79)
    for the large parity cases: the same float64 operations in the same order per node
    (sequential sum over the node's atoms, then one division), so the result is bit-identical
    (tests/test_oracle_golden.py::test_com_nodes_fast_is_bit_identical)."""
    coords = np.asarray(coords, dtype=np.float32)
    masses = np.asarray(masses, dtype=np.float64)
    node_ptr = np.asarray(node_ptr, dtype=np.int64)
    atom_idx = np.asarray(atom_idx, dtype=np.int64)
    align_idx = np.asarray(align_idx, dtype=np.int64)
    T = coords.shape[0]
    N = len(node_ptr) - 1
    out = np.empty((T, N, 3), dtype=np.float64)
    out_align = np.empty((T, len(align_idx), 3), dtype=np.float32)
    atomic = node_definition.upper() == "ATOMIC"
    m_al = masses[align_idx]
    m_nodes = masses[atom_idx]
    sizes = np.diff(node_ptr)
    kmax = int(sizes.max())
    for t0 in range(0, T, block):
        pos = coords[t0:t0 + block].astype(np.float64)                          # (b, A, 3)
        pa = pos[:, align_idx, :] * m_al[None, :, None]
        s = np.zeros((pos.shape[0], 3))
        for k in range(pa.shape[1]):                                             # np.sum(axis=0): row by row
            s += pa[:, k, :]
        shift = -(s / m_al.sum())                                               # :71
        pos = (pos + shift[:, None, :]).astype(np.float32)                      # translate, float32 store
        out_align[t0:t0 + block] = pos[:, align_idx, :]                         # :72
        if atomic:
            out[t0:t0 + block] = pos[:, atom_idx[node_ptr[:-1]], :]            # :78
            continue
        acc = np.zeros((pos.shape[0], N, 3))
        msum = np.zeros(N)
        for k in range(kmax):                                                    # k-th atom of every node
            live = sizes > k
            idx = atom_idx[node_ptr[:-1][live] + k]
            mk = m_nodes[node_ptr[:-1][live] + k]
            acc[:, live, :] += pos[:, idx, :].astype(np.float64) * mk[None, :, None]
            msum[live] += mk
        # m.sum() of the literal form is numpy's pairwise sum; recompute it exactly per node
        msum = np.array([masses[atom_idx[node_ptr[i]:node_ptr[i + 1]]].sum() for i in range(N)])
        out[t0:t0 + block] = acc / msum[None, :, None]                          # :76
    return out, out_align


# ----------------------------------------------------------------------------------------
# alignment (trajectory_analysis_functions.py:86-126) -- "next" row f1
# ----------------------------------------------------------------------------------------
def rotation_matrix(a, b):
    """MDAnalysis.analysis.align.rotation_matrix(a, b) restated: the proper rotation R
    minimising sum |R a_i - b_i|^2 (so that ``a . R^T`` overlays ``b``) and the RMSD.
    Kabsch/SVD in float64 (MDAnalysis uses QCP; same minimiser)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    H = a.T @ b                                   # 3x3, sum_i a_i b_i^T
    U, S, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(Vt.T @ U.T))
    D = np.diag([1.0, 1.0, d])
    R = Vt.T @ D @ U.T
    diff = a @ R.T - b
    rmsd = np.sqrt(np.sum(diff * diff) / a.shape[0])
    return R, rmsd


def _RMSD(x, y, n):
    """trajectory_analysis_functions.py:24-36"""
    return (np.sum(np.square(x - y)) / n) ** 0.5


def iterative_alignment(all_pos_Align, all_pos_Nodes, convergence_threshold=1e-5,
                        maximum_num_iterations=100, log=None):
    """trajectory_analysis_functions.py:86-126.  Arrays are modified in place like the
    reference does.  all_pos_Align keeps its dtype (float32 when it came from
    ``.positions``), all_pos_Nodes is float64."""
    nSteps = all_pos_Nodes.shape[0]
    nAlign = all_pos_Align.shape[1]
    nNodes = all_pos_Nodes.shape[1]
    avg_pos_Align = np.sum(all_pos_Align, axis=0) / nSteps                   # :86
    avg_pos_Nodes = np.sum(all_pos_Nodes, axis=0) / nSteps                   # :87
    iteration = 0
    residual = convergence_threshold + 9999.0                                # :93
    while residual > convergence_threshold and iteration < maximum_num_iterations:  # :96
        temp_avg_pos_Align = np.zeros((nAlign, 3), dtype=np.float64)
        temp_avg_pos_Nodes = np.zeros((nNodes, 3), dtype=np.float64)
        for ts in range(nSteps):                                             # :100
            R, d = rotation_matrix(all_pos_Align[ts], avg_pos_Align)         # :103
            all_pos_Align[ts] = np.dot(all_pos_Align[ts], R.T)               # :106
            all_pos_Nodes[ts] = np.dot(all_pos_Nodes[ts], R.T)               # :107
            temp_avg_pos_Align += all_pos_Align[ts]                          # :110
            temp_avg_pos_Nodes += all_pos_Nodes[ts]                          # :111
        temp_avg_pos_Align /= nSteps
        temp_avg_pos_Nodes /= nSteps
        residual = _RMSD(avg_pos_Align, temp_avg_pos_Align, nAlign)          # :118
        analysis_RMSD = _RMSD(avg_pos_Nodes, temp_avg_pos_Nodes, nNodes)
        iteration += 1
        avg_pos_Align = np.copy(temp_avg_pos_Align)
        avg_pos_Nodes = np.copy(temp_avg_pos_Nodes)
        if log is not None:
            log.append((iteration, residual, analysis_RMSD))
    return all_pos_Nodes, avg_pos_Nodes


def traj_alignment_and_averaging(coords, masses, node_ptr, atom_idx, align_idx,
                                 node_definition="COM", convergence_threshold=1e-5,
                                 maximum_num_iterations=100, step=1):
    """trajectory_analysis_functions.py:39-138 on raw arrays (file output omitted)."""
    coords = np.asarray(coords)[::step]
    nodes, align = com_nodes(coords, masses, node_ptr, atom_idx, align_idx, node_definition)
    return iterative_alignment(align, nodes, convergence_threshold, maximum_num_iterations)


# ----------------------------------------------------------------------------------------
# a2: cartesian covariance (trajectory_analysis_functions.py:141-184)
# ----------------------------------------------------------------------------------------
def cartesian_covariance(node_trajectory, avg_node_positions):
    """Literal: T rank-1 updates, /T, minus mean outer product (:153-171).
    (B6: the reference's undefined all_pos_Nodes/avg_pos_Nodes are the two arguments.)"""
    nSteps = node_trajectory.shape[0]
    nNodes = node_trajectory.shape[1]
    nCartCoords = 3 * nNodes
    cov = np.zeros((nCartCoords, nCartCoords), dtype=np.float64)             # :163
    for ts in range(nSteps):                                                 # :164
        f = node_trajectory[ts].flatten()
        cov += np.dot(f.reshape(nCartCoords, 1), f.reshape(1, nCartCoords))  # :166
    cov /= nSteps                                                            # :168
    a = avg_node_positions.flatten()
    cov -= np.dot(a.reshape(nCartCoords, 1), a.reshape(1, nCartCoords))      # :170
    return cov


def cartesian_covariance_fast(node_trajectory, avg_node_positions):
    """Same quantity as one GEMM (summation order differs: ~1e-16 relative per term)."""
    T, N, _ = node_trajectory.shape
    Y = np.ascontiguousarray(node_trajectory, dtype=np.float64).reshape(T, 3 * N)
    cov = (Y.T @ Y) / T
    a = np.asarray(avg_node_positions, dtype=np.float64).reshape(3 * N)
    cov -= np.outer(a, a)
    return cov


# ----------------------------------------------------------------------------------------
# a3: Pearson correlation (adjacency_matrix_functions.py:8-59)
# ----------------------------------------------------------------------------------------
def pearson_correlation_analysis(cartesian_covariance_matrix):
    """Literal double loop (:22-46).  Returns (node_variance, node_covariance,
    node_correlation) -- the reference returns only the last and writes the other two."""
    nCartCoords = cartesian_covariance_matrix.shape[0]
    nNodes = nCartCoords // 3                                                # :23 (B9)
    node_variance = np.zeros((nNodes), dtype=np.float64)
    node_covariance = np.zeros((nNodes, nNodes), dtype=np.float64)
    node_correlation = np.zeros((nNodes, nNodes), dtype=np.float64)
    cartesian_variance_matrix = np.diagonal(cartesian_covariance_matrix)     # :33
    for i in range(nNodes):
        node_variance[i] = np.sum(cartesian_variance_matrix[3 * i:3 * i + 3])  # :37
    for i in range(nNodes):                                                  # :39
        for j in range(i, nNodes):
            node_covariance[i, j] = node_covariance[j, i] = np.trace(
                cartesian_covariance_matrix[3 * i:3 * i + 3, 3 * j:3 * j + 3])  # :45
            node_correlation[i, j] = node_correlation[j, i] = (
                node_covariance[i, j] / np.sqrt(node_variance[i] * node_variance[j]))  # :46
    return node_variance, node_covariance, node_correlation


def pearson_correlation_fast(cartesian_covariance_matrix):
    """Vectorised, bit-identical to the literal loop: trace of a 3x3 = (xx + yy) + zz,
    upper triangle mirrored."""
    C = np.asarray(cartesian_covariance_matrix, dtype=np.float64)
    n3 = C.shape[0]
    N = n3 // 3
    B = C.reshape(N, 3, N, 3)
    ncov = (B[:, 0, :, 0] + B[:, 1, :, 1]) + B[:, 2, :, 2]
    iu = np.triu_indices(N)
    full = np.zeros((N, N))
    full[iu] = ncov[iu]
    full = full + np.triu(full, 1).T
    d = np.diagonal(C)
    var = (d[0::3] + d[1::3]) + d[2::3]
    corr = full / np.sqrt(var[:, None] * var[None, :])
    # mirror exactly like the reference (value computed for i<=j is copied to [j,i])
    corr = np.triu(corr) + np.triu(corr, 1).T
    return var, full, corr


def node_gram_truth(node_trajectory):
    """longdouble centred Gram matrix  G_ij = (1/T) sum_t (x_i-mu_i).(x_j-mu_j)  and the
    correlation from it -- the yardstick used to show that a centred GPU contraction is at
    least as close to the truth as the reference's E[xx^T]-mu mu^T."""
    X = np.asarray(node_trajectory, dtype=np.longdouble)
    T, N, _ = X.shape
    Xc = X - X.mean(axis=0, keepdims=True)
    G = np.zeros((N, N), dtype=np.longdouble)
    for c in range(3):
        Y = Xc[:, :, c]
        G += Y.T @ Y
    G /= T
    v = np.diagonal(G)
    C = G / np.sqrt(v[:, None] * v[None, :])
    return G.astype(np.float64), C.astype(np.float64)


# ----------------------------------------------------------------------------------------
# a4: contact map (trajectory_analysis_functions.py:187-236)
# ----------------------------------------------------------------------------------------
def _distance_array(a, b, mda_float32=False):
    """MDAnalysis distance_array(a, b) restated: Euclidean, no box.  With
    ``mda_float32`` the coordinates are cast to float32 first and differences are taken in
    float32 (what MDAnalysis >= 0.19 does internally); default is plain float64."""
    if mda_float32:
        a = np.asarray(a, dtype=np.float32)
        b = np.asarray(b, dtype=np.float32)
        d = (b[None, :, :] - a[:, None, :]).astype(np.float64)
    else:
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        d = b[None, :, :] - a[:, None, :]
    rsq = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    return np.sqrt(rsq)


def calc_contact_map(trajectory_data, distance_cutoff=9999.0, mda_float32=False):
    """Literal (:200-225): per-frame N x N distances summed, /T, strict '<' threshold on
    i<j mirrored, diagonal 0.  (B7: ``trajectory`` -> ``trajectory_data``; B10 np.int.)"""
    nSteps = trajectory_data.shape[0]
    nNodes = trajectory_data.shape[1]
    avg = np.zeros((nNodes, nNodes), dtype=np.float64)
    for ts in range(nSteps):                                                 # :207
        avg += _distance_array(trajectory_data[ts], trajectory_data[ts], mda_float32)
    avg /= nSteps                                                            # :215
    binary = np.zeros((nNodes, nNodes), int)                                 # :220
    for i in range(nNodes - 1):
        for j in range(i + 1, nNodes):
            if avg[i, j] < distance_cutoff:                                  # :224
                binary[i, j] = binary[j, i] = 1
    return binary, avg


def calc_contact_map_fast(trajectory_data, distance_cutoff=9999.0, block=None):
    """Frame-blocked vectorised form (same per-pair summation order over frames)."""
    X = np.asarray(trajectory_data, dtype=np.float64)
    T, N, _ = X.shape
    if block is None:
        block = max(1, min(64, int(3e7 // (N * N))))
    avg = np.zeros((N, N), dtype=np.float64)
    for t0 in range(0, T, block):
        x = X[t0:t0 + block]
        d = x[:, None, :, :] - x[:, :, None, :]
        rsq = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        r = np.sqrt(rsq)
        for k in range(r.shape[0]):     # keep sequential frame order of the literal loop
            avg += r[k]
    avg /= T
    binary = np.zeros((N, N), int)
    iu = np.triu_indices(N, 1)
    m = avg[iu] < distance_cutoff
    binary[iu[0][m], iu[1][m]] = 1
    binary[iu[1][m], iu[0][m]] = 1
    return binary, avg


# ----------------------------------------------------------------------------------------
# a5: -log|C| weighting (functionalize_adj_matrix_functions.py:10-24)
# ----------------------------------------------------------------------------------------
def func_pearson_correlation(adjacency_matrix, contact_map=None, Lambda=None):
    with np.errstate(divide="ignore"):
        out = -np.log(np.fabs(np.asarray(adjacency_matrix, dtype=np.float64)))   # :17
    if contact_map is not None:
        out *= contact_map                                                       # :19-20
    return out


# ----------------------------------------------------------------------------------------
# graph helpers restating the networkx-1.x semantics the reference relies on
# ----------------------------------------------------------------------------------------
def graph_edges(W):
    """networkx 1.x ``Graph(data=W)``: entry != 0 <=> edge (NaN / inf are edges, +-0.0 are
    not) (network_analysis_functions.py:20,108).  Returns sorted neighbour lists."""
    W = np.asarray(W)
    N = W.shape[0]
    nz = (W != 0) | (W.T != 0)
    return [np.nonzero(nz[u])[0].tolist() for u in range(N)]


def dijkstra(W, source, adj=None):
    """Single-source Dijkstra (binary heap, like networkx).  Returns (dist, pred) arrays;
    unreachable -> inf / -1."""
    W = np.asarray(W, dtype=np.float64)
    N = W.shape[0]
    if adj is None:
        adj = graph_edges(W)
    dist = np.full(N, np.inf)
    pred = np.full(N, -1, dtype=np.int64)
    done = np.zeros(N, dtype=bool)
    dist[source] = 0.0
    h = [(0.0, source)]
    while h:
        d, u = heapq.heappop(h)
        if done[u]:
            continue
        done[u] = True
        for v in adj[u]:
            nd = d + W[u, v]
            if nd < dist[v]:
                dist[v] = nd
                pred[v] = u
                heapq.heappush(h, (nd, v))
    return dist, pred


def dijkstra_path(W, source, sink, adj=None):
    dist, pred = dijkstra(W, source, adj)
    if not np.isfinite(dist[sink]) and pred[sink] < 0:
        raise ValueError("no path between %d and %d" % (source, sink))
    p = [sink]
    while p[-1] != source:
        p.append(int(pred[p[-1]]))
    return p[::-1]


def apsp_floyd_warshall(W):
    """Textbook Floyd-Warshall with predecessors on the graph of ``graph_edges(W)``.
    The reference has no APSP (it calls Dijkstra, network_analysis_functions.py:30,71-72);
    this is the oracle for the north-star APSP stage.  pred[i,j] = node before j on the
    shortest i->j path, -1 on the diagonal / unreachable."""
    W = np.asarray(W, dtype=np.float64)
    N = W.shape[0]
    D = np.where((W != 0), W, np.inf)
    np.fill_diagonal(D, 0.0)
    P = np.where((W != 0), np.arange(N)[:, None].repeat(N, 1), -1).astype(np.int64)
    np.fill_diagonal(P, -1)
    for k in range(N):
        cand = D[:, k, None] + D[None, k, :]
        better = cand < D
        D = np.where(better, cand, D)
        P = np.where(better, P[k][None, :].repeat(N, 0), P)
    return D, P


def apsp_dijkstra(W):
    """N Dijkstra runs -- literally what the reference's call sites compute, for all
    sources."""
    W = np.asarray(W, dtype=np.float64)
    N = W.shape[0]
    adj = graph_edges(W)
    D = np.empty((N, N))
    P = np.empty((N, N), dtype=np.int64)
    for s in range(N):
        D[s], P[s] = dijkstra(W, s, adj)
    return D, P


def lr_length(W, nodes):
    """Left-to-right float64 path length (network_analysis_functions.py:31-33,133)."""
    length = 0.0
    for t in range(len(nodes) - 1):
        length += W[nodes[t], nodes[t + 1]]
    return float(length)


# ----------------------------------------------------------------------------------------
# a6-a9: get_paths (network_analysis_functions.py:12-214)
# ----------------------------------------------------------------------------------------
def enumerate_paths_literal(W, adj_sub, source, sink, cutoff):
    """The frontier-list loop (:111-146) transliterated, including the membership test
    ``node in [length, n0, n1, ...]`` that makes node 0 collide with the initial 0.0
    length (B12)."""
    paths_growing = [[0.0, source]]                                          # :112
    full = []
    while len(paths_growing) > 0:                                            # :115
        for i in range(len(paths_growing)):
            if paths_growing[i][0] > cutoff:                                 # :120
                paths_growing.pop(i)
                break
            elif paths_growing[i][-1] == sink:                               # :123
                full.append(paths_growing[i])
                paths_growing.pop(i)
                break
            else:
                nbrs = adj_sub[paths_growing[i][-1]]                         # :128
                for j in range(len(nbrs)):
                    if not nbrs[j] in paths_growing[i]:                      # :130
                        temp = paths_growing[i][:]
                        temp.append(nbrs[j])
                        temp[0] += W[temp[-2], temp[-1]]                     # :133
                        paths_growing.insert(i + j + 1, temp)
                paths_growing.pop(i)
                break
    full.sort()                                                              # :145
    return full


def dedupe_paths(paths):
    """remove_redundant_paths (:157-174 and :189-207): two entries are the same path iff
    equal number of nodes and equal node sequences after orienting each so that the first
    node >= the last node; first occurrence wins.  (Set-based, same result as the O(P^2)
    loop.)"""
    if len(paths) == 1:
        return list(paths)
    seen = set()
    out = []
    for p in paths:
        nodes = tuple(p[1:])
        if nodes[0] < nodes[-1]:
            nodes = nodes[::-1]
        if nodes in seen:
            continue
        seen.add(nodes)
        out.append(p)
    return out


def _seed(W, sources, sinks, adj):
    """:25-36 -- first strict minimum over (source, sink) pairs of the LR length of a
    Dijkstra path."""
    shortest_length = 99999999.999
    shortest_path = []
    for source in sources:
        for sink in sinks:
            if source != sink:
                sp = dijkstra_path(W, source, sink, adj)
                length = 0
                for t in range(len(sp) - 1):
                    length += W[sp[t], sp[t + 1]]
                if length < shortest_length:
                    shortest_length = length
                    shortest_path = sp
    return float(shortest_length), [int(x) for x in shortest_path]


def get_paths(adjacency_matrix, sources, sinks, number_of_paths, max_passes=None,
              cutoff_increment=0.1, stats=None):
    """Literal get_paths (:12-214).  The networkx-1.x calls are replaced by the helpers
    above (B11: SSSP results are indexed by node).  ``max_passes`` only guards the
    infinite loop the reference has when fewer than K paths exist (App. A.6)."""
    W = np.asarray(adjacency_matrix, dtype=np.float64)
    nNodes = len(W)
    adj = graph_edges(W)                                                     # :20
    shortest_length, shortest_path = _seed(W, sources, sinks, adj)           # :25-36
    num_paths = 1
    paths = [[shortest_length] + shortest_path]                              # :43-46
    cutoff = shortest_length                                                 # :48
    passes = 0
    while num_paths < number_of_paths:                                       # :56
        temp_paths = []
        for source in sources:
            for sink in sinks:
                if source != sink:
                    source_lengths, sp = dijkstra(W, source, adj)            # :71
                    sink_lengths, tp = dijkstra(W, sink, adj)                # :72
                    if not (np.all(np.isfinite(source_lengths)) and np.all(np.isfinite(sink_lengths))):
                        raise IndexError("unreachable node (reference :84-85)")
                    # :88-99 keep nodes with d(s,i)+d(t,i) <= cutoff (the nodes on their
                    # tree paths satisfy the same inequality -- App. A.3)
                    keep = (source_lengths + sink_lengths) <= cutoff
                    # tree paths of kept nodes (literal :92-93) -- add them explicitly so
                    # the set is the reference's even under rounding
                    unique_nodes = set(np.nonzero(keep)[0].tolist())
                    for i in list(unique_nodes):
                        for pred, root in ((sp, source), (tp, sink)):
                            u = i
                            while u != root:
                                u = int(pred[u])
                                unique_nodes.add(u)
                    unique_nodes = sorted(unique_nodes)
                    keepmask = np.zeros(nNodes, dtype=bool)
                    keepmask[unique_nodes] = True
                    adj_sub = [[v for v in adj[u] if keepmask[v]] if keepmask[u] else []
                               for u in range(nNodes)]                       # :102-108
                    full = enumerate_paths_literal(W, adj_sub, source, sink, cutoff)
                    temp_paths.extend(full)
        if len(temp_paths) != 1:                                             # :157
            temp_paths = dedupe_paths(temp_paths)
        num_paths = len(temp_paths)                                          # :179
        paths.extend(temp_paths)
        if stats is not None:
            stats.append((cutoff, num_paths))
        cutoff += cutoff_increment                                           # :187
        passes += 1
        if max_passes is not None and passes >= max_passes and num_paths < number_of_paths:
            raise RuntimeError("fewer than K paths after %d passes" % passes)
    if len(paths) != 1:                                                      # :189
        paths = dedupe_paths(paths)
    paths.sort()                                                             # :209
    if num_paths != number_of_paths:                                         # :211
        paths = paths[:number_of_paths]
    return paths


def enumerate_paths_pruned(W, adj, dist_to_sink, source, sink, cutoff, node0_quirk=True,
                           slack=1e-9, counter=None):
    """All simple source->sink paths with LR length <= cutoff (sink only last), minus those
    whose 2nd node is 0 when source != 0 (App. A.3) -- same set as
    ``enumerate_paths_literal`` but with the admissible bound
    len + w(u,v) + d(v,sink) <= cutoff(1+slack) so dead ends are never expanded."""
    out = []
    bound = cutoff + slack * max(1.0, abs(cutoff))
    path = [source]
    onpath = np.zeros(len(adj), dtype=bool)
    onpath[source] = True

    def rec(u, length):
        for v in adj[u]:
            if onpath[v]:
                continue
            if node0_quirk and v == 0 and len(path) == 1:
                continue          # `0 in [0.0, s]` is True in the reference (:130)
            nl = length + W[u, v]
            if counter is not None:
                counter[0] += 1
            if not (nl + dist_to_sink[v] <= bound):
                continue
            if v == sink:
                if nl <= cutoff:
                    out.append([float(nl)] + path + [v])
                continue
            if nl > cutoff:
                continue
            path.append(v)
            onpath[v] = True
            rec(v, nl)
            path.pop()
            onpath[v] = False

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 10000))
    try:
        rec(source, 0.0)
    finally:
        sys.setrecursionlimit(old)
    out.sort()
    return out


def get_paths_pruned(adjacency_matrix, sources, sinks, number_of_paths, node0_quirk=True,
                     max_passes=100000, stats=None):
    """The Appendix-A model of get_paths: same seed, same cutoff schedule (repeated
    ``+= 0.1``), same per-pass dedupe / stop rule / K-vs-K+1 rule, with the pruned
    enumerator.  Result-identical to ``get_paths`` (tests/test_oracle_paths.py)."""
    W = np.asarray(adjacency_matrix, dtype=np.float64)
    adj = graph_edges(W)
    shortest_length, shortest_path = _seed(W, sources, sinks, adj)
    paths = [[shortest_length] + shortest_path]
    num_paths = 1
    cutoff = shortest_length
    sink_dist = {}
    for sink in sinks:
        sink_dist[sink] = dijkstra(W, sink, adj)[0]
    passes = 0
    while num_paths < number_of_paths:
        temp_paths = []
        for source in sources:
            for sink in sinks:
                if source != sink:
                    temp_paths.extend(enumerate_paths_pruned(
                        W, adj, sink_dist[sink], source, sink, cutoff, node0_quirk))
        if len(temp_paths) != 1:
            temp_paths = dedupe_paths(temp_paths)
        num_paths = len(temp_paths)
        paths.extend(temp_paths)
        if stats is not None:
            stats.append((cutoff, num_paths))
        cutoff += 0.1
        passes += 1
        if passes >= max_passes and num_paths < number_of_paths:
            raise RuntimeError("fewer than K paths after %d passes" % passes)
    if len(paths) != 1:
        paths = dedupe_paths(paths)
    paths.sort()
    if num_paths != number_of_paths:
        paths = paths[:number_of_paths]
    return paths


# ----------------------------------------------------------------------------------------
# linear mutual information adjacency and its functionalisation
# ----------------------------------------------------------------------------------------
def linear_mutual_information_analysis(cartesian_covariance_matrix):
    """adjacency_matrix_functions.py:62-103 literal: MI_ij = 0.5 ln(det C_ii det C_jj /
    det C_(ij)) with C_(ij) the 6x6 covariance of the two nodes; +inf on the diagonal."""
    C = np.asarray(cartesian_covariance_matrix, dtype=np.float64)
    nNodes = C.shape[0] // 3
    MI = np.zeros((nNodes, nNodes), dtype=np.float64)
    for i in range(nNodes - 1):
        iIndex = i * 3
        iIndex3 = iIndex + 3
        for j in range(i + 1, nNodes):
            jIndex = j * 3
            jIndex3 = jIndex + 3
            temp = np.linalg.det(C[iIndex:iIndex3, iIndex:iIndex3]) * np.linalg.det(C[jIndex:jIndex3, jIndex:jIndex3])
            idx = np.append(np.arange(iIndex, iIndex3, 1), np.arange(jIndex, jIndex3, 1))
            MI[i, j] = MI[j, i] = 0.5 * np.log(temp / np.linalg.det(C[np.ix_(idx, idx)]))   # :87
    MI += np.diagflat(np.full(nNodes, np.inf))                                               # :89
    return MI


def generalized_correlation_coefficient_calc(adjacency_matrix, contact_map=None, Lambda=None):
    """functionalize_adj_matrix_functions.py:27-46: r_MI = sqrt(1 - exp(-2 MI / 3)),
    optionally damped by exp(-d_ij / Lambda); returns (r_MI, -log r_MI)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        rMI = np.sqrt(1.0 - np.exp(-2.0 * np.asarray(adjacency_matrix, dtype=np.float64) / 3))   # :34
        if contact_map is not None and Lambda is not None:
            rMI = rMI * np.exp(-np.asarray(contact_map, dtype=np.float64) / Lambda)               # :38-40
        return rMI, -np.log(rMI)                                                                  # :42
