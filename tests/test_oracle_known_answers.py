"""The oracle on closed-form cases (SURVEY 8c item d).  CPU only."""
import numpy as np

import oracle
from known_answers import (LADDER_PATHS_1_TO_5, correlated_trajectory, ladder_graph, lattice_contact_case,
                           rigid_translation_trajectory)


def _corr(nodes):
    avg = nodes.sum(axis=0) / nodes.shape[0]
    return oracle.pearson_correlation_analysis(oracle.cartesian_covariance(nodes, avg))[2]


def test_correlated_anticorrelated_orthogonal():
    nodes, want = correlated_trajectory()
    corr = _corr(nodes)
    assert np.array_equal(corr, want)
    w = oracle.func_pearson_correlation(corr)
    # |C| = 1 -> weight 0 -> NO edge; C = 0 -> weight +inf -> an (infinitely long) edge
    assert np.all(w[:3, :3] == 0.0) and np.all(np.isinf(w[3, :3])) and w[3, 3] == 0.0
    assert [list(a) for a in oracle.graph_edges(w)] == [[3], [3], [3], [0, 1, 2]]


def test_rigid_translation_all_ones():
    nodes, base = rigid_translation_trajectory()
    corr = _corr(nodes)
    # E[xx^T] - mu mu^T cancels ~|mu|^2 / var digits: 1 up to that noise, and
    # exactly 1 on the diagonal
    np.testing.assert_allclose(corr, np.ones_like(corr), rtol=0, atol=1e-11)
    assert np.all(np.diagonal(corr) == 1.0)
    binary, avgdist = oracle.calc_contact_map(nodes, 7.5)
    want = np.sqrt(((base[:, None, :] - base[None, :, :]) ** 2).sum(-1))
    np.testing.assert_allclose(avgdist, want, rtol=1e-14, atol=0)
    assert np.array_equal(binary, ((want < 7.5) & ~np.eye(len(base), dtype=bool)).astype(binary.dtype))


def test_contact_cutoff_is_strict():
    nodes, cutoff, binary, d = lattice_contact_case()
    b, a = oracle.calc_contact_map(nodes, cutoff)
    assert np.array_equal(a, d)
    assert np.array_equal(b, binary) and b[0, 2] == 0 and b[0, 1] == 1
    b2, a2 = oracle.calc_contact_map_fast(nodes, cutoff)
    assert np.array_equal(a2, d) and np.array_equal(b2, binary)


def test_ladder_graph_paths_by_hand():
    W = ladder_graph()
    every = LADDER_PATHS_1_TO_5
    # K = 2: seed 1-0-5 (Dijkstra), enumeration never lists it (node-0 rule); the first pass that
    # reaches >= 2 enumerated paths has cutoff 0.65 and finds exactly 2 -> "num == K" -> nothing is
    # truncated and the seed rides along: K + 1 = 3 results
    got = oracle.get_paths(W, [1], [5], 2)
    assert got == every[:3]
    assert oracle.get_paths_pruned(W, [1], [5], 2) == every[:3]
    # K = 3: cutoff reaches 1.05 -> enumerated {0.5, 0.625, 1.0} = 3 == K -> again K + 1
    assert oracle.get_paths(W, [1], [5], 3) == every[:4]
    # from the other end node 0 is allowed as second node only if the SOURCE is 0
    rev = oracle.get_paths(W, [0], [5], 1)
    assert rev[0] == [0.125, 0, 5]
    # without the quirk the plain K shortest simple paths come back
    assert oracle.get_paths_pruned(W, [1], [5], 4, node0_quirk=False) == every[:4]


def test_reference_code_confirms_the_hand_enumeration(golden_dir):
    """tests/golden/get_paths.npz cases 11-14 are the ladder graph run through the reference's
    own get_paths (make_golden.py): its output is the list derived on paper."""
    import os
    g = np.load(os.path.join(golden_dir, "get_paths.npz"))

    def case(i):
        off, nodes = g["offsets%d" % i], g["nodes%d" % i]
        return [[float(g["lengths%d" % i][k])] + [int(x) for x in nodes[off[k]:off[k + 1]]]
                for k in range(len(off) - 1)]
    assert np.array_equal(g["W11"], ladder_graph())
    assert case(11) == LADDER_PATHS_1_TO_5[:3]          # K = 2 -> 3 results
    assert case(12) == LADDER_PATHS_1_TO_5[:4]          # K = 3 -> 4 results
    assert case(13) == [[0.125, 0, 5]]
    # 5 -> 1: same paths, stored in the orientation they were enumerated in
    assert [p[0] for p in case(14)] == [0.25, 0.5, 0.625]
    assert case(14)[0][1:] in ([5, 0, 1], [1, 0, 5])
