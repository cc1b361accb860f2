"""The oracle against golden vectors produced by the reference's own source
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

import oracle


@pytest.mark.parametrize("tag", ["traj_small", "traj_aligned"])
def test_trajectory_stages(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    nodes, avg = oracle.traj_alignment_and_averaging(
        g["coords"].copy(), g["masses"], g["node_ptr"], g["atom_idx"], g["align_idx"])
    # Kabsch rotations go through an SVD in both the shim and the oracle: identical code
    # path -> identical results; allow a few ulp for BLAS differences
    np.testing.assert_allclose(nodes, g["node_traj"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(avg, g["avg"], rtol=0, atol=1e-10)

    cov = oracle.cartesian_covariance(g["node_traj"], g["avg"])
    assert np.array_equal(cov, g["cov"])                       # same loop, same order: bit-exact
    cov_fast = oracle.cartesian_covariance_fast(g["node_traj"], g["avg"])
    np.testing.assert_allclose(cov_fast, g["cov"], rtol=0, atol=1e-9 * np.abs(g["cov"]).max())

    var, ncov, corr = oracle.pearson_correlation_analysis(g["cov"])
    assert np.array_equal(corr, g["corr"])
    # node_variance / node_covariance went through np.savetxt('%.18e') -> exact round trip
    assert np.array_equal(var, g["var"])
    assert np.array_equal(ncov, g["ncov"])
    var2, ncov2, corr2 = oracle.pearson_correlation_fast(g["cov"])
    assert np.array_equal(corr2, g["corr"]) and np.array_equal(var2, g["var"])
    assert np.array_equal(ncov2, g["ncov"])
    # diagonal is exactly 1
    assert np.all(np.diagonal(corr) == 1.0)

    binary, avgdist = oracle.calc_contact_map(g["node_traj"], float(g["cutoff"]))
    assert np.array_equal(binary, g["binary"])
    assert np.array_equal(avgdist, g["avgdist"])
    b2, a2 = oracle.calc_contact_map_fast(g["node_traj"], float(g["cutoff"]))
    assert np.array_equal(b2, g["binary"]) and np.array_equal(a2, g["avgdist"])

    assert np.array_equal(oracle.func_pearson_correlation(g["corr"]), g["w_plain"])
    assert np.array_equal(oracle.func_pearson_correlation(g["corr"], g["binary"]), g["w_binary"])
    assert np.array_equal(oracle.func_pearson_correlation(g["corr"], g["avgdist"]), g["w_avg"])


def _unpack(g, i):
    off = g["offsets%d" % i]
    nodes = g["nodes%d" % i]
    return [[float(g["lengths%d" % i][k])] + [int(x) for x in nodes[off[k]:off[k + 1]]]
            for k in range(len(off) - 1)]


def test_get_paths_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "get_paths.npz"))
    n = int(g["n_cases"])
    assert n >= 10
    saw_k_plus_1 = False
    for i in range(n):
        W = g["W%d" % i]
        so = [int(x) for x in g["sources%d" % i]]
        si = [int(x) for x in g["sinks%d" % i]]
        K = int(g["K%d" % i])
        want = _unpack(g, i)
        saw_k_plus_1 |= len(want) == K + 1
        lit = oracle.get_paths(W, so, si, K, max_passes=500)
        pru = oracle.get_paths_pruned(W, so, si, K)
        for got in (lit, pru):
            assert len(got) == len(want), (i, len(got), len(want))
            for a, b in zip(got, want):
                assert [int(x) for x in a[1:]] == b[1:], (i, a, b)
                assert float(a[0]) == b[0], (i, a, b)       # bit-exact float64 LR lengths
    assert saw_k_plus_1


def test_com_extraction_golden(golden_dir):
    """COM extraction alone: the reference run with maximum_num_iterations=0 returns the
    translated, un-rotated node COMs (trajectory_analysis_functions.py:70-87)."""
    for tag in ("traj_small", "traj_aligned"):
        g = np.load(os.path.join(golden_dir, tag + ".npz"))
        nodes, align = oracle.com_nodes(g["coords"], g["masses"], g["node_ptr"], g["atom_idx"],
                                        g["align_idx"])
        np.testing.assert_allclose(nodes, g["com_only"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(nodes.sum(axis=0) / nodes.shape[0], g["com_avg"], rtol=0,
                                   atol=1e-13)


@pytest.mark.parametrize("tag", ["traj_small", "traj_aligned"])
def test_lmi_golden(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    mi = oracle.linear_mutual_information_analysis(g["cov"])
    assert np.array_equal(mi, g["mi"])
    rmi, func = oracle.generalized_correlation_coefficient_calc(g["mi"])
    assert np.array_equal(rmi, g["rmi"]) and np.array_equal(func, g["func_mi"])
    rmi_d, func_d = oracle.generalized_correlation_coefficient_calc(g["mi"], g["avgdist"], 5.0)
    np.testing.assert_allclose(rmi_d, g["rmi_damped"], rtol=1e-15)
    np.testing.assert_allclose(func_d, g["func_mi_damped"], rtol=1e-13)


def test_pruned_enumerator_equals_literal_on_random_graphs():
    """The GPU parity tests use oracle.get_paths_pruned at sizes where the literal transliteration
    of the reference loop is too slow; here the two are held together on many small random graphs
    (single and multiple sources / sinks, node 0 in every role, sparse and dense)."""
    from wisp_mdanalysis_implementation_b200 import synth
    rng = np.random.default_rng(20260101)
    checked = k_plus_1 = 0
    for trial in range(60):
        n = int(rng.integers(6, 12))
        W = synth.random_cost_matrix(n, seed=int(rng.integers(1 << 30)), density=float(rng.uniform(0.3, 0.8)),
                                     lo=0.05, hi=0.7)
        ns, nt = (1, 1) if trial % 3 else (2, 2)
        so = [int(x) for x in rng.choice(n, size=ns, replace=False)]
        si = [int(x) for x in rng.choice(n, size=nt, replace=False)]
        if all(s == t for s in so for t in si):
            continue
        K = int(rng.integers(2, 9))
        D, _ = oracle.apsp_dijkstra(W)
        if not any(np.isfinite(D[s, t]) for s in so for t in si if s != t):
            continue
        try:
            lit = oracle.get_paths(W, so, si, K, max_passes=80)
        except RuntimeError:                     # fewer than K simple paths: the reference never returns
            continue
        pru = oracle.get_paths_pruned(W, so, si, K)
        assert len(pru) == len(lit), (trial, len(pru), len(lit))
        for a, b in zip(pru, lit):
            assert [int(x) for x in a[1:]] == [int(x) for x in b[1:]] and float(a[0]) == float(b[0]), (trial, a, b)
        checked += 1
        k_plus_1 += len(lit) == K + 1
    assert checked >= 30


def test_com_nodes_fast_is_bit_identical(golden_dir):
    """The vectorised COM form used by the large GPU parity cases equals the literal per-frame /
    per-node loop bit for bit: on the golden coordinates (which pin the literal form to the
    reference's own output) and on ragged / ATOMIC node maps."""
    g = np.load(os.path.join(golden_dir, "traj_small.npz"))
    a, al = oracle.com_nodes(g["coords"], g["masses"], g["node_ptr"], g["atom_idx"], g["align_idx"])
    b, bl = oracle.com_nodes_fast(g["coords"], g["masses"], g["node_ptr"], g["atom_idx"], g["align_idx"], block=7)
    assert np.array_equal(a, b) and np.array_equal(al, bl)
    rng = np.random.default_rng(0)
    coords = (rng.normal(size=(23, 90, 3)) * 20).astype(np.float32)
    masses = rng.uniform(1, 32, size=90)
    ptr, idx = [0], []
    for _ in range(17):
        k = int(rng.integers(1, 19))
        idx += list(rng.choice(90, size=k, replace=False))
        ptr.append(len(idx))
    align = rng.choice(90, size=31, replace=False)
    a, al = oracle.com_nodes(coords, masses, np.array(ptr), np.array(idx), align)
    b, bl = oracle.com_nodes_fast(coords, masses, np.array(ptr), np.array(idx), align)
    assert np.array_equal(a, b) and np.array_equal(al, bl)
    a, _ = oracle.com_nodes(coords, masses, np.arange(18), np.array(idx[:17]), align, "ATOMIC")
    b, _ = oracle.com_nodes_fast(coords, masses, np.arange(18), np.array(idx[:17]), align, "ATOMIC")
    assert np.array_equal(a, b)
