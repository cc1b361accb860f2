"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and against the
golden vectors produced by the reference's own code.  Tolerances (BASELINE.json north_star):
correlation / contact matrices 1e-6 relative (floor 1e-3 on |C|), path
lengths 1e-9 relative (bit-exact here), contact topology / predecessors / node sequences
exact outside ties."""
import os

import numpy as np
import pytest

import oracle
from wisp_mdanalysis_implementation_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from wisp_mdanalysis_implementation_b200 import _lib
    return _lib.default_context()


def assert_corr_close(got, want, rel=1e-6, floor=1e-3):
    err = np.abs(got - want)
    tol = rel * np.maximum(np.abs(want), floor)
    bad = err > tol
    assert not bad.any(), "max err %.3e (tol %.3e) at %s" % (err.max(), tol.flat[err.argmax()], np.argwhere(bad)[:5])


def assert_binary_equal_outside_ties(binary, avgdist_ref, cutoff, binary_ref, rel=1e-6):
    diff = binary != binary_ref
    if diff.any():
        near = np.abs(avgdist_ref - cutoff) <= rel * cutoff
        assert not (diff & ~near).any(), "contact topology differs away from the cutoff"


def _unpack(g, i):
    off = g["offsets%d" % i]
    nodes = g["nodes%d" % i]
    return [[float(g["lengths%d" % i][k])] + [int(x) for x in nodes[off[k]:off[k + 1]]]
            for k in range(len(off) - 1)]


@pytest.mark.parametrize("tag", ["traj_small", "traj_aligned"])
def test_golden_trajectory_stages(ctx, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    # K1: translate + COM (reference run with zero alignment iterations)
    nodes, avg = ctx.com_nodes(g["coords"], g["masses"], g["node_ptr"], g["atom_idx"], g["align_idx"])
    np.testing.assert_allclose(nodes, g["com_only"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(avg, g["com_avg"], rtol=0, atol=1e-12)
    # K1 + K7: full traj_alignment_and_averaging
    nodes_al, avg_al, log = ctx.traj_alignment_and_averaging(
        g["coords"], g["masses"], g["node_ptr"], g["atom_idx"], g["align_idx"])
    np.testing.assert_allclose(nodes_al, g["node_traj"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(avg_al, g["avg"], rtol=0, atol=1e-8)
    assert len(log) >= 1 and log[-1][1] <= 1e-5

    # K2 (stride 1) + finalize: 3N x 3N covariance
    cov = ctx.cartesian_covariance(g["node_traj"], g["avg"])
    scale = np.abs(g["cov"]).max()
    assert np.abs(cov - g["cov"]).max() <= 1e-9 * scale
    assert np.array_equal(cov, cov.T)
    # a3 on the reference's covariance: same IEEE operations -> bit-exact
    var, ncov, corr = ctx.pearson_from_covariance(g["cov"])
    assert np.array_equal(var, g["var"])
    assert np.array_equal(ncov, g["ncov"])
    assert np.array_equal(corr, g["corr"])
    # K3 + K4: contact maps
    binary, avgdist = ctx.contact_map(g["node_traj"], float(g["cutoff"]))
    np.testing.assert_allclose(avgdist, g["avgdist"], rtol=1e-6, atol=1e-9)
    assert np.all(np.diagonal(avgdist) == 0.0) and np.all(np.diagonal(binary) == 0)
    assert_binary_equal_outside_ties(binary, g["avgdist"], float(g["cutoff"]), g["binary"])
    assert binary.dtype == np.int64
    # a5: -log|C| (.) map
    for key, cm in (("w_plain", None), ("w_binary", g["binary"]), ("w_avg", g["avgdist"])):
        w = ctx.func_pearson(g["corr"], None if cm is None else cm.astype(np.float64))
        want = g[key]
        assert np.array_equal(np.isnan(w), np.isnan(want))
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(w), fin)
        np.testing.assert_allclose(w[fin], want[fin], rtol=1e-13, atol=1e-300)
        assert np.all(w[np.diag_indices_from(w)] == 0.0)


def test_golden_get_paths(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "get_paths.npz"))
    for i in range(int(g["n_cases"])):
        W = g["W%d" % i]
        so = [int(x) for x in g["sources%d" % i]]
        si = [int(x) for x in g["sinks%d" % i]]
        K = int(g["K%d" % i])
        want = _unpack(g, i)
        got = ctx.get_paths(W, so, si, K)
        assert len(got) == len(want), (i, len(got), len(want))
        for a, b in zip(got, want):
            assert a[1:] == b[1:], (i, a, b)
            assert a[0] == b[0], (i, a, b)          # bit-exact float64 LR lengths


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("n", [37, 64, 150, 333])
def test_apsp_vs_oracle(ctx, n, precision):
    W = synth.random_cost_matrix(n, seed=100 + n, density=0.08)
    D, P = oracle.apsp_dijkstra(W)
    dist, pred = ctx.apsp(W, precision=precision)
    tol = 1e-12 if precision == 64 else 2e-6
    np.testing.assert_allclose(dist, D, rtol=tol, atol=1e-300)
    if precision == 64:
        # random real weights: shortest paths are unique -> predecessors must agree exactly
        assert np.array_equal(pred.astype(np.int64), P)
    # predecessor consistency (holds in both precisions): dist[i][pred] + w[pred][j] ~ dist[i][j]
    ii, jj = np.nonzero(pred >= 0)
    pp = pred[ii, jj]
    np.testing.assert_allclose(dist[ii, pp] + W[pp, jj], dist[ii, jj], rtol=10 * tol, atol=1e-12)
    assert np.all(np.diagonal(dist) == 0) and np.all(np.diagonal(pred) == -1)


def test_apsp_disconnected_and_inf(ctx):
    W = synth.random_cost_matrix(40, seed=5, density=0.2)
    W[:, 17] = 0.0
    W[17, :] = 0.0                      # isolated node
    W[3, 4] = W[4, 3] = np.inf          # |C| = 0 -> +inf weight: an edge that is never useful
    dist, pred = ctx.apsp(W)
    D, P = oracle.apsp_dijkstra(W)
    assert np.array_equal(np.isinf(dist), np.isinf(D))
    fin = np.isfinite(D)
    np.testing.assert_allclose(dist[fin], D[fin], rtol=1e-12)
    assert np.all(pred[17, np.arange(40) != 17] == -1)


@pytest.mark.parametrize("n,t,apn", [(50, 100, 3), (128, 240, 4), (300, 1000, 16), (257, 333, 5)])
def test_pipeline_vs_oracle(ctx, n, t, apn):
    system, coords = synth.synth(n, t, apn, seed=7 + n)
    cutoff = 15.0
    out = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                       cutoff=cutoff, which_map=1, want_nodes=True)
    nodes, _ = oracle.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx,
                                system.align_idx)
    np.testing.assert_allclose(out["nodes"], nodes, rtol=0, atol=1e-11)
    avg = nodes.sum(axis=0) / t
    np.testing.assert_allclose(out["avg"], avg, rtol=0, atol=1e-11)
    cov = oracle.cartesian_covariance_fast(nodes, avg)
    _, _, corr = oracle.pearson_correlation_fast(cov)
    assert_corr_close(out["corr"], corr)
    # at least as close to the long-double truth as the reference formula is
    _, corr_true = oracle.node_gram_truth(nodes)
    assert np.abs(out["corr"] - corr_true).max() <= max(np.abs(corr - corr_true).max(), 1e-13) * 4
    assert np.all(np.diagonal(out["corr"]) == 1.0)
    assert np.array_equal(out["corr"], out["corr"].T)
    binary, avgdist = oracle.calc_contact_map_fast(nodes, cutoff)
    np.testing.assert_allclose(out["avgdist"], avgdist, rtol=1e-6, atol=1e-9)
    assert_binary_equal_outside_ties(out["binary"], avgdist, cutoff, binary)
    w = oracle.func_pearson_correlation(corr, binary)
    same_topology = np.array_equal(out["binary"], binary)
    if same_topology:
        nz = w != 0
        assert np.array_equal(out["w"] != 0, nz)
        # -log|C|: relative error of C (1e-6, floor) becomes absolute error of W
        np.testing.assert_allclose(out["w"][nz], w[nz], rtol=0, atol=2e-6 + 1e-3 * 0)


def test_cartesian_covariance_supplied_average(ctx):
    """E[xx^T] - a a^T with an average that is NOT the frame mean."""
    rng = np.random.default_rng(3)
    nodes = rng.normal(size=(200, 40, 3)) + rng.normal(scale=20, size=(1, 40, 3))
    a = nodes.mean(axis=0) + rng.normal(scale=0.05, size=(40, 3))
    cov = ctx.cartesian_covariance(nodes, a)
    want = oracle.cartesian_covariance(nodes, a)
    assert np.abs(cov - want).max() <= 1e-9 * np.abs(want).max()


def test_get_paths_c1_stagewise(ctx):
    """Config-1-like run: oracle cost matrix in, K=50 suboptimal paths out, bit-exact."""
    system, coords = synth.synth(300, 400, 4, seed=1001)
    nodes, _ = oracle.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx,
                                system.align_idx)
    avg = nodes.sum(axis=0) / nodes.shape[0]
    _, _, corr = oracle.pearson_correlation_fast(oracle.cartesian_covariance_fast(nodes, avg))
    binary, _ = oracle.calc_contact_map_fast(nodes, 12.0)
    W = oracle.func_pearson_correlation(corr, binary)
    D, _ = oracle.apsp_dijkstra(W)
    # first pair in index order that is 5-8 hops apart
    hops, _ = oracle.apsp_dijkstra((W != 0).astype(float))
    s, t = [(i, j) for i in range(300) for j in range(i + 1, 300) if 5 <= hops[i, j] <= 8][0]
    want = oracle.get_paths_pruned(W, [s], [t], 50)
    got = ctx.get_paths(W, [s], [t], 50)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a[1:] == [int(x) for x in b[1:]]
        assert a[0] == float(b[0])
    st = ctx.paths_stats()
    assert st["passes"] >= 1 and st["expansions"] > 0


def test_get_paths_multi_pair_and_quirk(ctx):
    rng = np.random.default_rng(11)
    for trial in range(6):
        n = int(rng.integers(14, 30))
        W = synth.random_cost_matrix(n, seed=int(rng.integers(1 << 30)), density=0.3, lo=0.05, hi=0.7)
        so = [int(x) for x in rng.choice(n, size=2, replace=False)]
        si = [int(x) for x in rng.choice(n, size=2, replace=False)]
        K = int(rng.integers(5, 40))
        want = oracle.get_paths_pruned(W, so, si, K)
        got = ctx.get_paths(W, so, si, K)
        assert [p[1:] for p in got] == [[int(x) for x in p[1:]] for p in want], trial
        assert [p[0] for p in got] == [float(p[0]) for p in want], trial
        # quirk off == plain K shortest simple paths
        want_nq = oracle.get_paths_pruned(W, so, si, K, node0_quirk=False)
        got_nq = ctx.get_paths(W, so, si, K, node0_quirk=False)
        assert [p[1:] for p in got_nq] == [[int(x) for x in p[1:]] for p in want_nq], trial


def test_get_paths_errors(ctx):
    from wisp_mdanalysis_implementation_b200 import _lib
    W = synth.random_cost_matrix(12, seed=2, density=0.3)
    W[:, 5] = 0
    W[5, :] = 0
    with pytest.raises(_lib.WispError):
        ctx.get_paths(W, [0], [5], 3)           # disconnected: reference raises NetworkXNoPath
    chain = np.zeros((5, 5))
    for i in range(4):
        chain[i, i + 1] = chain[i + 1, i] = 1.0
    with pytest.raises(_lib.WispError):
        ctx.get_paths(chain, [0], [4], 3)       # only one simple path exists: reference never returns


def test_properties_large(ctx):
    """Size-independent properties at a size the oracle would take minutes for."""
    n, t = 1000, 3000
    system = synth.SynthSystem(n, t, 4, seed=99)
    coords = system.coords()
    out = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                       cutoff=12.0, which_map=2)
    c = out["corr"]
    assert np.array_equal(c, c.T) and np.all(np.diagonal(c) == 1.0)
    assert np.abs(c).max() <= 1.0 + 1e-12
    ev = np.linalg.eigvalsh(c)
    assert ev.min() > -1e-8                       # a correlation matrix is PSD
    d = out["avgdist"]
    assert np.array_equal(d, d.T) and np.all(np.diagonal(d) == 0)
    # triangle inequality of mean distances on a sample
    rng = np.random.default_rng(0)
    i, j, k = rng.integers(0, n, size=(3, 2000))
    assert np.all(d[i, j] <= d[i, k] + d[k, j] + 1e-9)
    # linearity in frames: halves combine to the whole (frame-sharding identity)
    h = t // 2
    o1 = ctx.pipeline(coords[:h], system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=12.0)
    o2 = ctx.pipeline(coords[h:], system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=12.0)
    # (distances are FP32 per frame, FP64 accumulated: 1e-6 is the stated tolerance)
    np.testing.assert_allclose(0.5 * (o1["avgdist"] + o2["avgdist"]), d, rtol=1e-6)
    # APSP on the resulting cost matrix: triangle inequality + symmetric
    W = out["w"]
    dist, pred = ctx.apsp(W)
    assert np.allclose(dist, dist.T, rtol=1e-12)
    assert np.all(dist[i, j] <= dist[i, k] + dist[k, j] + 1e-9)


def test_com_general_node_maps(ctx):
    """K1 on node maps the streaming kernel cannot take: ATOMIC nodes, nodes made of
    non-contiguous atoms spanning two residues (make_node_selections.py:210), overlapping
    nodes, descending order -- the CSR gather kernel must match the oracle too."""
    system, coords = synth.synth(40, 90, 6, seed=21)
    A = system.n_atoms
    rng = np.random.default_rng(4)
    # (a) ATOMIC: one atom per node
    atoms = np.sort(rng.choice(A, size=70, replace=False)).astype(np.int32)
    ptr = np.arange(71, dtype=np.int32)
    nodes, avg = ctx.com_nodes(coords, system.masses, ptr, atoms, system.align_idx, atomic=True)
    want, _ = oracle.com_nodes(coords, system.masses, ptr, atoms, system.align_idx, "ATOMIC")
    np.testing.assert_allclose(nodes, want, rtol=0, atol=1e-12)
    # (b) ragged, non-contiguous, overlapping, unordered nodes
    idx, ptr = [], [0]
    for n in range(55):
        size = int(rng.integers(1, 23))
        start = int(rng.integers(0, A - 30))
        members = np.unique(np.concatenate([np.arange(start, start + size // 2 + 1),
                                            rng.choice(A, size=size // 2 + 1)]))
        rng.shuffle(members)
        idx.extend(members.tolist())
        ptr.append(len(idx))
    ptr = np.array(ptr, dtype=np.int32)
    idx = np.array(idx, dtype=np.int32)
    nodes, avg = ctx.com_nodes(coords, system.masses, ptr, idx, system.align_idx)
    want, _ = oracle.com_nodes(coords, system.masses, ptr, idx, system.align_idx)
    np.testing.assert_allclose(nodes, want, rtol=0, atol=1e-11)
    # (c) contiguous residues with unused atoms in between and a large gap (streaming pieces)
    ptr2 = [0]
    idx2 = []
    for start, size in [(0, 5), (5, 7), (20, 3), (23, 9), (150, 12), (162, 6), (200, 1), (201, 30)]:
        idx2.extend(range(start, start + size))
        ptr2.append(len(idx2))
    nodes, avg = ctx.com_nodes(coords, system.masses, np.array(ptr2, np.int32), np.array(idx2, np.int32),
                               system.align_idx[::3])
    want, _ = oracle.com_nodes(coords, system.masses, np.array(ptr2), np.array(idx2), system.align_idx[::3])
    np.testing.assert_allclose(nodes, want, rtol=0, atol=1e-12)


def test_com_streaming_many_pieces(ctx):
    """A frame larger than one 2048-atom piece, odd atom count (unaligned frame starts)."""
    system, coords = synth.synth(333, 37, 13, seed=8)     # 4329 atoms: 3 pieces, A % 4 == 1
    nodes, avg = ctx.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx)
    want, _ = oracle.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx)
    np.testing.assert_allclose(nodes, want, rtol=0, atol=1e-11)
    np.testing.assert_allclose(avg, want.sum(axis=0) / want.shape[0], rtol=0, atol=1e-11)


@pytest.mark.parametrize("n,t,apn", [(333, 37, 13), (150, 700, 16), (40, 3, 5)])
def test_com_fused_column_sums(ctx, n, t, apn):
    """wisp_dev_com with d_colsum: the column sums come out of the streaming COM kernel itself
    (per frame-class partials, fixed-order finish) and equal the sums of the rows it wrote."""
    import torch
    from wisp_mdanalysis_implementation_b200.pipeline import FramePipeline
    system, coords = synth.synth(n, t, apn, seed=5 + n)
    pipe = FramePipeline(ctx, n, system.n_atoms, system.masses, system.node_ptr, system.atom_idx,
                         system.align_idx, frames_local=t)
    assert pipe.plan_info["streaming"]
    pipe.d_colsum.fill_(float("nan"))
    pipe.d_colsum[3 * n:] = 0.0
    pipe.com(torch.from_numpy(coords).cuda(), colsum=True)
    torch.cuda.synchronize()
    x = pipe.d_x.cpu().numpy()[:t, :3 * n]
    want, _ = oracle.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx)
    np.testing.assert_allclose(x, want.reshape(t, 3 * n), rtol=0, atol=1e-11)
    got = pipe.d_colsum.cpu().numpy()
    np.testing.assert_allclose(got[:3 * n], x.sum(axis=0), rtol=1e-13, atol=1e-10)
    assert np.all(got[3 * n:] == 0.0)


def test_driver_end_to_end(ctx, tmp_path):
    """`python wisp_w_mdanalysis.py <config>` with every stage pointed at the B200 modules
    (config keys *_functions_file), on a stand-in universe; compared with the oracle."""
    import importlib
    import sys
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "wisp_mdanalysis_implementation_b200")
    n, t, apn = 48, 160, 5
    system, coords = synth.synth(n, t, apn, seed=31)
    np.save(tmp_path / "traj.npy", coords)
    cfg = tmp_path / "run.config"
    out = tmp_path / "out"
    cfg.write_text("""
output_directory = %r
node_selection_file = 'make_node_selections.py'
trajectory_functions_file = 'trajectory_analysis_functions.py'
adjacency_matrix_functions_file = 'adjacency_matrix_functions.py'
func_adjacency_matrix_functions_file = 'functionalize_adj_matrix_functions.py'
network_functions_file = 'network_analysis_functions.py'
visualization_functions_file = None
trajectory_analysis_boolean = True
weight_by_contact_map_boolean = True
summary_boolean = True
pdb = 'synthetic'
visualization_frame_pdb = 'synthetic'
universe_factory = 'wisp_mdanalysis_implementation_b200.mda_standin:synthetic_universe'
synthetic = dict(n_nodes=%d, n_frames=%d, atoms_per_node=%d, seed=31)
substrate_node_definition = 'COM'
substrate_selection_string = 'protein'
alignment_selection = 'protein and name CA'
traj_list = [%r]
contact_map_distance_cutoff = 12.0
which_contact_map = 'binary contact map'
adjacency_matrix_style = 'pearson correlation'
source_selection_string_list = ['ALA_3']
sink_selection_string_list = ['ALA_40']
number_of_paths = 12
""" % (str(out), n, t, apn, str(tmp_path / "traj.npy")))
    sys.path.insert(0, pkg)
    try:
        for m in ("IO", "wisp_w_mdanalysis", "trajectory_analysis_functions", "adjacency_matrix_functions",
                  "functionalize_adj_matrix_functions", "network_analysis_functions", "make_node_selections"):
            sys.modules.pop(m, None)
        driver = importlib.import_module("wisp_w_mdanalysis")
        paths = driver.main(str(cfg), argv=["wisp_w_mdanalysis.py", str(cfg)])
    finally:
        sys.path.remove(pkg)
    # oracle on the same inputs
    nodes, avg = oracle.traj_alignment_and_averaging(coords.copy(), system.masses, system.node_ptr,
                                                     system.atom_idx, system.align_idx)
    _, _, corr = oracle.pearson_correlation_analysis(oracle.cartesian_covariance(nodes, avg))
    binary, _ = oracle.calc_contact_map(nodes, 12.0)
    W = oracle.func_pearson_correlation(corr, binary)
    want = oracle.get_paths_pruned(W, [2], [39], 12)
    assert [p[1:] for p in paths] == [[int(x) for x in p[1:]] for p in want]
    np.testing.assert_allclose([p[0] for p in paths], [p[0] for p in want], rtol=1e-9)
    for name in ("average_node_positions.dat", "cartesian_covariance.dat", "binary_contact_map.dat",
                 "avg_distance_contact_map.dat", "node_correlation.dat", "func_pearson_correlation.dat",
                 "simply_formatted_paths.txt", "node_selections.txt", "summary.txt"):
        assert (out / name).exists(), name
    got_corr = np.loadtxt(out / "node_correlation.dat")
    assert_corr_close(got_corr, corr)
    lines = (out / "simply_formatted_paths.txt").read_text().splitlines()
    assert lines[0].startswith("# Pathway_Length") and len(lines) == len(paths) + 1
    # the fused device pass (config key fused_pipeline_boolean) gives the same paths
    cfg2 = tmp_path / "run_fused.config"
    cfg2.write_text(cfg.read_text().replace(str(out), str(tmp_path / "out_fused")) + "\nfused_pipeline_boolean = True\n")
    sys.path.insert(0, pkg)
    try:
        paths2 = driver.main(str(cfg2), argv=["wisp_w_mdanalysis.py", str(cfg2)])
    finally:
        sys.path.remove(pkg)
    assert [p[1:] for p in paths2] == [p[1:] for p in paths]
    np.testing.assert_allclose([p[0] for p in paths2], [p[0] for p in paths], rtol=1e-9)
    assert_corr_close(np.loadtxt(tmp_path / "out_fused" / "node_correlation.dat"), corr)


def test_frame_pipeline_matches_c_abi_pipeline(ctx):
    """pipeline.FramePipeline (device API + torch plumbing, what bench.py and the multi-GPU
    path run) against the single-call C-ABI wisp_pipeline and a host round trip."""
    import torch
    from wisp_mdanalysis_implementation_b200.pipeline import FramePipeline
    system, coords = synth.synth(300, 700, 5, seed=12)
    ref = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                       cutoff=13.0, which_map=1)
    pipe = FramePipeline(ctx, 300, system.n_atoms, system.masses, system.node_ptr, system.atom_idx,
                         system.align_idx, frames_local=700, cutoff=13.0, which_map=1)
    assert pipe.plan_info["streaming"]
    d_coords = torch.from_numpy(coords).cuda()
    pipe.run_device(d_coords)
    torch.cuda.synchronize()
    np.testing.assert_allclose(pipe.d_corr.cpu().numpy(), ref["corr"], rtol=0, atol=1e-13)
    # wisp_pipeline runs K3 block by block under the transfer (tile references = the first frame's positions,
    # FP32 sums promoted per block); the resident pass references the tiles to the mean: FP32-level agreement
    np.testing.assert_allclose(pipe.d_avgdist.cpu().numpy(), ref["avgdist"], rtol=1e-6)
    assert_binary_equal_outside_ties(pipe.d_binary.cpu().numpy(), ref["avgdist"], 13.0, ref["binary"])
    h = torch.from_numpy(coords).pin_memory()
    out = pipe.run_host(h)
    np.testing.assert_allclose(out["corr"], ref["corr"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["w"], ref["w"], rtol=1e-12, atol=1e-300)
    # streamed mode: blocks centred with a provisional reference + parallel-axis correction
    out = pipe.run_host_streamed(h, n_blocks=5)
    assert_corr_close(out["corr"], ref["corr"], rel=1e-9, floor=1e-3)
    assert np.all(np.diagonal(out["corr"]) == 1.0)
    np.testing.assert_allclose(out["avgdist"], ref["avgdist"], rtol=1e-6)
    assert_binary_equal_outside_ties(out["binary"], ref["avgdist"], 13.0, ref["binary"])
    np.testing.assert_allclose(pipe.d_mean.cpu().numpy()[:900], ref["avg"].reshape(-1), rtol=0, atol=1e-11)


def test_frame_pipeline_int8_engine_vs_oracle(ctx):
    """Trajectories of 4096+ frames per rank take the INT8 tensor-core K2 inside FramePipeline
    (device-resident, host-fed and streamed modes); all three against the oracle."""
    import torch
    from wisp_mdanalysis_implementation_b200.pipeline import FramePipeline
    n, t, apn = 140, 4500, 3
    system, coords = synth.synth(n, t, apn, seed=31)
    nodes, _ = oracle.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx)
    _, _, corr = oracle.pearson_correlation_fast(oracle.cartesian_covariance_fast(nodes, nodes.sum(axis=0) / t))
    binary, avgdist = oracle.calc_contact_map_fast(nodes, 13.0)
    pipe = FramePipeline(ctx, n, system.n_atoms, system.masses, system.node_ptr, system.atom_idx,
                         system.align_idx, frames_local=t, cutoff=13.0, which_map=1)
    pipe.run_device(torch.from_numpy(coords).cuda())
    torch.cuda.synchronize()
    got_int8 = pipe.d_corr.cpu().numpy()
    ctx.set_gram_mode("dmma")
    try:
        pipe.gram()
        pipe.reduce_and_epilogue()
        torch.cuda.synchronize()
        got_dmma = pipe.d_corr.cpu().numpy()
    finally:
        ctx.set_gram_mode("auto")
    assert not np.array_equal(got_int8, got_dmma)              # the INT8 engine really ran
    assert np.abs(got_int8 - got_dmma).max() < 1e-9
    h = torch.from_numpy(coords).pin_memory()
    for out in (dict(corr=got_int8, avgdist=pipe.d_avgdist.cpu().numpy(), binary=pipe.d_binary.cpu().numpy()),
                pipe.run_host(h), pipe.run_host_streamed(h, n_blocks=1)):
        assert_corr_close(out["corr"], corr)
        assert np.all(np.diagonal(out["corr"]) == 1.0) and np.array_equal(out["corr"], out["corr"].T)
        np.testing.assert_allclose(out["avgdist"], avgdist, rtol=1e-6, atol=1e-9)
        assert_binary_equal_outside_ties(out["binary"], avgdist, 13.0, binary)


def test_streamed_blocks_on_the_int8_engine(ctx):
    """run_host_streamed with two 4100-frame blocks: each block is centred with the provisional
    reference, goes through the INT8 K2 engine on the side stream, and the parallel-axis term is
    removed in the epilogue -- against the oracle on the whole trajectory."""
    import torch
    from wisp_mdanalysis_implementation_b200.pipeline import FramePipeline
    n, t, apn = 130, 8200, 3
    system, coords = synth.synth(n, t, apn, seed=44)
    nodes, _ = oracle.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx)
    _, _, corr = oracle.pearson_correlation_fast(oracle.cartesian_covariance_fast(nodes, nodes.sum(axis=0) / t))
    binary, avgdist = oracle.calc_contact_map_fast(nodes, 13.0)
    pipe = FramePipeline(ctx, n, system.n_atoms, system.masses, system.node_ptr, system.atom_idx,
                         system.align_idx, frames_local=t, cutoff=13.0, which_map=1)
    # the returned arrays are views of the pipeline's pinned output buffers: copy before the next run
    out = {k: v.copy() for k, v in pipe.run_host_streamed(torch.from_numpy(coords).pin_memory(), n_blocks=2).items()}
    assert_corr_close(out["corr"], corr)
    assert np.all(np.diagonal(out["corr"]) == 1.0)
    np.testing.assert_allclose(out["avgdist"], avgdist, rtol=1e-6, atol=1e-9)
    assert_binary_equal_outside_ties(out["binary"], avgdist, 13.0, binary)
    ctx.set_gram_mode("dmma")
    try:
        out_f64 = pipe.run_host_streamed(torch.from_numpy(coords).pin_memory(), n_blocks=2)
    finally:
        ctx.set_gram_mode("auto")
    assert not np.array_equal(out["corr"], out_f64["corr"])          # the INT8 engine really ran
    assert np.abs(out["corr"] - out_f64["corr"]).max() < 1e-9


def test_paths_prepare_query_stress(ctx):
    """configs[4] in miniature: one APSP, many independent (source, sink) queries."""
    W = synth.random_cost_matrix(60, seed=77, density=0.12, lo=0.05, hi=0.9)
    ctx.paths_prepare(W)
    rng = np.random.default_rng(5)
    for _ in range(12):
        s, t = (int(x) for x in rng.choice(60, size=2, replace=False))
        K = int(rng.integers(3, 60))
        want = oracle.get_paths_pruned(W, [s], [t], K)
        got = ctx.paths_query([s], [t], K)
        assert [p[1:] for p in got] == [[int(x) for x in p[1:]] for p in want]
        assert [p[0] for p in got] == [float(p[0]) for p in want]


def test_pair_sharded_queries(ctx):
    """paths_sharded.query_pairs (SURVEY 8(e), configs[4]) on one rank equals per-pair get_paths
    and the oracle; the multi-rank gather is covered on gloo in test_sharding_gloo.py."""
    from wisp_mdanalysis_implementation_b200 import paths_sharded
    W = synth.random_cost_matrix(50, seed=12, density=0.15, lo=0.05, hi=0.9)
    pairs = [(3, 41), (0, 17), (29, 5), (11, 48)]
    got = paths_sharded.query_pairs(ctx, W, pairs, 7)
    for (s, t), paths in zip(pairs, got):
        assert paths == ctx.get_paths(W, [s], [t], 7)
        want = oracle.get_paths_pruned(W, [s], [t], 7)
        assert [p[1:] for p in paths] == [[int(x) for x in p[1:]] for p in want]


def test_apsp_fp32_dense_predecessors(ctx):
    """Dense graph: the FP32 mode recovers predecessors with the tiled argmin product."""
    W = synth.random_cost_matrix(150, seed=3, density=0.95)
    D, _ = oracle.apsp_dijkstra(W)
    dist, pred = ctx.apsp(W, precision=32)
    np.testing.assert_allclose(dist, D, rtol=2e-6)
    ii, jj = np.nonzero(pred >= 0)
    pp = pred[ii, jj]
    assert np.all(W[pp, jj] != 0)
    np.testing.assert_allclose(dist[ii, pp] + W[pp, jj], dist[ii, jj], rtol=2e-5, atol=1e-12)
    off = ~np.eye(150, dtype=bool)
    assert np.all(pred[off] >= 0) and np.all(np.diagonal(pred) == -1)


@pytest.mark.parametrize("tag", ["traj_small", "traj_aligned"])
def test_golden_lmi(ctx, golden_dir, tag):
    """SURVEY 8f row 4 against the reference's own LMI output."""
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    mi = ctx.lmi_from_covariance(g["cov"])
    assert np.all(np.isinf(np.diagonal(mi))) and np.array_equal(mi, mi.T)
    off = ~np.eye(mi.shape[0], dtype=bool)
    np.testing.assert_allclose(mi[off], g["mi"][off], rtol=1e-9, atol=1e-12)
    rmi, func = ctx.func_generalized_correlation(g["mi"])
    np.testing.assert_allclose(rmi, g["rmi"], rtol=1e-12)
    np.testing.assert_allclose(func[off], g["func_mi"][off], rtol=1e-11, atol=1e-15)
    assert np.all(np.diagonal(func) == 0.0)
    rmi_d, func_d = ctx.func_generalized_correlation(g["mi"], g["avgdist"], 5.0)
    np.testing.assert_allclose(rmi_d, g["rmi_damped"], rtol=1e-12)
    np.testing.assert_allclose(func_d[off], g["func_mi_damped"][off], rtol=1e-11, atol=1e-15)


def test_tiny_and_ragged_inputs(ctx):
    """Smallest legal sizes and sizes that straddle every tile boundary."""
    from wisp_mdanalysis_implementation_b200 import _lib
    for n, t, apn in [(1, 1, 1), (2, 3, 1), (3, 2, 2), (127, 23, 1), (129, 25, 2)]:
        system, coords = synth.synth(n, t, apn, seed=100 + n)
        out = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                           cutoff=8.0, which_map=2, want_nodes=True)
        nodes, _ = oracle.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx)
        np.testing.assert_allclose(out["nodes"], nodes, rtol=0, atol=1e-11)
        binary, avgdist = oracle.calc_contact_map_fast(nodes, 8.0)
        np.testing.assert_allclose(out["avgdist"], avgdist, rtol=1e-6, atol=1e-9)
        if t > 2:
            avg = nodes.sum(axis=0) / t
            _, _, corr = oracle.pearson_correlation_fast(oracle.cartesian_covariance_fast(nodes, avg))
            ok = np.isfinite(corr)
            assert_corr_close(out["corr"][ok], corr[ok], rel=1e-6, floor=1e-3)
    # error paths fail loudly with a message (no silent fallback)
    system, coords = synth.synth(4, 5, 2, seed=1)
    with pytest.raises((_lib.WispError, ValueError)):
        ctx.pipeline(coords, system.masses[:-1], system.node_ptr, system.atom_idx, system.align_idx)
    with pytest.raises(_lib.WispError):
        bad_ptr = system.node_ptr.copy(); bad_ptr[2] = bad_ptr[1]          # empty node
        ctx.com_nodes(coords, system.masses, bad_ptr, system.atom_idx, system.align_idx)
    with pytest.raises(_lib.WispError):
        bad_idx = system.atom_idx.copy(); bad_idx[0] = 999                 # atom out of range
        ctx.com_nodes(coords, system.masses, system.node_ptr, bad_idx, system.align_idx)
    with pytest.raises(_lib.WispError):
        W = synth.random_cost_matrix(6, seed=1); W[1, 2] = np.nan
        ctx.get_paths(W, [0], [5], 2)
    with pytest.raises(_lib.WispError):
        ctx.get_paths(synth.random_cost_matrix(6, seed=1), [0], [9], 2)    # sink out of range


def test_apsp_fp32_multi_tile_vs_fp64(ctx):
    """FP32 throughput mode (128-wide rounds, FADD2 + FMNMX3) against the FP64 parity mode on a
    graph spanning many tiles."""
    W = synth.random_cost_matrix(1100, seed=9, density=0.02)
    d64, p64 = ctx.apsp(W, precision=64)
    d32, p32 = ctx.apsp(W, precision=32)
    np.testing.assert_allclose(d32, d64, rtol=3e-6)
    rng = np.random.default_rng(0)
    i, j, k = rng.integers(0, 1100, size=(3, 5000))
    assert np.all(d64[i, j] <= d64[i, k] + d64[k, j] + 1e-9)
    # walking FP32 predecessors reproduces the FP32 distance
    for s, t in zip(i[:50], j[:50]):
        if s == t:
            continue
        node, length, hops = t, 0.0, 0
        while node != s:
            prev = p32[s, node]
            assert prev >= 0
            length += W[prev, node]
            node = prev
            hops += 1
            assert hops < 1100
        assert abs(length - d64[s, t]) <= 1e-4 * max(1.0, d64[s, t])


@pytest.mark.parametrize("world", [1, 2, 3])
def test_sharded_apsp_protocol(ctx, world):
    """Row-block sharded FP32 Floyd-Warshall: the per-round pivot-panel protocol driven rank by
    rank inside one process (the broadcast replaced by a device copy) equals the single-GPU APSP."""
    import torch
    from wisp_mdanalysis_implementation_b200.apsp_sharded import ShardedAPSP, TILE
    n = 700
    W = synth.random_cost_matrix(n, seed=44, density=0.03)
    want, _ = ctx.apsp(W, precision=32, want_pred=False)
    d_W = torch.from_numpy(W).cuda()
    ranks = [ShardedAPSP(ctx, n, rank=r, world=world) for r in range(world)]
    for sh in ranks:
        sh.load_rows(d_W[sh.t0 * TILE:min(n, sh.t1 * TILE)].contiguous())
    for kb in range(ranks[0].n_tiles):
        owner = ranks[ranks[0].owner(kb)]
        owner.round_pivot(kb)
        for sh in ranks:
            sh.round_update(kb, panel=owner.d_panel)
    torch.cuda.synchronize()
    got = torch.cat([sh.local_rows() for sh in ranks]).cpu().numpy().astype(np.float64)
    assert got.shape == (n, n)
    assert np.array_equal(got, want)


def test_get_paths_fp32_apsp_bounds(ctx):
    """FP32 APSP is only used for pruning bounds: the enumerated set stays exact."""
    W = synth.random_cost_matrix(80, seed=123, density=0.1, lo=0.05, hi=0.9)
    for s, t, K in [(3, 70, 25), (10, 11, 40)]:
        want = oracle.get_paths_pruned(W, [s], [t], K)
        got = ctx.get_paths(W, [s], [t], K, apsp_fp32=True)
        assert [p[1:] for p in got] == [[int(x) for x in p[1:]] for p in want]
        assert [p[0] for p in got] == [float(p[0]) for p in want]


def _device_traj(n, t, seed, spread=True):
    """Centred node trajectory in the device layout [frames_pad][node_ld] (float64)."""
    import torch
    from wisp_mdanalysis_implementation_b200 import _lib
    ld, tp = _lib.node_ld(n), _lib.frames_pad(t)
    rng = np.random.default_rng(seed)
    amp = np.exp(rng.normal(size=3 * n)) if spread else np.ones(3 * n)
    xh = rng.normal(size=(t, 3 * n)) * amp
    xh -= xh.mean(axis=0)
    x = torch.zeros(tp, ld, dtype=torch.float64, device="cuda")
    x[:t, :3 * n] = torch.from_numpy(xh).cuda()
    return x, xh, ld


def _gram(ctx, mode, x, t, ld, n, forced=False):
    import torch
    from wisp_mdanalysis_implementation_b200 import _lib
    npad = _lib.gram_ld(n)
    g = torch.full((npad, npad), np.nan, dtype=torch.float64, device="cuda")
    ctx.set_gram_mode(mode)
    try:
        if forced:
            ctx.dev_gram_i8(x.data_ptr(), t, ld, n, g.data_ptr(), stream=1)
        else:
            ctx.dev_gram(x.data_ptr(), t, ld, n, 3, g.data_ptr(), stream=1)
        torch.cuda.synchronize()
    finally:
        ctx.set_gram_mode("auto")
    return np.triu(g.cpu().numpy()[:n, :n])


@pytest.mark.parametrize("n,t", [(130, 64), (300, 5000), (640, 12007)])
def test_gram_int8_tensor_core_path(ctx, n, t):
    """K2 on tcgen05 INT8 (digit planes, exact integer GEMMs) against float64 truth and the DMMA
    kernel: |dG| <= 1e-9 sqrt(G_ii G_jj), i.e. 1e-9 absolute on the correlation -- the 1e-6
    relative bar at the |C| floor of 1e-3 (module docstring)."""
    x, xh, ld = _device_traj(n, t, seed=n + t)
    g_dmma = _gram(ctx, "dmma", x, t, ld, n)
    g_i8 = _gram(ctx, "i8", x, t, ld, n, forced=True)
    nodes = xh.reshape(t, n, 3)
    truth = np.triu(np.einsum("tic,tjc->ij", nodes, nodes))
    dg = np.sqrt(np.outer(np.diag(truth), np.diag(truth)))
    assert (np.abs(g_dmma - truth) / dg).max() < 1e-12
    assert (np.abs(g_i8 - truth) / dg).max() < 1e-9
    np.testing.assert_allclose(np.diag(g_i8), np.diag(truth), rtol=1e-12)     # diagonal is plain FP64
    # deterministic: same bits on a second run
    assert np.array_equal(g_i8, _gram(ctx, "i8", x, t, ld, n, forced=True))


def test_gram_int8_accuracy_gate(ctx):
    """Dispatcher: long smooth trajectories take the INT8 kernel, spiky data fails the on-device
    accuracy bound and the DMMA kernel behind the gate produces the result instead."""
    n, t = 200, 4500
    x, xh, ld = _device_traj(n, t, seed=3, spread=False)
    g_dmma = _gram(ctx, "dmma", x, t, ld, n)
    g_auto = _gram(ctx, "auto", x, t, ld, n)
    g_i8 = _gram(ctx, "i8", x, t, ld, n, forced=True)
    assert np.array_equal(g_auto, g_i8) and not np.array_equal(g_auto, g_dmma)
    x[11, 7] = 1e6                                   # one wild frame in one coordinate
    g_dmma = _gram(ctx, "dmma", x, t, ld, n)
    assert np.array_equal(_gram(ctx, "auto", x, t, ld, n), g_dmma)
    assert np.array_equal(_gram(ctx, "i8", x, t, ld, n), g_dmma)
    # short trajectories stay on DMMA in auto mode
    x2, _, ld2 = _device_traj(64, 500, seed=4)
    assert np.array_equal(_gram(ctx, "auto", x2, 500, ld2, 64), _gram(ctx, "dmma", x2, 500, ld2, 64))


def test_pipeline_int8_gram_vs_oracle(ctx):
    """Whole stage-4..6 pipeline with K2 forced onto the INT8 tensor cores, against the oracle."""
    n, t, apn = 257, 1500, 5
    system, coords = synth.synth(n, t, apn, seed=21)
    ctx.set_gram_mode("i8")
    try:
        out = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                           cutoff=15.0, which_map=1, want_nodes=True)
    finally:
        ctx.set_gram_mode("auto")
    nodes, _ = oracle.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx)
    _, _, corr = oracle.pearson_correlation_fast(oracle.cartesian_covariance_fast(nodes, nodes.sum(axis=0) / t))
    assert_corr_close(out["corr"], corr)
    assert np.all(np.diagonal(out["corr"]) == 1.0) and np.array_equal(out["corr"], out["corr"].T)


def test_known_answers_through_the_c_abi(ctx):
    """Closed-form cases (tests/known_answers.py): perfectly correlated / mirrored /
    orthogonal nodes, a rigid translation, an exact 3-4-5 contact geometry with the cutoff on a
    distance, and a hand-enumerated graph with the node-0 rule and the K+1 return."""
    from known_answers import (LADDER_PATHS_1_TO_5, correlated_trajectory, ladder_graph,
                               lattice_contact_case, rigid_translation_trajectory)
    nodes, want = correlated_trajectory()
    avg = nodes.sum(axis=0) / nodes.shape[0]
    _, _, corr = ctx.pearson_from_covariance(ctx.cartesian_covariance(nodes, avg))
    np.testing.assert_allclose(corr, want, rtol=0, atol=1e-15)
    assert np.all(np.diagonal(corr) == 1.0)
    w = ctx.func_pearson(np.where(np.abs(want) == 1.0, want, 0.0), None)
    assert np.all(w[:3, :3] == 0.0) and np.all(np.isinf(w[3, :3])) and w[3, 3] == 0.0

    nodes, base = rigid_translation_trajectory()
    avg = nodes.sum(axis=0) / nodes.shape[0]
    _, _, corr = ctx.pearson_from_covariance(ctx.cartesian_covariance(nodes, avg))
    np.testing.assert_allclose(corr, np.ones_like(corr), rtol=0, atol=1e-11)
    binary, avgdist = ctx.contact_map(nodes, 7.5)
    dist = np.sqrt(((base[:, None, :] - base[None, :, :]) ** 2).sum(-1))
    np.testing.assert_allclose(avgdist, dist, rtol=1e-6, atol=1e-9)
    assert_binary_equal_outside_ties(binary, dist, 7.5, ((dist < 7.5) & ~np.eye(len(base), dtype=bool)).astype(np.int64))

    nodes, cutoff, binary_want, d = lattice_contact_case()
    binary, avgdist = ctx.contact_map(nodes, cutoff)
    np.testing.assert_allclose(avgdist, d, rtol=1e-6, atol=1e-9)
    assert_binary_equal_outside_ties(binary, d, cutoff, binary_want)     # |p0 - p2| == cutoff is the tie
    assert binary[0, 1] == 1 and binary[1, 2] == 1 and binary[0, 3] == 0 and np.all(np.diagonal(binary) == 0)

    W = ladder_graph()
    assert ctx.get_paths(W, [1], [5], 2) == LADDER_PATHS_1_TO_5[:3]       # K + 1: the seed rides along
    assert ctx.get_paths(W, [1], [5], 3) == LADDER_PATHS_1_TO_5[:4]
    assert ctx.get_paths(W, [0], [5], 1)[0] == [0.125, 0, 5]
    assert ctx.get_paths(W, [1], [5], 4, node0_quirk=False) == LADDER_PATHS_1_TO_5[:4]


# ---------------------------------------------------------------------------------------
# parity at the sizes bench.py publishes numbers for (VERDICT r1 "What's weak" 1-3)
# ---------------------------------------------------------------------------------------
def _oracle_on_node_subset(system, coords, sel, cutoff):
    """Oracle correlation / mean distances for the nodes `sel` only: every entry (i, j) of
    both matrices depends on nodes i and j alone, so the oracle run on a node subset IS the
    corresponding sub-block of the full result."""
    ptr = np.concatenate([[0], np.cumsum(np.diff(system.node_ptr)[sel])]).astype(np.int64)
    idx = np.concatenate([system.atom_idx[system.node_ptr[i]:system.node_ptr[i + 1]] for i in sel])
    nodes, _ = oracle.com_nodes_fast(coords, system.masses, ptr, idx, system.align_idx)
    avg = nodes.sum(axis=0) / nodes.shape[0]
    _, _, corr = oracle.pearson_correlation_fast(oracle.cartesian_covariance_fast(nodes, avg))
    binary, avgdist = oracle.calc_contact_map_fast(nodes, cutoff)
    return nodes, corr, binary, avgdist


def _check_subset(out, sel, corr, binary, avgdist, cutoff):
    ix = np.ix_(sel, sel)
    assert_corr_close(out["corr"][ix], corr)
    np.testing.assert_allclose(out["avgdist"][ix], avgdist, rtol=1e-6, atol=1e-9)
    assert_binary_equal_outside_ties(out["binary"][ix], avgdist, cutoff, binary)


def test_published_config_c2_vs_oracle(ctx):
    """configs[1] geometry -- 2,000 nodes x 16 atoms, 16 node tiles with a partial last tile --
    at 8,200 frames: the tcgen05 INT8 K2 engine, the K3 frame segments and the tile-referenced
    FP32 copy all run exactly as in the bench.  wisp_pipeline, FramePipeline.run_device and
    run_host_streamed against the oracle on 288 nodes drawn from every tile (82,944 of the
    4,000,000 matrix entries; each entry depends on its two nodes only)."""
    import torch
    from wisp_mdanalysis_implementation_b200.pipeline import FramePipeline
    n, t, apn, cutoff = 2000, 8200, 16, 15.0
    system = synth.SynthSystem(n, t, apn, seed=1002)
    coords = system.coords()
    rng = np.random.default_rng(5)
    sel = np.unique(np.concatenate([rng.choice(n, size=250, replace=False),
                                    [0, 1, 127, 128, 129, 1919, 1920, 1921, 1998, 1999]]))
    _, corr, binary, avgdist = _oracle_on_node_subset(system, coords, sel, cutoff)
    out = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                       cutoff=cutoff, which_map=1)
    _check_subset(out, sel, corr, binary, avgdist, cutoff)
    assert np.all(np.diagonal(out["corr"]) == 1.0) and np.array_equal(out["corr"], out["corr"].T)
    pipe = FramePipeline(ctx, n, system.n_atoms, system.masses, system.node_ptr, system.atom_idx,
                         system.align_idx, frames_local=t, cutoff=cutoff, which_map=1)
    h = torch.from_numpy(coords).pin_memory()
    d_coords = h.cuda()
    pipe.run_device(d_coords)
    torch.cuda.synchronize()
    dev = dict(corr=pipe.d_corr.cpu().numpy(), avgdist=pipe.d_avgdist.cpu().numpy(),
               binary=pipe.d_binary.cpu().numpy())
    _check_subset(dev, sel, corr, binary, avgdist, cutoff)
    # the INT8 engine really ran: the FP64 DMMA engine gives different bits, same numbers
    ctx.set_gram_mode("dmma")
    try:
        pipe.gram()
        pipe.reduce_and_epilogue()
        torch.cuda.synchronize()
        f64 = pipe.d_corr.cpu().numpy()
    finally:
        ctx.set_gram_mode("auto")
    assert not np.array_equal(f64, dev["corr"]) and np.abs(f64 - dev["corr"]).max() < 1e-9
    del d_coords
    _check_subset(pipe.run_host_streamed(h, n_blocks=2), sel, corr, binary, avgdist, cutoff)


@pytest.mark.parametrize("kind", ["bimodal", "heavy_tailed", "drift"])
def test_gram_int8_gate_on_md_like_columns(ctx, kind):
    """The INT8 engine's accuracy gate on what real MD columns look like:
    side chains flipping between two rotamers (bimodal), heavy-tailed excursions (Student t,
    3 degrees of freedom) and slow drift.  Whatever the dispatcher decides, the result must meet
    the bar |dG| <= 1e-9 sqrt(G_ii G_jj); when the gate keeps the tensor-core path the bits differ
    from DMMA, when it trips they are DMMA's."""
    import torch
    from wisp_mdanalysis_implementation_b200 import _lib
    n, t = 260, 6000
    rng = np.random.default_rng({"bimodal": 1, "heavy_tailed": 2, "drift": 3}[kind])
    if kind == "bimodal":
        state = np.cumsum(rng.random((t, 3 * n)) < 0.002, axis=0) % 2          # rare flips
        xh = (2.0 * state - 1.0) * rng.uniform(0.5, 3.0, size=3 * n) + rng.normal(scale=0.3, size=(t, 3 * n))
        xh[:, ::7] = rng.normal(scale=0.4, size=(t, len(range(0, 3 * n, 7))))  # and some quiet columns
    elif kind == "heavy_tailed":
        xh = rng.standard_t(3, size=(t, 3 * n)) * rng.uniform(0.2, 2.0, size=3 * n)
    else:
        xh = np.cumsum(rng.normal(scale=0.05, size=(t, 3 * n)), axis=0) + rng.normal(scale=0.3, size=(t, 3 * n))
    xh -= xh.mean(axis=0)
    ld, tp = _lib.node_ld(n), _lib.frames_pad(t)
    x = torch.zeros(tp, ld, dtype=torch.float64, device="cuda")
    x[:t, :3 * n] = torch.from_numpy(xh).cuda()
    nodes = xh.reshape(t, n, 3)
    truth = np.triu(np.einsum("tic,tjc->ij", nodes, nodes))
    dg = np.sqrt(np.outer(np.diag(truth), np.diag(truth)))
    g_dmma = _gram(ctx, "dmma", x, t, ld, n)
    g_auto = _gram(ctx, "auto", x, t, ld, n)
    assert (np.abs(g_dmma - truth) / dg).max() < 1e-12
    assert (np.abs(g_auto - truth) / dg).max() < 1e-9, kind
    # the un-gated tensor-core kernel on the same data: how far the bound is from being needed
    g_forced = _gram(ctx, "i8", x, t, ld, n, forced=True)
    err_forced = (np.abs(g_forced - truth) / dg).max()
    took_i8 = not np.array_equal(g_auto, g_dmma)
    if not took_i8:
        assert np.array_equal(g_auto, g_dmma)
    else:
        assert np.array_equal(g_auto, g_forced)
    print("%s: engine=%s, forced-int8 error %.2e" % (kind, "int8" if took_i8 else "dmma (gate)", err_forced))


@pytest.mark.parametrize("precision,tol", [(64, 1e-12), (32, 2e-6)])
def test_apsp_c3_size_vs_scipy_dijkstra(ctx, precision, tol):
    """configs[2] size: 10,000 nodes, sparse contact-map-like graph (mean degree 10), against
    scipy's C Dijkstra from 300 sources (rows of the all-pairs result)."""
    import scipy.sparse
    import scipy.sparse.csgraph
    n = 10000
    W = synth.random_cost_matrix(n, seed=1003, density=10.0 / n, lo=0.05, hi=3.0)
    rng = np.random.default_rng(0)
    src = np.sort(rng.choice(n, size=300, replace=False))
    want, want_pred = scipy.sparse.csgraph.dijkstra(scipy.sparse.csr_matrix(W), directed=True, indices=src,
                                                    return_predecessors=True)
    dist, pred = ctx.apsp(W, precision=precision)
    assert np.all(np.isfinite(dist))
    np.testing.assert_allclose(dist[src], want, rtol=tol, atol=1e-300)
    assert np.all(np.diagonal(dist) == 0) and np.all(np.diagonal(pred) == -1)
    # predecessors: every pred is a real edge on a shortest path; in FP64 (unique shortest paths
    # on random real weights) they are the same tree scipy found
    jj = np.arange(n)
    for r, s in enumerate(src[:40]):
        p = pred[s]
        ok = jj != s
        assert np.all(p[ok] >= 0)
        assert np.all(W[p[ok], jj[ok]] != 0)
        np.testing.assert_allclose(dist[s, p[ok]] + W[p[ok], jj[ok]], dist[s, ok], rtol=10 * tol, atol=1e-12)
        if precision == 64:
            assert np.array_equal(p[ok], want_pred[r][ok])


# ---------------------------------------------------------------------------------------
# wisp_pipeline on a multi-GPU context (wisp_ctx_create_multi): frames sharded inside the one C
# call, K7 alignment with per-iteration host reduction, partial sums combined inside the fused
# peer-memory epilogue.  On a one-GPU box the shards share the device (same code path: host
# threads, barriers, reductions, row-sliced epilogue reading every shard's partials).
# ---------------------------------------------------------------------------------------
def _multi_ctx(n_shards):
    import torch
    from wisp_mdanalysis_implementation_b200 import _lib
    n_dev = torch.cuda.device_count()
    return _lib.Context(gpus=[k % n_dev for k in range(n_shards)])


@pytest.mark.parametrize("shards,n,t,apn", [(2, 150, 700, 5), (3, 257, 1001, 4), (8, 300, 999, 3)])
def test_pipeline_multi_shard_with_alignment(ctx, shards, n, t, apn):
    system, coords = synth.synth(n, t, apn, seed=50 + shards)
    # a slow global tumble so that the alignment has something to do
    ang = np.linspace(0.0, 0.6, t)
    c, s = np.cos(ang), np.sin(ang)
    rot = np.zeros((t, 3, 3)); rot[:, 0, 0] = c; rot[:, 0, 1] = -s; rot[:, 1, 0] = s; rot[:, 1, 1] = c; rot[:, 2, 2] = 1
    centre = coords.mean(axis=1, keepdims=True)
    coords = (np.einsum("tac,tdc->tad", coords - centre, rot) + centre).astype(np.float32)
    cutoff = 13.0
    mctx = _multi_ctx(shards)
    assert mctx.n_gpus == shards
    one = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=cutoff,
                       which_map=1, want_nodes=True, maximum_num_iterations=100)
    got = mctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=cutoff,
                        which_map=1, want_nodes=True, maximum_num_iterations=100)
    assert len(got["alignment_log"]) == len(one["alignment_log"]) >= 2
    assert got["alignment_log"][-1][1] <= 1e-5
    # sharded vs single GPU: same arithmetic up to the order of the frame sums
    # (the float32 initial landmark sum is chained over the shards in frame order: bit-identical start)
    np.testing.assert_allclose(got["nodes"], one["nodes"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(got["avg"], one["avg"], rtol=0, atol=1e-9)
    assert_corr_close(got["corr"], one["corr"], rel=1e-8)
    # K3 sums FP32 per-frame distances in FP32 over 128-frame groups before promoting to FP64; the groups
    # fall differently in a shard, so the mean distances agree to the stated 1e-6, not to FP64 rounding
    np.testing.assert_allclose(got["avgdist"], one["avgdist"], rtol=1e-6)
    assert np.all(np.diagonal(got["corr"]) == 1.0) and np.array_equal(got["corr"], got["corr"].T)
    # and against the oracle's traj_alignment_and_averaging -> covariance -> Pearson / contact map
    nodes, avg = oracle.traj_alignment_and_averaging(coords.copy(), system.masses, system.node_ptr,
                                                     system.atom_idx, system.align_idx)
    np.testing.assert_allclose(got["nodes"], nodes, rtol=0, atol=1e-6)
    _, _, corr = oracle.pearson_correlation_fast(oracle.cartesian_covariance_fast(nodes, avg))
    binary, avgdist = oracle.calc_contact_map_fast(nodes, cutoff)
    assert_corr_close(got["corr"], corr)
    np.testing.assert_allclose(got["avgdist"], avgdist, rtol=1e-6, atol=1e-9)
    assert_binary_equal_outside_ties(got["binary"], avgdist, cutoff, binary)
    w = oracle.func_pearson_correlation(corr, binary)
    if np.array_equal(got["binary"], binary):
        nz = w != 0
        assert np.array_equal(got["w"] != 0, nz)
        np.testing.assert_allclose(got["w"][nz], w[nz], rtol=0, atol=2e-6)
    mctx.close()


def test_pipeline_multi_shard_without_alignment_and_errors(ctx):
    from wisp_mdanalysis_implementation_b200 import _lib
    system, coords = synth.synth(130, 500, 4, seed=77)
    mctx = _multi_ctx(4)
    one = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=12.0, which_map=2)
    got = mctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=12.0, which_map=2)
    assert got["alignment_log"] == []
    assert_corr_close(got["corr"], one["corr"], rel=1e-9)
    np.testing.assert_allclose(got["avgdist"], one["avgdist"], rtol=1e-6)
    np.testing.assert_allclose(got["w"], one["w"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(got["avg"], one["avg"], rtol=0, atol=1e-11)
    # fewer frames than shards: the call still works (shards beyond the frame count stay idle)
    tiny = mctx.pipeline(coords[:3], system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=12.0)
    ref = ctx.pipeline(coords[:3], system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=12.0)
    np.testing.assert_allclose(tiny["avgdist"], ref["avgdist"], rtol=1e-6)
    # a failing shard fails the call loudly instead of hanging the others
    with pytest.raises(_lib.WispError):
        bad = system.atom_idx.copy(); bad[5] = 10 ** 6
        mctx.pipeline(coords, system.masses, system.node_ptr, bad, system.align_idx)
    # every other entry point of a multi context runs on its first device
    W = synth.random_cost_matrix(40, seed=3, density=0.2)
    assert mctx.get_paths(W, [1], [30], 5) == ctx.get_paths(W, [1], [30], 5)
    mctx.close()


def test_published_config_c2_multi_shard_with_alignment(ctx):
    """configs[1] geometry through the ONE call a drop-in user makes -- wisp_pipeline on a multi-GPU
    context, alignment included -- against the oracle (alignment on all landmarks, correlation /
    contacts on 200 nodes drawn from every tile).  Two shards of 4,100 frames: each runs the INT8
    tensor-core K2 engine on its own frames."""
    n, t, apn, cutoff = 2000, 8200, 16, 15.0
    system = synth.SynthSystem(n, t, apn, seed=1002)
    coords = system.coords()
    rng = np.random.default_rng(6)
    sel = np.unique(np.concatenate([rng.choice(n, size=190, replace=False), [0, 127, 128, 1919, 1920, 1999]]))
    ptr = np.concatenate([[0], np.cumsum(np.diff(system.node_ptr)[sel])]).astype(np.int64)
    idx = np.concatenate([system.atom_idx[system.node_ptr[i]:system.node_ptr[i + 1]] for i in sel])
    nodes, align = oracle.com_nodes_fast(coords, system.masses, ptr, idx, system.align_idx)
    log = []
    nodes, avg = oracle.iterative_alignment(align, nodes, log=log)
    _, _, corr = oracle.pearson_correlation_fast(oracle.cartesian_covariance_fast(nodes, avg))
    binary, avgdist = oracle.calc_contact_map_fast(nodes, cutoff)
    mctx = _multi_ctx(2)
    out = mctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                        cutoff=cutoff, which_map=1, maximum_num_iterations=100)
    assert len(out["alignment_log"]) == len(log)
    np.testing.assert_allclose(out["avg"][sel], avg, rtol=0, atol=1e-6)
    _check_subset(out, sel, corr, binary, avgdist, cutoff)
    mctx.close()


def test_paths_query_pairs_batched(ctx):
    """wisp_paths_query_pairs: many independent get_paths([s], [t], K) queries served by the same kernel
    passes (configs[4]) -- every query must equal its stand-alone answer and the oracle, including the
    K + 1 returns and K = 1."""
    W = synth.random_cost_matrix(70, seed=78, density=0.12, lo=0.05, hi=0.9)
    ctx.paths_prepare(W)
    rng = np.random.default_rng(9)
    pairs = [tuple(int(x) for x in rng.choice(70, size=2, replace=False)) for _ in range(17)]
    pairs += [pairs[0], (pairs[1][1], pairs[1][0])]           # a repeated and a reversed query
    for K in (1, 7, 60):
        got = ctx.paths_query_pairs(pairs, K)
        assert len(got) == len(pairs)
        for (s, t), paths in zip(pairs, got):
            assert paths == ctx.paths_query([s], [t], K), (s, t, K)
        for (s, t), paths in list(zip(pairs, got))[:5]:
            want = oracle.get_paths_pruned(W, [s], [t], K)
            assert [p[1:] for p in paths] == [[int(x) for x in p[1:]] for p in want]
            assert [p[0] for p in paths] == [float(p[0]) for p in want]
    qoff, lengths, noff, nodes = ctx.paths_query_pairs(pairs, 7, packed=True)
    assert len(qoff) == len(pairs) + 1 and qoff[-1] == len(lengths) and noff[-1] == len(nodes)
    from wisp_mdanalysis_implementation_b200 import _lib
    with pytest.raises(_lib.WispError):
        ctx.paths_query_pairs([(3, 3)], 5)


def test_get_paths_duplicate_pairs(ctx):
    """A repeated index in sources / sinks makes identical paths inside one pass; the reference removes
    them per pass before comparing with K (remove_redundant_paths), so the ratchet must too."""
    W = synth.random_cost_matrix(40, seed=31, density=0.2, lo=0.05, hi=0.9)
    for so, si, K in [([3, 3], [30], 9), ([5], [22, 22], 14), ([3, 3], [30, 30], 6)]:
        want = oracle.get_paths_pruned(W, so, si, K)
        got = ctx.get_paths(W, so, si, K)
        assert [p[1:] for p in got] == [[int(x) for x in p[1:]] for p in want], (so, si, K)
        assert [p[0] for p in got] == [float(p[0]) for p in want]


def test_multirank_nccl_paths():
    """The NCCL multi-rank paths on real ranks (needs >= 2 GPUs; skipped on a one-GPU box, where the same
    host logic is covered by the emulated-shard tests above and, on the CPU, by test_sharding_gloo.py):
    tools/multirank_check.py asserts sharded APSP == single GPU, FramePipeline == wisp_pipeline with K7
    on, query_pairs == serial get_paths, in-process multi-GPU wisp_pipeline == single GPU."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(root, "tools", "multirank_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def test_get_paths_longer_than_the_initial_record(ctx):
    """Paths of 150+ nodes (a chain with three detours): the path record / per-warp path buffer starts at
    96 nodes and grows on demand (it used to be a hard error)."""
    n = 153
    W = np.zeros((n, n))
    for i in range(149):
        W[i, i + 1] = W[i + 1, i] = 0.5
    for k, i in enumerate((10, 50, 90)):
        d = 150 + k
        W[i, d] = W[d, i] = 0.3
        W[d, i + 1] = W[i + 1, d] = 0.3
    for K in (3, 8):
        want = oracle.get_paths_pruned(W, [0], [149], K)
        got = ctx.get_paths(W, [0], [149], K)
        assert [p[1:] for p in got] == [[int(x) for x in p[1:]] for p in want]
        assert [p[0] for p in got] == [float(p[0]) for p in want]
        assert min(len(p) for p in got) > 150
    assert ctx.paths_query_pairs([(0, 149), (149, 0)], 4) == [ctx.get_paths(W, [0], [149], 4), ctx.get_paths(W, [149], [0], 4)]


@pytest.mark.parametrize("n,n_align_step,t", [(70, 1, 45), (257, 3, 130)])
def test_k7_deferred_node_rotation_equals_in_place(ctx, n, n_align_step, t):
    """K7 building blocks: one iteration with the node rotation deferred (wisp_dev_align_iter: landmarks
    rotated in place, nodes only read, rotations composed) + the final in-place rotation
    (wisp_dev_align_nodes) or the fused centring (wisp_dev_center_rotated) must equal the literal data
    flow of trajectory_analysis_functions.py:102-111 (wisp_dev_align_rotate rotates both arrays every
    iteration) -- and both must equal numpy on the same rotations."""
    import torch
    from wisp_mdanalysis_implementation_b200 import _lib
    rng = np.random.default_rng(n)
    n_align = (n + n_align_step - 1) // n_align_step
    ld = _lib.node_ld(n)
    npad = _lib.gram_ld(n)
    base = rng.normal(size=(n, 3)) * np.array([12.0, 7.0, 4.0])
    nodes0 = np.zeros((t, n, 3))
    for f in range(t):
        ang = rng.normal(size=3) * 0.3
        cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
        R = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
             @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
        nodes0[f] = (base + rng.normal(size=(n, 3)) * 0.4) @ R.T
    align0 = nodes0[:, ::n_align_step, :].astype(np.float32)
    assert align0.shape[1] == n_align

    def dev_nodes(a):
        x = torch.zeros((_lib.frames_pad(t), ld), dtype=torch.float64, device="cuda")
        x[:t, :3 * n] = torch.from_numpy(a.reshape(t, 3 * n)).cuda()
        return x

    avg = align0.astype(np.float64).mean(axis=0).reshape(-1)
    # (a) literal: both arrays rotated in place, twice
    xa, aa = dev_nodes(nodes0), torch.from_numpy(align0.copy()).cuda()
    d_avg = torch.from_numpy(avg).cuda()
    sums_a = torch.zeros(3 * (n_align + n), dtype=torch.float64, device="cuda")
    ctx.dev_align_rotate(aa.data_ptr(), n_align, d_avg.data_ptr(), xa.data_ptr(), t, n, d_sums=sums_a.data_ptr(), stream=1)
    avg2 = (sums_a[:3 * n_align] / t).clone()
    sums_a1 = sums_a.clone()
    ctx.dev_align_rotate(aa.data_ptr(), n_align, avg2.data_ptr(), xa.data_ptr(), t, n, d_sums=sums_a.data_ptr(), stream=1)
    # (b) deferred
    xb, ab = dev_nodes(nodes0), torch.from_numpy(align0.copy()).cuda()
    rot = torch.empty((t, 9), dtype=torch.float64, device="cuda")
    sums_b = torch.zeros_like(sums_a)
    ctx.dev_align_iter(ab.data_ptr(), n_align, d_avg.data_ptr(), xb.data_ptr(), t, n, rot.data_ptr(), True,
                       d_sums=sums_b.data_ptr(), stream=1)
    torch.cuda.synchronize()
    assert torch.equal(xb[:t, :3 * n].cpu(), torch.from_numpy(nodes0.reshape(t, 3 * n)))     # nodes untouched
    np.testing.assert_allclose(sums_b.cpu().numpy(), sums_a1.cpu().numpy(), rtol=0, atol=1e-9)
    rot1 = rot.cpu().numpy().reshape(t, 3, 3).copy()
    ctx.dev_align_iter(ab.data_ptr(), n_align, avg2.data_ptr(), xb.data_ptr(), t, n, rot.data_ptr(), False,
                       d_sums=sums_b.data_ptr(), stream=1)
    torch.cuda.synchronize()
    assert torch.equal(ab.cpu(), aa.cpu())                       # landmarks: identical arithmetic
    np.testing.assert_allclose(sums_b.cpu().numpy(), sums_a.cpu().numpy(), rtol=0, atol=1e-9)
    rot2 = rot.cpu().numpy().reshape(t, 3, 3)
    assert np.abs(np.einsum("tij,tkj->tik", rot2, rot2) - np.eye(3)).max() < 1e-13
    assert np.abs(np.linalg.det(rot2) - 1.0).max() < 1e-13
    # numpy: first-iteration rotations are the Kabsch optimum of every frame against avg
    want1 = np.einsum("tij,tnj->tni", rot1, nodes0)
    ref_al = avg.reshape(n_align, 3)
    for f in (0, t // 2, t - 1):
        a = align0[f].astype(np.float64)
        u, s, vt = np.linalg.svd(a.T @ ref_al)
        d = np.sign(np.linalg.det(u @ vt))
        Rk = (u @ np.diag([1, 1, d]) @ vt).T
        np.testing.assert_allclose(rot1[f], Rk, rtol=0, atol=1e-10)
    # final rotation in place ...
    xc = xb.clone()
    ctx.dev_align_nodes(xc.data_ptr(), t, n, rot.data_ptr(), stream=1)
    torch.cuda.synchronize()
    want2 = np.einsum("tij,tnj->tni", rot2, nodes0).reshape(t, 3 * n)
    np.testing.assert_allclose(xc[:t, :3 * n].cpu().numpy(), want2, rtol=0, atol=1e-12)
    np.testing.assert_allclose(xc[:t, :3 * n].cpu().numpy(), xa[:t, :3 * n].cpu().numpy(), rtol=0, atol=1e-12)
    assert float(xc[:, 3 * n:].abs().max()) == 0.0 if ld > 3 * n else True
    del want1
    # ... or fused into the centring pass (with the FP32 tile copy)
    mean = torch.zeros(ld, dtype=torch.float64, device="cuda")
    mean[:3 * n] = sums_b[3 * n_align:] / t
    p_f = torch.zeros((t, npad // 128, 3, 128), dtype=torch.float32, device="cuda")
    p_s = torch.zeros_like(p_f)
    xd = xb.clone()
    ctx.dev_center_rotated(xd.data_ptr(), t, n, mean.data_ptr(), rot.data_ptr(), d_p32=p_f.data_ptr(), stream=1)
    ctx.dev_center(xc.data_ptr(), t, n, mean.data_ptr(), d_p32=p_s.data_ptr(), stream=1)
    torch.cuda.synchronize()
    np.testing.assert_allclose(xd[:t].cpu().numpy(), xc[:t].cpu().numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(p_f.cpu().numpy(), p_s.cpu().numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(xd[:t, :3 * n].cpu().numpy().sum(axis=0), 0.0, atol=1e-9 * t)
