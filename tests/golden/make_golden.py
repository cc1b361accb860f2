#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REFERENCE'S OWN source files.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

The reference (rbdavid/WISP_mdanalysis_implementation) is Python 2 + networkx 1.x +
MDAnalysis and cannot be imported as is.  This script reads its source files where they
lie, adapts them IN MEMORY in a purely mechanical way and executes them -- no reference
source is copied into this repository:

  * tabs expanded to 8 columns (Python 2 tab semantics), ``print x`` -> ``print(x)``;
  * every ``a / b`` is routed through ``_py2div`` (Python-2 semantics: floor division
    when both operands are ints, true division otherwise) -- SURVEY App. B9;
  * ``np.int`` is provided as ``int`` (B10);
  * module-level names the reference uses without defining them are injected with the
    obvious argument (B6: all_pos_Nodes / avg_pos_Nodes, B7: trajectory);
  * ``import MDAnalysis`` resolves to a shim exposing ``rotation_matrix`` (Kabsch) and
    ``distance_array`` (plain Euclidean), and the Universe is the in-repo stand-in
    (wisp_mdanalysis_implementation_b200/mda_standin.py);
  * ``import networkx`` resolves to a shim giving networkx-1.x behaviour on top of the
    installed networkx 3.x: ``Graph(data=A)`` makes an edge of every non-zero entry,
    ``neighbors()`` returns a list, ``G.edge[u][v]``, and SSSP dicts ordered by node
    (Python-2 small-int dict order, B11).

Everything else -- loops, summation order, thresholds, the path enumeration, the dedupe,
the K / K+1 rule -- is the reference's code running unmodified.
"""
import ast
import os
import re
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("WISP_REFERENCE", "/root/reference")

from wisp_mdanalysis_implementation_b200 import mda_standin, synth  # noqa: E402


def _py2div(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        return a // b
    return a / b


class _Div(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(
                ast.Call(func=ast.Name(id="_py2div", ctx=ast.Load()),
                         args=[node.left, node.right], keywords=[]), node)
        return node


def _kabsch(a, b, weights=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    H = a.T @ b
    U, S, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(Vt.T @ U.T))
    R = Vt.T @ np.diag([1.0, 1.0, d]) @ U.T
    diff = a @ R.T - b
    return R, np.sqrt((diff * diff).sum() / len(a))


def _distance_array(a, b, box=None, result=None, backend="serial"):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = b[None, :, :] - a[:, None, :]
    r = np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])
    if result is not None:
        result[...] = r
        return result
    return r


def _mda_shim():
    m = types.ModuleType("MDAnalysis")
    m.Universe = mda_standin.Universe
    m.version = types.ModuleType("MDAnalysis.version")
    m.version.__version__ = "standin"
    m.analysis = types.ModuleType("MDAnalysis.analysis")
    m.analysis.align = types.ModuleType("MDAnalysis.analysis.align")
    m.analysis.align.rotation_matrix = _kabsch
    m.analysis.distances = types.ModuleType("MDAnalysis.analysis.distances")
    m.analysis.distances.distance_array = _distance_array
    return {"MDAnalysis": m, "MDAnalysis.version": m.version, "MDAnalysis.analysis": m.analysis,
            "MDAnalysis.analysis.align": m.analysis.align,
            "MDAnalysis.analysis.distances": m.analysis.distances}


def _nx_shim():
    import networkx as real

    class Graph(real.Graph):
        def __init__(self, data=None, **attr):
            super().__init__(**attr)
            if data is not None:
                A = np.asarray(data)
                self.add_nodes_from(range(A.shape[0]))
                for u, v in zip(*A.nonzero()):
                    self.add_edge(int(u), int(v), weight=A[u, v])

        def neighbors(self, n):
            return list(super().neighbors(n))

        @property
        def edge(self):
            return self._adj

    def single_source_dijkstra(G, source, target=None, cutoff=None, weight="weight"):
        lengths, paths = real.single_source_dijkstra(G, source, target=target, cutoff=cutoff,
                                                     weight=weight)
        keys = sorted(lengths.keys())
        return {k: lengths[k] for k in keys}, {k: paths[k] for k in keys}

    m = types.ModuleType("networkx")
    m.Graph = Graph
    m.dijkstra_path = real.dijkstra_path
    m.single_source_dijkstra = single_source_dijkstra
    return {"networkx": m}


def load_reference(name, shims):
    path = os.path.join(REF, name + ".py")
    src = open(path).read().expandtabs(8)
    src = re.sub(r"^(\s*)print (.*)$", r"\1print(\2)", src, flags=re.M)
    tree = _Div().visit(ast.parse(src, path))
    ast.fix_missing_locations(tree)
    mod = types.ModuleType("reference_" + name)
    mod.__dict__["_py2div"] = _py2div
    saved = {k: sys.modules.get(k) for k in shims}
    sys.modules.update(shims)
    had_int = "int" in np.__dict__
    np.int = int
    try:
        exec(compile(tree, path, "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mod.__dict__["sys"] = sys          # B8
    return mod


def quiet(fn, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def golden_trajectory(tag, n_nodes, n_frames, apn, seed, cutoff, rigid_motion):
    shims = {}
    shims.update(_mda_shim())
    traj_mod = load_reference("trajectory_analysis_functions", shims)
    sel_mod = load_reference("make_node_selections", shims)
    adj_mod = load_reference("adjacency_matrix_functions", shims)
    fun_mod = load_reference("functionalize_adj_matrix_functions", shims)

    system, coords = synth.synth(n_nodes, n_frames, apn, seed)
    if rigid_motion:
        # add per-frame rigid rotation + translation so the alignment loop has work to do
        rng = np.random.default_rng(seed + 17)
        for t in range(n_frames):
            ang = rng.normal(0.0, 0.05, size=3)
            cx, cy, cz = np.cos(ang)
            sx, sy, sz = np.sin(ang)
            Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
            Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
            Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
            R = Rz @ Ry @ Rx
            coords[t] = (coords[t].astype(np.float64) @ R.T + rng.normal(0, 2.0, size=3)).astype(np.float32)
    coords_in = coords.copy()
    u = mda_standin.Universe.from_synth(system)
    out = tempfile.mkdtemp() + os.sep
    selection_list, src, snk = quiet(sel_mod.make_selections, u, out + "node_selections.txt",
                                     "COM", "protein", ["ALA_2"], ["ALA_%d" % n_nodes])
    # COM extraction alone: zero alignment iterations (the while loop at :96 never runs)
    com_only, com_avg = quiet(traj_mod.traj_alignment_and_averaging, u, "protein and name CA",
                              selection_list, "COM", [coords_in.copy()], out,
                              maximum_num_iterations=0)
    com_only = np.array(com_only)
    node_traj, avg = quiet(traj_mod.traj_alignment_and_averaging, u, "protein and name CA",
                           selection_list, "COM", [coords], out)
    node_traj = np.array(node_traj)
    # B6 / B7: names the reference uses without defining
    traj_mod.all_pos_Nodes = node_traj
    traj_mod.avg_pos_Nodes = avg
    traj_mod.trajectory = node_traj
    cov = quiet(traj_mod.cartesian_covariance, node_traj, avg, out)
    binary, avgdist = quiet(traj_mod.calc_contact_map, node_traj, out, distance_cutoff=cutoff)
    corr = quiet(adj_mod.pearson_correlation_analysis, cov, avg, out)
    var = np.loadtxt(out + "node_variance.dat")
    ncov = np.loadtxt(out + "node_covariance.dat")
    w_plain = quiet(fun_mod.func_pearson_correlation, corr, out)
    w_binary = quiet(fun_mod.func_pearson_correlation, corr, out, contact_map=binary)
    w_avg = quiet(fun_mod.func_pearson_correlation, corr, out, contact_map=avgdist)
    # linear mutual information style
    mi = quiet(adj_mod.linear_mutual_information_analysis, cov, avg, out)
    with np.errstate(divide="ignore", invalid="ignore"):
        func_mi = quiet(fun_mod.generalized_correlation_coefficient_calc, mi, out)
        rmi = np.loadtxt(out + "gen_corr_coefficients.dat")
        func_mi_damped = quiet(fun_mod.generalized_correlation_coefficient_calc, mi, out,
                               contact_map=avgdist, Lambda=5.0)
        rmi_damped = np.loadtxt(out + "gen_corr_coefficients.dat")
    node_ptr = np.array([0] + list(np.cumsum([len(s.indices) for s in selection_list])), dtype=np.int32)
    atom_idx = np.concatenate([s.indices for s in selection_list]).astype(np.int32)
    align_idx = u.select_atoms("protein and name CA").indices.astype(np.int32)
    np.savez_compressed(
        os.path.join(HERE, tag + ".npz"),
        coords=coords_in, masses=system.masses, node_ptr=node_ptr, atom_idx=atom_idx,
        align_idx=align_idx, com_only=com_only, com_avg=com_avg, node_traj=node_traj, avg=avg, cov=cov, binary=binary,
        avgdist=avgdist, corr=corr, var=var, ncov=ncov, w_plain=w_plain, w_binary=w_binary,
        w_avg=w_avg, mi=mi, func_mi=func_mi, rmi=rmi, func_mi_damped=func_mi_damped,
        rmi_damped=rmi_damped, cutoff=np.float64(cutoff), sources=np.array(src), sinks=np.array(snk))
    print(tag, "nodes", n_nodes, "frames", n_frames, "contacts", int(binary.sum()) // 2)


def golden_paths():
    net = load_reference("network_analysis_functions", _nx_shim())
    cases = []
    rng = np.random.default_rng(4242)
    specs = [
        # (N, density, sources, sinks, K)
        (10, 0.45, [3], [8], 6),
        (12, 0.40, [0], [11], 8),          # source is node 0: quirk vacuous
        (12, 0.50, [5], [9], 10),          # node 0 adjacent to source: quirk bites
        (11, 0.50, [4, 7], [2, 9], 9),     # multi source / multi sink
        (11, 0.50, [1, 6], [6, 1], 7),     # overlapping sets: reversed duplicates + skipped pairs
        (14, 0.30, [2], [13], 12),
        (9, 0.70, [8], [1], 15),
        (13, 0.35, [12], [0], 5),          # sink is node 0
    ]
    for n, dens, so, si, k in specs:
        W = synth.random_cost_matrix(n, seed=int(rng.integers(1 << 30)), density=dens,
                                     lo=0.05, hi=0.6)
        paths = quiet(net.get_paths, W, so, si, k)
        cases.append((W, so, si, k, paths))
    # hand-built K+1 case: node 0 is the 2nd node of the optimal path, so the seed path is
    # never re-enumerated and K+1 entries come back when the last pass counts exactly K.
    W = np.zeros((6, 6))

    def e(a, b, w):
        W[a, b] = W[b, a] = w
    e(1, 0, 0.10); e(0, 2, 0.10); e(1, 3, 0.30); e(3, 2, 0.30); e(1, 4, 0.50); e(4, 2, 0.55)
    e(3, 4, 0.05); e(4, 5, 0.4); e(5, 2, 0.4)
    for k in (2, 3, 4):
        paths = quiet(net.get_paths, W, [1], [2], k)
        cases.append((W.copy(), [1], [2], k, paths))
    # the hand-enumerated ladder graph of tests/known_answers.py: the reference's own code confirms
    # the answers derived on paper (node-0 rule, K+1 returns)
    sys.path.insert(0, os.path.dirname(HERE))
    from known_answers import ladder_graph
    for so, si, k in (([1], [5], 2), ([1], [5], 3), ([0], [5], 1), ([5], [1], 2)):
        cases.append((ladder_graph(), so, si, k, quiet(net.get_paths, ladder_graph(), so, si, k)))
    blob = {}
    for i, (W, so, si, k, paths) in enumerate(cases):
        blob["W%d" % i] = W
        blob["sources%d" % i] = np.array(so)
        blob["sinks%d" % i] = np.array(si)
        blob["K%d" % i] = np.int64(k)
        blob["lengths%d" % i] = np.array([p[0] for p in paths], dtype=np.float64)
        blob["offsets%d" % i] = np.cumsum([0] + [len(p) - 1 for p in paths]).astype(np.int64)
        blob["nodes%d" % i] = np.array([n for p in paths for n in p[1:]], dtype=np.int64)
        print("paths case", i, "N", len(W), "K", k, "returned", len(paths))
    blob["n_cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "get_paths.npz"), **blob)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference not found at %s (golden vectors can only be regenerated in the "
                 "build container)" % REF)
    golden_trajectory("traj_small", n_nodes=12, n_frames=60, apn=4, seed=11, cutoff=9.0,
                      rigid_motion=False)
    golden_trajectory("traj_aligned", n_nodes=20, n_frames=48, apn=6, seed=12, cutoff=10.0,
                      rigid_motion=True)
    golden_paths()
