"""Host-side boundary code (no GPU): config parsing, the stand-in Universe, node selections and
the selection -> CSR adaptor the C ABI is fed from."""
import numpy as np
import pytest

from wisp_mdanalysis_implementation_b200 import IO, _adapt, make_node_selections, mda_standin, synth


def _universe(n_res=5, apn=3, frames=4, seed=0):
    rng = np.random.default_rng(seed)
    names = np.array((["CA"] + ["X%d" % k for k in range(1, apn)]) * n_res)
    resnames = np.repeat(np.array(["ALA", "GLY", "SER", "ALA", "LYS", "VAL", "THR"])[:n_res], apn)
    resids = np.repeat(np.arange(10, 10 + n_res), apn)
    masses = rng.uniform(1.0, 16.0, size=n_res * apn)
    coords = rng.normal(scale=5.0, size=(frames, n_res * apn, 3)).astype(np.float32)
    return mda_standin.Universe(names, resnames, resids, masses, coords), coords, masses


def test_selection_language():
    u, _, _ = _universe()
    assert list(u.select_atoms("name CA").indices) == [0, 3, 6, 9, 12]
    assert list(u.select_atoms("resid 11:12 and name CA").indices) == [3, 6]
    assert list(u.select_atoms("resname ALA and not name CA").indices) == [1, 2, 10, 11]
    assert list(u.select_atoms("(resid 10 or resid 14) and name X1 X2").indices) == [1, 2, 13, 14]
    assert u.select_atoms("protein").n_atoms == 15 and u.select_atoms("all").n_residues == 5
    assert list(u.select_atoms("index 2:4").indices) == [2, 3, 4]
    with pytest.raises(ValueError):
        u.select_atoms("(name CA")
    with pytest.raises(ValueError):
        u.select_atoms("segid A")


def test_atomgroup_matches_reference_semantics():
    u, coords, masses = _universe()
    res = u.select_atoms("resid 12")
    want = (coords[0, 6:9] * masses[6:9, None]).sum(0) / masses[6:9].sum()
    np.testing.assert_allclose(res.center_of_mass(), want, rtol=1e-15)
    # translate(): float32 positions += float64 vector, rounded back to float32, in place
    shift = np.array([0.1, -0.2, 0.3])
    before = coords[0].copy()
    u.atoms.translate(shift)
    after = u.trajectory.ts.positions
    assert after.dtype == np.float32
    assert np.array_equal(after, (before.astype(np.float64) + shift).astype(np.float32))
    # iteration walks the frames and ts.positions is a view of the stored block
    seen = [ts.frame for ts in u.trajectory[::2]]
    assert seen == [0, 2]
    assert np.shares_memory(u.trajectory[1].positions, u.trajectory.get_array())


def test_make_selections_and_csr(tmp_path):
    u, coords, masses = _universe()
    sel, src, snk = make_node_selections.make_selections(
        u, str(tmp_path / "nodes.txt"), "COM", "protein", ["GLY_11"], ["LYS_14"])
    assert (src, snk) == ([1], [4]) and len(sel) == 5
    lines = (tmp_path / "nodes.txt").read_text().splitlines()
    assert lines[1].endswith("# SOURCE") and lines[4].endswith("# SINK") and lines[0] == "00   ALA   10"
    ptr, idx = _adapt.selection_to_csr(sel)
    assert ptr.dtype == np.int32 and list(ptr) == [0, 3, 6, 9, 12, 15] and list(idx) == list(range(15))
    # ATOMIC: one node per atom of the selection; Atom objects go through the same adaptor
    sel_a, src_a, snk_a = make_node_selections.make_selections(
        u, str(tmp_path / "atoms.txt"), "atomic", "name CA", ["ALA_10"], ["VAL_15"])
    ptr_a, idx_a = _adapt.selection_to_csr(sel_a)
    assert list(ptr_a) == [0, 1, 2, 3, 4, 5] and list(idx_a) == [0, 3, 6, 9, 12] and src_a == [0] and snk_a == []
    # overlapping / non-contiguous hand-made groups are legal node definitions (general CSR gather)
    groups = [u.select_atoms("index 0:4"), u.select_atoms("index 3:5"), u.select_atoms("name CA")]
    ptr_g, idx_g = _adapt.selection_to_csr(groups)
    assert list(ptr_g) == [0, 5, 8, 13] and list(idx_g[5:8]) == [3, 4, 5]
    with pytest.raises(ValueError):
        _adapt.selection_to_csr([u.select_atoms("resid 99")])
    with pytest.raises(TypeError):
        _adapt.selection_to_csr([object()])
    with pytest.raises(SystemExit):
        make_node_selections.make_selections(u, str(tmp_path / "x.txt"), "spheres", "protein", [], [])


def test_universe_coordinates_and_masses():
    u, coords, masses = _universe(frames=7)
    block = _adapt.universe_coordinates(u, step=3)
    assert block.dtype == np.float32 and block.flags["C_CONTIGUOUS"]
    assert np.array_equal(block, coords[::3])
    assert np.array_equal(_adapt.universe_masses(u), masses)
    system = synth.SynthSystem(6, 5, 4, seed=3)
    us = mda_standin.Universe.from_synth(system, system.coords())
    sel, _, _ = make_node_selections.make_selections(us, "/dev/null", "COM", "protein", [], [])
    ptr, idx = _adapt.selection_to_csr(sel)
    assert np.array_equal(ptr, system.node_ptr) and np.array_equal(idx, system.atom_idx)
    assert list(us.select_atoms("protein and name CA").indices) == list(system.align_idx)


CONFIG = """
output_directory = 'out'
node_selection_file = 'make_node_selections.py'
trajectory_functions_file = 'trajectory_analysis_functions.py'
adjacency_matrix_functions_file = 'adjacency_matrix_functions.py'
func_adjacency_matrix_functions_file = 'functionalize_adj_matrix_functions.py'
network_functions_file = 'network_analysis_functions.py'
visualization_functions_file = 'visualization_functions.py'
trajectory_analysis_boolean = True
adjacency_matrix_style = 'pearson correlation'
pdb = 'system.pdb'
visualization_frame_pdb = 'system.pdb'
substrate_node_definition = 'COM'
substrate_selection_string = 'protein'
source_selection_string_list = ['ALA_10']
sink_selection_string_list = ['LYS_14']
number_of_paths = 25
contact_map_distance_cutoff = 12.5
"""


def test_config_parser(tmp_path, capsys):
    cfg = tmp_path / "run.config"
    cfg.write_text(CONFIG)
    parameters = {}
    IO.config_parser(str(cfg), parameters)
    assert parameters["number_of_paths"] == 25 and parameters["contact_map_distance_cutoff"] == 12.5
    assert parameters["trajectory_step"] == 1 and parameters["alignment_selection"] == "protein and name CA"
    assert parameters["which_contact_map"] == "average contact map"          # reference default (IO.py:39-40)
    assert parameters["fused_pipeline_boolean"] is False and parameters["universe_factory"] is None
    # a mandatory key left out: message per key, then exit -- the reference's convention
    broken = tmp_path / "broken.config"
    broken.write_text(CONFIG.replace("number_of_paths = 25\n", "").replace("pdb = 'system.pdb'\n", ""))
    with pytest.raises(SystemExit):
        IO.config_parser(str(broken), {})
    out = capsys.readouterr().out
    assert "number_of_paths has not been assigned a value" in out and "pdb has not been assigned a value" in out
    IO.summary(str(tmp_path / "summary.txt"), ["wisp_w_mdanalysis.py", "run.config"], parameters)
    text = (tmp_path / "summary.txt").read_text()
    assert "wisp_w_mdanalysis.py run.config" in text and "number_of_paths = 25" in text


def test_k7_rotation_solver_on_the_cpu(tmp_path):
    """csrc/align_solve.cuh (the K7 eigen-solver: Newton on the characteristic quartic of Horn's
    quaternion matrix + adjugate column + Rayleigh polish, Jacobi for degenerate spectra) is
    __host__ __device__: tools/align_solve_check.cpp compiles the very header the kernels include with
    g++ and holds it against cyclic Jacobi on 200,000 random / near-aligned / planar / collinear /
    180-degree / single-point cases (proper rotation, optimal objective, |dR| < 1e-11 on generic input).
    Replaces MDAnalysis' rotation_matrix as called at trajectory_analysis_functions.py:102."""
    import os
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "align_solve_check")
    subprocess.run(["g++", "-O2", "-o", exe, os.path.join(root, "tools", "align_solve_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Jacobi fallbacks" in r.stdout


def _k3_deal(diag, valid_i, valid_j):
    """Python model of the block dealing in contact_sum_ws_kernel (csrc/k3_contact.cu, `balance`): returns
    per warp (task block or None, frame parity: None = all frames / 0 = even / 1 = odd)."""
    sub_row = [0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 1, 2, 2, 3, 3, 3]
    sub_col = [0, 1, 2, 3, 1, 2, 3, 2, 3, 3, 0, 0, 1, 0, 1, 2]

    def active(ow):
        if diag:
            return ow < 10 and 32 * sub_row[ow] < valid_i and 32 * sub_col[ow] < valid_j
        return 8 * ow < valid_j

    act = [ow for ow in range(16) if active(ow)]
    n_active = len(act)
    if n_active == 0 or n_active == 16:
        return [(w if active(w) else None, None) for w in range(16)], act
    n_own = 2 * n_active - 16 if n_active > 8 else 0
    deal = []
    for w in range(16):
        if w < n_own:
            deal.append((act[w], None))
        else:
            h = w - n_own
            k = n_own + (h >> 1)
            deal.append((act[k], h & 1) if k < n_active else (None, None))
    return deal, act


def test_k3_block_dealing_model():
    """Every active block of a diagonal / padded K3 tile is covered exactly once -- by one warp taking all
    frames or by two consecutive warps taking the even and the odd frames -- and the four schedulers
    (warp index mod 4) carry the same work to within half a block (10 active blocks: 2.5 each, where
    idling six warps leaves 3, 3, 2, 2)."""
    for diag in (False, True):
        for valid in range(1, 129):
            deal, act = _k3_deal(diag, valid if diag else 128, valid)
            cover = {b: [] for b in act}
            load = [0.0] * 4
            for w, (task, parity) in enumerate(deal):
                if task is None:
                    continue
                cover[task].append(parity)
                load[w % 4] += 1.0 if parity is None else 0.5
            for b, parts in cover.items():
                assert parts == [None] or sorted(parts) == [0, 1], (diag, valid, b, parts)
            for w in range(15):                      # the two halves of a block sit in consecutive warps
                if deal[w][1] == 0:
                    assert deal[w + 1] == (deal[w][0], 1)
            assert abs(sum(load) - len(act)) < 1e-12
            if 0 < len(act) < 16:
                assert max(load) - min(load) <= 0.5 + 1e-12, (diag, valid, load)
    # the C2 cases: 10 active blocks -> 2.5 per scheduler
    deal, act = _k3_deal(True, 128, 128)
    assert len(act) == 10 and sorted(p is None for _, p in deal if _ is not None).count(True) == 4
