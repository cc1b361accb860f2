"""Drop-in replacement for the reference's ``trajectory_analysis_functions.py``.

Same function names, argument meaning, return values, output files and error behaviour
(print + sys.exit) as rbdavid/WISP_mdanalysis_implementation; the arithmetic runs in
hand-written sm_100a CUDA kernels behind libwisp_b200.so.  Point the config key
``trajectory_functions_file`` at this module (wisp_w_mdanalysis.config:8).
"""
import sys

import numpy as np

from wisp_mdanalysis_implementation_b200 import _adapt, _lib, textio


def traj_alignment_and_averaging(universe, alignment_selection, selection_list, node_definition,
                                 trajectory_list, output_directory, convergence_threshold=1E-5,
                                 maximum_num_iterations=100, step=1):
    """trajectory_analysis_functions.py:39-138.  K1 (translate + node COM) and K7 (iterative
    alignment to the running average) run on the GPU; returns the aligned node trajectory
    (T, N, 3) float64 and the average node positions (N, 3) float64, and writes
    ``average_node_positions.dat``."""
    average_file_name = output_directory + 'average_node_positions.dat'
    ctx = _lib.default_context()
    u_alignment = universe.select_atoms(alignment_selection)
    align_idx = np.asarray(u_alignment.indices, dtype=np.int32)
    nNodes = len(selection_list)
    if node_definition.upper() not in ('COM', 'ATOMIC'):
        # the reference appends nothing for other values and then fails on the array shape
        print('The substrate_node_definition parameter is not understood.')
        sys.exit()
    node_ptr, atom_idx = _adapt.selection_to_csr(selection_list)
    masses = _adapt.universe_masses(universe)

    blocks = []
    nSteps = 0
    for traj in trajectory_list:
        print('Loading trajectory', traj)
        universe.load_new(traj)
        if len(universe.trajectory) % step != 0:
            print('User defined step size is not a factor of the number of frames in ', traj,
                  '. The user is not getting the desired step size.')
            sys.exit()
        nSteps += len(universe.trajectory) // step
        blocks.append(_adapt.universe_coordinates(universe, step))
    print('Analyzed', nSteps, 'frames.')
    coords = blocks[0] if len(blocks) == 1 else np.concatenate(blocks)

    print('Beginning the iterative process of aligning to the average alignment positions, '
          'calculating new positions, and recalculating the average positions')
    all_pos_Nodes, avg_pos_Nodes, log = ctx.traj_alignment_and_averaging(
        coords, masses, node_ptr, atom_idx, align_idx,
        atomic=node_definition.upper() == 'ATOMIC', convergence_threshold=convergence_threshold,
        maximum_num_iterations=maximum_num_iterations)
    for iteration, residual, analysis_RMSD in log:
        print('Iteration ', iteration, ': RMSD btw alignment landmarks: ', residual,
              ', RMSD btw Node Positions: ', analysis_RMSD)
    print('Finished calculating the average structure using the iterative averaging approach. '
          'Outputting the average node positions to file.')
    textio.savetxt(average_file_name, avg_pos_Nodes, header='Shape: nNodes x 3 (%d x 3)' % (nNodes))
    return all_pos_Nodes, avg_pos_Nodes


def cartesian_covariance(node_trajectory, avg_node_positions, output_directory):
    """trajectory_analysis_functions.py:141-184: (3N, 3N) covariance
    (1/T) sum_t x x^T - avg avg^T via the FP64 tensor-core contraction (K2, stride 1)."""
    variance_file_name = output_directory + 'cartesian_variance.dat'
    covariance_file_name = output_directory + 'cartesian_covariance.dat'
    print('Beginning cartesian covariance analysis.')
    ctx = _lib.default_context()
    xyz_node_covariance = ctx.cartesian_covariance(node_trajectory, avg_node_positions)
    xyz_node_variance = np.diag(xyz_node_covariance)
    print('Finished calculating the variance and covariance of node cartesian coordinates. '
          'Outputting the two arrays to file. The covariance matrix file can be reused in '
          'subsequent analyses of various adjacency matrices, assuming studying the same range '
          'of frames, selections, etc.')
    textio.savetxt(variance_file_name, xyz_node_variance)
    textio.savetxt(covariance_file_name, xyz_node_covariance)
    return xyz_node_covariance


def calc_contact_map(trajectory_data, output_directory, distance_cutoff=9999.):
    """trajectory_analysis_functions.py:187-236: mean pairwise node distance over frames
    (K3) and the strict '<' binary contact map (K4)."""
    binary_contact_map_file_name = output_directory + 'binary_contact_map.dat'
    avg_node_node_distance_file_name = output_directory + 'avg_distance_contact_map.dat'
    print('Starting to calculate the average distance between nodes.')
    ctx = _lib.default_context()
    binary, avg = ctx.contact_map(trajectory_data, distance_cutoff)
    textio.savetxt(binary_contact_map_file_name, binary, fmt='%i')
    textio.savetxt(avg_node_node_distance_file_name, avg)
    return binary, avg
