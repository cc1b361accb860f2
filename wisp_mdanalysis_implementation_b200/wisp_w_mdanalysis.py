"""Driver: ``python wisp_w_mdanalysis.py <config>`` -- the reference's command line
(wisp_w_mdanalysis.py:5,54) restated for Python 3 with the repairs of SURVEY.md Appendix B
(B1 output file names, B2 stray mkdir, B3 adjacency call arity, B4 undeclared keys).

Stage implementations are looked up BY NAME from the module files named in the config
(wisp_w_mdanalysis.config:7-12, wisp_w_mdanalysis.py:233-267), exactly like the reference,
so a config can point any stage at the reference's own module or at the B200 module of the
same name in this directory.  Boundary code: nothing is computed here.
"""
import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (_HERE, os.path.dirname(_HERE)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from IO import config_parser, summary  # noqa: E402
from wisp_mdanalysis_implementation_b200 import textio  # noqa: E402


def _stage(parameters, key, name):
    module = importlib.import_module(parameters[key].split('.')[0], package=None)
    return getattr(module, name)


def _universe(parameters):
    if parameters['universe_factory']:
        mod, fn = parameters['universe_factory'].split(':')
        return getattr(importlib.import_module(mod), fn)(parameters)
    import MDAnalysis
    return MDAnalysis.Universe(parameters['pdb'])


def main(config_file, argv=None):
    parameters = {}
    config_parser(config_file, parameters)                                          # :213-214
    if parameters['output_directory'][-1] != os.sep:
        parameters['output_directory'] += os.sep
    if os.path.exists(parameters['output_directory']):                              # :222-224
        print('The output directory, ', parameters['output_directory'],
              'already exists. Please select a different directory name for output.')
        sys.exit()
    os.mkdir(parameters['output_directory'])
    out = parameters['output_directory']

    # ---- 2) stage functions by name (:233-267) ----
    make_selections = _stage(parameters, 'node_selection_file', 'make_selections')
    if parameters['trajectory_analysis_boolean']:
        traj_alignment_and_averaging = _stage(parameters, 'trajectory_functions_file', 'traj_alignment_and_averaging')
        cartesian_covariance = _stage(parameters, 'trajectory_functions_file', 'cartesian_covariance')
        calc_contact_map = _stage(parameters, 'trajectory_functions_file', 'calc_contact_map')
    style = parameters['adjacency_matrix_style'].upper()
    if style in ('PEARSON', 'PEARSON CORRELATION'):
        adjacency_matrix_analysis = _stage(parameters, 'adjacency_matrix_functions_file', 'pearson_correlation_analysis')
        functionalize_adjacency_matrix = _stage(parameters, 'func_adjacency_matrix_functions_file', 'func_pearson_correlation')
    elif style in ('LMI', 'LINEAR MUTUAL INFORMATION'):
        adjacency_matrix_analysis = _stage(parameters, 'adjacency_matrix_functions_file', 'linear_mutual_information_analysis')
        functionalize_adjacency_matrix = _stage(parameters, 'func_adjacency_matrix_functions_file', 'generalized_correlation_coefficient_calc')
    else:
        print("The user has not read in an accepted style of adjacency matrix. Accelerated here: "
              "'PEARSON' / 'PEARSON CORRELATION'.")
        sys.exit()
    get_paths = _stage(parameters, 'network_functions_file', 'get_paths')
    create_vis_state = None
    if parameters['visualization_functions_file'] not in (None, 'None', 'none'):
        create_vis_state = _stage(parameters, 'visualization_functions_file', 'create_vis_state')

    # B1: the three names the reference left commented out (:65-67)
    simply_formatted_paths_file_name = out + 'simply_formatted_paths.txt'
    pathways_vis_state_file_name = out + 'pathways_vis_state.vmd'
    summary_file_name = out + 'summary.txt'

    # ---- 3) universe and node selections (:72-75) ----
    u = _universe(parameters)
    selection_list, source_indices, sink_indices = make_selections(
        u, out + 'node_selections.txt', parameters['substrate_node_definition'],
        parameters['substrate_selection_string'], parameters['source_selection_string_list'],
        parameters['sink_selection_string_list'],
        nonstandard_substrates_selection=parameters['nonstandard_substrates_selection'],
        homemade_selections=parameters['homemade_selections'])
    print('Number of nodes:', len(selection_list), '\nsource node indices:', source_indices,
          '\nsink node indices:', sink_indices, '\nNode selections written out to', out + 'node_selections.txt')

    contact_map = None
    # ---- 4) trajectory analysis (:80-123) ----
    if parameters['trajectory_analysis_boolean']:
        print('Beginning trajectory analysis.')
        Node_trajectory, avg_Node_positions = traj_alignment_and_averaging(
            u, parameters['alignment_selection'], selection_list, parameters['substrate_node_definition'],
            parameters['traj_list'], out, step=parameters['trajectory_step'], convergence_threshold=1E-5)
        Node_cart_covariance = cartesian_covariance(Node_trajectory, avg_Node_positions, out)
        print('Beginning node pair distance calculation.')
        binary_contact_map, avg_node_node_distances = calc_contact_map(
            Node_trajectory, out, distance_cutoff=parameters['contact_map_distance_cutoff'])
        if parameters['which_contact_map'].upper() == 'AVERAGE CONTACT MAP':
            contact_map = avg_node_node_distances
        elif parameters['which_contact_map'].upper() == 'BINARY CONTACT MAP':
            contact_map = binary_contact_map
        else:
            print("User has not defined which contact map should be used in subsequent weighting of the "
                  "adjacency matrix. Acceptable values for the 'which_contact_map' parameter are "
                  "'AVERAGE CONTACT MAP' or 'BINARY CONTACT MAP'.")
            sys.exit()
    else:
        n = len(selection_list)
        print('Loading in the user specified cartesian covariance data file. Note: no node trajectory has been created.')
        Node_cart_covariance = textio.loadtxt(parameters['user_input_cartesian_covariance_matrix'])
        if Node_cart_covariance.shape[0] % 3 != 0 or Node_cart_covariance.shape[0] != Node_cart_covariance.shape[1]:
            print('User has read in a cartesian covariance matrix that does not have the correct shape '
                  '(expected: %d x %d, got: %d x %d).' % (3 * n, 3 * n, Node_cart_covariance.shape[0], Node_cart_covariance.shape[1]))
            sys.exit()
        avg_Node_positions = None
        if parameters['user_input_average_node_positions']:
            avg_Node_positions = textio.loadtxt(parameters['user_input_average_node_positions'])
            if avg_Node_positions.shape != (n, 3):
                print('User has read in an average node positions file that does not have the correct shape.')
                sys.exit()
        if parameters['user_input_contact_map']:
            contact_map = textio.loadtxt(parameters['user_input_contact_map'])
            if contact_map.shape != (n, n):
                print('User has read in a contact map that does not have the correct shape.')
                sys.exit()

    # ---- 5) adjacency matrix (:128-136) ----
    if parameters['adjacency_matrix_analysis_boolean']:
        print('Beginning the calculation of the adjacency matrix (' + parameters['adjacency_matrix_style'] + ').')
        adjacency_matrix = adjacency_matrix_analysis(Node_cart_covariance, avg_Node_positions, out)   # B3
    else:
        adjacency_matrix = textio.loadtxt(parameters['user_input_adjacency_matrix'])
        if adjacency_matrix.shape != (len(selection_list), len(selection_list)):
            print('User has read in an adjacency matrix that does not have the correct shape.')
            sys.exit()

    # ---- 6) cost matrix (:141-144) ----
    if parameters['weight_by_contact_map_boolean']:
        functionalized_adjacency_matrix = functionalize_adjacency_matrix(
            adjacency_matrix, out, contact_map=contact_map, Lambda=parameters['Lambda'])
    else:
        functionalized_adjacency_matrix = functionalize_adjacency_matrix(adjacency_matrix, out)
    print('Finished functionalizing and weighting the adjacency matrix.')

    # ---- 7) paths (:154-165; B2: the stray os.mkdir at :152 is dropped) ----
    paths = get_paths(functionalized_adjacency_matrix, source_indices, sink_indices, parameters['number_of_paths'])
    print('Finished calculating the pathways that connect the source and sink selections.')
    with open(simply_formatted_paths_file_name, 'w') as W:
        W.write('# Pathway_Length Node_Indices (zero indexed)\n')
        for path in paths:
            path_string = '%f ' % (path[0])
            for node in path[1:]:
                path_string += '%d ' % (node)
            path_string += '\n'
            W.write(path_string)

    # ---- 9) VMD state (out of the accelerated path; optional) ----
    if create_vis_state is not None:
        u.load_new(parameters['visualization_frame_pdb'])
        create_vis_state(parameters['visualization_frame_pdb'], selection_list, paths, pathways_vis_state_file_name,
                         node_sphere_radius=parameters['node_sphere_radius'], node_sphere_rgb=parameters['node_sphere_rgb'],
                         shortest_path_radius=parameters['shortest_path_radius'], shortest_path_rgb=parameters['shortest_path_rgb'],
                         longest_path_radius=parameters['longest_path_radius'], longest_path_rgb=parameters['longest_path_rgb'],
                         node_sphere_color_index=parameters['node_sphere_color_index'],
                         VMD_color_index_range=parameters['VMD_color_index_range'], VMD_resolution=parameters['VMD_resolution'],
                         VMD_spline_smoothness=parameters['VMD_spline_smoothness'])
        print('Finished creating the vis state file.')

    # ---- 11) summary (:207-208) ----
    if parameters['summary_boolean']:
        summary(summary_file_name, argv if argv is not None else sys.argv, parameters)
    return paths


if __name__ == '__main__':
    main(sys.argv[1])
