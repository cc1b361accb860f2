"""Python-3 restatement of the reference's ``IO.py`` (config_parser / summary) with the
repairs listed in SURVEY.md Appendix B4.  Boundary code only: nothing here is accelerated.

The config file is Python source executed into the ``parameters`` dict (IO.py:54); the 16
mandatory keys are initialised to '' and checked (IO.py:24,28-29,55-58); optional keys get
the reference's defaults (IO.py:32-51) plus the keys the reference's driver reads but never
declares (wisp_w_mdanalysis.py:114,128,141,142).
"""
import sys

NECESSARY = ['output_directory', 'node_selection_file', 'trajectory_functions_file',
             'adjacency_matrix_functions_file', 'func_adjacency_matrix_functions_file',
             'network_functions_file', 'visualization_functions_file', 'trajectory_analysis_boolean',
             'adjacency_matrix_style', 'pdb', 'visualization_frame_pdb', 'substrate_node_definition',
             'substrate_selection_string', 'source_selection_string_list', 'sink_selection_string_list',
             'number_of_paths']


def config_parser(config_file, parameters):
    for key in NECESSARY:
        parameters[key] = ''
    parameters['calc_contact_map_boolean'] = False
    parameters['summary_boolean'] = False
    parameters['nonstandard_substrates_selection'] = None
    parameters['homemade_selections'] = None
    parameters['alignment_selection'] = 'protein and name CA'
    parameters['traj_list'] = None
    parameters['trajectory_step'] = 1
    parameters['contact_map_distance_cutoff'] = 99999.9
    parameters['which_contact_map'] = 'average contact map'
    parameters['user_input_cartesian_covariance_matrix'] = None
    parameters['user_input_adjacency_matrix'] = None
    parameters['user_input_contact_map'] = None
    parameters['node_sphere_radius'] = 0.0
    parameters['node_sphere_rgb'] = (1.0, 1.0, 1.0)
    parameters['shortest_path_rgb'] = (1.0, 1.0, 1.0)
    parameters['longest_path_rgb'] = (1.0, 1.0, 1.0)
    parameters['shortest_path_radius'] = 0.0
    parameters['longest_path_radius'] = 0.0
    parameters['node_sphere_color_index'] = 34
    parameters['VMD_color_index_range'] = (35, 1056)
    parameters['VMD_resolution'] = 10
    parameters['VMD_spline_smoothness'] = 1
    # keys the reference's driver uses without declaring them (App. B4)
    parameters['adjacency_matrix_analysis_boolean'] = True
    parameters['weight_by_contact_map_boolean'] = False
    parameters['Lambda'] = None
    parameters['user_input_average_node_positions'] = None
    # extensions of this implementation (all optional)
    parameters['universe_factory'] = None      # 'module:function' returning a Universe-like object
    parameters['fused_pipeline_boolean'] = False   # coordinates -> cost matrix in one device pass

    with open(config_file) as f:
        exec(compile(f.read(), config_file, 'exec'), parameters)          # IO.py:54 (execfile)
    for key, value in parameters.items():
        if isinstance(value, str) and value == '' and key in NECESSARY:
            print('%s has not been assigned a value. This variable is necessary for the script to run. '
                  'Please declare this variable within the config file.' % (key))
            sys.exit()


def summary(summary_file_name, arguments, parameters):
    try:
        import MDAnalysis
        version = MDAnalysis.version.__version__
    except Exception:
        version = 'not installed (in-memory stand-in universe)'
    with open(summary_file_name, 'w') as f:
        f.write('Using MDAnalysis version: %s\n' % (version))
        f.write('\nAtom selections analyzed have been written out to node_selections.txt\n')
        f.write('To recreate this analysis, run this line:\n')
        for a in arguments:
            f.write('%s ' % (a))
        f.write('\n\n')
        f.write('Parameters used:\n')
        for key, value in parameters.items():
            if key == '__builtins__':
                continue
            if type(value) == int or type(value) == float:
                f.write("%s = %s\n" % (key, value))
            else:
                f.write("%s = '%s'\n" % (key, value))
