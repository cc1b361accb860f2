"""Config handling and run summary for the driver (boundary code; nothing is accelerated).

Behaviour follows the reference's ``IO.py``: the config file is Python source executed into
the ``parameters`` dict (IO.py:54), sixteen keys are mandatory (IO.py:24,55-58), the rest have
defaults (IO.py:32-51).  Keys the reference's driver reads without ever declaring
(wisp_w_mdanalysis.py:114,128,141,142 -- SURVEY.md App. B4) get defaults here as well, and two
extension keys exist: ``universe_factory``, ``fused_pipeline_boolean`` and ``gpus`` (GPUs of this
box the fused pipeline shards the frames over; 0 = all of them).
"""
import sys

MANDATORY = (
    'output_directory', 'node_selection_file', 'trajectory_functions_file',
    'adjacency_matrix_functions_file', 'func_adjacency_matrix_functions_file',
    'network_functions_file', 'visualization_functions_file', 'trajectory_analysis_boolean',
    'adjacency_matrix_style', 'pdb', 'visualization_frame_pdb', 'substrate_node_definition',
    'substrate_selection_string', 'source_selection_string_list', 'sink_selection_string_list',
    'number_of_paths',
)

DEFAULTS = dict(
    # the reference's optional keys
    calc_contact_map_boolean=False, summary_boolean=False, nonstandard_substrates_selection=None,
    homemade_selections=None, alignment_selection='protein and name CA', traj_list=None,
    trajectory_step=1, contact_map_distance_cutoff=99999.9, which_contact_map='average contact map',
    user_input_cartesian_covariance_matrix=None, user_input_adjacency_matrix=None,
    user_input_contact_map=None, node_sphere_radius=0.0, node_sphere_rgb=(1.0, 1.0, 1.0),
    shortest_path_rgb=(1.0, 1.0, 1.0), longest_path_rgb=(1.0, 1.0, 1.0), shortest_path_radius=0.0,
    longest_path_radius=0.0, node_sphere_color_index=34, VMD_color_index_range=(35, 1056),
    VMD_resolution=10, VMD_spline_smoothness=1,
    # read by the reference's driver but never declared there (App. B4)
    adjacency_matrix_analysis_boolean=True, weight_by_contact_map_boolean=False, Lambda=None,
    user_input_average_node_positions=None,
    # extensions of this implementation
    universe_factory=None, fused_pipeline_boolean=False, gpus=1,
)


def config_parser(config_file, parameters):
    """Fill ``parameters`` from ``config_file``; print + exit when a mandatory key is unset."""
    parameters.update({key: '' for key in MANDATORY})
    parameters.update(DEFAULTS)
    with open(config_file) as handle:
        exec(compile(handle.read(), config_file, 'exec'), parameters)
    unset = [key for key in MANDATORY if isinstance(parameters[key], str) and parameters[key] == '']
    for key in unset:
        print('%s has not been assigned a value. This variable is necessary for the script to run. '
              'Please declare this variable within the config file.' % key)
    if unset:
        sys.exit()


def summary(summary_file_name, arguments, parameters):
    """Re-run recipe: MDAnalysis version, the command line, every parameter (IO.py:60-84)."""
    try:
        import MDAnalysis
        version = MDAnalysis.version.__version__
    except Exception:
        version = 'not installed (in-memory stand-in universe)'
    lines = ['Using MDAnalysis version: %s' % version, '',
             'Atom selections analyzed have been written out to node_selections.txt',
             'To recreate this analysis, run this line:', ' '.join(str(a) for a in arguments) + ' ', '',
             'Parameters used:']
    for key, value in parameters.items():
        if key == '__builtins__':
            continue
        numeric = type(value) in (int, float)
        lines.append(("%s = %s" if numeric else "%s = '%s'") % (key, value))
    with open(summary_file_name, 'w') as handle:
        handle.write('\n'.join(lines) + '\n')
