"""``python wisp_w_mdanalysis.py <config>`` -- the reference's command line and stage order
(wisp_w_mdanalysis.py:5,54,60-208), for Python 3, with the defects of SURVEY.md Appendix B
repaired (B1 output names, B2 stray mkdir, B3 adjacency call arity, B4 undeclared keys).

Like the reference, stage implementations are looked up BY NAME from the module files the
config names (wisp_w_mdanalysis.config:7-12), so any stage can point at the reference's own
module or at the B200 module of the same name in this directory.  Nothing is computed here.
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (_HERE, os.path.dirname(_HERE)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from IO import config_parser, summary  # noqa: E402
from wisp_mdanalysis_implementation_b200 import textio  # noqa: E402

STYLES = {   # adjacency_matrix_style -> (adjacency function, functionalisation function)
    'PEARSON': ('pearson_correlation_analysis', 'func_pearson_correlation'),
    'PEARSON CORRELATION': ('pearson_correlation_analysis', 'func_pearson_correlation'),
    'LMI': ('linear_mutual_information_analysis', 'generalized_correlation_coefficient_calc'),
    'LINEAR MUTUAL INFORMATION': ('linear_mutual_information_analysis', 'generalized_correlation_coefficient_calc'),
}


def _fail(message):
    """The reference's error convention: print, then sys.exit()."""
    print(message)
    sys.exit()


class WispRun:
    def __init__(self, config_file, argv=None):
        self.argv = list(argv) if argv is not None else list(sys.argv)
        self.p = {}
        config_parser(config_file, self.p)
        out = self.p['output_directory']
        self.out = out if out.endswith(os.sep) else out + os.sep
        self.p['output_directory'] = self.out
        if os.path.exists(self.out):
            _fail('The output directory, %s already exists. Please select a different directory name '
                  'for output.' % self.out)
        os.mkdir(self.out)
        self._bind_stages()

    # -- stage lookup (wisp_w_mdanalysis.py:233-267) ---------------------------------------
    def _load(self, key, name):
        module = importlib.import_module(self.p[key].split('.')[0], package=None)
        return getattr(module, name)

    def _bind_stages(self):
        p = self.p
        self.make_selections = self._load('node_selection_file', 'make_selections')
        if p['trajectory_analysis_boolean']:
            self.align = self._load('trajectory_functions_file', 'traj_alignment_and_averaging')
            self.covariance = self._load('trajectory_functions_file', 'cartesian_covariance')
            self.contact = self._load('trajectory_functions_file', 'calc_contact_map')
        style = p['adjacency_matrix_style'].upper()
        if style not in STYLES:
            _fail("adjacency_matrix_style %r is not available on the accelerated path; use one of %s"
                  % (p['adjacency_matrix_style'], sorted(STYLES)))
        adj_name, func_name = STYLES[style]
        self.adjacency = self._load('adjacency_matrix_functions_file', adj_name)
        self.functionalize = self._load('func_adjacency_matrix_functions_file', func_name)
        self.get_paths = self._load('network_functions_file', 'get_paths')
        vis = p['visualization_functions_file']
        self.vis_state = None if vis in (None, 'None', 'none') else self._load('visualization_functions_file', 'create_vis_state')

    def _universe(self):
        factory = self.p['universe_factory']
        if factory:
            module, function = factory.split(':')
            return getattr(importlib.import_module(module), function)(self.p)
        import MDAnalysis
        return MDAnalysis.Universe(self.p['pdb'])

    # -- stages -----------------------------------------------------------------------------
    def select_nodes(self):
        p = self.p
        self.universe = self._universe()
        self.nodes, self.sources, self.sinks = self.make_selections(
            self.universe, self.out + 'node_selections.txt', p['substrate_node_definition'],
            p['substrate_selection_string'], p['source_selection_string_list'],
            p['sink_selection_string_list'],
            nonstandard_substrates_selection=p['nonstandard_substrates_selection'],
            homemade_selections=p['homemade_selections'])
        print('Number of nodes: %d\nsource node indices: %s\nsink node indices: %s'
              % (len(self.nodes), self.sources, self.sinks))

    def analyze_trajectory(self):
        """Stage 4 with a trajectory (reference :80-104)."""
        p = self.p
        trajectory, self.avg = self.align(self.universe, p['alignment_selection'], self.nodes,
                                          p['substrate_node_definition'], p['traj_list'], self.out,
                                          step=p['trajectory_step'], convergence_threshold=1E-5)
        self.cov = self.covariance(trajectory, self.avg, self.out)
        binary, mean_distance = self.contact(trajectory, self.out,
                                             distance_cutoff=p['contact_map_distance_cutoff'])
        which = p['which_contact_map'].upper()
        if which == 'AVERAGE CONTACT MAP':
            self.contact_map = mean_distance
        elif which == 'BINARY CONTACT MAP':
            self.contact_map = binary
        else:
            _fail("which_contact_map must be 'AVERAGE CONTACT MAP' or 'BINARY CONTACT MAP'.")

    def analyze_trajectory_fused(self):
        """Extension (config key ``fused_pipeline_boolean``): stages 4a-6 in ONE device pass --
        K1, K7, K2, K3, K4 -- the node trajectory never leaves the GPU and the 3N x 3N
        covariance is never materialised.  Writes the N x N stage files only."""
        import numpy as np
        from wisp_mdanalysis_implementation_b200 import _adapt, _lib
        p = self.p
        # config key ``gpus``: frames sharded over that many GPUs inside the one C call (0 = all)
        ctx = _lib.default_context(gpus=p['gpus']) if p['gpus'] != 1 else _lib.default_context()
        align_idx = np.asarray(self.universe.select_atoms(p['alignment_selection']).indices, dtype=np.int32)
        node_ptr, atom_idx = _adapt.selection_to_csr(self.nodes)
        blocks = []
        for traj in p['traj_list']:
            self.universe.load_new(traj)
            if len(self.universe.trajectory) % p['trajectory_step'] != 0:
                _fail('User defined step size is not a factor of the number of frames in %s.' % traj)
            blocks.append(_adapt.universe_coordinates(self.universe, p['trajectory_step']))
        coords = blocks[0] if len(blocks) == 1 else np.concatenate(blocks)
        which = {'AVERAGE CONTACT MAP': 2, 'BINARY CONTACT MAP': 1}.get(p['which_contact_map'].upper())
        if which is None:
            _fail("which_contact_map must be 'AVERAGE CONTACT MAP' or 'BINARY CONTACT MAP'.")
        out = ctx.pipeline(coords, _adapt.universe_masses(self.universe), node_ptr, atom_idx, align_idx,
                           cutoff=p['contact_map_distance_cutoff'],
                           which_map=which if p['weight_by_contact_map_boolean'] else 0,
                           atomic=p['substrate_node_definition'].upper() == 'ATOMIC',
                           convergence_threshold=1E-5, maximum_num_iterations=100)
        for iteration, residual, analysis_RMSD in out['alignment_log']:
            print('Iteration ', iteration, ': RMSD btw alignment landmarks: ', residual,
                  ', RMSD btw Node Positions: ', analysis_RMSD)
        textio.savetxt(self.out + 'average_node_positions.dat', out['avg'],
                       header='Shape: nNodes x 3 (%d x 3)' % len(self.nodes))
        textio.savetxt(self.out + 'binary_contact_map.dat', out['binary'], fmt='%i')
        textio.savetxt(self.out + 'avg_distance_contact_map.dat', out['avgdist'])
        textio.savetxt(self.out + 'node_correlation.dat', out['corr'])
        textio.savetxt(self.out + 'func_pearson_correlation.dat', out['w'])
        self.cost = out['w']

    def load_matrices(self):
        """Stage 4 from user-supplied text matrices (reference :106-123)."""
        p, n = self.p, len(self.nodes)
        self.cov = textio.loadtxt(p['user_input_cartesian_covariance_matrix'])
        if self.cov.ndim != 2 or self.cov.shape[0] % 3 or self.cov.shape[0] != self.cov.shape[1]:
            _fail('The supplied cartesian covariance matrix has shape %s, expected (%d, %d).'
                  % (self.cov.shape, 3 * n, 3 * n))
        self.avg = self.contact_map = None
        if p['user_input_average_node_positions']:
            self.avg = textio.loadtxt(p['user_input_average_node_positions'])
            if self.avg.shape != (n, 3):
                _fail('The supplied average node positions have shape %s, expected (%d, 3).' % (self.avg.shape, n))
        if p['user_input_contact_map']:
            self.contact_map = textio.loadtxt(p['user_input_contact_map'])
            if self.contact_map.shape != (n, n):
                _fail('The supplied contact map has shape %s, expected (%d, %d).' % (self.contact_map.shape, n, n))

    def build_cost_matrix(self):
        p, n = self.p, len(self.nodes)
        if p['adjacency_matrix_analysis_boolean']:
            adjacency = self.adjacency(self.cov, self.avg, self.out)          # B3: three arguments
        else:
            adjacency = textio.loadtxt(p['user_input_adjacency_matrix'])
            if adjacency.shape != (n, n):
                _fail('The supplied adjacency matrix has shape %s, expected (%d, %d).' % (adjacency.shape, n, n))
        if p['weight_by_contact_map_boolean']:
            self.cost = self.functionalize(adjacency, self.out, contact_map=self.contact_map, Lambda=p['Lambda'])
        else:
            self.cost = self.functionalize(adjacency, self.out)

    def find_paths(self):
        self.paths = self.get_paths(self.cost, self.sources, self.sinks, self.p['number_of_paths'])
        with open(self.out + 'simply_formatted_paths.txt', 'w') as handle:   # reference :157-165
            handle.write('# Pathway_Length Node_Indices (zero indexed)\n')
            for path in self.paths:
                handle.write('%f ' % path[0] + ''.join('%d ' % node for node in path[1:]) + '\n')

    def write_extras(self):
        p = self.p
        if self.vis_state is not None:      # VMD state: outside the accelerated path, optional
            self.universe.load_new(p['visualization_frame_pdb'])
            keys = ('node_sphere_radius', 'node_sphere_rgb', 'shortest_path_radius', 'shortest_path_rgb',
                    'longest_path_radius', 'longest_path_rgb', 'node_sphere_color_index',
                    'VMD_color_index_range', 'VMD_resolution', 'VMD_spline_smoothness')
            self.vis_state(p['visualization_frame_pdb'], self.nodes, self.paths,
                           self.out + 'pathways_vis_state.vmd', **{k: p[k] for k in keys})
        if p['summary_boolean']:
            summary(self.out + 'summary.txt', self.argv, p)

    def run(self):
        self.select_nodes()
        fused = (self.p['fused_pipeline_boolean'] and self.p['trajectory_analysis_boolean']
                 and self.p['adjacency_matrix_style'].upper() in ('PEARSON', 'PEARSON CORRELATION'))
        if fused:
            self.analyze_trajectory_fused()
        else:
            if self.p['trajectory_analysis_boolean']:
                self.analyze_trajectory()
            else:
                self.load_matrices()
            self.build_cost_matrix()
        self.find_paths()
        self.write_extras()
        return self.paths


def main(config_file, argv=None):
    return WispRun(config_file, argv).run()


if __name__ == '__main__':
    main(sys.argv[1])
