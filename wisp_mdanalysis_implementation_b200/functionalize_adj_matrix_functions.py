"""Drop-in replacement for ``func_pearson_correlation`` of the reference's
``functionalize_adj_matrix_functions.py`` (config key
``func_adjacency_matrix_functions_file``)."""
import numpy as np

from wisp_mdanalysis_implementation_b200 import _lib, textio


def func_pearson_correlation(adjacency_matrix, output_directory, contact_map=None, Lambda=None):
    """functionalize_adj_matrix_functions.py:10-24: -log|C|, optionally multiplied
    element-wise by the (binary or average-distance) contact map.  ``Lambda`` is ignored, as
    in the reference."""
    functionalized_correlation_file_name = output_directory + 'func_pearson_correlation.dat'
    ctx = _lib.default_context()
    cm = None if contact_map is None else np.asarray(contact_map, dtype=np.float64)
    functionalized_correlation = ctx.func_pearson(adjacency_matrix, cm)
    textio.savetxt(functionalized_correlation_file_name, functionalized_correlation)
    return functionalized_correlation


def generalized_correlation_coefficient_calc(adjacency_matrix, output_directory, contact_map=None, Lambda=None):
    """functionalize_adj_matrix_functions.py:27-46: r_MI = sqrt(1 - exp(-2 MI / 3)), damped by
    exp(-d_ij / Lambda) when both a contact map and Lambda are given; returns -log r_MI and writes
    both matrices."""
    ctx = _lib.default_context()
    cm = None if contact_map is None else np.asarray(contact_map, dtype=np.float64)
    if cm is not None and Lambda is not None:
        print('Applying a node-node distance-dependent exponential decay to the generalized correlation '
              'coefficient matrix (r_{MI}). The exponential damping factor (%.3f angstroms) is being used.' % Lambda)
    rMI, functionalized_rMI = ctx.func_generalized_correlation(adjacency_matrix, cm, Lambda)
    textio.savetxt(output_directory + 'gen_corr_coefficients.dat', rMI)
    textio.savetxt(output_directory + 'func_gen_corr_coefficients.dat', functionalized_rMI)
    return functionalized_rMI


def do_nothing(adjacency_matrix, output_directory, contact_map=None, Lambda=None):
    """functionalize_adj_matrix_functions.py:68-71."""
    return adjacency_matrix
