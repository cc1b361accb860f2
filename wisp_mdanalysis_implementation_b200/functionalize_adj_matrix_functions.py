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


def do_nothing(adjacency_matrix, output_directory, contact_map=None, Lambda=None):
    """functionalize_adj_matrix_functions.py:68-71."""
    return adjacency_matrix
