"""Drop-in replacement for the Pearson part of the reference's
``adjacency_matrix_functions.py`` (config key ``adjacency_matrix_functions_file``).  The hENM
style is out of scope (it does not run in the reference either)."""
from wisp_mdanalysis_implementation_b200 import _lib, textio


def pearson_correlation_analysis(cartesian_covariance_matrix, average_node_positions=None,
                                 output_directory=None, distance_cutoff=9999., threshold=1e-4,
                                 max_iterations=100, guess=None, alpha=0.1):
    """adjacency_matrix_functions.py:8-59.  The reference's driver passes only
    ``(covariance, output_directory)`` (wisp_w_mdanalysis.py:130); both call
    forms are accepted."""
    if output_directory is None and isinstance(average_node_positions, str):
        output_directory, average_node_positions = average_node_positions, None
    print('Beginning Pearson correlation coefficient analysis.')
    ctx = _lib.default_context()
    node_variance, node_covariance, node_correlation = ctx.pearson_from_covariance(
        cartesian_covariance_matrix)
    print('Pearson correlation coefficients have been calculated from the cartesian covariance.')
    if output_directory is not None:
        textio.savetxt(output_directory + 'node_variance.dat', node_variance)
        textio.savetxt(output_directory + 'node_covariance.dat', node_covariance)
        textio.savetxt(output_directory + 'node_correlation.dat', node_correlation)
    return node_correlation


def linear_mutual_information_analysis(cartesian_covariance_matrix, average_node_positions=None,
                                       output_directory=None, distance_cutoff=9999., threshold=1e-4,
                                       max_iterations=100, guess=None, alpha=0.1):
    """adjacency_matrix_functions.py:62-103: linear mutual information from
    the 3x3 / 6x6 covariance determinants of every node pair, +inf on the diagonal."""
    if output_directory is None and isinstance(average_node_positions, str):
        output_directory, average_node_positions = average_node_positions, None
    ctx = _lib.default_context()
    MI = ctx.lmi_from_covariance(cartesian_covariance_matrix)
    print('LMI has been calculated from the cartesian covariance.')
    if output_directory is not None:
        textio.savetxt(output_directory + 'linear_mutual_information.dat', MI)
    return MI
