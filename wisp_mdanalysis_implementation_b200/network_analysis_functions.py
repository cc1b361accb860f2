"""Drop-in replacement for the reference's ``network_analysis_functions.py`` (config key
``network_functions_file``): ``get_paths`` with the exact contract of SURVEY.md Appendix A,
computed by the blocked Floyd-Warshall (K5) and bounded-DFS enumeration (K6) kernels."""
import sys

import numpy as np

from wisp_mdanalysis_implementation_b200 import _lib


def get_paths(adjacency_matrix, sources, sinks, number_of_paths):
    """network_analysis_functions.py:12-214.  Returns ``[[length, n0, n1, ...], ...]`` sorted
    ascending, K entries (K+1 in the case the reference returns K+1)."""
    ctx = _lib.default_context()
    W = np.asarray(adjacency_matrix, dtype=np.float64)
    try:
        paths = ctx.get_paths(W, list(sources), list(sinks), number_of_paths)
    except _lib.WispError as e:
        # the reference's error convention is print + sys.exit
        print('get_paths failed:', e)
        sys.exit()
    print('Shortest pathway information:', paths[0][0], paths[0][1:])
    return paths
