"""Stage-boundary text files (np.savetxt '%.18e' / '%i', np.loadtxt) -- the reference's
checkpoint format.  Set WISP_B200_SKIP_TEXT=1 to skip writing the
N^2 text matrices (they dominate wall time after the GPU path at large N)."""
import os

import numpy as np


def savetxt(path, array, fmt='%.18e', header=''):
    if os.environ.get('WISP_B200_SKIP_TEXT') == '1':
        return
    np.savetxt(path, array, fmt=fmt, header=header)


def loadtxt(path):
    return np.loadtxt(path)
