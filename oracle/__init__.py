"""CPU oracle for the WISP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package, and there only as the
checker (or as the CPU baseline being timed), never as the thing shipped.  The product
package ``wisp_mdanalysis_implementation_b200`` never imports it and fails loudly when its
CUDA library is missing.

What it is: a Python-3 / numpy restatement of the reference's algorithm for the path
(rbdavid/WISP_mdanalysis_implementation, Python 2 + networkx 1.x + MDAnalysis, which
cannot be imported here).  Every function cites the reference ``file:line`` it follows.

How it is pinned: ``tests/golden/make_golden.py`` executes the reference's OWN source files
(read from /root/reference at generation time, mechanically adapted in memory: tabs
expanded, ``print`` statements parenthesised, Python-2 ``/`` emulated, MDAnalysis /
networkx-1.x call sites served by thin shims) on seeded inputs and commits the outputs as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this oracle against those
fixtures.  The third-party arithmetic the reference borrows (MDAnalysis ``center_of_mass``,
``distance_array``, ``rotation_matrix``; networkx Dijkstra) is NOT vendored in the
reference and its versions are unpinned there -- for those call sites parity is pinned to
the published semantics restated in ``oracle/wisp_oracle.py`` ("parity unpinned" beyond
that).
"""
from .wisp_oracle import *  # noqa: F401,F403
