"""Adaptors between the reference's Python objects and the raw arrays the C ABI takes.

* ``selection_to_csr``: the reference's ``selection_list`` (MDAnalysis AtomGroups for 'COM'
  nodes, Atoms for 'ATOMIC' nodes -- make_node_selections.py:39-70, and arbitrary
  sub-residue groups for nucleic acids :121-388) -> CSR ``(node_ptr, atom_idx)``.
* ``universe_coordinates``: the (T, A, 3) float32 block of a Universe's current trajectory.
  A MemoryReader exposes it directly (zero copy); any other reader is iterated once.

Works with real MDAnalysis objects and with ``mda_standin`` (same attribute names).
"""
from __future__ import annotations

import numpy as np


def selection_to_csr(selection_list):
    ptr = [0]
    idx = []
    for sel in selection_list:
        if hasattr(sel, "indices"):                 # AtomGroup
            ix = np.asarray(sel.indices, dtype=np.int64)
        elif hasattr(sel, "index"):                 # Atom
            ix = np.asarray([sel.index], dtype=np.int64)
        elif hasattr(sel, "ix"):
            ix = np.atleast_1d(np.asarray(sel.ix, dtype=np.int64))
        else:
            raise TypeError("selection_list entry %r is neither an AtomGroup nor an Atom" % (sel,))
        if len(ix) == 0:
            raise ValueError("a node selection is empty")
        idx.append(ix)
        ptr.append(ptr[-1] + len(ix))
    return np.asarray(ptr, dtype=np.int32), np.concatenate(idx).astype(np.int32)


def universe_masses(universe):
    return np.asarray(universe.atoms.masses, dtype=np.float64)


def universe_coordinates(universe, step=1):
    """(T, A, 3) float32, C order, frames [::step] of the loaded trajectory."""
    traj = universe.trajectory
    arr = None
    if hasattr(traj, "get_array"):                  # MemoryReader / stand-in
        arr = traj.get_array()
        order = getattr(traj, "stored_order", "fac")
        if order != "fac":
            arr = np.transpose(arr, [order.index(c) for c in "fac"])
    if arr is not None:
        arr = arr[::step]
        return np.ascontiguousarray(arr, dtype=np.float32)
    n = len(traj)
    frames = range(0, n, step)
    out = np.empty((len(frames), universe.atoms.n_atoms, 3), dtype=np.float32)
    for k, ts in enumerate(traj[::step]):
        out[k] = ts.positions
    return out
