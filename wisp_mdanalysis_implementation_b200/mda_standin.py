"""Minimal stand-in for the slice of MDAnalysis the WISP hot path touches.

MDAnalysis is not installable in the build container, so the drop-in modules are written
against the *attributes* the reference uses (``select_atoms``, ``AtomGroup.positions /
center_of_mass / translate / n_atoms / n_residues / residues / resnames / resids``,
``Atom.position``, ``Universe.load_new``, ``len(u.trajectory)``, ``u.trajectory[::step]``)
and accept either a real ``MDAnalysis.Universe`` or this stand-in.  The stand-in is an
in-memory ("MemoryReader"-like) universe: coordinates are a float32 ``(T, A, 3)`` array in
'fac' order and ``ts.positions`` is a *view* of the current frame, exactly like
MDAnalysis' MemoryReader, so ``translate`` edits the trajectory in place.

Selection language: ``all``, ``protein`` (every atom of the stand-in), ``name A B ..``,
``resname A B ..``, ``resid 3 7:12 ..``, ``index i:j``, combined with ``and`` / ``or`` /
``not`` and parentheses -- enough for the reference's config strings
(wisp_w_mdanalysis.config:29,31) on synthetic systems.
"""
from __future__ import annotations

import re
import numpy as np


class Timestep:
    def __init__(self, positions, frame):
        self.positions = positions
        self.frame = frame


class MemoryTrajectory:
    """(T, A, 3) float32 block; iteration sets ``universe.trajectory.ts``."""

    def __init__(self, coordinate_array):
        arr = np.asarray(coordinate_array)
        if arr.dtype != np.float32 or arr.ndim != 3 or arr.shape[2] != 3:
            raise TypeError("MemoryTrajectory needs a float32 (T, A, 3) array")
        self.coordinate_array = arr
        self.n_frames = arr.shape[0]
        self.ts = Timestep(arr[0], 0)

    def get_array(self):
        return self.coordinate_array

    def __len__(self):
        return self.n_frames

    def _goto(self, i):
        self.ts = Timestep(self.coordinate_array[i], i)
        return self.ts

    def __iter__(self):
        for i in range(self.n_frames):
            yield self._goto(i)
        self._goto(0)

    def __getitem__(self, item):
        if isinstance(item, slice):
            def gen():
                for i in range(*item.indices(self.n_frames)):
                    yield self._goto(i)
            return gen()
        return self._goto(int(item) % self.n_frames)


class Atom:
    def __init__(self, universe, ix):
        self.universe = universe
        self.ix = self.index = int(ix)

    @property
    def position(self):
        return self.universe.trajectory.ts.positions[self.ix]

    @property
    def mass(self):
        return self.universe._masses[self.ix]

    @property
    def resname(self):
        return self.universe._resnames[self.ix]

    @property
    def resid(self):
        return int(self.universe._resids[self.ix])

    @property
    def name(self):
        return self.universe._names[self.ix]


class Residue:
    def __init__(self, universe, ixs):
        self.universe = universe
        self.atoms = AtomGroup(universe, ixs)
        self.resname = universe._resnames[ixs[0]]
        self.resid = int(universe._resids[ixs[0]])


class AtomGroup:
    def __init__(self, universe, indices):
        self.universe = universe
        self.indices = self.ix = np.asarray(indices, dtype=np.int64)

    def __len__(self):
        return len(self.indices)

    @property
    def n_atoms(self):
        return len(self.indices)

    @property
    def atoms(self):
        return _AtomList(self)

    @property
    def masses(self):
        return self.universe._masses[self.indices]

    @property
    def resnames(self):
        return self.universe._resnames[self.indices]

    @property
    def resids(self):
        return self.universe._resids[self.indices]

    @property
    def names(self):
        return self.universe._names[self.indices]

    @property
    def positions(self):
        return self.universe.trajectory.ts.positions[self.indices]

    @property
    def residues(self):
        key = self.universe._resindex[self.indices]
        order = []
        seen = {}
        for k, i in zip(key, self.indices):
            if k not in seen:
                seen[k] = []
                order.append(k)
            seen[k].append(i)
        return [Residue(self.universe, np.array(seen[k])) for k in order]

    @property
    def n_residues(self):
        return len(np.unique(self.universe._resindex[self.indices]))

    def center_of_mass(self):
        m = self.masses
        p = self.positions
        return np.sum(p * m[:, None], axis=0) / m.sum()

    def translate(self, t):
        vector = np.asarray(t)
        pos = self.universe.trajectory.ts.positions
        tmp = pos[self.indices]
        tmp += vector                      # float32 += float64 -> rounded to float32
        pos[self.indices] = tmp
        return self

    def select_atoms(self, sel):
        mask = np.zeros(self.universe.n_atoms, dtype=bool)
        mask[self.indices] = True
        full = self.universe._select_mask(sel)
        return AtomGroup(self.universe, np.nonzero(mask & full)[0])


class _AtomList:
    def __init__(self, group):
        self.group = group

    def __getitem__(self, i):
        return Atom(self.group.universe, self.group.indices[i])

    def __len__(self):
        return len(self.group.indices)

    @property
    def indices(self):
        return self.group.indices

    @property
    def resnames(self):
        return self.group.resnames

    @property
    def resids(self):
        return self.group.resids

    def __getattr__(self, name):
        return getattr(self.group, name)


class Universe:
    """In-memory universe.  ``topology`` = dict(names, resnames, resids, masses)."""

    def __init__(self, names, resnames, resids, masses, coordinates=None):
        self._names = np.asarray(names)
        self._resnames = np.asarray(resnames)
        self._resids = np.asarray(resids, dtype=np.int64)
        self._masses = np.asarray(masses, dtype=np.float64)
        self.n_atoms = len(self._masses)
        change = np.ones(self.n_atoms, dtype=bool)
        change[1:] = (self._resids[1:] != self._resids[:-1]) | (self._resnames[1:] != self._resnames[:-1])
        self._resindex = np.cumsum(change) - 1
        if coordinates is None:
            coordinates = np.zeros((1, self.n_atoms, 3), dtype=np.float32)
        self.trajectory = MemoryTrajectory(coordinates)
        self.atoms = AtomGroup(self, np.arange(self.n_atoms))

    @classmethod
    def from_synth(cls, system, coordinates=None):
        """Universe for a ``synth.SynthSystem``: residue r = node r, first atom 'CA'."""
        apn = system.atoms_per_node
        names = np.array((["CA"] + ["X%d" % k for k in range(1, apn)]) * system.n_nodes)
        resids = np.repeat(np.arange(1, system.n_nodes + 1), apn)
        resnames = np.array(["ALA"] * system.n_atoms)
        return cls(names, resnames, resids, system.masses, coordinates)

    def load_new(self, trajectory):
        if isinstance(trajectory, str):
            trajectory = np.load(trajectory, mmap_mode=None)
        self.trajectory = MemoryTrajectory(trajectory)
        return self

    def select_atoms(self, sel):
        return AtomGroup(self, np.nonzero(self._select_mask(sel))[0])

    # --- tiny selection parser -------------------------------------------------------
    def _select_mask(self, sel):
        tokens = re.findall(r"\(|\)|[^\s()]+", sel)
        pos = [0]

        def peek():
            return tokens[pos[0]] if pos[0] < len(tokens) else None

        def take():
            if pos[0] >= len(tokens):
                raise ValueError("selection %r ends unexpectedly" % sel)
            tok = tokens[pos[0]]
            pos[0] += 1
            return tok

        keywords = {"and", "or", "not", "(", ")", "all", "protein", "name", "resname",
                    "resid", "index"}

        def values():
            vals = []
            while peek() is not None and peek() not in keywords:
                vals.append(take())
            return vals

        def ranges(vals, arr):
            m = np.zeros(self.n_atoms, dtype=bool)
            for v in vals:
                if ":" in v or "-" in v[1:]:
                    a, b = re.split(r"[:\-]", v, maxsplit=1)
                    m |= (arr >= int(a)) & (arr <= int(b))
                else:
                    m |= arr == int(v)
            return m

        def atom():
            tok = take()
            if tok == "(":
                m = expr()
                if take() != ")":
                    raise ValueError("unbalanced parentheses in selection %r" % sel)
                return m
            if tok == "not":
                return ~atom()
            if tok in ("all", "protein"):
                return np.ones(self.n_atoms, dtype=bool)
            if tok == "name":
                return np.isin(self._names, values())
            if tok == "resname":
                return np.isin(self._resnames, values())
            if tok == "resid":
                return ranges(values(), self._resids)
            if tok == "index":
                return ranges(values(), np.arange(self.n_atoms))
            raise ValueError("selection token %r not supported by the stand-in" % tok)

        def conj():
            m = atom()
            while peek() == "and":
                take()
                m = m & atom()
            return m

        def expr():
            m = conj()
            while peek() == "or":
                take()
                m = m | conj()
            return m

        mask = expr()
        if pos[0] != len(tokens):
            raise ValueError("could not parse selection %r" % sel)
        return mask


def synthetic_universe(parameters):
    """``universe_factory`` hook of the driver (IO.py extension key): builds the stand-in
    universe of a ``synth.SynthSystem`` described by ``parameters['synthetic']`` =
    dict(n_nodes, n_frames, atoms_per_node, seed)."""
    from . import synth
    spec = parameters['synthetic']
    system = synth.SynthSystem(spec['n_nodes'], spec['n_frames'], spec.get('atoms_per_node', 16),
                               spec.get('seed', 1001))
    return Universe.from_synth(system)
