"""Synthetic trajectories for the WISP hot path.

The reference ships no sample data (its example config points at the author's private
files, wisp_w_mdanalysis.config:20-21,32), so tests and ``bench.py`` use this generator:

* node centres on a jittered cubic lattice filled in boustrophedon (chain-contiguous) order,
  spacing 5.5 A, jitter N(0, 0.5 A);
* ``atoms_per_node`` atoms per node at N(0, 1.2 A) offsets, masses drawn from
  {H, C, N, O, S} with protein-like frequencies;
* dynamics = sum of 16 smooth collective modes (amplitude 1.5 exp(-m/4) A, iid N(0,1)
  time coefficients) + white noise N(0, 0.3 A) per atom; no global rotation/translation;
* coordinates float32, shape (T, A, 3), C order (MDAnalysis MemoryReader 'fac' layout).

``SynthSystem`` holds the static part (small); ``frames(t0, t1)`` materialises a block of
frames with numpy so arbitrarily long trajectories can be streamed.  The CUDA library has a
device-side generator of the same model for bench-scale inputs
(``wisp_dev_synth_coords``, csrc/synth.cu).
"""
from __future__ import annotations

import numpy as np

_MASSES = np.array([1.008, 12.011, 14.007, 15.999, 32.06])
_MASS_P = np.array([0.50, 0.32, 0.08, 0.09, 0.01])
N_MODES = 16

CONFIG_SEEDS = {"C1": 1001, "C2": 1002, "C3": 1003, "C4": 1004, "C5": 1005}


class SynthSystem:
    """Static description of a synthetic system plus its per-frame mode coefficients."""

    def __init__(self, n_nodes, n_frames, atoms_per_node=16, seed=1001):
        rng = np.random.default_rng(seed)
        self.seed = int(seed)
        self.n_nodes = N = int(n_nodes)
        self.n_frames = T = int(n_frames)
        self.atoms_per_node = apn = int(atoms_per_node)
        self.n_atoms = A = N * apn
        side = int(np.ceil(N ** (1.0 / 3.0)))
        # boustrophedon lattice walk: consecutive nodes are always lattice neighbours
        cells = []
        for z in range(side):
            ys = range(side) if z % 2 == 0 else range(side - 1, -1, -1)
            for yi, y in enumerate(ys):
                fwd = (yi % 2 == 0) ^ (z % 2 == 1)
                xs = range(side) if fwd else range(side - 1, -1, -1)
                for x in xs:
                    cells.append((x, y, z))
        cells = np.array(cells[:N], dtype=np.float64)
        self.node_base = cells * 5.5 + rng.normal(0.0, 0.5, size=(N, 3))
        self.node_of_atom = np.repeat(np.arange(N, dtype=np.int32), apn)
        self.atom_base = (self.node_base[self.node_of_atom]
                          + rng.normal(0.0, 1.2, size=(A, 3)))
        self.masses = rng.choice(_MASSES, size=A, p=_MASS_P).astype(np.float64)
        # smooth vector modes: cos(k.r + phase) per component
        kdir = rng.normal(size=(N_MODES, 3))
        kdir /= np.linalg.norm(kdir, axis=1, keepdims=True)
        wavelength = rng.uniform(25.0, 80.0, size=(N_MODES, 1))
        self.kvec = 2.0 * np.pi * kdir / wavelength
        self.phase = rng.uniform(0.0, 2.0 * np.pi, size=(N_MODES, 3))
        self.amp = 1.5 * np.exp(-np.arange(N_MODES) / 4.0)
        arg = self.node_base @ self.kvec.T                      # (N, M)
        self.modes = (np.cos(arg[:, :, None] + self.phase[None, :, :])
                      * self.amp[None, :, None])                # (N, M, 3)
        self.coeff = rng.normal(size=(T, N_MODES))              # a_m(t)
        self.noise_sigma = 0.3
        # CSR node -> atoms map (contiguous residues, like make_node_selections.py:39-58)
        self.node_ptr = (np.arange(N + 1, dtype=np.int32) * apn).astype(np.int32)
        self.atom_idx = np.arange(A, dtype=np.int32)
        # alignment landmarks: first atom of every node ("name CA")
        self.align_idx = self.node_ptr[:-1].copy()

    def frames(self, t0=0, t1=None, block=256):
        """Float32 coordinates (t1-t0, A, 3) of frames [t0, t1) -- deterministic per frame."""
        t1 = self.n_frames if t1 is None else int(t1)
        out = np.empty((t1 - t0, self.n_atoms, 3), dtype=np.float32)
        for b0 in range(t0, t1, block):
            b1 = min(b0 + block, t1)
            disp = np.einsum("tm,nmc->tnc", self.coeff[b0:b1], self.modes)   # (b, N, 3)
            x = self.atom_base[None] + disp[:, self.node_of_atom, :]
            for t in range(b0, b1):
                r = np.random.default_rng([self.seed, 7919, t])
                x[t - b0] += r.normal(0.0, self.noise_sigma, size=(self.n_atoms, 3))
            out[b0 - t0:b1 - t0] = x.astype(np.float32)
        return out

    def coords(self):
        return self.frames(0, self.n_frames)


def synth(n_nodes, n_frames, atoms_per_node=16, seed=1001):
    """Convenience: (SynthSystem, coords float32 (T, A, 3))."""
    s = SynthSystem(n_nodes, n_frames, atoms_per_node, seed)
    return s, s.coords()


def random_cost_matrix(n_nodes, seed=0, density=0.15, lo=0.05, hi=3.0):
    """Symmetric non-negative cost matrix with zero (= no edge) entries, plus a Hamiltonian
    chain so the graph is connected -- a stand-in for -log|C| (.) binary contact map."""
    rng = np.random.default_rng(seed)
    N = int(n_nodes)
    W = rng.uniform(lo, hi, size=(N, N))
    mask = rng.random((N, N)) < density
    W = np.where(mask, W, 0.0)
    W = np.triu(W, 1)
    idx = np.arange(N - 1)
    chain = rng.uniform(lo, hi, size=N - 1)
    W[idx, idx + 1] = chain
    W = W + W.T
    return W
