"""Closed-form cases: inputs with answers known without running any implementation.
Shared by the CPU oracle tests and the GPU parity tests."""
import numpy as np


def correlated_trajectory(T=64):
    """4 nodes: 1 follows 0 exactly (C = 1), 2 mirrors 0 (C = -1), 3 moves orthogonally to 0
    with zero cross moment (C = 0 exactly in floating point: every product sums to +-pairs)."""
    t = np.arange(T)
    a = np.where(t % 2 == 0, 1.0, -1.0)             # +1, -1, +1, ...
    b = np.where((t // 2) % 2 == 0, 1.0, -1.0)      # +1, +1, -1, -1, ...   sum(a*b) == 0 exactly
    nodes = np.zeros((T, 4, 3))
    nodes[:, 0, 0] = 5.0 + 0.5 * a
    nodes[:, 1, 0] = 9.0 + 0.5 * a                   # same displacement, other place
    nodes[:, 2, 0] = -3.0 - 0.5 * a                  # mirrored
    nodes[:, 3, 1] = 2.0 + 0.25 * b                  # orthogonal direction AND uncorrelated sign
    want = np.array([[1.0, 1.0, -1.0, 0.0],
                     [1.0, 1.0, -1.0, 0.0],
                     [-1.0, -1.0, 1.0, 0.0],
                     [0.0, 0.0, 0.0, 1.0]])
    return nodes, want


def rigid_translation_trajectory(T=40, N=6, seed=2):
    """Every node gets the same displacement each frame: all correlations are 1, all distances
    constant in time."""
    rng = np.random.default_rng(seed)
    base = rng.normal(scale=6.0, size=(N, 3))
    # displacements on a dyadic grid: node + shift is exact in float64, so the centred
    # trajectories of all nodes are bit-identical
    shift = rng.integers(-64, 65, size=(T, 1, 3)) / 64.0
    base = np.round(base * 64.0) / 64.0
    return base[None, :, :] + shift, base


def lattice_contact_case():
    """Static 3-4-5 geometry: distances are exactly representable, one of them equals the cutoff
    -> pins the strict '<' of trajectory_analysis_functions.py:222."""
    pos = np.array([[0.0, 0.0, 0.0], [3.0, 0.0, 0.0], [3.0, 4.0, 0.0], [0.0, 4.0, 12.0]])
    nodes = np.repeat(pos[None], 5, axis=0)
    d = np.sqrt(((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1))
    cutoff = 5.0                                       # |p0 - p2| == 5 exactly: NOT a contact
    binary = ((d < cutoff) & ~np.eye(4, dtype=bool)).astype(np.int64)
    return nodes, cutoff, binary, d


def ladder_graph():
    """6-node weighted graph whose s -> t paths can be listed by hand (dyadic weights: every
    left-to-right float64 sum is exact).  Node 0 sits on the cheapest route so that the
    reference's node-0 rule changes the answer for source 1."""
    W = np.zeros((6, 6))

    def edge(u, v, w):
        W[u, v] = W[v, u] = w
    edge(1, 0, 0.125)
    edge(0, 5, 0.125)      # 1-0-5     = 0.25   (excluded from enumeration: second node is 0, s != 0)
    edge(1, 2, 0.25)
    edge(2, 5, 0.25)       # 1-2-5     = 0.5
    edge(1, 3, 0.25)
    edge(3, 4, 0.25)
    edge(4, 5, 0.125)      # 1-3-4-5   = 0.625
    edge(2, 3, 0.5)        # 1-2-3-4-5 = 1.125, 1-3-2-5 = 1.0
    return W


LADDER_PATHS_1_TO_5 = [            # every simple path 1 -> 5, ascending (length, nodes)
    [0.25, 1, 0, 5],
    [0.5, 1, 2, 5],
    [0.625, 1, 3, 4, 5],
    [1.0, 1, 3, 2, 5],
    [1.125, 1, 2, 3, 4, 5],
]
