"""ctypes binding of libwisp_b200.so (the C ABI declared in include/wisp_b200.h).

There is NO CPU fallback: importing this module without the built library raises, and
creating a context without a B200 raises.  numpy arrays are passed as plain pointers; the
only state is an opaque ``wisp_ctx*`` per device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwisp_b200.so")

# (name, restype, argtypes) -- must list every symbol of include/wisp_b200.h
_c_ctx = C.c_void_p
_P = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_dbl = C.c_double
SIGNATURES = {
    "wisp_ctx_create": (_int, [_int, C.POINTER(_c_ctx)]),
    "wisp_ctx_create_multi": (_int, [_int, _P, C.POINTER(_c_ctx)]),
    "wisp_ctx_gpus": (_int, [_c_ctx]),
    "wisp_ctx_ingest_info": (_int, [_c_ctx, _P, _P]),
    "wisp_host_register": (_int, [_P, C.c_size_t]),
    "wisp_host_unregister": (_int, [_P]),
    "wisp_ctx_destroy": (None, [_c_ctx]),
    "wisp_last_error": (C.c_char_p, []),
    "wisp_abi_version": (_int, []),
    "wisp_sync": (_int, [_c_ctx]),
    "wisp_node_ld": (_i64, [_int]),
    "wisp_frames_pad": (_i64, [_i64]),
    "wisp_gram_ld": (_i64, [_int]),
    "wisp_kernel_launches": (_i64, [_c_ctx]),
    "wisp_set_gram_mode": (_int, [_c_ctx, _int]),
    "wisp_com_nodes": (_int, [_c_ctx, _P, _i64, _int, _P, _P, _P, _int, _P, _int, _int, _P, _P]),
    "wisp_traj_alignment_and_averaging": (_int, [_c_ctx, _P, _i64, _int, _P, _P, _P, _int, _P, _int,
                                                 _int, _dbl, _int, _P, _P, _P, C.POINTER(_int)]),
    "wisp_cartesian_covariance": (_int, [_c_ctx, _P, _i64, _int, _P, _P]),
    "wisp_pearson_from_covariance": (_int, [_c_ctx, _P, _int, _P, _P, _P]),
    "wisp_contact_map": (_int, [_c_ctx, _P, _i64, _int, _dbl, _P, _P]),
    "wisp_func_pearson": (_int, [_c_ctx, _P, _int, _P, _P]),
    "wisp_lmi_from_covariance": (_int, [_c_ctx, _P, _int, _P]),
    "wisp_func_generalized_correlation": (_int, [_c_ctx, _P, _int, _P, _dbl, _int, _P, _P]),
    "wisp_pipeline": (_int, [_c_ctx, _P, _i64, _int, _P, _P, _P, _int, _P, _int, _int, _dbl, _int,
                             _dbl, _int, _P, _P, _P, _P, _P, _P]),
    "wisp_pipeline_log": (_int, [_c_ctx, _P, _i64, _int, _P, _P, _P, _int, _P, _int, _int, _dbl, _int,
                                 _dbl, _int, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_int)]),
    "wisp_pipeline_stats": (_int, [_c_ctx, _P]),
    "wisp_apsp": (_int, [_c_ctx, _P, _int, _int, _P, _P]),
    "wisp_get_paths": (_int, [_c_ctx, _P, _int, _P, _int, _P, _int, _int, _int]),
    "wisp_paths_prepare": (_int, [_c_ctx, _P, _int, _int]),
    "wisp_paths_query": (_int, [_c_ctx, _P, _int, _P, _int, _int, _int]),
    "wisp_paths_query_pairs": (_int, [_c_ctx, _P, _int, _int, _int]),
    "wisp_paths_query_offsets": (_int, [_c_ctx, C.POINTER(_i64), _P]),
    "wisp_paths_size": (_int, [_c_ctx, C.POINTER(_i64), C.POINTER(_i64)]),
    "wisp_paths_fetch": (_int, [_c_ctx, _P, _P, _P]),
    "wisp_paths_stats": (_int, [_c_ctx, _P]),
    "wisp_savetxt_f64": (_int, [C.c_char_p, _P, _i64, _i64, C.c_char_p]),
    "wisp_savetxt_i64": (_int, [C.c_char_p, _P, _i64, _i64, C.c_char_p]),
    "wisp_loadtxt_f64": (_int, [C.c_char_p, _P, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "wisp_com_plan_create": (_int, [_c_ctx, _int, _P, _P, _P, _int, _P, _int, _int, C.POINTER(_P)]),
    "wisp_com_plan_destroy": (None, [_c_ctx, _P]),
    "wisp_com_plan_info": (_int, [_P, C.POINTER(_int), C.POINTER(_int)]),
    "wisp_dev_com": (_int, [_c_ctx, _P, _P, _i64, _int, _P, _P, _P]),
    "wisp_dev_com_align": (_int, [_c_ctx, _P, _P, _i64, _int, _P, _P, _P, _P]),
    "wisp_dev_align_sums": (_int, [_c_ctx, _P, _int, _P, _i64, _int, _int, _P, _P]),
    "wisp_dev_align_rotate": (_int, [_c_ctx, _P, _int, _P, _P, _i64, _int, _P, _P]),
    "wisp_dev_align_iter": (_int, [_c_ctx, _P, _int, _P, _P, _i64, _int, _P, _int, _P, _P, _P]),
    "wisp_dev_align_update_gated": (_int, [_c_ctx, _P, _i64, _int, _int, _P, _P, _dbl, _P, _P]),
    "wisp_dev_align_nodes": (_int, [_c_ctx, _P, _i64, _int, _P, _P]),
    "wisp_dev_center_rotated": (_int, [_c_ctx, _P, _i64, _int, _P, _P, _P, _P]),
    "wisp_dev_align_update": (_int, [_c_ctx, _P, _i64, _int, _int, _P, _P, _P]),
    "wisp_dev_center": (_int, [_c_ctx, _P, _i64, _int, _P, _P, _P]),
    "wisp_dev_colsum": (_int, [_c_ctx, _P, _i64, _int, _P, _P]),
    "wisp_dev_gram": (_int, [_c_ctx, _P, _i64, _i64, _int, _int, _P, _P]),
    "wisp_dev_gram_i8": (_int, [_c_ctx, _P, _i64, _i64, _int, _P, _P]),
    "wisp_dev_contact_sum": (_int, [_c_ctx, _P, _i64, _int, _P, _P, _P]),
    "wisp_dev_epilogue": (_int, [_c_ctx, _P, _P, _i64, _int, _dbl, _int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "wisp_dev_reduce_partials": (_int, [_c_ctx, _P, _int, _int, _P, _P]),
    "wisp_dev_apsp": (_int, [_c_ctx, _P, _int, _int, _P, _P, _P]),
    "wisp_dev_fw32_init_rows": (_int, [_c_ctx, _P, _int, _i64, _i64, _P, _P]),
    "wisp_dev_fw32_pivot_panel": (_int, [_c_ctx, _P, _int, _int, _P]),
    "wisp_dev_fw32_update_rows": (_int, [_c_ctx, _P, _i64, _P, _int, _int, _int, _P]),
    "wisp_dev_fw32_update_rows_part": (_int, [_c_ctx, _P, _i64, _P, _int, _int, _int, _int, _int, _P]),
    "wisp_dev_synth_coords": (_int, [_c_ctx, _P, _i64, _i64, _int, _P, _P, _P, _int, _P, _dbl,
                                     C.c_uint64, _P]),
}


class WispError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise WispError(
            "libwisp_b200.so not found at %s -- build it with "
            "`python -m wisp_mdanalysis_implementation_b200.build` (there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


def _arr(a, dtype, name, shape=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("%s: expected shape %s, got %s" % (name, tuple(shape), a.shape))
    return a


class Context:
    """One ``wisp_ctx`` (streams + scratch memory) on one B200 -- or, with ``gpus``, on several GPUs
    of the box: ``gpus`` = a count (devices 0..gpus-1; 0 = every visible GPU) or an explicit device
    list.  ``pipeline`` then shards the frames over them inside the one C call; every other method
    runs on the first device."""

    def __init__(self, device=0, gpus=None):
        self._lib = lib()
        self._ctx = _c_ctx()
        if gpus is None:
            rc = self._lib.wisp_ctx_create(int(device), C.byref(self._ctx))
            self.device = int(device)
        else:
            if isinstance(gpus, int):
                n, devs = int(gpus), None
            else:
                devs = np.ascontiguousarray(list(gpus), dtype=np.int32)
                n = len(devs)
            rc = self._lib.wisp_ctx_create_multi(n, _ptr(devs), C.byref(self._ctx))
            self.device = int(devs[0]) if devs is not None else 0
        if rc != 0:
            raise WispError(self._lib.wisp_last_error().decode())
        self.n_gpus = int(self._lib.wisp_ctx_gpus(self._ctx))

    def ingest_info(self):
        """How the coordinates reach the GPUs (see wisp_ctx_ingest_info)."""
        n = self.n_gpus
        via = np.zeros(n, dtype=np.int32)
        rates = np.zeros(2 * n + 2)
        self._check(self._lib.wisp_ctx_ingest_info(self._ctx, _ptr(via), _ptr(rates)))
        return dict(via=via.tolist(), h2d_gbs_all_pulling=rates[:n].round(1).tolist(),
                    h2d_gbs_fast_set_pulling=rates[n:2 * n].round(1).tolist(),
                    aggregate_all=round(float(rates[2 * n]), 1), aggregate_fast_set=round(float(rates[2 * n + 1]), 1))

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.wisp_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise WispError(self._lib.wisp_last_error().decode())

    @property
    def handle(self):
        return self._ctx

    def kernel_launches(self):
        return int(self._lib.wisp_kernel_launches(self._ctx))

    def set_gram_mode(self, mode):
        """K2 engine: 'auto' | 'dmma' | 'i8' (see include/wisp_b200.h)."""
        self._check(self._lib.wisp_set_gram_mode(self._ctx, {"auto": 0, "dmma": 1, "i8": 2}[mode]))

    def sync(self):
        self._check(self._lib.wisp_sync(self._ctx))

    # ---- host-buffer API ---------------------------------------------------------------
    def com_nodes(self, coords, masses, node_ptr, atom_idx, align_idx, atomic=False,
                  want_nodes=True):
        coords = _arr(coords, np.float32, "coords")
        T, A, _ = coords.shape
        masses = _arr(masses, np.float64, "masses", (A,))
        node_ptr = _arr(node_ptr, np.int32, "node_ptr")
        atom_idx = _arr(atom_idx, np.int32, "atom_idx")
        align_idx = _arr(align_idx, np.int32, "align_idx")
        N = len(node_ptr) - 1
        nodes = np.empty((T, N, 3), dtype=np.float64) if want_nodes else None
        avg = np.empty((N, 3), dtype=np.float64)
        self._check(self._lib.wisp_com_nodes(
            self._ctx, _ptr(coords), T, A, _ptr(masses), _ptr(node_ptr), _ptr(atom_idx), N,
            _ptr(align_idx), len(align_idx), int(bool(atomic)), _ptr(nodes), _ptr(avg)))
        return nodes, avg

    def traj_alignment_and_averaging(self, coords, masses, node_ptr, atom_idx, align_idx,
                                     atomic=False, convergence_threshold=1e-5,
                                     maximum_num_iterations=100):
        coords = _arr(coords, np.float32, "coords")
        T, A, _ = coords.shape
        masses = _arr(masses, np.float64, "masses", (A,))
        node_ptr = _arr(node_ptr, np.int32, "node_ptr")
        atom_idx = _arr(atom_idx, np.int32, "atom_idx")
        align_idx = _arr(align_idx, np.int32, "align_idx")
        N = len(node_ptr) - 1
        nodes = np.empty((T, N, 3), dtype=np.float64)
        avg = np.empty((N, 3), dtype=np.float64)
        log = np.zeros((max(1, int(maximum_num_iterations)), 3), dtype=np.float64)
        n_iter = _int(0)
        self._check(self._lib.wisp_traj_alignment_and_averaging(
            self._ctx, _ptr(coords), T, A, _ptr(masses), _ptr(node_ptr), _ptr(atom_idx), N,
            _ptr(align_idx), len(align_idx), int(bool(atomic)), float(convergence_threshold),
            int(maximum_num_iterations), _ptr(nodes), _ptr(avg), _ptr(log), C.byref(n_iter)))
        return nodes, avg, [(int(r[0]), float(r[1]), float(r[2])) for r in log[:n_iter.value]]

    def cartesian_covariance(self, nodes, avg):
        nodes = _arr(nodes, np.float64, "node_trajectory")
        T, N, _ = nodes.shape
        avg = _arr(avg, np.float64, "avg_node_positions", (N, 3))
        cov = np.empty((3 * N, 3 * N), dtype=np.float64)
        self._check(self._lib.wisp_cartesian_covariance(self._ctx, _ptr(nodes), T, N, _ptr(avg),
                                                        _ptr(cov)))
        return cov

    def pearson_from_covariance(self, cov):
        cov = _arr(cov, np.float64, "cartesian_covariance_matrix")
        n3 = cov.shape[0]
        if cov.ndim != 2 or cov.shape[1] != n3 or n3 % 3:
            raise ValueError("covariance must be (3N, 3N), got %s" % (cov.shape,))
        N = n3 // 3
        var = np.empty(N)
        ncov = np.empty((N, N))
        corr = np.empty((N, N))
        self._check(self._lib.wisp_pearson_from_covariance(self._ctx, _ptr(cov), N, _ptr(var),
                                                           _ptr(ncov), _ptr(corr)))
        return var, ncov, corr

    def contact_map(self, nodes, cutoff):
        nodes = _arr(nodes, np.float64, "trajectory_data")
        T, N, _ = nodes.shape
        binary = np.empty((N, N), dtype=np.int64)
        avgdist = np.empty((N, N), dtype=np.float64)
        self._check(self._lib.wisp_contact_map(self._ctx, _ptr(nodes), T, N, float(cutoff),
                                               _ptr(binary), _ptr(avgdist)))
        return binary, avgdist

    def func_pearson(self, corr, contact_map=None):
        corr = _arr(corr, np.float64, "adjacency_matrix")
        N = corr.shape[0]
        cm = None if contact_map is None else _arr(contact_map, np.float64, "contact_map", (N, N))
        w = np.empty((N, N), dtype=np.float64)
        self._check(self._lib.wisp_func_pearson(self._ctx, _ptr(corr), N, _ptr(cm), _ptr(w)))
        return w

    def lmi_from_covariance(self, cov):
        cov = _arr(cov, np.float64, "cartesian_covariance_matrix")
        n3 = cov.shape[0]
        if cov.ndim != 2 or cov.shape[1] != n3 or n3 % 3:
            raise ValueError("covariance must be (3N, 3N), got %s" % (cov.shape,))
        N = n3 // 3
        mi = np.empty((N, N))
        self._check(self._lib.wisp_lmi_from_covariance(self._ctx, _ptr(cov), N, _ptr(mi)))
        return mi

    def func_generalized_correlation(self, mi, contact_map=None, Lambda=None):
        mi = _arr(mi, np.float64, "adjacency_matrix")
        N = mi.shape[0]
        cm = None if contact_map is None else _arr(contact_map, np.float64, "contact_map", (N, N))
        rmi = np.empty((N, N))
        func = np.empty((N, N))
        self._check(self._lib.wisp_func_generalized_correlation(
            self._ctx, _ptr(mi), N, _ptr(cm), float(Lambda) if Lambda is not None else 0.0,
            int(Lambda is not None), _ptr(rmi), _ptr(func)))
        return rmi, func

    def pipeline(self, coords, masses, node_ptr, atom_idx, align_idx, cutoff=9999.0,
                 which_map=0, atomic=False, want_nodes=False, convergence_threshold=1e-5,
                 maximum_num_iterations=0, out=None):
        """coords -> dict(avg, corr, avgdist, binary, w[, nodes], alignment_log) in one fused pass
        (K1, optional K7 alignment, K2, K3, K4) on every GPU of this context.  ``out`` may supply
        preallocated (e.g. page-locked) result arrays under the same keys."""
        coords = _arr(coords, np.float32, "coords")
        T, A, _ = coords.shape
        masses = _arr(masses, np.float64, "masses", (A,))
        node_ptr = _arr(node_ptr, np.int32, "node_ptr")
        atom_idx = _arr(atom_idx, np.int32, "atom_idx")
        align_idx = _arr(align_idx, np.int32, "align_idx")
        N = len(node_ptr) - 1
        given = out or {}
        out = {}
        for key, shape, dt in (("avg", (N, 3), np.float64), ("corr", (N, N), np.float64),
                               ("avgdist", (N, N), np.float64), ("binary", (N, N), np.int64),
                               ("w", (N, N), np.float64)):
            buf = given.get(key)
            if buf is None:
                buf = np.empty(shape, dtype=dt)
            elif buf.shape != shape or buf.dtype != dt or not buf.flags.c_contiguous:
                raise ValueError("pipeline: out[%r] must be a C-contiguous %s array of shape %s" % (key, dt.__name__, shape))
            out[key] = buf
        nodes = np.empty((T, N, 3)) if want_nodes else None
        log = np.zeros((max(1, int(maximum_num_iterations)), 3), dtype=np.float64)
        n_iter = _int(0)
        self._check(self._lib.wisp_pipeline_log(
            self._ctx, _ptr(coords), T, A, _ptr(masses), _ptr(node_ptr), _ptr(atom_idx), N,
            _ptr(align_idx), len(align_idx), int(bool(atomic)), float(convergence_threshold),
            int(maximum_num_iterations), float(cutoff), int(which_map),
            _ptr(out["avg"]), _ptr(out["corr"]), _ptr(out["avgdist"]), _ptr(out["binary"]),
            _ptr(out["w"]), _ptr(nodes), _ptr(log), C.byref(n_iter)))
        if want_nodes:
            out["nodes"] = nodes
        out["alignment_log"] = [(int(r[0]), float(r[1]), float(r[2])) for r in log[:n_iter.value]]
        return out

    def pipeline_stats(self):
        s = np.empty(8)
        self._check(self._lib.wisp_pipeline_stats(self._ctx, _ptr(s)))
        return dict(h2d_k1_k3_ms=float(s[0]), k7_ms=float(s[1]), center_k2_ms=float(s[2]),
                    wait_for_slowest_gpu_ms=float(s[3]), epilogue_d2h_ms=float(s[4]), total_ms=float(s[5]),
                    gpus=int(s[7]))

    def apsp(self, w, precision=64, want_pred=True):
        w = _arr(w, np.float64, "cost matrix")
        N = w.shape[0]
        dist = np.empty((N, N), dtype=np.float64)
        pred = np.empty((N, N), dtype=np.int32) if want_pred else None
        self._check(self._lib.wisp_apsp(self._ctx, _ptr(w), N, int(precision), _ptr(dist), _ptr(pred)))
        return dist, pred

    def get_paths(self, w, sources, sinks, number_of_paths, node0_quirk=True, apsp_fp32=False):
        w = _arr(w, np.float64, "adjacency_matrix")
        N = w.shape[0]
        so = _arr(sources, np.int32, "sources")
        si = _arr(sinks, np.int32, "sinks")
        flags = (1 if node0_quirk else 0) | (2 if apsp_fp32 else 0)
        self._check(self._lib.wisp_get_paths(self._ctx, _ptr(w), N, _ptr(so), len(so), _ptr(si),
                                             len(si), int(number_of_paths), flags))
        return self._fetch_paths()

    def paths_prepare(self, w, apsp_fp32=False):
        """Upload a cost matrix and run the APSP once; follow with any number of paths_query."""
        w = _arr(w, np.float64, "adjacency_matrix")
        self._check(self._lib.wisp_paths_prepare(self._ctx, _ptr(w), w.shape[0], 2 if apsp_fp32 else 0))

    def paths_query(self, sources, sinks, number_of_paths, node0_quirk=True):
        so = _arr(sources, np.int32, "sources")
        si = _arr(sinks, np.int32, "sinks")
        self._check(self._lib.wisp_paths_query(self._ctx, _ptr(so), len(so), _ptr(si), len(si),
                                               int(number_of_paths), 1 if node0_quirk else 0))
        return self._fetch_paths()

    def paths_query_pairs(self, pairs, number_of_paths, node0_quirk=True, packed=False):
        """``[get_paths(w, [s], [t], K) for (s, t) in pairs]`` on the prepared graph, all queries served
        by the same kernel passes.  packed=True returns the raw arrays (query_offsets, lengths,
        node_offsets, nodes) instead of Python lists."""
        pr = _arr(pairs, np.int32, "pairs").reshape(-1, 2)
        self._check(self._lib.wisp_paths_query_pairs(self._ctx, _ptr(pr), len(pr), int(number_of_paths),
                                                     1 if node0_quirk else 0))
        n_paths, n_nodes, nq = _i64(), _i64(), _i64()
        self._check(self._lib.wisp_paths_size(self._ctx, C.byref(n_paths), C.byref(n_nodes)))
        self._check(self._lib.wisp_paths_query_offsets(self._ctx, C.byref(nq), None))
        qoff = np.empty(nq.value + 1, dtype=np.int64)
        self._check(self._lib.wisp_paths_query_offsets(self._ctx, C.byref(nq), _ptr(qoff)))
        lengths = np.empty(n_paths.value, dtype=np.float64)
        offsets = np.empty(n_paths.value + 1, dtype=np.int64)
        nodes = np.empty(n_nodes.value, dtype=np.int32)
        self._check(self._lib.wisp_paths_fetch(self._ctx, _ptr(lengths), _ptr(offsets), _ptr(nodes)))
        if packed:
            return qoff, lengths, offsets, nodes
        return unpack_paths(qoff, lengths, offsets, nodes)

    def _fetch_paths(self):
        n_paths, n_nodes = _i64(), _i64()
        self._check(self._lib.wisp_paths_size(self._ctx, C.byref(n_paths), C.byref(n_nodes)))
        lengths = np.empty(n_paths.value, dtype=np.float64)
        offsets = np.empty(n_paths.value + 1, dtype=np.int64)
        nodes = np.empty(n_nodes.value, dtype=np.int32)
        self._check(self._lib.wisp_paths_fetch(self._ctx, _ptr(lengths), _ptr(offsets), _ptr(nodes)))
        return [[float(lengths[k])] + [int(x) for x in nodes[offsets[k]:offsets[k + 1]]]
                for k in range(n_paths.value)]

    def paths_stats(self):
        s = np.empty(6)
        self._check(self._lib.wisp_paths_stats(self._ctx, _ptr(s)))
        return dict(passes=int(s[0]), expansions=int(s[1]), final_cutoff=float(s[2]),
                    apsp_ms=float(s[3]), enumerate_ms=float(s[4]), host_ms=float(s[5]))

    # ---- device-buffer API (pointers are ints, e.g. torch.Tensor.data_ptr()) -----------
    def com_plan(self, n_atoms, masses, node_ptr, atom_idx, align_idx, atomic=False):
        """Upload the node map once; returns an opaque plan handle for dev_com."""
        masses = _arr(masses, np.float64, "masses", (n_atoms,))
        node_ptr = _arr(node_ptr, np.int32, "node_ptr")
        atom_idx = _arr(atom_idx, np.int32, "atom_idx")
        align_idx = _arr(align_idx, np.int32, "align_idx")
        plan = _P()
        self._check(self._lib.wisp_com_plan_create(
            self._ctx, int(n_atoms), _ptr(masses), _ptr(node_ptr), _ptr(atom_idx), len(node_ptr) - 1,
            _ptr(align_idx), len(align_idx), int(bool(atomic)), C.byref(plan)))
        return plan

    def com_plan_info(self, plan):
        s, p = _int(0), _int(0)
        self._check(self._lib.wisp_com_plan_info(plan, C.byref(s), C.byref(p)))
        return dict(streaming=bool(s.value), pieces=p.value)

    def com_plan_destroy(self, plan):
        self._lib.wisp_com_plan_destroy(self._ctx, plan)

    def dev_com(self, plan, d_coords, T, N, d_nodes, d_colsum=None, stream=None):
        self._check(self._lib.wisp_dev_com(self._ctx, plan, _ptr(d_coords), T, N, _ptr(d_nodes),
                                           _ptr(d_colsum), _ptr(stream)))

    def dev_com_align(self, plan, d_coords, T, N, d_nodes, d_colsum=None, d_align_out=None, stream=None):
        self._check(self._lib.wisp_dev_com_align(self._ctx, plan, _ptr(d_coords), T, N, _ptr(d_nodes),
                                                 _ptr(d_colsum), _ptr(d_align_out), _ptr(stream)))

    def dev_align_sums(self, d_align, n_align, d_nodes, T, N, first, d_sums, stream=None):
        self._check(self._lib.wisp_dev_align_sums(self._ctx, _ptr(d_align), n_align, _ptr(d_nodes), T, N,
                                                  int(bool(first)), _ptr(d_sums), _ptr(stream)))

    def dev_align_rotate(self, d_align, n_align, d_avg_align, d_nodes, T, N, d_sums=None, stream=None):
        self._check(self._lib.wisp_dev_align_rotate(self._ctx, _ptr(d_align), n_align, _ptr(d_avg_align),
                                                    _ptr(d_nodes), T, N, _ptr(d_sums), _ptr(stream)))

    def dev_align_iter(self, d_align, n_align, d_avg_align, d_nodes, T, N, d_rot_total, first, d_sums=None,
                       d_gate=None, stream=None):
        """One alignment iteration with the node rotation deferred (nodes are only read); a no-op while
        the device flag d_gate is set."""
        self._check(self._lib.wisp_dev_align_iter(self._ctx, _ptr(d_align), n_align, _ptr(d_avg_align),
                                                  _ptr(d_nodes), T, N, _ptr(d_rot_total), int(bool(first)),
                                                  _ptr(d_sums), _ptr(d_gate), _ptr(stream)))

    def dev_align_update_gated(self, d_sums, T_total, n_align, N, d_avg, d_res3, threshold, d_gate, stream=None):
        self._check(self._lib.wisp_dev_align_update_gated(self._ctx, _ptr(d_sums), T_total, n_align, N, _ptr(d_avg),
                                                          _ptr(d_res3), float(threshold), _ptr(d_gate), _ptr(stream)))

    def dev_align_nodes(self, d_nodes, T, N, d_rot_total, stream=None):
        self._check(self._lib.wisp_dev_align_nodes(self._ctx, _ptr(d_nodes), T, N, _ptr(d_rot_total),
                                                   _ptr(stream)))

    def dev_center_rotated(self, d_nodes, T, N, d_mean, d_rot_total, d_p32=None, stream=None):
        self._check(self._lib.wisp_dev_center_rotated(self._ctx, _ptr(d_nodes), T, N, _ptr(d_mean),
                                                      _ptr(d_rot_total), _ptr(d_p32), _ptr(stream)))

    def dev_align_update(self, d_sums, T_total, n_align, N, d_avg, d_res2, stream=None):
        self._check(self._lib.wisp_dev_align_update(self._ctx, _ptr(d_sums), T_total, n_align, N, _ptr(d_avg),
                                                    _ptr(d_res2), _ptr(stream)))

    def dev_center(self, d_nodes, T, N, d_mean, d_p32=None, stream=None):
        self._check(self._lib.wisp_dev_center(self._ctx, _ptr(d_nodes), T, N, _ptr(d_mean),
                                              _ptr(d_p32), _ptr(stream)))

    def dev_colsum(self, d_nodes, T, N, d_colsum, stream=None):
        self._check(self._lib.wisp_dev_colsum(self._ctx, _ptr(d_nodes), T, N, _ptr(d_colsum), _ptr(stream)))

    def dev_gram(self, d_x, T, ld, n_out, stride, d_out, stream=None):
        self._check(self._lib.wisp_dev_gram(self._ctx, _ptr(d_x), T, ld, n_out, stride, _ptr(d_out),
                                            _ptr(stream)))

    def dev_gram_i8(self, d_x, T, ld, N, d_out, stream=None):
        self._check(self._lib.wisp_dev_gram_i8(self._ctx, _ptr(d_x), T, ld, N, _ptr(d_out), _ptr(stream)))

    def dev_contact_sum(self, d_p32, T, N, d_mean, d_out, stream=None):
        self._check(self._lib.wisp_dev_contact_sum(self._ctx, _ptr(d_p32), T, N, _ptr(d_mean),
                                                   _ptr(d_out), _ptr(stream)))

    def dev_epilogue(self, d_gram, d_dsum, T_total, N, cutoff, which_map, d_var=None, d_ncov=None,
                     d_corr=None, d_avgdist=None, d_binary=None, d_w=None, d_offset_sum=None,
                     stream=None):
        self._check(self._lib.wisp_dev_epilogue(self._ctx, _ptr(d_gram), _ptr(d_dsum), T_total, N,
                                                float(cutoff), int(which_map), _ptr(d_var),
                                                _ptr(d_ncov), _ptr(d_corr), _ptr(d_avgdist),
                                                _ptr(d_binary), _ptr(d_w), _ptr(d_offset_sum),
                                                _ptr(stream)))

    def dev_reduce_partials(self, d_parts, n_parts, N, d_out, stream=None):
        self._check(self._lib.wisp_dev_reduce_partials(self._ctx, _ptr(d_parts), n_parts, N,
                                                       _ptr(d_out), _ptr(stream)))

    def dev_apsp(self, d_w, N, precision, d_dist, d_pred, stream=None):
        self._check(self._lib.wisp_dev_apsp(self._ctx, _ptr(d_w), N, int(precision), _ptr(d_dist),
                                            _ptr(d_pred), _ptr(stream)))

    def dev_fw32_init_rows(self, d_w_rows, N, row0, rows_pad, d_dist_local, stream=None):
        self._check(self._lib.wisp_dev_fw32_init_rows(self._ctx, _ptr(d_w_rows), N, row0, rows_pad,
                                                      _ptr(d_dist_local), _ptr(stream)))

    def dev_fw32_pivot_panel(self, d_panel, N, kb, stream=None):
        self._check(self._lib.wisp_dev_fw32_pivot_panel(self._ctx, _ptr(d_panel), N, kb, _ptr(stream)))

    def dev_fw32_update_rows(self, d_dist_local, rows_pad, d_panel, N, kb, skip_local_tile, stream=None):
        self._check(self._lib.wisp_dev_fw32_update_rows(self._ctx, _ptr(d_dist_local), rows_pad,
                                                        _ptr(d_panel), N, kb, skip_local_tile, _ptr(stream)))

    def dev_fw32_update_rows_part(self, d_dist_local, rows_pad, d_panel, N, kb, skip_local_tile,
                                  next_local_tile, part, stream=None):
        self._check(self._lib.wisp_dev_fw32_update_rows_part(self._ctx, _ptr(d_dist_local), rows_pad,
                                                             _ptr(d_panel), N, kb, skip_local_tile,
                                                             next_local_tile, part, _ptr(stream)))

    def dev_synth_coords(self, d_coords, t0, T, A, d_atom_base, d_node_of_atom, d_modes, n_modes,
                         d_coeff, noise_sigma, seed, stream=None):
        self._check(self._lib.wisp_dev_synth_coords(self._ctx, _ptr(d_coords), t0, T, A,
                                                    _ptr(d_atom_base), _ptr(d_node_of_atom),
                                                    _ptr(d_modes), n_modes, _ptr(d_coeff),
                                                    float(noise_sigma), int(seed), _ptr(stream)))


def unpack_paths(query_offsets, lengths, node_offsets, nodes):
    """Packed path arrays -> list (per query) of lists of ``[length, n0, n1, ...]``."""
    nl = nodes.tolist()
    ll = lengths.tolist()
    no = node_offsets.tolist()
    return [[[ll[k]] + nl[no[k]:no[k + 1]] for k in range(int(query_offsets[q]), int(query_offsets[q + 1]))]
            for q in range(len(query_offsets) - 1)]


def savetxt(path, array, fmt='%.18e', header=''):
    """np.savetxt-compatible writer for the two formats the reference uses."""
    a = np.asarray(array)
    shape = a.shape if a.ndim == 2 else (a.size, 1)
    hdr = header.encode() if header else None
    if fmt == '%.18e':
        a = np.ascontiguousarray(a, dtype=np.float64)
        rc = lib().wisp_savetxt_f64(os.fsencode(path), _ptr(a), shape[0], shape[1], hdr)
    elif fmt in ('%i', '%d'):
        a = np.ascontiguousarray(a, dtype=np.int64)
        rc = lib().wisp_savetxt_i64(os.fsencode(path), _ptr(a), shape[0], shape[1], hdr)
    else:
        raise ValueError("savetxt: unsupported format %r" % fmt)
    if rc != 0:
        raise WispError(lib().wisp_last_error().decode())


def loadtxt(path):
    rows, cols = _i64(), _i64()
    if lib().wisp_loadtxt_f64(os.fsencode(path), None, 0, C.byref(rows), C.byref(cols)) != 0:
        raise WispError(lib().wisp_last_error().decode())
    out = np.empty((rows.value, cols.value), dtype=np.float64)
    if lib().wisp_loadtxt_f64(os.fsencode(path), _ptr(out), out.size, C.byref(rows), C.byref(cols)) != 0:
        raise WispError(lib().wisp_last_error().decode())
    if cols.value == 1:
        return out[:, 0].copy()
    return out[0].copy() if rows.value == 1 else out      # np.loadtxt squeezes a single row too


def node_ld(n_nodes):
    return int(lib().wisp_node_ld(int(n_nodes)))


def frames_pad(n_frames):
    return int(lib().wisp_frames_pad(int(n_frames)))


def gram_ld(n_cols):
    return int(lib().wisp_gram_ld(int(n_cols)))


def host_register(array):
    """Page-lock a numpy array in place (cudaHostRegister, portable) so that the pipeline's H2D
    copies run asynchronously at full PCIe rate from every GPU; pair with host_unregister."""
    if lib().wisp_host_register(_ptr(array), array.nbytes) != 0:
        raise WispError(lib().wisp_last_error().decode())


def host_unregister(array):
    if lib().wisp_host_unregister(_ptr(array)) != 0:
        raise WispError(lib().wisp_last_error().decode())


_default_ctx = {}


def default_context(device=None, gpus=None):
    """Process-wide context: one per device (device defaults to $WISP_DEVICE, $LOCAL_RANK or 0), or
    ONE over several GPUs when ``gpus`` (or $WISP_GPUS; 0 = every visible GPU) is given -- what the
    config key ``gpus`` of wisp_w_mdanalysis.py selects."""
    if gpus is None and device is None and os.environ.get("WISP_GPUS") is not None:
        gpus = int(os.environ["WISP_GPUS"])
    if gpus is not None and gpus != 1:
        key = ("multi", gpus if isinstance(gpus, int) else tuple(gpus))
        if key not in _default_ctx:
            _default_ctx[key] = Context(gpus=gpus)
        return _default_ctx[key]
    if device is None:
        device = int(os.environ.get("WISP_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
