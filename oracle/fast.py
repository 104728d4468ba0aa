"""C-accelerated execution of the oracle FMM passes (test infrastructure only).

Takes an ``oracle.bbfmm.FmmTree`` (tree, lists and operators built by the numpy restatement),
flattens it to CSR arrays and runs the same five passes with the hot loops in
``oracle/csrc/oracle_passes.c`` (OpenMP over cells / leaves, like the reference's rayon loops,
bbfmm.rs:669-1159).  Used for larger parity cases and as the CPU baseline of bench.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import chebyshev, morton
from .kernels import SPHEROIDAL_CONSTANTS

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "csrc", "liboracle_passes.so")
_lib = None


class _OrcKernel(C.Structure):
    _fields_ = [("kernel_type", C.c_int), ("pw", C.c_int), ("s2", C.c_double), ("ip2", C.c_double),
                ("near_slope", C.c_double), ("far_coef", C.c_double), ("total_sill", C.c_double)]


def build():
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _ck(kernel):
    k = _OrcKernel()
    k.kernel_type = kernel.kernel_type
    k.total_sill = kernel.total_sill
    if kernel.kernel_type in SPHEROIDAL_CONSTANTS:
        k.pw, k.s2, k.ip2, k.near_slope, k.far_coef = kernel.pow, kernel.s2, kernel.ip2, kernel.near_slope, kernel.far_coef
    return k


class FastFmm:
    def __init__(self, tree):
        self.t = tree
        L = tree.lists
        self.dim, self.P, self.p = tree.dim, tree.P, tree.p
        self.keys = sorted(L.tree)
        self.index = {k: i for i, k in enumerate(self.keys)}
        nc = len(self.keys)
        self.nc = nc
        self.level = np.array([k & morton.LEVEL_MASK for k in self.keys], dtype=np.int64)
        cen = np.zeros((nc, self.dim))
        half = np.zeros(nc)
        for i, k in enumerate(self.keys):
            c, length = morton.get_center_length(k, tree.center, tree.radius, self.dim)
            cen[i] = c
            half[i] = 0.5 * length
        self.center, self.half = cen, half
        self.kernel = _ck(tree.kernel)
        self.nodes_nd = np.ascontiguousarray(tree.ops.nodes_nd, dtype=np.float64)
        self.src = np.ascontiguousarray(tree.source_points)
        # leaves with sources
        self.leaf_keys = [k for k in self.keys if k in L.leaf_source_indices]
        # per-leaf CSR (targets = sources of the leaf), U sources concatenated, W cells
        t_ptr, t_idx, u_ptr, u_idx, w_ptr, w_cell = [0], [], [0], [], [0], []
        for k in self.leaf_keys:
            t_idx.append(np.asarray(L.leaf_source_indices[k], dtype=np.int64))
            t_ptr.append(t_ptr[-1] + len(t_idx[-1]))
            us = [np.asarray(L.leaf_source_indices[u], dtype=np.int64) for u in sorted(L.u_lists.get(k, ()))
                  if u in L.leaf_source_indices]
            u_idx.extend(us)
            u_ptr.append(u_ptr[-1] + sum(len(a) for a in us))
            ws = sorted((L.w_lists or {}).get(k, ()))
            w_cell.extend(self.index[w] for w in ws)
            w_ptr.append(w_ptr[-1] + len(ws))
        cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)
        self.t_ptr, self.t_idx = np.array(t_ptr, dtype=np.int64), cat(t_idx)
        self.u_ptr, self.u_idx = np.array(u_ptr, dtype=np.int64), cat(u_idx)
        self.w_ptr, self.w_cell = np.array(w_ptr, dtype=np.int64), np.array(w_cell, dtype=np.int64)
        # X lists
        xc, x_ptr, x_idx = [], [0], []
        for k, xl in (L.x_lists or {}).items():
            xs = [np.asarray(L.leaf_source_indices[x], dtype=np.int64) for x in sorted(xl) if x in L.leaf_source_indices]
            if not xs:
                continue
            xc.append(self.index[k])
            x_idx.extend(xs)
            x_ptr.append(x_ptr[-1] + sum(len(a) for a in xs))
        self.x_cells, self.x_ptr, self.x_idx = np.array(xc, dtype=np.int64), np.array(x_ptr, dtype=np.int64), cat(x_idx)
        # M2L entries by target cell
        ops = tree.ops
        self.op_ids = {}
        self.op_u, self.op_vt, self.op_rank = [], [], []
        for lvl in sorted(ops.u):
            for r in sorted(ops.u[lvl]):
                self.op_ids[(lvl, r)] = len(self.op_u)
                u = np.asfortranarray(ops.u[lvl][r])
                self.op_u.append(u)
                vt = np.asfortranarray(ops.vt[lvl][r]) if r in ops.vt.get(lvl, {}) else None
                self.op_vt.append(vt)
                self.op_rank.append(u.shape[1])
        tgt, e_ptr, e_src, e_op, e_perm = [], [0], [], [], []
        for k in self.keys:
            vl = L.v_lists.get(k)
            if not vl:
                continue
            lvl = k & morton.LEVEL_MASK
            a, _ = morton.decode_key(k, self.dim)
            for v in sorted(vl):
                b, _ = morton.decode_key(v, self.dim)
                tix = 0
                for d in range(self.dim):
                    tix = tix * 7 + (a[d] - b[d] + 3)
                e_src.append(self.index[v])
                e_op.append(self.op_ids[(lvl, ops.reference_vector_lookups[tix])])
                e_perm.append(ops.permutation_lookups[tix])
            tgt.append(self.index[k])
            e_ptr.append(len(e_src))
        self.m2l_tgt = np.array(tgt, dtype=np.int64)
        self.m2l_ptr = np.array(e_ptr, dtype=np.int64)
        self.m2l_src = np.array(e_src, dtype=np.int64)
        self.m2l_op = np.array(e_op, dtype=np.int32)
        self.m2l_perm = np.array(e_perm, dtype=np.int32)
        self.perm_tab = np.ascontiguousarray(np.array(ops.permutation_indices, dtype=np.int32))
        self.inv_tab = np.ascontiguousarray(np.array(ops.inverse_permutations, dtype=np.int32))
        n_ops = len(self.op_u)
        self._u_ptrs = (C.POINTER(C.c_double) * n_ops)(*[_p(u, C.c_double) for u in self.op_u])
        self._vt_ptrs = (C.POINTER(C.c_double) * n_ops)(
            *[(_p(v, C.c_double) if v is not None else C.POINTER(C.c_double)()) for v in self.op_vt])
        self._ranks = np.array(self.op_rank, dtype=np.int32)
        # children for M2M / L2L
        self.children = {self.index[k]: [self.index[c] for c in ch] for k, ch in L.children.items() if ch}
        self.child_slot = np.array([morton.get_child_index(k, self.dim) for k in self.keys], dtype=np.int64)
        self.m2m = [np.ascontiguousarray(m) for m in ops.m2m]
        # C transfers (oracle_passes.c: orc_p2m / orc_l2p / orc_transfer_level): per-leaf cell ids, T_k(node_m) table,
        # per level the parents with their children as CSR
        self.leaf_cell = np.array([self.index[k] for k in self.leaf_keys], dtype=np.int64)
        self.tn = np.ascontiguousarray(tree.ops.polynomial_nodes, dtype=np.float64)
        self.m2m_all = np.ascontiguousarray(np.stack(self.m2m), dtype=np.float64)
        self.level_transfers = {}
        for lvl, cells in L.level_cells_map.items():
            par = [self.index[k] for k in cells if self.children.get(self.index[k])]
            ptr, idx = [0], []
            for pi in par:
                idx.extend(self.children[pi])
                ptr.append(len(idx))
            self.level_transfers[lvl] = (np.array(par, dtype=np.int64), np.array(ptr, dtype=np.int64),
                                         np.array(idx, dtype=np.int64))

    # ------------------------------------------------------------------------------------
    def upward(self, w):
        nrhs = w.shape[1]
        M = np.zeros((self.nc, nrhs, self.P))
        l = lib()
        wc = np.ascontiguousarray(w)
        l.orc_p2m(self.p, self.dim, self.P, nrhs, len(self.leaf_keys), _p(self.leaf_cell, C.c_int64),
                  _p(self.t_ptr, C.c_int64), _p(self.t_idx, C.c_int64), _p(self.src, C.c_double), _p(wc, C.c_double),
                  _p(self.center, C.c_double), _p(self.half, C.c_double), _p(self.tn, C.c_double), _p(M, C.c_double))
        for lvl in range(self.t.depth - 1, 0, -1):
            self._transfer(lvl, M, up=1, nrhs=nrhs)
        return M

    def _transfer(self, lvl, X, up, nrhs):
        par, ptr, idx = self.level_transfers.get(lvl, (None, None, None))
        if par is None or len(par) == 0:
            return
        lib().orc_transfer_level(self.P, nrhs, up, len(par), _p(par, C.c_int64), _p(ptr, C.c_int64),
                                 _p(idx, C.c_int64), _p(self.child_slot, C.c_int64), _p(self.m2m_all, C.c_double),
                                 _p(X, C.c_double))

    def upward_numpy(self, w):
        """the same pass with numpy per cell (the round-1 implementation; kept as the cross-check of the C version)"""
        t = self.t
        nrhs = w.shape[1]
        M = np.zeros((self.nc, nrhs, self.P))
        for k in self.leaf_keys:
            idx = t.lists.leaf_source_indices[k]
            i = self.index[k]
            S, _ = chebyshev.get_approximation_coefficients(self.p, self.src[idx], self.center[i], 2 * self.half[i],
                                                            t.ops.polynomial_nodes, self.dim)
            M[i] += (S.T @ w[idx]).T
        for lvl in range(t.depth - 1, 0, -1):
            for k in t.lists.level_cells_map.get(lvl, []):
                i = self.index[k]
                for c in self.children.get(i, []):
                    M[i] += M[c] @ self.m2m[self.child_slot[c]].T
        return M

    def downward(self, w, M):
        nrhs = w.shape[1]
        Lc = np.zeros((self.nc, nrhs, self.P))
        l = lib()
        wc = np.ascontiguousarray(w)
        l.orc_m2l(self.P, nrhs, len(self.m2l_tgt), _p(self.m2l_tgt, C.c_int64), _p(self.m2l_ptr, C.c_int64),
                  _p(self.m2l_src, C.c_int64), _p(self.m2l_op, C.c_int32), _p(self.m2l_perm, C.c_int32),
                  _p(self.perm_tab, C.c_int32), _p(self.inv_tab, C.c_int32), self._u_ptrs, self._vt_ptrs,
                  _p(self._ranks, C.c_int32), _p(M, C.c_double), _p(Lc, C.c_double))
        if len(self.x_cells):
            l.orc_p2l(C.byref(self.kernel), self.dim, nrhs, len(self.x_cells), _p(self.x_cells, C.c_int64),
                      _p(self.x_ptr, C.c_int64), _p(self.x_idx, C.c_int64), _p(self.src, C.c_double),
                      _p(wc, C.c_double), _p(self.center, C.c_double), _p(self.half, C.c_double), self.P,
                      _p(self.nodes_nd, C.c_double), _p(Lc, C.c_double))
        for lvl in range(1, self.t.depth + 1):
            self._transfer(lvl, Lc, up=0, nrhs=nrhs)
        return Lc

    def leaf_pass(self, w, M, Lc, leaf_subset=None):
        """returns out (n x nrhs); leaf_subset = indices into self.leaf_keys (None = all leaves)"""
        t = self.t
        nrhs = w.shape[1]
        out = np.zeros((self.src.shape[0], nrhs))
        sel = np.arange(len(self.leaf_keys)) if leaf_subset is None else np.asarray(leaf_subset)
        # build sub-CSR for the selected leaves
        def sub(ptr, idx):
            lens = ptr[sel + 1] - ptr[sel]
            nptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            parts = [idx[ptr[s]:ptr[s + 1]] for s in sel]
            return nptr, (np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64))
        tp, ti = sub(self.t_ptr, self.t_idx)
        up, ui = sub(self.u_ptr, self.u_idx)
        wp, wi = sub(self.w_ptr, self.w_cell)
        wc = np.ascontiguousarray(w)
        Mc = np.ascontiguousarray(M)
        lib().orc_leaf_pass(C.byref(self.kernel), self.dim, nrhs, len(sel), _p(tp, C.c_int64), _p(ti, C.c_int64),
                            _p(self.src, C.c_double), _p(up, C.c_int64), _p(ui, C.c_int64), _p(self.src, C.c_double),
                            _p(wc, C.c_double), _p(wp, C.c_int64), _p(wi, C.c_int64), _p(self.center, C.c_double),
                            _p(self.half, C.c_double), _p(Mc, C.c_double), self.P, _p(self.nodes_nd, C.c_double),
                            _p(out, C.c_double))
        Lcc = np.ascontiguousarray(Lc)
        lcell = np.ascontiguousarray(self.leaf_cell[sel])
        lib().orc_l2p(self.p, self.dim, self.P, nrhs, len(sel), _p(lcell, C.c_int64), _p(tp, C.c_int64), _p(ti, C.c_int64),
                      _p(self.src, C.c_double), _p(self.center, C.c_double), _p(self.half, C.c_double),
                      _p(self.tn, C.c_double), _p(Lcc, C.c_double), _p(out, C.c_double))
        return out

    def matvec(self, w):
        w = np.asarray(w, dtype=np.float64)
        if w.ndim == 1:
            w = w[:, None]
        M = self.upward(w)
        Lc = self.downward(w, M)
        return self.leaf_pass(w, M, Lc)

    # ------------------------------------------------------------------------------------
    def timed_matvec_estimate(self, w, leaf_fraction=0.02, seed=0):
        """CPU-baseline helper: wall time of one full matvec on this host, with the leaf pass (P2P+M2P,
        the dominant cost) timed on a random sample of target leaves and scaled by the pair counts;
        upward pass, M2L, P2L and L2L are timed in full.  Returns (seconds, detail dict)."""
        import time
        w = np.asarray(w, dtype=np.float64)
        if w.ndim == 1:
            w = w[:, None]
        t0 = time.perf_counter()
        M = self.upward(w)
        t1 = time.perf_counter()
        Lc = self.downward(w, M)
        t2 = time.perf_counter()
        nl = len(self.leaf_keys)
        rng = np.random.default_rng(seed)
        k = max(1, min(nl, int(round(nl * leaf_fraction))))
        sel = np.sort(rng.choice(nl, size=k, replace=False))
        nt = (self.t_ptr[1:] - self.t_ptr[:-1]).astype(np.float64)
        work = nt * ((self.u_ptr[1:] - self.u_ptr[:-1]) + (self.w_ptr[1:] - self.w_ptr[:-1]) * self.P)
        t3 = time.perf_counter()
        self.leaf_pass(w, M, Lc, sel)
        t4 = time.perf_counter()
        scale = float(work.sum() / max(work[sel].sum(), 1.0))
        total = (t1 - t0) + (t2 - t1) + (t4 - t3) * scale
        return total, {"upward_s": t1 - t0, "downward_s": t2 - t1, "leaf_sample_s": t4 - t3,
                       "leaf_scale": scale, "sample_leaves": int(k), "leaves": int(nl)}
