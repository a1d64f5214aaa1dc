"""TEST INFRASTRUCTURE (oracle): ctypes wrapper of oracle/cfem.c -- P1 atom assembly, SpMV and (node-block-)Jacobi PCG in
plain C with OpenMP, for the sizes at which the NumPy restatement (oracle/fem.py: 95 s for one elasticity atom on a
32^3 box) cannot serve as the checker or as bench.py's CPU baseline.  ``build()`` compiles it with gcc into
oracle/_build/ (git-ignored; travels to the GPU box with the snapshot).  Only tests/, __graft_entry__ and bench.py's
CPU legs import this module."""
import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cfem.c")
LIB = os.path.join(HERE, "_build", "libcfem.so")
_lib = None

c_i32, c_i64, c_dbl, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = ctypes.CDLL(LIB)
        L.cfem_p1_bilinear.restype = c_i64
        L.cfem_p1_bilinear.argtypes = [c_i32, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
        L.cfem_spmv.restype = None
        L.cfem_spmv.argtypes = [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]
        L.cfem_pcg.restype = c_i32
        L.cfem_pcg.argtypes = [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_dbl, c_i32, ctypes.POINTER(c_dbl)]
        L.cfem_threads.restype = c_i32
        _lib = L
    return _lib


def threads():
    return int(lib().cfem_threads())


def _p(a):
    return a.ctypes.data_as(c_vp)


def pattern(cell_nodes, n_nodes, bs):
    """CSR pattern (rowptr, colidx int32) of a node-blocked Lagrange space: union of per-cell dof cliques, columns
    ascending (node-level unique + block expansion: the dof-level np.unique of oracle/fem.sparsity needs 9x the keys)."""
    cn = np.asarray(cell_nodes, dtype=np.int64)
    nd = cn.shape[1]
    r = np.repeat(cn, nd, axis=1).ravel()
    c = np.tile(cn, (1, nd)).ravel()
    key = np.unique(r * n_nodes + c)
    nr, nc = key // n_nodes, key % n_nodes
    cnt = np.bincount(nr, minlength=n_nodes)  # node neighbours per node row
    rowptr_n = np.concatenate([[0], np.cumsum(cnt)])
    # dof rows: every node row repeated bs times, each with bs*cnt columns (bs*c .. bs*c+bs-1)
    rowlen = np.repeat(cnt * bs, bs)
    rowptr = np.concatenate([[0], np.cumsum(rowlen)]).astype(np.int64)
    cols_node = (nc[:, None] * bs + np.arange(bs)[None, :]).reshape(-1)  # per node row: bs*cnt dof columns, ascending
    seg = np.repeat(np.arange(n_nodes), bs)  # node of every dof row
    start = rowptr_n[seg] * bs
    idx = np.repeat(start, rowlen) + (np.arange(rowptr[-1]) - np.repeat(rowptr[:-1], rowlen))
    colidx = cols_node[idx]
    if rowptr[-1] >= 2**31:
        raise ValueError("pattern too large for int32 indices")
    return rowptr.astype(np.int32), colidx.astype(np.int32)


def assemble_bilinear_p1(space, T, weight_cell=None, pat=None):
    """scipy CSR of the atom with form tensor T [bs, g+1, bs, g+1] on a P1 space (oracle.fem.Space); weight_cell: one
    coefficient per cell (degree-0 Expression) or None."""
    g, bs = space.gdim, space.bs
    assert space.degree == 1 and space.tdim == g, "cfem covers affine P1 simplices"
    T = np.ascontiguousarray(np.asarray(T, dtype=np.float64).reshape(bs, g + 1, bs, g + 1))
    coords = np.ascontiguousarray(space.coords, dtype=np.float64)
    cells = np.ascontiguousarray(space.cells, dtype=np.int32)
    cn = np.ascontiguousarray(space.cell_nodes, dtype=np.int64)
    if pat is None:
        pat = getattr(space, "_cfem_pattern", None)
        if pat is None:
            pat = pattern(cn, space.n_nodes, bs)
            space._cfem_pattern = pat
    rowptr, colidx = pat
    vals = np.zeros(len(colidx))
    w = None if weight_cell is None else np.ascontiguousarray(weight_cell, dtype=np.float64)
    missed = lib().cfem_p1_bilinear(g, bs, len(cells), _p(coords), _p(cells), _p(cn), _p(T), _p(w) if w is not None else None,
                                    _p(rowptr), _p(colidx), _p(vals))
    assert missed == 0, "%d contributions outside the pattern" % missed
    n = space.n_nodes * bs
    return sp.csr_matrix((vals, colidx.copy(), rowptr.copy()), shape=(n, n))


def spmv(A, x):
    A = A.tocsr()
    y = np.empty(A.shape[0])
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib().cfem_spmv(A.shape[0], _p(A.indptr.astype(np.int32)), _p(A.indices.astype(np.int32)), _p(A.data), _p(x), _p(y))
    return y


def pcg(A, b, block=1, rtol=1e-13, max_iters=100000, x0=None):
    """(x, iterations, relative residual) of the Jacobi / node-block-Jacobi PCG on a scipy CSR matrix."""
    A = A.tocsr()
    A.sort_indices()
    rp, ci = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    va = np.ascontiguousarray(A.data, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(len(b)) if x0 is None else np.array(x0, dtype=np.float64)
    rr = c_dbl(0.0)
    it = lib().cfem_pcg(len(b), _p(rp), _p(ci), _p(va), _p(b), _p(x), int(block), float(rtol), int(max_iters), ctypes.byref(rr))
    if it < 0:
        raise MemoryError("cfem_pcg")
    return x, int(it), float(rr.value)
