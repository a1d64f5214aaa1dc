"""TEST INFRASTRUCTURE ONLY: a NumPy/SciPy stand-in for the ctypes wrappers of libpgdb200.so.

The product (pgdrome_b200) has no CPU path: every wrapper in ``pgdrome_b200._lib`` raises without
a CUDA device.  To exercise the *host logic* (mini-UFL capture, form compiler, lazy functionals,
PGDProblem bookkeeping, PGD.evaluate argument handling) in the ``-m "not gpu"`` suite, the fixture
``cpu_abi`` monkeypatches those wrappers with the reference implementations below for the duration
of one test.  Nothing under pgdrome_b200/ imports this module.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

F64, I32, I64 = torch.float64, torch.int32, torch.int64


def _t(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype)


def _n(t):
    return t.detach().cpu().numpy()


def pattern_build(cell_dofs, n_dofs):
    cd = _n(cell_dofs).astype(np.int64)
    nc, ndl = cd.shape
    r = np.repeat(cd, ndl, axis=1).ravel()
    c = np.tile(cd, (1, ndl)).ravel()
    key = r * n_dofs + c
    order = np.argsort(key, kind="stable")
    ks = key[order]
    head = np.concatenate([[True], ks[1:] != ks[:-1]])
    gptr = np.concatenate([np.nonzero(head)[0], [len(ks)]])
    uk = ks[head]
    rows, cols = uk // n_dofs, uk % n_dofs
    rowptr = np.zeros(n_dofs + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return _t(np.cumsum(rowptr), I32), _t(cols, I32), _t(gptr, I64), _t(order, I32)


def vecmap_build(cell_dofs, n_dofs):
    cd = _n(cell_dofs).astype(np.int64).ravel()
    order = np.argsort(cd, kind="stable")
    cnt = np.bincount(cd, minlength=n_dofs)
    return _t(np.concatenate([[0], np.cumsum(cnt)]), I64), _t(order, I32)


def _geometry(coords, cv, tdim, gdim, dphi):
    X = coords.reshape(-1, gdim)[cv]
    J = np.swapaxes(X[:, 1:, :] - X[:, :1, :], 1, 2)
    if tdim == gdim:
        det = np.abs(np.linalg.det(J))
        grad = np.einsum("qat,etg->eqag", dphi, np.linalg.inv(J))
    else:
        G = np.einsum("egt,egs->ets", J, J)
        det = np.sqrt(np.abs(np.linalg.det(G)))
        grad = np.zeros((len(cv), dphi.shape[0], dphi.shape[1], gdim))
    return grad, det


def _slots(coords, cv, tdim, gdim, nd, phi, dphi, qw, wq):
    coords, cv = _n(coords), _n(cv).astype(np.int64)
    phi, dphi, qw = _n(phi).reshape(-1, nd), _n(dphi).reshape(-1, nd, tdim), _n(qw)
    grad, det = _geometry(coords, cv, tdim, gdim, dphi)
    D = np.zeros((len(cv), len(qw), nd, gdim + 1))
    D[..., 0] = phi[None]
    D[..., 1:] = grad
    W = det[:, None] * qw[None, :]
    if wq is not None:
        W = W * _n(wq).reshape(len(cv), len(qw))
    return D, W


def elem_bilinear(coords, cell_verts, tdim, gdim, bs, nd, phi, dphi, qw, wq, T, out=None):
    D, W = _slots(coords, cell_verts, tdim, gdim, nd, phi, dphi, qw, wq)
    T = _n(T).reshape(bs, gdim + 1, bs, gdim + 1)
    Ae = np.einsum("eq,ijkl,eqaj,eqbl->eaibk", W, T, D, D, optimize=True)
    return _t(Ae.reshape(-1))


def elem_linear(coords, cell_verts, tdim, gdim, bs, nd, phi, dphi, qw, wq, L, out=None):
    D, W = _slots(coords, cell_verts, tdim, gdim, nd, phi, dphi, qw, wq)
    L = _n(L).reshape(bs, gdim + 1)
    return _t(np.einsum("eq,ij,eqaj->eai", W, L, D, optimize=True).reshape(-1))


def gather_values(src, gptr, gidx, n_out, out=None):
    s, gp, gi = _n(src), _n(gptr), _n(gidx)
    seg = np.repeat(np.arange(n_out), np.diff(gp))  # empty groups (dofs without contributions) stay 0
    res = np.bincount(seg, weights=s[gi], minlength=n_out) if len(gi) else np.zeros(n_out)
    res[np.diff(gp) == 0] = 0.0
    r = _t(res)
    if out is not None:
        out.copy_(r)
        return out
    return r


def assemble_p1(coords, cell_verts, gdim, c_mass, c_stiff, c_adv, gptr, gidx, nnz, out=None):
    g = gdim
    phi = np.array([[1.0 / (g + 1)] * (g + 1)])
    raise NotImplementedError("cpu_abi: fused P1 kernel is covered by the gpu tests only")


def bsr_plan(rowptr, colidx, bs):
    """stand-in: block columns of the first row of every node (host logic only needs the shapes)"""
    rp, ci = _n(rowptr).astype(np.int64), _n(colidx).astype(np.int64)
    n = len(rp) - 1
    if bs < 2 or n % bs:
        return None
    cols = [ci[rp[r]:rp[r + 1]][::bs] // bs for r in range(0, n, bs)]
    return _t(np.concatenate(cols), I32), int(max(len(c) for c in cols))


def p1_rowplan_build(rowptr, colidx, cell_dofs, vptr, vidx, n_nodes):
    return cell_dofs  # the stand-in below only needs the cell -> dof table


def assemble_p1_rows(coords, cell_verts, gdim, c_mass, c_stiff, c_adv, rowptr, vptr, vent, n_nodes, out=None, coords_soa=None, nnz=None):
    """closed-form P1 simplex matrices (mass, stiffness, advection) summed into CSR order."""
    import scipy.sparse as sp

    g = gdim
    # (cell_verts None: coords are NODE coordinates and the stand-in plan `vent` = cell -> node table indexes them)
    X = _n(coords).reshape(-1, g)[_n(cell_verts if cell_verts is not None else vent).astype(np.int64)]
    J = np.swapaxes(X[:, 1:, :] - X[:, :1, :], 1, 2)
    det = np.linalg.det(J)
    Jinv = np.linalg.inv(J)  # [e, t, g]
    grad = np.concatenate([-Jinv.sum(axis=1, keepdims=True), Jinv], axis=1)  # [e, nv, g]
    vol = np.abs(det) / {1: 1.0, 2: 2.0, 3: 6.0}[g]
    nv = g + 1
    mass = vol[:, None, None] * (np.ones((nv, nv)) + np.eye(nv))[None] / ((g + 1) * (g + 2))
    stiff = vol[:, None, None] * np.einsum("eam,ebm->eab", grad, grad)
    cadv = np.zeros(g) if c_adv is None else np.asarray(list(c_adv)[:g], dtype=np.float64)
    adv = (vol / (g + 1))[:, None, None] * np.einsum("m,ebm->eb", cadv, grad)[:, None, :] * np.ones((1, nv, 1))
    Ae = c_mass * mass + c_stiff * stiff + adv
    cv = _n(vent).astype(np.int64)  # the stand-in "plan" is the cell -> dof table
    rows = np.repeat(cv[:, :, None], nv, axis=2).ravel()
    cols = np.repeat(cv[:, None, :], nv, axis=1).ravel()
    A = sp.coo_matrix((Ae.ravel(), (rows, cols)), shape=(n_nodes, n_nodes)).tocsr()
    A.sort_indices()
    assert np.array_equal(A.indptr, _n(rowptr))
    r = _t(A.data, F64)
    if out is not None:
        out.copy_(r)
        return out
    return r


def assemble_p1_tensor(gdim, bs, T, coords_soa, cell_verts, rowptr, node_vptr, node_vent, n_rows, nnz, w_cell=None, out=None):
    """general-tensor P1 atom: exact quadrature through the element stand-in, summed into CSR order (node_vent is the
    stand-in plan = the cell -> node table)"""
    from pgdrome_b200.fem import simplex_quadrature, tabulate_lagrange

    g = gdim
    pts, qw = simplex_quadrature(g, 2)
    phi, dphi = tabulate_lagrange(g, 1, pts)
    coords = _t(np.ascontiguousarray(_n(coords_soa).T))
    nq = len(qw)
    wq = None if w_cell is None else _t(np.repeat(_n(w_cell)[:, None], nq, axis=1))
    Ae = _n(elem_bilinear(coords, cell_verts, g, g, bs, g + 1, _t(phi), _t(dphi), _t(qw), wq, _t(np.asarray(T, dtype=np.float64))))
    cn = _n(node_vent).astype(np.int64)
    cd = (cn[:, :, None] * bs + np.arange(bs)[None, None, :]).reshape(len(cn), -1)
    ndl = cd.shape[1]
    rows = np.repeat(cd[:, :, None], ndl, axis=2).ravel()
    cols = np.repeat(cd[:, None, :], ndl, axis=1).ravel()
    n = int(n_rows)
    A = sp.coo_matrix((Ae.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    A.sort_indices()
    rp, ci = _n(rowptr), None
    # explicit zeros of the pattern are dropped by COO -> CSR: scatter into the full pattern
    vals = np.zeros(int(nnz))
    full_rows = np.repeat(np.arange(n), np.diff(rp))
    import pgdrome_b200._lib as L  # pattern columns are not passed: rebuild them from the cell cliques (same union)
    r2 = np.repeat(cd, ndl, axis=1).ravel()
    c2 = np.tile(cd, (1, ndl)).ravel()
    key = np.unique(r2 * n + c2)
    assert len(key) == int(nnz)
    pos = np.searchsorted(key, A.tocoo().row.astype(np.int64) * n + A.tocoo().col)
    vals[pos] = A.tocoo().data
    r = _t(vals)
    if out is not None:
        out.copy_(r)
        return out
    return r


def lincomb(xs, coefs, out=None, accumulate=False):
    n = (out if out is not None else xs[0]).numel()
    acc = out.clone() if (accumulate and out is not None) else torch.zeros(n, dtype=F64)
    for x, c in zip(xs, coefs):
        acc = acc + float(c) * x
    if out is not None:
        out.copy_(acc)
        return out
    return acc


def _csr(rowptr, colidx, values, ncols=None):
    """rows = len(rowptr) - 1 (an element-partitioned space passes the owned rows of a wider local pattern)"""
    n = rowptr.numel() - 1
    rp = _n(rowptr)
    nnz = int(rp[-1])
    return sp.csr_matrix((_n(values)[:nnz], _n(colidx)[:nnz], rp), shape=(n, n if ncols is None else ncols))


def apply_dirichlet(rowptr, colidx, values, b, bc_dofs, bc_vals=None):
    if bc_dofs is None or bc_dofs.numel() == 0:
        return
    bc = _n(bc_dofs).astype(np.int64)
    g = np.zeros(rowptr.numel() - 1)
    g[bc] = _n(bc_vals) if bc_vals is not None else 0.0
    if values is not None:
        A = _csr(rowptr, colidx, values)
        if b is not None:
            b -= _t(A @ g)
        mask = np.zeros(A.shape[0], dtype=bool)
        mask[bc] = True
        rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
        kill = mask[rows] | mask[A.indices]
        v = _n(values).copy()
        v[kill] = 0.0
        v[kill & (rows == A.indices)] = 1.0
        values.copy_(_t(v))
    if b is not None:
        b[torch.as_tensor(bc)] = _t(g[bc])


def set_entries(x, idx, vals=None):
    if idx is None or idx.numel() == 0:
        return
    x[idx.long()] = vals if vals is not None else 0.0


def spmv(rowptr, colidx, values, x, y=None, lpr=0):
    r = _t(_csr(rowptr, colidx, values, x.numel()) @ _n(x))
    if y is not None:
        y[: r.numel()].copy_(r)
        return y
    return r


def spmv_dot(rowptr, colidx, values, x, w, y=None, out=None, lpr=0):
    y = spmv(rowptr, colidx, values, x, y)
    d = torch.dot(w, y).reshape(1)
    if out is not None:
        out.copy_(d)
        return y, out
    return y, d


def bilinear(rowptr, colidx, values, x, y, out=None, lpr=0):
    M = _csr(rowptr, colidx, values, y.numel())
    d = _t(np.array([_n(x)[: M.shape[0]] @ (M @ _n(y))]))
    if out is not None:
        out.copy_(d)
        return out
    return d


def dot(x, y, out=None):
    d = torch.dot(x, y).reshape(1)
    if out is not None:
        out.copy_(d)
        return out
    return d


def panel_dots(P, n_vecs, x, out=None):
    d = P[:n_vecs, : x.numel()] @ x
    if out is not None:
        out.copy_(d)
        return out
    return d


def pcg(rowptr, colidx, values, b, x=None, rtol=1e-12, atol=0.0, maxit=20000, check_every=50, block=1, lpr=0, work=None,
        x0=None):
    A = _csr(rowptr, colidx, values)
    bb = _n(b)
    if not np.any(bb):
        return torch.zeros_like(b), 0, 0.0
    sol = spla.spsolve(A.tocsc(), bb)
    res = np.linalg.norm(A @ sol - bb) / np.linalg.norm(bb)
    return _t(sol), 1, float(res)


def banded_solve(rowptr, colidx, values, b, perm, kl, ku, x=None, work=None, info=None):
    A = _csr(rowptr, colidx, values)
    return _t(spla.spsolve(A.tocsc(), _n(b))), torch.zeros(1, dtype=I32)


def eval_weights(xs, cds, Phis, degs, R, points, out=None):
    P = _n(points)
    C = P.shape[0]
    W = np.ones((R, C))
    flag = 0
    for i, (x, cd, Ph, dg) in enumerate(zip(xs, cds, Phis, degs)):
        x, cd, Ph = _n(x), _n(cd), _n(Ph)
        p = P[:, i]
        tol = 1e-12 * max(abs(x[-1] - x[0]), 1e-300)
        if flag == 0 and np.any((p < x[0] - tol) | (p > x[-1] + tol)):
            flag = 1 + i
        p = np.clip(p, x[0], x[-1])
        e = np.clip(np.searchsorted(x, p, side="right") - 1, 0, len(x) - 2)
        xi = (p - x[e]) / (x[e + 1] - x[e])
        if dg == 1:
            v = Ph[:R][:, cd[e, 0]] * (1 - xi) + Ph[:R][:, cd[e, 1]] * xi
        else:
            v = (Ph[:R][:, cd[e, 0]] * ((1 - xi) * (1 - 2 * xi)) + Ph[:R][:, cd[e, 1]] * (xi * (2 * xi - 1))
                 + Ph[:R][:, cd[e, 2]] * (4 * xi * (1 - xi)))
        W *= v
    return _t(W), torch.tensor([flag], dtype=I32)


def eval_gemv(X, R, w, out=None):
    return X[:R].T @ w[:R]


def eval_gemm(W, X, R, out=None):
    r = W[:R].T @ X[:R]
    if out is not None:
        out.copy_(r)
        return out
    return r


def row_stats(U, F=None):
    u = _n(U)
    f = _n(F) if F is not None else None
    out = np.zeros((u.shape[0], 7))
    out[:, 0], out[:, 1] = u.min(axis=1), u.max(axis=1)
    out[:, 2], out[:, 3] = np.abs(u).min(axis=1), np.abs(u).max(axis=1)
    out[:, 4] = (u * u).sum(axis=1)
    if f is not None:
        out[:, 5], out[:, 6] = ((u - f) ** 2).sum(axis=1), (f * f).sum(axis=1)
    return _t(out)


def locate_points(coords, cells, points, tol=1e-10):
    from oracle.evaluate import locate_points as ref

    c, b = ref(_n(coords), _n(cells), _n(points), tol)
    return torch.as_tensor(c.astype(np.int32)), _t(b)


def probe_modes(X, R, dofs, w):
    Xn, d, wn = _n(X)[:R], _n(dofs), _n(w)
    return _t(np.einsum("rj,krj->kr", wn, Xn[:, d]))


def scalar_programs(programs, consts, pool, out):
    P = _n(pool) if pool is not None else None
    for g, code in enumerate(programs):
        st = []
        for ins in code:
            op, arg = ins >> 24, ins & 0xffffff
            if op == 0:
                st.append(np.float64(consts[arg]))
            elif op == 1:
                st.append(np.float64(P[arg]))
            elif op == 6:
                st[-1] = -st[-1]
            else:
                y, x = st.pop(), st.pop()
                st.append(x * y if op == 2 else x + y if op == 3 else x - y if op == 4 else x / y)
        assert len(st) == 1
        out[g] = float(st[0])
    return out


def pcg_start(*a, **k):
    return None  # the stand-in has no resident solver: the host logic falls back to pcg()


def pcg_finish(device=None):
    return -1, 0.0


NAMES = ["pattern_build", "vecmap_build", "elem_bilinear", "elem_linear", "gather_values", "assemble_p1",
         "p1_rowplan_build", "assemble_p1_rows", "assemble_p1_tensor", "bsr_plan", "lincomb",
         "apply_dirichlet", "set_entries", "spmv", "spmv_dot", "bilinear", "dot", "panel_dots", "pcg", "banded_solve",
         "eval_weights", "eval_gemv", "eval_gemm", "row_stats", "locate_points", "probe_modes", "pcg_start", "pcg_finish", "scalar_programs"]


def install(monkeypatch):
    """Patch pgdrome_b200._lib (and the cached device helpers) for one test."""
    import sys

    from pgdrome_b200 import _lib, assembly, functions

    me = sys.modules[__name__]
    for name in NAMES:
        monkeypatch.setattr(_lib, name, getattr(me, name))
    monkeypatch.setattr(_lib, "require_cuda", lambda: None)
    monkeypatch.setattr(_lib, "to_device", lambda a, dtype=None: torch.as_tensor(
        a if isinstance(a, torch.Tensor) else np.array(a, copy=True, order="C"), dtype=dtype).clone())
    monkeypatch.setattr(_lib, "to_host", lambda t: t.detach().cpu().numpy())
    cpu = torch.device("cpu")
    monkeypatch.setattr(assembly, "_dev", lambda: cpu)
    monkeypatch.setattr(functions, "_device", lambda: cpu)
    for mod in ("pgdrome_b200.forms", "pgdrome_b200.model", "pgdrome_b200.solver"):
        if mod in sys.modules and hasattr(sys.modules[mod], "_device"):
            monkeypatch.setattr(sys.modules[mod], "_device", lambda: cpu)
