"""-m gpu: every C-ABI kernel of libpgdb200.so against the CPU oracle on the same seeded inputs.

Bars (north_star): sparsity pattern bit-exact; assembled matrices 1e-12 relative; solves / modes
1e-8 relative (tighter here because single solves are compared, not whole enrichments).
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

from oracle import fem as ofem
from oracle import meshes as omesh

pytestmark = pytest.mark.gpu

MAT_RTOL = 1e-12


def _product_space(kind, degree, bs):
    from pgdrome_b200 import fem

    if kind == "interval":
        m = fem.IntervalMesh(37, 0.2, 2.0)
        o = omesh.interval_mesh(37, 0.2, 2.0)
    elif kind == "tri":
        m = fem.RectangleMesh((0.0, 0.0), (3.0, 1.0), 10, 5, "crossed")
        o = omesh.rectangle_mesh(0.0, 0.0, 3.0, 1.0, 10, 5, "crossed")
    elif kind == "tri_right":
        m = fem.UnitSquareMesh(17, 13)
        o = omesh.rectangle_mesh(0.0, 0.0, 1.0, 1.0, 17, 13, "right")
    else:
        m = fem.BoxMesh((0.0, 0.0, 0.0), (1.0, 2.0, 1.5), 5, 4, 3)
        o = omesh.box_mesh(0.0, 0.0, 0.0, 1.0, 2.0, 1.5, 5, 4, 3)
    V = fem.FunctionSpace(m, "P", degree, bs)
    S = ofem.Space(o[0], o[1], degree, bs)
    assert np.array_equal(V.cell_dofs, S.cell_dofs)
    return V, S


def _csr(ds, values):
    rowptr, colidx, _, _ = ds.pattern
    n = ds.n_dofs
    return sp.csr_matrix((values.cpu().numpy(), colidx.cpu().numpy(), rowptr.cpu().numpy()), shape=(n, n))


def _relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


SPACES = [("interval", 1, 1), ("interval", 2, 1), ("tri", 1, 1), ("tri", 2, 2), ("tri_right", 1, 1), ("tet", 1, 1),
          ("tet", 1, 3), ("tet", 2, 1)]


@pytest.mark.parametrize("kind,degree,bs", SPACES)
def test_pattern_bit_exact(kind, degree, bs):
    from pgdrome_b200.assembly import device_space

    V, S = _product_space(kind, degree, bs)
    ds = device_space(V)
    rowptr, colidx, gptr, gidx = ds.pattern
    rp, ci = ofem.sparsity(S.cell_dofs, S.n_dofs)
    assert np.array_equal(rowptr.cpu().numpy(), rp)
    assert np.array_equal(colidx.cpu().numpy(), ci)
    # gather lists: every contribution appears once, in ascending order inside a group
    gi = gidx.cpu().numpy()
    assert np.array_equal(np.sort(gi), np.arange(gi.size))
    gp = gptr.cpu().numpy()
    assert gp[0] == 0 and gp[-1] == gi.size and np.all(np.diff(gp) > 0)


@pytest.mark.parametrize("kind,degree,bs", SPACES)
def test_atoms_match_oracle(kind, degree, bs):
    from pgdrome_b200.assembly import device_space

    V, S = _product_space(kind, degree, bs)
    ds = device_space(V)
    g = S.gdim
    atoms = {"mass": ofem.T_mass(bs, g), "stiff": ofem.T_stiff(bs, g)}
    if bs == 1:
        atoms["adv"] = ofem.T_adv(g, g - 1)
    if bs == g and g >= 2:
        C = ofem.isotropic_C(1.3, 0.7, g) if g == 3 else np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 0.0]])
        atoms["voigt"] = ofem.T_voigt(C, g)
    for name, T in atoms.items():
        A = _csr(ds, ds.assemble_bilinear(T))
        Ao = ofem.assemble_bilinear(S, T)
        assert _relerr(A.data, Ao.data) < MAT_RTOL, name
    # weighted mass, coefficient interpolated into P_p per cell (Expression(degree=p) semantics)
    if bs == 1:
        if kind == "interval":
            w, p = (lambda x: 1.0 / (2.0 * (1.0 + x[..., 0]) * (1.0 + 0.3 * x[..., 0]))), 10
        else:
            w, p = (lambda x: 1.0 + x[..., 0] * x[..., 1]), 2
        A = _csr(ds, ds.assemble_bilinear(ofem.T_mass(1, g), weight=w, wdeg=p))
        Ao = ofem.assemble_bilinear(S, ofem.T_mass(1, g), weight=w, weight_degree=p)
        assert _relerr(A.data, Ao.data) < MAT_RTOL
        L = np.zeros((1, g + 1))
        L[0, 0] = 1.0
        b = ds.assemble_linear(L, weight=w, wdeg=p).cpu().numpy()
        bo = ofem.assemble_linear(S, L, weight=w, weight_degree=p)
        assert _relerr(b, bo) < MAT_RTOL
        L[0, 1] = 0.5
        b = ds.assemble_linear(L).cpu().numpy()
        assert _relerr(b, ofem.assemble_linear(S, L)) < MAT_RTOL


@pytest.mark.parametrize("kind", ["interval", "tri_right", "tet"])
def test_fused_p1_operator(kind):
    from pgdrome_b200 import _lib
    from pgdrome_b200.assembly import device_space

    V, S = _product_space(kind, 1, 1)
    ds = device_space(V)
    g = S.gdim
    cm, ck = 0.37, 2.25
    cadv = [0.5, -0.25, 0.125][:g]
    rowptr, colidx, gptr, gidx = ds.pattern
    vals = _lib.assemble_p1(ds.coords, ds.cell_verts, g, cm, ck, cadv, gptr, gidx, colidx.numel())
    Ao = cm * ofem.assemble_bilinear(S, ofem.T_mass(1, g)) + ck * ofem.assemble_bilinear(S, ofem.T_stiff(1, g))
    for m in range(g):
        Ao = Ao + cadv[m] * ofem.assemble_bilinear(S, ofem.T_adv(g, m))
    assert _relerr(_csr(ds, vals).data, Ao.tocsr().data) < MAT_RTOL
    # row-owner kernel (one thread per mesh node): same operator, bitwise reproducible
    assert ds.rowplan is not False
    v2 = _lib.assemble_p1_rows(ds.coords, ds.cell_verts, g, cm, ck, cadv, rowptr, ds.vecmap[0], ds.rowplan, ds.n_dofs)
    assert _relerr(_csr(ds, v2).data, Ao.tocsr().data) < MAT_RTOL
    v3 = _lib.assemble_p1_rows(ds.coords, ds.cell_verts, g, cm, ck, cadv, rowptr, ds.vecmap[0], ds.rowplan, ds.n_dofs)
    assert torch.equal(v2, v3)
    # component-major coordinates (the layout the atom assembly uses) and the no-advection instantiation
    v4 = _lib.assemble_p1_rows(ds.coords, ds.cell_verts, g, cm, ck, cadv, rowptr, ds.vecmap[0], ds.rowplan, ds.n_dofs,
                               coords_soa=ds.coords_soa)
    assert torch.equal(v2, v4)
    v5 = _lib.assemble_p1_rows(ds.coords, ds.cell_verts, g, cm, ck, None, rowptr, ds.vecmap[0], ds.rowplan, ds.n_dofs,
                               coords_soa=ds.coords_soa)
    Ao0 = cm * ofem.assemble_bilinear(S, ofem.T_mass(1, g)) + ck * ofem.assemble_bilinear(S, ofem.T_stiff(1, g))
    assert _relerr(_csr(ds, v5).data, Ao0.tocsr().data) < MAT_RTOL
    # and it is what the atom assembly of a scalar P1 space uses
    K = ds.assemble_bilinear(ofem.T_stiff(1, g))
    assert _relerr(_csr(ds, K).data, ofem.assemble_bilinear(S, ofem.T_stiff(1, g)).tocsr().data) < MAT_RTOL


@pytest.mark.parametrize("kind,bs", [("interval", 1), ("interval", 2), ("tri", 1), ("tri", 2), ("tri_right", 3), ("tet", 1),
                                     ("tet", 2), ("tet", 3)])
def test_tensor_p1_kernel(kind, bs):
    """pgd_assemble_p1_tensor: ANY constant form tensor T[iv,jv,iu,ju] on an affine P1 space (scalar or vector), with and
    without a cell-wise coefficient, straight into the CSR pattern; it is the route assemble_bilinear takes for P1."""
    from pgdrome_b200 import _lib
    from pgdrome_b200.assembly import device_space

    V, S = _product_space(kind, 1, bs)
    ds = device_space(V)
    g = S.gdim
    assert ds.node_plan is not False
    rng = np.random.default_rng(7 + bs)
    T = rng.uniform(-1.0, 1.0, (bs, g + 1, bs, g + 1))
    rowptr = ds.pattern[0]
    vals = _lib.assemble_p1_tensor(g, bs, T, ds.coords_soa, ds.cell_verts, rowptr, ds.node_plan[0], ds.node_plan[1], ds.n_dofs, ds.nnz)
    Ao = ofem.assemble_bilinear(S, T)
    assert _relerr(_csr(ds, vals).data, Ao.data) < MAT_RTOL
    again = _lib.assemble_p1_tensor(g, bs, T, ds.coords_soa, ds.cell_verts, rowptr, ds.node_plan[0], ds.node_plan[1], ds.n_dofs, ds.nnz)
    assert torch.equal(vals, again)  # one thread owns one row, fixed cell order
    assert torch.equal(ds.assemble_bilinear(T), vals)  # the public route uses it
    # cell-wise (degree-0) coefficient, e.g. a material zone indicator
    w = lambda x: 1.0 + 2.0 * (x[..., 0] > 0.45) + 0.25 * x[..., 0]  # noqa: E731
    Aw = ds.assemble_bilinear(T, weight=w, wdeg=0)
    Aow = ofem.assemble_bilinear(S, T, weight=w, weight_degree=0)
    assert _relerr(_csr(ds, Aw).data, Aow.data) < MAT_RTOL
    if bs == g:
        C = ofem.isotropic_C(1.3, 0.7, g) if g == 3 else (np.array([[2.0, 0.6, 0.0], [0.6, 2.0, 0.0], [0.0, 0.0, 0.7]]) if g == 2 else None)
        if C is not None:
            Tv = ofem.T_voigt(C, g)
            assert _relerr(_csr(ds, ds.assemble_bilinear(Tv, weight=w, wdeg=0)).data,
                           ofem.assemble_bilinear(S, Tv, weight=w, weight_degree=0).data) < MAT_RTOL


def test_fused_p1_rows_large_mesh():
    """Row blocks whose value slice exceeds the shared-memory buffer fall back to global accumulation;
    a 20^3 box exercises full interior rows (15 entries) over many CTAs."""
    from pgdrome_b200 import _lib, fem
    from pgdrome_b200.assembly import device_space

    m = fem.UnitCubeMesh(20, 20, 20)
    V = fem.FunctionSpace(m, "P", 1)
    o = omesh.box_mesh(0.0, 0.0, 0.0, 1.0, 1.0, 1.0, 20, 20, 20)
    S = ofem.Space(o[0], o[1], 1, 1)
    ds = device_space(V)
    rowptr, colidx, gptr, gidx = ds.pattern
    v_old = _lib.assemble_p1(ds.coords, ds.cell_verts, 3, 0.3, 1.7, None, gptr, gidx, colidx.numel())
    v_new = _lib.assemble_p1_rows(ds.coords, ds.cell_verts, 3, 0.3, 1.7, None, rowptr, ds.vecmap[0], ds.rowplan, ds.n_dofs)
    Ao = 0.3 * ofem.assemble_bilinear(S, ofem.T_mass(1, 3)) + 1.7 * ofem.assemble_bilinear(S, ofem.T_stiff(1, 3))
    assert _relerr(_csr(ds, v_new).data, Ao.tocsr().data) < MAT_RTOL
    assert _relerr(v_new.cpu().numpy(), v_old.cpu().numpy()) < MAT_RTOL


def test_facet_load_matches_oracle():
    from pgdrome_b200.assembly import device_space

    V, S = _product_space("tri", 2, 2)
    ds = device_space(V)
    m = V.mesh()
    cell, loc = m.boundary_facets()
    fv = m.facet_vertices(cell, loc)
    X = m.coordinates()
    top_left = np.all(np.abs(X[fv][:, :, 1] - 1.0) < 1e-12, axis=1) & np.all(X[fv][:, :, 0] < 1.5 + 1e-9, axis=1)
    b = ds.assemble_facet_linear("tl", cell[top_left], loc[top_left], (0.0, -0.5)).cpu().numpy()
    fs = ofem.facet_space(S, lambda x: abs(x[1] - 1.0) < 1e-12 and x[0] < 1.5 + 1e-9)
    Lt = np.zeros((2, 3))
    Lt[:, 0] = (0.0, -0.5)
    bo = ofem.assemble_linear(fs, Lt)
    assert abs(bo.sum() + 0.5 * 1.5) < 1e-12
    assert _relerr(b, bo) < MAT_RTOL


def _random_system(kind="tet", degree=1, bs=1, seed=0):
    from pgdrome_b200.assembly import device_space

    V, S = _product_space(kind, degree, bs)
    ds = device_space(V)
    g = S.gdim
    K1 = ofem.assemble_bilinear(S, ofem.T_stiff(bs, g))
    K2 = ofem.assemble_bilinear(S, ofem.T_mass(bs, g))
    # add the value arrays: scipy's sparse '+' would drop the pattern's explicit zeros
    K = sp.csr_matrix((K1.data + 0.3 * K2.data, K1.indices, K1.indptr), shape=K1.shape)
    rng = np.random.default_rng(seed)
    return V, S, ds, K, rng


def test_spmv_bilinear_dots_lincomb():
    from pgdrome_b200 import _lib

    V, S, ds, K, rng = _random_system()
    rowptr, colidx, _, _ = ds.pattern
    dev = rowptr.device
    vals = torch.as_tensor(K.data).to(dev)
    x = rng.uniform(-1, 1, S.n_dofs)
    y = rng.uniform(-1, 1, S.n_dofs)
    xd, yd = torch.as_tensor(x).to(dev), torch.as_tensor(y).to(dev)
    for lpr in (0, 2, 4, 8, 16, 32):
        out = _lib.spmv(rowptr, colidx, vals, xd, lpr=lpr).cpu().numpy()
        assert _relerr(out, K @ x) < 1e-14
        yy, d = _lib.spmv_dot(rowptr, colidx, vals, xd, yd, lpr=lpr)
        assert _relerr(yy.cpu().numpy(), K @ x) < 1e-14
        assert abs(d.item() - y @ (K @ x)) < 1e-12 * abs(y @ (K @ x)) + 1e-13
        s = _lib.bilinear(rowptr, colidx, vals, xd, yd, lpr=lpr).item()
        assert abs(s - x @ (K @ y)) < 1e-12 * abs(x @ (K @ y)) + 1e-13
    assert abs(_lib.dot(xd, yd).item() - x @ y) < 1e-12
    # bitwise reproducible reductions
    a = _lib.bilinear(rowptr, colidx, vals, xd, yd).item()
    assert all(_lib.bilinear(rowptr, colidx, vals, xd, yd).item() == a for _ in range(5))
    P = rng.uniform(-1, 1, (7, S.n_dofs))
    Pd = torch.as_tensor(P).to(dev)
    assert _relerr(_lib.panel_dots(Pd, 7, xd).cpu().numpy(), P @ x) < 1e-13
    coefs = rng.uniform(-2, 2, 30)
    vecs = [torch.as_tensor(rng.uniform(-1, 1, 1001)).to(dev) for _ in range(30)]
    ref = sum(c * v.cpu().numpy() for c, v in zip(coefs, vecs))
    out = _lib.lincomb(vecs, coefs)
    assert _relerr(out.cpu().numpy(), ref) < 1e-14
    out2 = _lib.lincomb(vecs[:3], coefs[:3], out=out.clone(), accumulate=True)
    assert _relerr(out2.cpu().numpy(), ref + sum(c * v.cpu().numpy() for c, v in zip(coefs[:3], vecs[:3]))) < 1e-14


@pytest.mark.parametrize("kind,degree,bs,block", [("tet", 1, 1, 1), ("tri", 2, 2, 1), ("tri", 2, 2, 2), ("tet", 1, 3, 3)])
def test_dirichlet_and_pcg(kind, degree, bs, block):
    from pgdrome_b200 import _lib

    V, S, ds, K, rng = _random_system(kind, degree, bs)
    rowptr, colidx, _, _ = ds.pattern
    dev = rowptr.device
    bc = ofem.dirichlet_dofs(S, lambda x, ob: x[0] < 1e-12)
    b = rng.uniform(-1, 1, S.n_dofs)
    Ao, bo = ofem.apply_dirichlet_sym(K, b, bc)
    vals = torch.as_tensor(K.data.copy()).to(dev)
    bd = torch.as_tensor(b.copy()).to(dev)
    _lib.apply_dirichlet(rowptr, colidx, vals, bd, torch.as_tensor(bc.astype(np.int32)).to(dev))
    assert _relerr(vals.cpu().numpy(), Ao.data) == 0.0
    assert _relerr(bd.cpu().numpy(), bo) == 0.0
    # inhomogeneous values: lifting
    gvals = rng.uniform(-1, 1, bc.size)
    Ao2, bo2 = ofem.apply_dirichlet_sym(K, b, bc, gvals)
    vals2 = torch.as_tensor(K.data.copy()).to(dev)
    bd2 = torch.as_tensor(b.copy()).to(dev)
    _lib.apply_dirichlet(rowptr, colidx, vals2, bd2, torch.as_tensor(bc.astype(np.int32)).to(dev),
                         torch.as_tensor(gvals).to(dev))
    assert _relerr(bd2.cpu().numpy(), bo2) < 1e-13
    xo = spla.spsolve(Ao.tocsc(), bo)
    its = {}
    for resident in (1, 0):  # single-kernel SM-resident PCG and the multi-kernel (HBM-streaming) PCG
        _lib.set_option("pcg_resident", resident)
        _lib.stats(reset=True)
        x, iters, relres = _lib.pcg(rowptr, colidx, vals, bd, rtol=1e-13, maxit=5000, check_every=25, block=block)
        assert _lib.stats()["pcg_resident_solves"] == resident
        assert relres <= 1e-13 and 0 < iters < 5000
        assert np.linalg.norm(x.cpu().numpy() - xo) / np.linalg.norm(xo) < 1e-10
        its[resident] = iters
        # bitwise reproducible
        x2, iters2, _ = _lib.pcg(rowptr, colidx, vals, bd, rtol=1e-13, maxit=5000, check_every=25, block=block)
        assert iters2 == iters and torch.equal(x, x2)
        # iteration cap is honoured
        _, itc, rc = _lib.pcg(rowptr, colidx, vals, bd, rtol=1e-13, maxit=7, check_every=3, block=block)
        assert itc == 7 and rc > 1e-13
        # zero right-hand side converges immediately to zero
        x0, it0, _ = _lib.pcg(rowptr, colidx, vals, torch.zeros_like(bd), rtol=1e-13, maxit=10, block=block)
        assert it0 == 0 and float(x0.abs().max()) == 0.0
        # warm start: from the solution itself -> 0 iterations; from a perturbed solution -> fewer iterations, same answer
        xw, itw, rw = _lib.pcg(rowptr, colidx, vals, bd, rtol=1e-13, maxit=5000, check_every=25, block=block, x0=x)
        assert itw <= 1 and rw <= 1e-13 and np.linalg.norm(xw.cpu().numpy() - xo) / np.linalg.norm(xo) < 1e-10
        pert = x * (1.0 + 1e-6 * torch.sin(torch.arange(x.numel(), device=x.device, dtype=x.dtype)))
        xw, itw, rw = _lib.pcg(rowptr, colidx, vals, bd, rtol=1e-13, maxit=5000, check_every=25, block=block, x0=pert)
        assert 0 < itw < iters and rw <= 1e-13
        assert np.linalg.norm(xw.cpu().numpy() - xo) / np.linalg.norm(xo) < 1e-10
    _lib.set_option("pcg_resident", 1)
    assert abs(its[0] - its[1]) <= max(3, its[0] // 20)
    # started / finished solve: only after a blocking solve showed that the system fits the SM-resident solver
    other = rowptr.clone()
    assert _lib.pcg_start(other, colidx, vals, bd, rtol=1e-13, maxit=5000, block=block) is None
    assert _lib.pcg_finish() == (-1, 0.0)
    xs, it_s, _ = _lib.pcg(rowptr, colidx, vals, bd, rtol=1e-13, maxit=5000, check_every=25, block=block)
    _lib.stats(reset=True)
    xa = _lib.pcg_start(rowptr, colidx, vals, bd, rtol=1e-13, maxit=5000, block=block)
    assert xa is not None
    it_a, rr_a = _lib.pcg_finish()
    assert it_a == it_s and rr_a <= 1e-13 and torch.equal(xa, xs)  # same kernel, bitwise the same answer
    assert _lib.pcg_finish() == (-1, 0.0)
    pert = xs * (1.0 + 1e-6 * torch.cos(torch.arange(xs.numel(), device=xs.device, dtype=xs.dtype)))
    xw = _lib.pcg_start(rowptr, colidx, vals, bd, rtol=1e-13, maxit=5000, block=block, x0=pert)
    # a blocking call while a started solve is pending collects it first (its counters are kept)
    xz, itz, _ = _lib.pcg(rowptr, colidx, vals, torch.zeros_like(bd), rtol=1e-13, maxit=10, block=block)
    assert itz == 0 and _lib.pcg_finish() == (-1, 0.0)
    st = _lib.stats()
    assert st["pcg_solves"] == 3 and 0 < st["pcg_iters"] - it_a < it_s
    assert np.linalg.norm(xw.cpu().numpy() - xo) / np.linalg.norm(xo) < 1e-10


@pytest.mark.parametrize("degree", [1, 2])
def test_banded_solve_nonsymmetric(degree):
    from pgdrome_b200 import _lib
    from pgdrome_b200.assembly import device_space

    V, S = _product_space("interval", degree, 1)
    ds = device_space(V)
    rowptr, colidx, _, _ = ds.pattern
    dev = rowptr.device
    # rho c int u' v + k int u v : non-symmetric, needs pivoting-safe LU
    A = (2.0 * ofem.assemble_bilinear(S, ofem.T_adv(1)) + 0.05 * ofem.assemble_bilinear(S, ofem.T_mass(1, 1))).tocsr()
    rng = np.random.default_rng(3)
    b = rng.uniform(-1, 1, S.n_dofs)
    bc = ofem.dirichlet_dofs(S, lambda x, ob: x[0] < 0.2 + 1e-9)
    Ao, bo = ofem.apply_dirichlet_sym(A, b, bc)
    xo = spla.spsolve(Ao.tocsc(), bo)
    perm, bw = V.band_permutation()
    x, info = _lib.banded_solve(rowptr, colidx, torch.as_tensor(Ao.data).to(dev), torch.as_tensor(bo).to(dev),
                                torch.as_tensor(perm).to(dev), bw, bw)
    assert int(info.item()) == 0
    assert np.linalg.norm(x.cpu().numpy() - xo) / np.linalg.norm(xo) < 1e-11
    # a matrix that needs row interchanges (tiny diagonal)
    n = 64
    T = sp.diags([np.full(n - 1, 1.0), np.full(n, 1e-14), np.full(n - 1, -0.7)], [-1, 0, 1]).tocsr()
    T.sort_indices()
    bb = rng.uniform(-1, 1, n)
    xt, info = _lib.banded_solve(torch.as_tensor(T.indptr.astype(np.int32)).to(dev), torch.as_tensor(T.indices.astype(np.int32)).to(dev),
                                 torch.as_tensor(T.data).to(dev), torch.as_tensor(bb).to(dev),
                                 torch.arange(n, dtype=torch.int32, device=dev), 1, 1)
    xs = spla.spsolve(T.tocsc(), bb)
    assert int(info.item()) == 0
    assert np.linalg.norm(xt.cpu().numpy() - xs) / np.linalg.norm(xs) < 1e-10
    # wide band (kl = ku = 6 > 3: the warp-parallel recurrences), random entries, weak diagonal => interchanges
    n, w = 150, 6
    diags = [rng.uniform(-1, 1, n - abs(o)) * (0.05 if o == 0 else 1.0) for o in range(-w, w + 1)]
    Wm = sp.diags(diags, list(range(-w, w + 1))).tocsr()
    Wm.sort_indices()
    bb = rng.uniform(-1, 1, n)
    xw, info = _lib.banded_solve(torch.as_tensor(Wm.indptr.astype(np.int32)).to(dev), torch.as_tensor(Wm.indices.astype(np.int32)).to(dev),
                                 torch.as_tensor(Wm.data).to(dev), torch.as_tensor(bb).to(dev),
                                 torch.arange(n, dtype=torch.int32, device=dev), w, w)
    xs = spla.spsolve(Wm.tocsc(), bb)
    assert int(info.item()) == 0
    assert np.linalg.norm(Wm @ xw.cpu().numpy() - bb) / np.linalg.norm(bb) < 1e-10
    assert np.linalg.norm(xw.cpu().numpy() - xs) / np.linalg.norm(xs) < 1e-7
    # singular matrix: zero pivot reported, no exception
    Z = sp.diags([np.ones(9), np.zeros(10), np.ones(9)], [-1, 0, 1]).tolil()
    Z[3, :] = 0.0
    Z[:, 3] = 0.0
    Z = Z.tocsr()
    Z.sort_indices()
    pat = sp.diags([np.ones(9), np.ones(10), np.ones(9)], [-1, 0, 1]).tocsr()
    dat = np.array([Z[i, j] for i, j in zip(*pat.nonzero())])
    _, info = _lib.banded_solve(torch.as_tensor(pat.indptr.astype(np.int32)).to(dev), torch.as_tensor(pat.indices.astype(np.int32)).to(dev),
                                torch.as_tensor(dat).to(dev), torch.ones(10, dtype=torch.float64, device=dev),
                                torch.arange(10, dtype=torch.int32, device=dev), 1, 1)
    assert int(info.item()) > 0


def test_evaluate_kernels():
    from pgdrome_b200 import _lib

    dev = torch.device("cuda")
    rng = np.random.default_rng(7)
    R, N = 13, 1003
    spaces = [ofem.Space(*omesh.interval_mesh(11, 0.5, 2.0), degree=1), ofem.Space(*omesh.interval_mesh(6, -1.0, 1.0), degree=2)]
    Phi = [rng.normal(size=(R, s.n_dofs)) for s in spaces]
    X = rng.normal(size=(R, N))
    C = 300
    pts = np.column_stack([rng.uniform(0.5, 2.0, C), rng.uniform(-1.0, 1.0, C)])
    pts[0] = (0.5, -1.0)
    pts[1] = (2.0, 1.0)
    xs, cds = [], []
    for s in spaces:
        v = s.coords[:, 0]
        order = np.argsort(v)
        xs.append(v[order])
        cell_lo = np.minimum(s.cells[:, 0], s.cells[:, 1])
        corder = np.argsort(v[cell_lo])
        cd = []
        for e in corder:
            a, b = s.cells[e]
            if v[a] > v[b]:
                a, b = b, a
            row = [s.vertex_to_node[a], s.vertex_to_node[b]]
            if s.degree == 2:
                row.append(s.cell_nodes[e, 2])
            cd.append(row)
        cds.append(np.array(cd, dtype=np.int32))
    from oracle.evaluate import point_eval

    Wo = np.ones((R, C))
    for c in range(C):
        for i, s in enumerate(spaces):
            for k in range(R):
                Wo[k, c] *= point_eval(s, Phi[i][k], pts[c, i])
    W, flag = _lib.eval_weights([torch.as_tensor(a).to(dev) for a in xs], [torch.as_tensor(a).to(dev) for a in cds],
                                [torch.as_tensor(a).to(dev) for a in Phi], [1, 2], R, torch.as_tensor(pts).to(dev))
    assert int(flag.item()) == 0
    assert _relerr(W.cpu().numpy(), Wo) < 1e-13
    # out-of-range coordinate is reported (interp1d ValueError / test_pgdclass.py:319-326)
    bad = pts.copy()
    bad[5, 1] = 1.5
    _, flag = _lib.eval_weights([torch.as_tensor(a).to(dev) for a in xs], [torch.as_tensor(a).to(dev) for a in cds],
                                [torch.as_tensor(a).to(dev) for a in Phi], [1, 2], R, torch.as_tensor(bad).to(dev))
    assert int(flag.item()) == 2
    Xd = torch.as_tensor(X).to(dev)
    u = _lib.eval_gemv(Xd, R, W[:, 3].contiguous()).cpu().numpy()
    assert _relerr(u, Wo[:, 3] @ X) < 1e-13
    U = _lib.eval_gemm(W, Xd, R).cpu().numpy()
    assert _relerr(U, Wo.T @ X) < 1e-13
    # ragged sizes around the 128x128 tile and K > 64 (two K chunks)
    # row-sharded evaluation: a column slab of X / U passed as a strided view (evaluate_batch(rows=...))
    Xv = Xd[:, 100:612]
    Us = _lib.eval_gemm(W, Xv, R).cpu().numpy()
    assert _relerr(Us, (Wo.T @ X)[:, 100:612]) < 1e-13
    Xv = Xd[:, 101:600]  # odd offset / width: the general kernel
    Us = _lib.eval_gemm(W, Xv, R).cpu().numpy()
    assert _relerr(Us, (Wo.T @ X)[:, 101:600]) < 1e-13
    # (even leading dimensions take the cp.async double-buffered kernel, odd ones / K > 64 the general one)
    for (R2, C2, N2) in [(1, 1, 1), (50, 129, 257), (70, 128, 128), (5, 7, 1000), (50, 130, 2050), (13, 300, 1004),
                         (64, 128, 64), (3, 2, 1026), (52, 258, 1090)]:
        W2, X2 = rng.normal(size=(R2, C2)), rng.normal(size=(R2, N2))
        U2 = _lib.eval_gemm(torch.as_tensor(W2).to(dev), torch.as_tensor(X2).to(dev), R2).cpu().numpy()
        assert _relerr(U2, W2.T @ X2) < 1e-13, (R2, C2, N2)


def _spd_random_csr(n, rng, long_row=None):
    """Random sparse SPD matrix with ascending columns; `long_row` makes row 0 / column 0 dense over
    the first `long_row` entries so that one row spans several ST_TILE buffers of the stream kernel."""
    m = 9
    rows = np.repeat(np.arange(n), m)
    cols = (rows + rng.integers(-60, 61, size=rows.size)) % n
    A = sp.coo_matrix((rng.uniform(-1, 1, rows.size), (rows, cols)), shape=(n, n)).tocsr()
    if long_row:
        A = A.tolil()
        A[0, :long_row] = rng.uniform(-1, 1, long_row)
        A = A.tocsr()
    A = A + A.T
    A = A + sp.diags(np.abs(A).sum(axis=1).A1 + 1.0)
    A = A.tocsr()
    A.sort_indices()
    return A


@pytest.mark.parametrize("n,long_row", [(8192, None), (20011, 9500), (70001, None), (40003, 21000), (131075, None)])
def test_stream_spmv_and_pcg(n, long_row):
    """Row-block streaming SpMV / PCG (n >= 8192 rows) against SciPy and against the sub-warp kernels."""
    from pgdrome_b200 import _lib

    rng = np.random.default_rng(n)
    A = _spd_random_csr(n, rng, long_row)
    dev = torch.device("cuda")
    rowptr = torch.as_tensor(A.indptr.astype(np.int32)).to(dev)
    colidx = torch.as_tensor(A.indices.astype(np.int32)).to(dev)
    vals = torch.as_tensor(A.data).to(dev)
    x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    xd, yd = torch.as_tensor(x).to(dev), torch.as_tensor(y).to(dev)
    ref = A @ x
    scale = np.abs(A).dot(np.abs(x)).max()
    res = {}
    try:
        for stream in (2, 1, 0):  # TMA-pipelined (n >= 32768) / register-staged row blocks / sub-warp per row
            _lib.set_option("spmv_stream", stream)
            out = _lib.spmv(rowptr, colidx, vals, xd, lpr=8).cpu().numpy()
            assert np.abs(out - ref).max() < 1e-14 * scale
            yy, d = _lib.spmv_dot(rowptr, colidx, vals, xd, yd, lpr=8)
            assert np.abs(yy.cpu().numpy() - ref).max() < 1e-14 * scale
            assert abs(d.item() - y @ ref) < 1e-12 * np.abs(y) @ np.abs(ref)
            s = _lib.bilinear(rowptr, colidx, vals, xd, yd, lpr=8).item()
            assert abs(s - x @ (A @ y)) < 1e-12 * np.abs(x) @ np.abs(A @ y)
            assert all(_lib.bilinear(rowptr, colidx, vals, xd, yd, lpr=8).item() == s for _ in range(3))
            _lib.set_option("pcg_resident", 0)
            b = torch.as_tensor(ref).to(dev)
            xs, iters, relres = _lib.pcg(rowptr, colidx, vals, b, rtol=1e-13, maxit=2000, check_every=20, lpr=8)
            assert relres <= 1e-13 and 0 < iters < 2000
            assert np.linalg.norm(xs.cpu().numpy() - x) / np.linalg.norm(x) < 1e-10
            xs2, iters2, _ = _lib.pcg(rowptr, colidx, vals, b, rtol=1e-13, maxit=2000, check_every=20, lpr=8)
            assert iters2 == iters and torch.equal(xs, xs2)
            res[stream] = iters
    finally:
        _lib.set_option("spmv_stream", 2)
        _lib.set_option("pcg_resident", 1)
    assert max(res.values()) - min(res.values()) <= max(2, res[0] // 20)


def test_sharded_pcg_single_rank_matches_pcg():
    """The sharded-PCG kernel set (pgd_spcg_*) on one rank (no ghosts, no collectives) against pgd_pcg_sync."""
    from pgdrome_b200 import _lib, partition as pt

    for block, n in ((1, 50000), (3, 30003)):
        rng = np.random.default_rng(n)
        A = _spd_random_csr(n, rng)
        dev = torch.device("cuda")
        rowptr = torch.as_tensor(A.indptr.astype(np.int32)).to(dev)
        colidx = torch.as_tensor(A.indices.astype(np.int32)).to(dev)
        vals = torch.as_tensor(A.data).to(dev)
        x = rng.uniform(-1, 1, n)
        b = torch.as_tensor(A @ x).to(dev)
        S = pt.shard_csr(rowptr, colidx, vals, pt.RowPartition(n, 1, block), 0)
        for graph in (False, True):
            xs, it, rr = pt.sharded_pcg(S, b, rtol=1e-13, maxit=2000, check_every=10, block=block, use_graph=graph)
            assert rr <= 1e-13 and 0 < it < 2000
            assert np.linalg.norm(xs.cpu().numpy() - x) / np.linalg.norm(x) < 1e-10
        _lib.set_option("pcg_resident", 0)
        try:
            x1, it1, _ = _lib.pcg(rowptr, colidx, vals, b, rtol=1e-13, maxit=2000, check_every=10, block=block, lpr=8)
        finally:
            _lib.set_option("pcg_resident", 1)
        assert abs(it - it1) <= max(2, it1 // 20)
        assert np.linalg.norm(xs.cpu().numpy() - x1.cpu().numpy()) / np.linalg.norm(x) < 1e-10


def test_edge_cases_and_error_codes():
    """Empty and ragged inputs, argument errors reported through the C ABI's return code + pgd_last_error."""
    from pgdrome_b200 import _lib

    dev = torch.device("cuda")
    # 1 x 1 system, a row without entries, and an isolated (empty) last row
    rp = torch.tensor([0, 1, 1, 3, 3], dtype=torch.int32, device=dev)
    ci = torch.tensor([0, 0, 2], dtype=torch.int32, device=dev)
    va = torch.tensor([2.0, -1.0, 4.0], dtype=torch.float64, device=dev)
    x = torch.tensor([1.0, 5.0, 0.5, 7.0], dtype=torch.float64, device=dev)
    y = _lib.spmv(rp, ci, va, x).cpu().numpy()
    assert np.array_equal(y, [2.0, 0.0, 1.0, 0.0])
    assert _lib.bilinear(rp, ci, va, x, x).item() == 2.0 + 0.5 * 1.0
    # zero-length vectors / zero vectors are no-ops, not errors
    e = torch.empty(0, dtype=torch.float64, device=dev)
    assert _lib.lincomb([e], [1.0]).numel() == 0
    assert _lib.panel_dots(torch.empty((0, 4), dtype=torch.float64, device=dev), 0, x).numel() == 0
    assert _lib.eval_gemm(torch.empty((3, 0), dtype=torch.float64, device=dev), torch.ones((3, 5), dtype=torch.float64, device=dev),
                          3).shape == (0, 5)
    # 1 x 1 PCG
    x1, it, rr = _lib.pcg(torch.tensor([0, 1], dtype=torch.int32, device=dev), torch.tensor([0], dtype=torch.int32, device=dev),
                          torch.tensor([4.0], dtype=torch.float64, device=dev), torch.tensor([2.0], dtype=torch.float64, device=dev),
                          rtol=1e-14, maxit=5)
    assert abs(x1.item() - 0.5) < 1e-15 and it <= 1
    # argument errors: negative code and a message naming the function
    with pytest.raises(_lib.PGDB200Error, match="block must be"):
        _lib.pcg(rp, ci, va, x, block=3)  # 4 rows are not a multiple of 3
    with pytest.raises(_lib.PGDB200Error, match="pgd_set_option"):
        _lib.set_option("no_such_option", 1)
    with pytest.raises(_lib.PGDB200Error, match="CUDA tensor"):
        _lib.spmv(rp, ci, va, x.cpu())
    # a non-SPD matrix is reported (NaN / breakdown), never silently "solved"
    bad = torch.tensor([0.0], dtype=torch.float64, device=dev)
    with pytest.raises(_lib.PGDB200Error):
        _lib.pcg(torch.tensor([0, 1], dtype=torch.int32, device=dev), torch.tensor([0], dtype=torch.int32, device=dev), bad,
                 torch.tensor([1.0], dtype=torch.float64, device=dev), maxit=3)


def test_scalar_programs_bitwise_and_lincomb_dev():
    """pgd_scalar_programs evaluates postfix expressions over device scalars with separately rounded IEEE operations in
    program order: bitwise equal to the same expression evaluated by Python floats; pgd_lincomb_dev with those
    coefficients equals pgd_lincomb with the host values bit for bit.  More programs than one launch takes (32) and an
    argument error (stack underflow) are covered."""
    from pgdrome_b200 import _lib

    dev = torch.device("cuda")
    rng = np.random.default_rng(21)
    pool_h = rng.standard_normal(500) * 10.0 ** rng.integers(-8, 8, 500)
    pool = torch.as_tensor(pool_h).to(dev)
    consts = [float(v) for v in rng.standard_normal(20)] + [1.0, -1.0, 0.5]
    C, L, MUL, ADD, SUB, DIV, NEG = (_lib.OP_CONST << 24, _lib.OP_LOAD << 24, _lib.OP_MUL << 24, _lib.OP_ADD << 24,
                                     _lib.OP_SUB << 24, _lib.OP_DIV << 24, _lib.OP_NEG << 24)
    programs, expect = [], []
    for g in range(75):
        # ((c * p_a) * p_b) [- / +] (p_c * p_d) ... a left-deep chain like the separated-form coefficients, random ops
        ia = rng.integers(0, 500, 6)
        ic = int(rng.integers(0, len(consts)))
        code = [C | ic, L | int(ia[0]), MUL]
        val = consts[ic] * pool_h[ia[0]]
        for k in range(1, 6):
            op = int(rng.integers(0, 4))
            code += [L | int(ia[k]), (MUL, ADD, SUB, DIV)[op]]
            x = float(pool_h[ia[k]])
            val = (val * x, val + x, val - x, val / x)[op]
            if rng.random() < 0.3:
                code.append(NEG)
                val = -val
        programs.append(code)
        expect.append(val)
    out = torch.empty(75, dtype=torch.float64, device=dev)
    _lib.scalar_programs(programs, consts, pool, out)
    assert np.array_equal(out.cpu().numpy(), np.array(expect))  # bitwise
    # nested (right operand is itself an expression): stack depth 3
    deep = [[L | 1, L | 2, L | 3, L | 4, ADD, MUL, SUB]]
    o1 = torch.empty(1, dtype=torch.float64, device=dev)
    _lib.scalar_programs(deep, consts, pool, o1)
    assert float(o1.item()) == pool_h[1] - pool_h[2] * (pool_h[3] + pool_h[4])
    with pytest.raises(_lib.PGDB200Error):
        _lib.scalar_programs([[L | 1, MUL]], consts, pool, o1)
    # device coefficients in the linear combination == host coefficients, bit for bit
    xs = [torch.as_tensor(rng.standard_normal(100003)).to(dev) for _ in range(30)]  # more terms than one lincomb launch
    coefs_dev = out[:30].contiguous()
    a = _lib.lincomb(xs, coefs_dev)
    b = _lib.lincomb(xs, expect[:30])
    assert torch.equal(a, b)
    acc = _lib.lincomb(xs[:3], coefs_dev[:3], out=a.clone(), accumulate=True)
    assert torch.equal(acc, _lib.lincomb(xs[:3], expect[:3], out=b.clone(), accumulate=True))


@pytest.mark.parametrize("bs", [1, 3])
def test_persistent_pcg_variants(bs):
    """pgd_pcg_persist_sync on one GPU: plain CSR ring, node-block walk and the single-reduction form of the iteration
    against a direct solve; identical iteration counts (within 1 %), warm start, zero right-hand side, row statistics."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    from pgdrome_b200 import _lib, fem
    from pgdrome_b200.assembly import device_space

    m = fem.UnitCubeMesh(14, 14, 14)
    V = fem.FunctionSpace(m, "P", 1) if bs == 1 else fem.VectorFunctionSpace(m, "P", 1)
    ds = device_space(V)
    g = 3
    T = np.zeros((bs, g + 1, bs, g + 1))
    for i in range(bs):
        T[i, 0, i, 0] = 0.3
        for j in range(bs):
            T[i, 1 + i, j, 1 + j] += 1.3 if bs > 1 else 0.0
            T[i, 1 + j, j, 1 + i] += 0.7 if bs > 1 else 0.0
            T[i, 1 + j, i, 1 + j] += 0.7 if bs > 1 else 0.0
    if bs == 1:
        for k in range(1, g + 1):
            T[0, k, 0, k] = 1.7
    vals = ds.assemble_bilinear(T)
    rowptr, colidx, _, _ = ds.pattern
    n = ds.n_dofs
    A = sp.csr_matrix((vals.cpu().numpy(), colidx.cpu().numpy(), rowptr.cpu().numpy()), shape=(n, n))
    rng = np.random.default_rng(1)
    xs = rng.uniform(-1, 1, n)
    b = torch.as_tensor(A @ xs, device=vals.device)
    ref = spla.spsolve(A.tocsc(), A @ xs)
    plan = _lib.bsr_plan(rowptr, colidx, bs) if bs > 1 else None
    assert (plan is not None) == (bs > 1)
    if plan is not None:
        assert plan[0].numel() * bs * bs == colidx.numel() and plan[1] <= 27
    results = {}
    try:
        for name, sr, pl, walk in (("csr", 0, None, 1), ("csr_sr", 2, None, 1), ("bsr", 0, plan, 1), ("bsr_sr", 2, plan, 1),
                                   ("bsr_direct", 0, plan, 2), ("bsr_direct_sr", 2, plan, 2)):
            if pl is None and name.startswith("bsr"):
                continue
            _lib.set_option("single_reduction", sr)
            _lib.set_option("bsr", walk)
            x, it, rr = _lib.pcg_persist(rowptr, colidx, vals, b, block=bs, rtol=1e-13, maxit=5000, bsr=pl)
            assert rr <= 1e-13 and 0 < it < 5000, (name, it, rr)
            assert np.linalg.norm(x.cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-10, name
            results[name] = it
            # warm start from the solution: no iteration; zero right-hand side: the solution is zero whatever x0 was
            x2, it2, _ = _lib.pcg_persist(rowptr, colidx, vals, b, block=bs, rtol=1e-12, maxit=50, x0=x, bsr=pl)
            assert it2 == 0 and torch.equal(x2, x)
            x3, it3, rr3 = _lib.pcg_persist(rowptr, colidx, vals, torch.zeros_like(b), block=bs, rtol=1e-12, maxit=50,
                                            x0=x.clone(), bsr=pl)
            assert it3 == 0 and rr3 == 0.0 and float(x3.abs().max()) == 0.0
            # bitwise reproducible
            x4, it4, _ = _lib.pcg_persist(rowptr, colidx, vals, b, block=bs, rtol=1e-13, maxit=5000, bsr=pl)
            assert it4 == it and torch.equal(x4, x), name
    finally:
        _lib.set_option("single_reduction", 0)
        _lib.set_option("bsr", 1)
    base = results["csr"]
    assert all(abs(v - base) <= max(2, base // 100) for v in results.values()), results
    # the multi-launch solver on the same system: same counts
    _lib.set_option("pcg_resident", 0)
    try:
        x5, it5, _ = _lib.pcg(rowptr, colidx, vals, b, rtol=1e-13, maxit=5000, check_every=10, block=bs)
    finally:
        _lib.set_option("pcg_resident", 1)
    assert abs(it5 - base) <= max(2, base // 100)


def test_row_stats_kernel():
    from pgdrome_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(3)
    U = torch.randn((37, 1001), dtype=torch.float64, device="cuda", generator=g)
    F = U + 1e-3 * torch.randn((37, 1001), dtype=torch.float64, device="cuda", generator=g)
    st = _lib.row_stats(U, F).cpu().numpy()
    u, f = U.cpu().numpy(), F.cpu().numpy()
    assert np.array_equal(st[:, 0], u.min(axis=1)) and np.array_equal(st[:, 1], u.max(axis=1))
    assert np.array_equal(st[:, 2], np.abs(u).min(axis=1)) and np.array_equal(st[:, 3], np.abs(u).max(axis=1))
    assert np.allclose(st[:, 4], (u * u).sum(axis=1), rtol=1e-13)
    assert np.allclose(st[:, 5], ((u - f) ** 2).sum(axis=1), rtol=1e-12) and np.allclose(st[:, 6], (f * f).sum(axis=1), rtol=1e-13)
    st2 = _lib.row_stats(U[:, :999]).cpu().numpy()  # strided rows, no reference block
    assert np.array_equal(st2[:, 1], u[:, :999].max(axis=1)) and np.all(st2[:, 5] == 0.0)
    assert torch.equal(_lib.row_stats(U, F), _lib.row_stats(U, F))
