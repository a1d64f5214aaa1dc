"""-m "not gpu": the plain-C pieces of the oracle (oracle/cfem.c: closed-form P1 atom assembly, SpMV, Jacobi /
node-block-Jacobi PCG) against the NumPy restatement they stand in for at large sizes (oracle/fem.py quadrature
assembly, SciPy solves)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import cfem, fem
from oracle.meshes import box_mesh, interval_mesh, rectangle_mesh


def _spaces():
    yield "interval", fem.Space(*interval_mesh(17, 0.0, 2.0)), 1
    yield "square", fem.Space(*rectangle_mesh(0.0, 0.0, 1.0, 2.0, 5, 4)), 1
    yield "square-vector", fem.Space(*rectangle_mesh(0.0, 0.0, 1.0, 1.0, 4, 3), degree=1, bs=2), 2
    yield "box", fem.Space(*box_mesh(0.0, 0.0, 0.0, 1.0, 1.5, 0.5, 3, 4, 2)), 1
    yield "box-vector", fem.Space(*box_mesh(0.0, 0.0, 0.0, 1.0, 1.0, 1.0, 3, 3, 3), degree=1, bs=3), 3


@pytest.mark.parametrize("name,space,bs", list(_spaces()), ids=[s[0] for s in _spaces()])
def test_c_assembly_matches_quadrature_assembly(name, space, bs):
    g = space.gdim
    rng = np.random.default_rng(3)
    old = fem.C_FAST_MIN_CELLS
    try:
        for T, weight in ((fem.T_mass(bs, g), None), (fem.T_stiff(bs, g), None),
                          (rng.uniform(-1, 1, (bs, g + 1, bs, g + 1)), None),  # every slot combination at once
                          (fem.T_stiff(bs, g), lambda x: np.where(x[..., 0] < 0.5, 1.0, 3.0))):
            fem.C_FAST_MIN_CELLS = None
            A = fem.assemble_bilinear(space, T, weight=weight, weight_degree=0 if weight else None)
            fem.C_FAST_MIN_CELLS = 0
            B = fem.assemble_bilinear(space, T, weight=weight, weight_degree=0 if weight else None)
            # same pattern (explicit zeros kept by the C path, dropped by COO->CSR: compare as matrices), values to round-off
            d = abs(A - B).max()
            assert d <= 2e-14 * abs(A).max(), (name, d)
            rp, ci = fem.sparsity(space.cell_dofs, space.n_dofs)
            assert np.array_equal(B.indptr, rp) and np.array_equal(B.indices, ci)  # union of cell cliques, ascending
    finally:
        fem.C_FAST_MIN_CELLS = old


def test_c_voigt_elasticity_atom():
    space = fem.Space(*box_mesh(0.0, 0.0, 0.0, 1.0, 1.0, 1.0, 4, 3, 3), degree=1, bs=3)
    T = fem.T_voigt(fem.isotropic_C(0.6, 0.4, 3), 3)
    old = fem.C_FAST_MIN_CELLS
    try:
        fem.C_FAST_MIN_CELLS = None
        A = fem.assemble_bilinear(space, T)
        fem.C_FAST_MIN_CELLS = 0
        B = fem.assemble_bilinear(space, T)
    finally:
        fem.C_FAST_MIN_CELLS = old
    assert abs(A - B).max() <= 2e-14 * abs(A).max()
    assert abs(B - B.T).max() <= 1e-15 * abs(B).max()


@pytest.mark.parametrize("block", [1, 3])
def test_c_pcg_and_spmv(block):
    space = fem.Space(*box_mesh(0.0, 0.0, 0.0, 1.0, 1.0, 1.0, 5, 5, 5), degree=1, bs=block)
    g = 3
    A = (fem.assemble_bilinear(space, fem.T_stiff(block, g)) + 0.5 * fem.assemble_bilinear(space, fem.T_mass(block, g))).tocsr()
    rng = np.random.default_rng(0)
    xs = rng.uniform(-1, 1, A.shape[0])
    assert np.allclose(cfem.spmv(A, xs), A @ xs, rtol=1e-14, atol=1e-14)
    b = A @ xs
    x, it, rr = cfem.pcg(A, b, block=block, rtol=1e-13)
    assert rr <= 1e-13 and 0 < it < 500
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-10
    ref = spla.spsolve(A.tocsc(), b)
    assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-10
    # iteration cap (bench.py's bounded CPU sample) and warm start
    x2, it2, rr2 = cfem.pcg(A, b, block=block, rtol=1e-13, max_iters=5)
    assert it2 == 5 and rr2 > 1e-13
    x3, it3, _ = cfem.pcg(A, b, block=block, rtol=1e-12, x0=x)
    assert it3 == 0
    assert cfem.threads() >= 1
