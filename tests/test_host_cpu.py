"""Host logic on the CPU (-m "not gpu"): the user-callback path of PGDProblem / PGD.evaluate with
the C-ABI wrappers replaced by the NumPy stand-in of tests/cpu_abi.py, checked against the oracle.
The kernels themselves are checked by the -m gpu tests."""
import numpy as np
import pytest

from oracle import fem as ofem
from oracle import pgd as opgd
from oracle import problems as oprob
from tests import cpu_abi


@pytest.fixture
def cpu(monkeypatch):
    import pgdrome_b200.forms  # noqa: F401
    import pgdrome_b200.model  # noqa: F401
    import pgdrome_b200.solver  # noqa: F401

    cpu_abi.install(monkeypatch)
    from pgdrome_b200 import lazy

    lazy._pending.clear()
    yield


def _ospaces(p):
    return [ofem.Space(v.mesh().coordinates(), v.mesh().cells(), v.degree, v.bs) for v in p.V]


def _mode_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b)) / np.linalg.norm(b)


def _compare(p, o, tol=1e-8):
    assert p.PGD_modes == o.PGD_modes
    assert p.num_fp_it == o.num_fp_it
    for d in range(len(p.V)):
        for k in range(p.PGD_modes):
            assert _mode_err(p.PGD_func[d][k].vector()[:], o.PGD_func[d][k]) < tol, (d, k)
    assert np.allclose(p.amplitude, o.amplitude, rtol=1e-8, atol=0)


def test_poisson1d_k_matches_oracle(cpu):
    from pgdrome_b200 import configs

    p = configs.poisson1d_k(nx=60, nk=12, PGD_nmax=4)
    p.solve_PGD(_problem="linear")
    o, _ = oprob.poisson1d_k(nx=60, nk=12, PGD_nmax=4, spaces=_ospaces(p))
    opgd.solve_pgd(o)
    _compare(p, o)
    # Newton path == linear path (tests/integration/test_solver_problem.py:748-752)
    q = configs.poisson1d_k(nx=60, nk=12, PGD_nmax=4)
    q.solve_PGD()
    assert np.allclose(q.amplitude, p.amplitude, rtol=1e-10, atol=0)


def test_heat2d_tk_matches_oracle(cpu):
    from pgdrome_b200 import configs

    p = configs.heat2d_tk(n=8, nt=12, nk=5, PGD_nmax=3)
    p.solve_PGD(_problem="linear")
    o, _ = oprob.heat2d_tk(n=8, nt=12, nk=5, PGD_nmax=3, spaces=_ospaces(p))
    opgd.solve_pgd(o)
    _compare(p, o)


def test_elasticity3d_matches_oracle(cpu):
    """configs[2] reduced: vector P1 tetrahedra, two-material Voigt stiffness (degree-0 indicator weights),
    facet traction, clamped face."""
    from pgdrome_b200 import configs

    p = configs.elasticity3d(n=3, nE=6, nF=2, PGD_nmax=3)
    p.solve_PGD(_problem="linear")
    o, _ = oprob.elasticity3d(n=3, nE=6, nF=2, PGD_nmax=3, spaces=_ospaces(p))
    opgd.solve_pgd(o)
    _compare(p, o)
    assert p.PGD_modes == 3


def test_thermal3d_matches_oracle(cpu):
    """configs[3] reduced: P1 tetrahedra x FD time x P x v; the moving source separated into 3 terms (the oracle gets
    the same separated tables) and the way-point surrogate."""
    from pgdrome_b200 import configs

    p = configs.thermal3d(n=4, nt=12, nP=3, nv=3, n_src=3, PGD_nmax=3)
    assert p.source_terms is not None and len(p.source_terms["G"]) == 3 and np.all(np.diff(p.source_terms["rel_err"]) < 0)
    p.solve_PGD(_problem="linear")
    o, _ = oprob.thermal3d(n=4, nt=12, nP=3, nv=3, n_src=3, PGD_nmax=3, spaces=_ospaces(p), source_terms=p.source_terms)
    opgd.solve_pgd(o)
    _compare(p, o)
    q = configs.thermal3d(n=4, nt=12, nP=3, nv=3, n_src=3, PGD_nmax=2, source="waypoints")
    q.solve_PGD(_problem="linear")
    o2, _ = oprob.thermal3d(n=4, nt=12, nP=3, nv=3, n_src=3, PGD_nmax=2, spaces=_ospaces(q))
    opgd.solve_pgd(o2)
    _compare(q, o2)


# ---- host logic against the reference-generated golden vectors (same bodies as tests/test_gpu_golden.py,
# with the NumPy ABI stand-in): FD-mode enrichment loop, normalisations, stopping criteria, evaluate
def test_golden_laplace_fd_host_logic(cpu):
    from tests import test_gpu_golden as tg

    tg.test_laplace_fd_reference_test_on_device()


@pytest.mark.parametrize("key,opts", [
    ("v_stiff", dict(norm_modes="stiff", PGD_nmax=6)),
    ("v_l2", dict(norm_modes="l2", PGD_nmax=6)),
    ("v_no", dict(norm_modes="no", PGD_nmax=4)),
    ("v_delta", dict(stop_fp="delta", PGD_nmax=4, tol_fp_it=1e-6)),
])
def test_golden_laplace_fd_variants_host_logic(cpu, key, opts):
    from tests import test_gpu_golden as tg

    tg.test_laplace_fd_variants_on_device(key, opts)


def test_golden_pgdclass_host_logic(cpu):
    from tests import test_gpu_golden as tg

    tg.test_pgdclass_evaluate_on_device()


# ---- solver options on the FEM path (host logic): normalisations, "delta" criterion, randomised start
@pytest.mark.parametrize("opts", [dict(norm_modes="l2"), dict(norm_modes="no"), dict(stop_fp="delta", tol_fp_it=1e-6),
                                  dict(fp_init="randomized")])
def test_solver_options_match_oracle(cpu, opts):
    from pgdrome_b200 import configs

    np.random.seed(7)  # fp_init="randomized" draws np.random.rand in dimension order, like solver.py:191-196
    p = configs.heat2d_tk(n=8, nt=12, nk=5, PGD_nmax=3, **opts)
    p.solve_PGD(_problem="linear")
    o, _ = oprob.heat2d_tk(n=8, nt=12, nk=5, PGD_nmax=3, spaces=_ospaces(p), **opts)
    np.random.seed(7)
    opgd.solve_pgd(o)
    _compare(p, o, tol=1e-7 if opts.get("fp_init") else 1e-8)
    assert np.allclose(p.alpha, o.alpha, rtol=1e-8, atol=0)


def test_unknown_options_raise(cpu):
    from pgdrome_b200 import configs

    p = configs.poisson1d_k(nx=20, nk=4, PGD_nmax=2, stop_fp="chady")
    with pytest.raises(ValueError):  # solver.py:873-879
        p.solve_PGD(_problem="linear")
    q = configs.poisson1d_k(nx=20, nk=4, PGD_nmax=2)
    with pytest.raises(ValueError):
        q.solve_PGD(_problem="quadratic")
    pgd = configs.poisson1d_k(nx=20, nk=4, PGD_nmax=2).solve_PGD(_problem="linear").return_PGD()
    with pytest.raises(ValueError):  # model.py:737-772
        pgd.evaluate(0, [1], [1.0, 2.0], 0)
    with pytest.raises(ValueError):
        pgd.evaluate(0, [1], [1.0], 5)
    with pytest.raises(ValueError):  # outside the parameter range (test_pgdclass.py:319-326)
        pgd.evaluate(0, [1], [9.0], 0)


def test_direct_solve_mode(cpu):
    """solve_modes "direct" (solver.py:909-925): scalar problem b / a broadcast into the dofs."""
    from pgdrome_b200 import dolfin as df
    from pgdrome_b200.solver import PGDProblem

    V = df.FunctionSpace(df.IntervalMesh(4, 0.0, 1.0), "P", 1)
    p = PGDProblem(name="d", name_coord=["X"], modes_info=["U", "Node", "Scalar"], Vs=[V], bc_fct=lambda Vs, dom, param: [0],
                   load=[], param={}, rhs_fct=None, lhs_fct=None, probs=["r"], PGD_nmax=1)
    f = p.direct_solve(4.0, 2.0, 0)
    assert np.array_equal(f.vector()[:], np.full(V.n_dofs, 0.5))


def test_parked_heap_restores_collector_state():
    """The enrichment step parks the existing heap in the collector's permanent generation and puts it back."""
    import gc

    from pgdrome_b200.solver import _ParkedHeap

    before = gc.get_freeze_count()
    with _ParkedHeap(True) as outer:
        assert outer.mine and gc.get_freeze_count() > before
        with _ParkedHeap(True) as inner:  # nested steps (normalisation callbacks) do not freeze twice
            assert not inner.mine
        assert gc.get_freeze_count() > before
    after = gc.get_freeze_count()
    assert after <= before and _ParkedHeap.depth == 0
    with _ParkedHeap(False) as off:
        assert not off.mine and gc.get_freeze_count() == after
    gc.freeze()  # a heap frozen by the caller is left alone
    try:
        n = gc.get_freeze_count()
        with _ParkedHeap(True) as p:
            assert not p.mine
        assert gc.get_freeze_count() == n
    finally:
        gc.unfreeze()
        _ParkedHeap.base = gc.get_freeze_count()


@pytest.mark.parametrize("make", [
    lambda c: c.heat2d_tk(n=10, nt=20, nk=6, PGD_nmax=2, PGD_tol=0.0),
    lambda c: c.elasticity3d(n=3, nE=5, nF=2, PGD_nmax=2),
    lambda c: c.thermal3d(n=4, nt=10, nP=3, nv=3, n_src=2, PGD_nmax=2),
    lambda c: c.poisson1d_k(nx=40, nk=10, PGD_nmax=2),
])
def test_dropped_problem_is_freed_by_reference_counting(cpu, make):
    """No reference cycles through the package's objects: dropping a solved problem (and its PGD model) releases its
    spaces, device arrays and modes immediately -- device memory must not wait for the cycle collector."""
    import gc
    import weakref

    from pgdrome_b200 import configs

    gc.collect()
    gc.disable()
    try:
        def run():
            p = make(configs)
            p.solve_PGD(_problem="linear")
            pgd = p.return_PGD()
            pgd.evaluate_batch(0, list(range(1, len(p.V))), [[float(v.mesh().coordinates()[1, 0]) for v in p.V[1:]]], 0)
            return [weakref.ref(o) for o in (p, pgd, p.V[0], p.V[0]._dev["device_space"], p.PGD_func[0][0])]

        refs = run()
        assert [r() is None for r in refs] == [True] * 5
        gc.set_debug(gc.DEBUG_SAVEALL)
        gc.collect()
        mine = [o for o in gc.garbage if type(o).__module__.startswith("pgdrome_b200")]
        assert not mine, sorted({type(o).__name__ for o in mine})
    finally:
        gc.set_debug(0)
        gc.garbage.clear()
        gc.enable()


def test_device_coefficients_equal_host_coefficients(cpu):
    """The separated form's coefficients evaluated by scalar programs (device path, here through the NumPy stand-in) give
    bitwise the modes of the host-float path; unsupported scalar operations fall back to the host path."""
    from pgdrome_b200 import configs, forms, lazy
    from pgdrome_b200 import dolfin as df

    runs = []
    for dc in (True, False):
        p = configs.heat2d_tk(n=8, nt=16, nk=6, PGD_nmax=3, PGD_tol=0.0)
        p.solve_PGD(_problem="linear", settings={"device_coefficients": dc, "linear_solver": "mumps"})
        runs.append(p)
    a, b = runs
    assert a.num_fp_it == b.num_fp_it and a.PGD_modes == b.PGD_modes == 3
    for d in range(3):
        for k in range(3):
            assert np.array_equal(a.PGD_func[d][k].vector().get_local(), b.PGD_func[d][k].vector().get_local())
    # fewer host synchronisations on the device path: values are read back once per sweep, not once per sub-problem
    V = df.FunctionSpace(df.UnitIntervalMesh(7), "CG", 1)
    f = df.interpolate(df.Expression("1.0+x[0]", degree=1), V)
    u, v = df.TrialFunction(V), df.TestFunction(V)
    forms.device_coefficients[0] = True
    try:
        ok = forms.compile_form(df.Constant(df.assemble(f * f * df.dx) * 2.0) * u * v * df.dx)
        c_dev = forms._coefs_dev(ok)
        assert c_dev is not None and abs(float(c_dev[0]) - 2.0 * float(df.assemble(f * f * df.dx))) < 1e-15
        unsupported = forms.compile_form(df.Constant(df.assemble(f * f * df.dx).sqrt()) * u * v * df.dx)
        assert forms._coefs_dev(unsupported) is None  # sqrt is evaluated on the host
        A = forms.assemble_matrix(unsupported)  # ... and the assembly still works through the host path
        assert np.isfinite(A.array()).all()
        lazy.fetch()
    finally:
        forms.device_coefficients[0] = False


def test_scalar_pool_wraps_around(cpu, monkeypatch):
    """A pool far smaller than a step's functionals: when it is full the launched values are read back and the pool
    is reused -- same modes, bit for bit, as with the large pool."""
    from pgdrome_b200 import configs, forms, lazy

    def run():
        p = configs.heat2d_tk(n=8, nt=16, nk=6, PGD_nmax=3, PGD_tol=0.0)
        p.solve_PGD(_problem="linear")
        return p

    ref = run()
    f0 = lazy.stats["fetches"]
    monkeypatch.setattr(forms, "POOL_DOUBLES", 24)
    monkeypatch.setattr(forms, "_pool", {})
    small = run()
    assert lazy.stats["fetches"] - f0 > 0
    assert forms._pool and next(iter(forms._pool.values()))[0].numel() == 24
    assert small.num_fp_it == ref.num_fp_it
    for d in range(3):
        for k in range(3):
            assert np.array_equal(small.PGD_func[d][k].vector().get_local(), ref.PGD_func[d][k].vector().get_local())


def test_negated_operator_shares_the_atom_and_long_forms_take_the_fast_path(cpu):
    """`-Constant(c) * op(G, v)` (every old-mode term of a right-hand side) must reuse the assembled atom of `op`:
    the sign goes into the coefficient, not into a second operator -K.  Functionals of long forms (a Voigt elasticity
    integrand expands to 33 monomials) take the memoised fast path and agree with the general path."""
    import pgdrome_b200.dolfin as df
    from pgdrome_b200 import forms, lazy
    from pgdrome_b200.assembly import device_space

    m = df.UnitCubeMesh(3, 3, 3)
    V = df.VectorFunctionSpace(m, "P", 1)
    rng = np.random.default_rng(5)
    f, g = df.Function(V), df.Function(V)
    f.vector()[:] = rng.uniform(-1, 1, V.n_dofs)
    g.vector()[:] = rng.uniform(-1, 1, V.n_dofs)
    v = df.TestFunction(V)
    C = np.zeros((6, 6))
    C[:3, :3] = 0.6
    C[np.arange(3), np.arange(3)] += 0.8
    C[np.arange(3, 6), np.arange(3, 6)] = 0.4
    Cm = df.as_matrix(C)
    chi = df.Expression("x[0] < 0.5 ? 1.0 : 0.25", degree=0)

    def eps(w):
        return df.as_vector([w[0].dx(0), w[1].dx(1), w[2].dx(2), w[1].dx(2) + w[2].dx(1), w[0].dx(2) + w[2].dx(0),
                             w[0].dx(1) + w[1].dx(0)])

    k = lambda a, b: chi * df.inner(Cm * eps(a), eps(b)) * df.dx(m)  # noqa: E731
    ds = device_space(V)
    plus = df.assemble(k(f, v)).tensor().clone()
    n_atoms = len(ds.atoms)
    minus = df.assemble(-k(f, v)).tensor().clone()
    assert len(ds.atoms) == n_atoms, "the negated form assembled a second atom"
    assert np.array_equal(minus.numpy(), -plus.numpy())
    # rank 0: fast path (one integral, <= 64 monomials) against the general path (two integrals of the same form)
    fast = float(df.assemble(k(f, g)))
    fast_neg = float(df.assemble(-1.0 * k(f, g)))
    general = float(df.assemble(k(f, g) + k(g, f)))
    assert len(ds.atoms) == n_atoms
    assert fast_neg == -fast
    assert abs(general - 2.0 * fast) <= 1e-12 * abs(general)
    A = ofem.assemble_bilinear(ofem.Space(m.coordinates(), m.cells(), 1, 3), ofem.T_voigt(C, 3),
                               weight=lambda x: np.where(x[..., 0] < 0.5, 1.0, 0.25), weight_degree=0)
    ref = float(g.vector()[:] @ (A @ f.vector()[:]))
    assert abs(fast - ref) <= 1e-11 * abs(ref)
    lazy._pending.clear()
    forms.functional_memo[0] = None


def test_capture_cache_is_scoped_to_the_enrichment_step(cpu):
    """Inside an enrichment step `v[i]` of a vector Function is ONE shared expression that remembers its derivatives;
    outside a step nothing is cached (no leaf -> expression -> leaf cycle survives the step)."""
    import pgdrome_b200.dolfin as df
    from pgdrome_b200 import ufl

    V = df.VectorFunctionSpace(df.UnitSquareMesh(2, 2), "P", 1)
    f = df.Function(V)
    assert ufl.capture_cache[0] is None
    assert f[0] is not f[0] and f[0]._dx is None
    ufl.capture_cache[0] = {}
    try:
        a, b = f[0], f[0]
        assert a is b and f[1] is not a
        assert a.dx(1) is b.dx(1) and a.dx(0) is not a.dx(1)
        (m,) = a.dx(1).comps[0]
        assert m.factors == (ufl.Factor(f, 0, 1),)
        g = df.Function(V)
        assert g[0] is not a  # another leaf, another entry
    finally:
        ufl.capture_cache[0] = None
    p = __import__("pgdrome_b200.configs", fromlist=["x"]).poisson1d_k(nx=20, nk=5, PGD_nmax=2)
    p.solve_PGD(_problem="linear")
    assert ufl.capture_cache[0] is None


def test_vector_space_pattern_from_the_node_pattern(cpu):
    """Vector spaces build the pattern of their NODES and expand it: row pointers, columns and the (lazily built) gather
    lists must equal the dof-level build bit for bit, the block-column list must equal the plan derived from the CSR
    arrays -- replicated spaces and the local spaces of an element partition alike."""
    from pgdrome_b200 import fem, sharding
    from pgdrome_b200.assembly import ShardedDeviceSpace, device_space

    cases = [(fem.UnitCubeMesh(3, 4, 2), 1, 3), (fem.UnitSquareMesh(5, 4), 2, 2), (fem.UnitSquareMesh(5, 4), 1, 2),
             (fem.IntervalMesh(7, 0.0, 1.0), 1, 2), (fem.UnitCubeMesh(2, 2, 2), 1, 2)]
    for mesh, degree, bs in cases:
        V = fem.FunctionSpace(mesh, "P", degree, bs)
        ds = device_space(V)
        rowptr, colidx, gptr, gidx = ds.pattern
        rp, ci = ofem.sparsity(V.cell_dofs, V.n_dofs)
        assert np.array_equal(rowptr.numpy(), rp) and np.array_equal(colidx.numpy(), ci)
        _, _, g2, i2 = cpu_abi.pattern_build(ds.cell_dofs, V.n_dofs)
        assert np.array_equal(gptr.numpy(), g2.numpy()) and np.array_equal(gidx.numpy(), i2.numpy())
        ref = cpu_abi.bsr_plan(rowptr, colidx, bs)
        assert np.array_equal(ds.bsr[0].numpy(), ref[0].numpy()) and ds.bsr[1] == ref[1]
    for world in (2, 3):
        for rank in range(world):
            V = fem.FunctionSpace(fem.UnitCubeMesh(5, 4, 3), "P", 1, 3)
            sh = sharding.SpaceShard(V, rank, world)
            ds = ShardedDeviceSpace(V, sh)
            rp, ci = ofem.sparsity(sh.local.cell_dofs, sh.local.n_dofs)
            assert np.array_equal(ds.pattern[0].numpy(), rp) and np.array_equal(ds.pattern[1].numpy(), ci)
            ref = cpu_abi.bsr_plan(ds.rowptr_owned, ds.pattern[1][: ds.nnz_owned], 3)
            assert np.array_equal(ds.bsr[0].numpy(), ref[0].numpy()) and ds.bsr[1] == ref[1], (world, rank)
            assert ds.node_plan is not False
