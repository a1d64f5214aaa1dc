"""Sensor evaluation: PGD.eval_fixed_modes / evaluate_sensor_response
(pgdrome/model.py:107-130, 862-953).

Pinning: tests/golden/sensor.npz comes from the UNMODIFIED reference method run with the Probes result injected
into its own cache (tests/golden/make_golden.py::sensor) -> combination + return shapes are pinned.  The Probes
part itself (point location, basis evaluation; fenicstools is not in /root/reference) is checked against the
oracle restatement and against exact polynomial fields -- parity unpinned, said so in oracle/evaluate.py."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold():
    return np.load(os.path.join(GOLD, "sensor.npz"))


# ----------------------------------------------------------------------------- oracle (CPU)
def test_oracle_sensor_response_matches_reference_golden():
    from oracle.evaluate import evaluate_interp1d, sensor_response

    g = _gold()
    px = np.load(os.path.join(GOLD, "pxdmf.npz"))
    free_x = [px["g1_x"], px["g2_x"]]
    free_data = [px["g1_a0_data"][:, :, 0], px["g2_a0_data"][:, :, 0]]

    def factors(c):
        return [[float(evaluate_interp1d([np.ones(1)], [free_x[i]], [[free_data[i][k]]], [c[i]])[0]) for k in range(3)]
                for i in range(2)]

    for tag in ("scalar", "vector"):
        for c, r3, r2 in zip(g["coords"], g["resp_" + tag], g["resp2_" + tag]):
            assert np.allclose(sensor_response(g["E_" + tag], factors(c), 3, 3), r3, rtol=1e-13, atol=1e-13 * np.abs(r3).max())
            assert np.allclose(sensor_response(g["E_" + tag], factors(c), 3, 2), r2, rtol=1e-13, atol=1e-13 * np.abs(r2).max())
    fx = [g["one_x1"], g["one_x2"]]
    fd = [g["one_data1"][:, :, 0], g["one_data2"][:, :, 0]]
    fac = [[float(evaluate_interp1d([np.ones(1)], [fx[i]], [[fd[i][0]]], [g["one_coord"][i]])[0])] for i in range(2)]
    for tag in ("scalar", "vector"):
        r = sensor_response(g["one_E_" + tag], fac, 1, 1)
        assert r.shape == g["one_resp_" + tag].shape
        assert np.allclose(r, g["one_resp_" + tag], rtol=1e-13, atol=0)


def _points(rng, lo, hi, n, coords, cells):
    """generic interior points + mesh vertices + edge midpoints + cell centroids (several cells contain them)"""
    g = coords.shape[1]
    P = [lo + (hi - lo) * rng.random((n, g)), coords[rng.integers(0, len(coords), 5)]]
    c = cells[rng.integers(0, len(cells), 5)]
    P.append(0.5 * (coords[c[:, 0]] + coords[c[:, 1]]))
    P.append(coords[c].mean(axis=1))
    return np.concatenate(P)


def test_oracle_probe_exact_fields():
    from oracle import meshes
    from oracle.fem import Space
    from oracle.evaluate import locate_points, probe_modes

    rng = np.random.default_rng(0)
    co, ce = meshes.box_mesh(0, 0, 0, 2, 1, 1, 4, 3, 2)
    for deg, bs in ((1, 1), (2, 1), (1, 3)):
        sp = Space(co, ce, degree=deg, bs=bs)
        X = sp.node_coords
        if deg == 1:
            f = lambda x: 1.0 + 2 * x[:, 0] - 3 * x[:, 1] + 0.5 * x[:, 2]
        else:
            f = lambda x: 1.0 + x[:, 0] * x[:, 1] - x[:, 2] ** 2 + 2 * x[:, 0]
        modes = []
        for k in range(2):
            v = np.stack([(k + 1) * f(X) * (c + 1) for c in range(bs)], axis=1).reshape(-1)
            modes.append(v)
        P = _points(rng, np.zeros(3), np.array([2.0, 1.0, 1.0]), 20, co, ce)
        E = probe_modes(sp, modes, P)
        for k in range(2):
            ex = (k + 1) * f(P)
            if bs == 1:
                assert np.allclose(E[:, k], ex, atol=1e-12)
            else:
                for c in range(bs):
                    assert np.allclose(E[:, c, k], ex * (c + 1), atol=1e-12)
    cell, bary = locate_points(co, ce, np.array([[2.5, 0.5, 0.5], [1.0, 0.5, 0.5]]))
    assert cell[0] == -1 and cell[1] >= 0 and np.allclose(bary[0], 0) and abs(bary[1].sum() - 1) < 1e-14


# ----------------------------------------------------------------------------- product
def _product_case(kind):
    from pgdrome_b200 import dolfin as df

    if kind == "2d_p1":
        mesh = df.RectangleMesh(df.Point(0, 0), df.Point(2, 1), 9, 7)
        V = df.FunctionSpace(mesh, "CG", 1)
    elif kind == "2d_p2":
        mesh = df.UnitSquareMesh(6, 5, "crossed")
        V = df.FunctionSpace(mesh, "CG", 2)
    elif kind == "3d_p1":
        mesh = df.BoxMesh(df.Point(0, 0, 0), df.Point(2, 1, 1), 6, 5, 4)
        V = df.FunctionSpace(mesh, "CG", 1)
    elif kind == "3d_vec":
        mesh = df.UnitCubeMesh(5, 4, 3)
        V = df.VectorFunctionSpace(mesh, "CG", 1)
    else:
        mesh = df.IntervalMesh(17, -1.0, 3.0)
        V = df.FunctionSpace(mesh, "CG", 2)
    return mesh, V


def _build_pgd(V, R, rng):
    """A PGD object as return_PGD() builds it: fixed dimension on V, two 1-D P1 free dimensions."""
    from pgdrome_b200 import dolfin as df
    from pgdrome_b200.model import PGD

    vs = [V, df.FunctionSpace(df.IntervalMesh(11, 0.0, 2.0), "CG", 1), df.FunctionSpace(df.IntervalMesh(7, 1.0, 3.0), "CG", 1)]
    modes = []
    for v in vs:
        ms = []
        for k in range(R):
            f = df.Function(v)
            f.vector()[:] = rng.standard_normal(v.dim())
            ms.append(f)
        modes.append(ms)
    pgd = PGD(name="sens", n_modes=R, fmeshes=[v.mesh() for v in vs], pgd_modes=modes, name_coord=["X", "E", "F"],
              modes_info=["U", "Node", "Scalar"])
    return pgd, vs, modes


def _sensor_checks(kind, R):
    from oracle.evaluate import probe_modes
    from oracle.fem import Space

    rng = np.random.default_rng(3)
    mesh, V = _product_case(kind)
    pgd, vs, modes = _build_pgd(V, R, rng)
    co, ce = mesh.coordinates(), mesh.cells()
    P = _points(rng, co.min(axis=0), co.max(axis=0), 12, co, ce)
    sp = Space(co, ce, degree=V.degree, bs=V.bs)
    # the oracle numbers its nodes independently: compare through point values, which are numbering-free
    E = pgd.eval_fixed_modes(P, 0, 0)
    exp_shape = (len(P),) + ((V.bs,) if V.bs > 1 else ()) + ((R,) if R > 1 else ())
    assert E.shape == exp_shape
    for k in range(R):
        vals = np.array([np.atleast_1d(modes[0][k](p)) for p in P])  # product host point evaluation (functions.py)
        Ek = E[..., k] if R > 1 else E
        assert np.allclose(Ek.reshape(len(P), -1), vals, rtol=0, atol=1e-12 * np.abs(vals).max())
    # oracle: same dof values transplanted through node coordinates
    order_o = np.lexsort(sp.node_coords.T[::-1])
    order_p = np.lexsort(V.node_coords.T[::-1])
    perm = np.empty(V.n_nodes, dtype=np.int64)
    perm[order_o] = order_p  # oracle node -> product node
    omodes = [modes[0][k].vector()[:].reshape(V.n_nodes, V.bs)[perm].reshape(-1) for k in range(R)]
    Eo = probe_modes(sp, omodes, P)
    Eo = Eo[..., 0] if R == 1 else Eo
    assert np.allclose(E, Eo, rtol=0, atol=1e-12 * np.abs(Eo).max())
    # response = sum_k E[..., k] * prod_i phi_ik(coord_i)   (model.py:904-953)
    coord = [0.731, 2.25]
    w = np.ones(R)
    for i in range(2):
        # oracle 1-D numbering differs as well: evaluate the product's own mode function on the host
        w *= np.array([modes[1 + i][k](coord[i]) for k in range(R)])
    resp = pgd.evaluate_sensor_response(0, [1, 2], coord, 0, P)
    Efull = E[..., None] if R == 1 else E
    ref = np.sum(Efull * w, axis=-1)
    assert resp.shape == ref.shape
    assert np.allclose(resp, ref, rtol=0, atol=1e-12 * np.abs(ref).max())
    # cache: the same sensor set is not located again; a permuted set with the same coordinate sum is not aliased
    assert pgd.eval_fixed_modes(P, 0, 0) is E
    E2 = pgd.eval_fixed_modes(P[::-1].copy(), 0, 0)
    assert np.allclose(E2, E[::-1], rtol=0, atol=1e-13 * np.abs(E).max())
    # truncated evaluation
    if R > 1:
        pgd.used_numModes = R - 1
        r2 = pgd.evaluate_sensor_response(0, [1, 2], coord, 0, P[::-1].copy())
        assert np.allclose(r2, np.sum(E2[..., :R - 1] * w[:R - 1], axis=-1), rtol=0, atol=1e-12 * np.abs(ref).max())
        pgd.used_numModes = R
    outside = P.copy()
    outside[0] = co.max(axis=0) + 1.0
    with pytest.raises(RuntimeError):
        pgd.eval_fixed_modes(outside, 0, 0)
    with pytest.raises(ValueError):
        pgd.evaluate_sensor_response(0, [1, 2], [0.5], 0, P)


@pytest.mark.parametrize("kind,R", [("2d_p1", 3), ("1d_p2", 2), ("3d_vec", 1)])
def test_sensor_host_logic(monkeypatch, kind, R):
    from tests import cpu_abi

    cpu_abi.install(monkeypatch)
    _sensor_checks(kind, R)


def test_injected_probe_cache_reproduces_reference(monkeypatch):
    """The reference's own numbers: load the PXDMF fixture, inject E under the reference's cache key, evaluate."""
    from tests import cpu_abi

    cpu_abi.install(monkeypatch)
    _golden_checks()


def _golden_checks():
    from pgdrome_b200.model import PGD

    g = _gold()
    pgd = PGD().load_pxdmf(os.path.join(GOLD, "pxdmf", "PGDsolution.pxdmf"))
    for d in (1, 2):
        pgd.mesh[d].attributes[0].interpolationInfo = {"name": 0, "kind": "linear"}
    pgd.create_interpolation_fcts([1, 2], 0)
    pts = g["points"]
    key = (float(np.sum(pts.flatten())), 0, 0)
    for tag in ("scalar", "vector"):
        pgd._eval_fixed_modes = {key: g["E_" + tag]}
        pgd.invalidate_device_cache()
        for used, name in ((3, "resp_"), (2, "resp2_")):
            pgd.used_numModes = used
            for c, ref in zip(g["coords"], g[name + tag]):
                r = pgd.evaluate_sensor_response(0, [1, 2], list(c), 0, pts)
                assert r.shape == ref.shape
                assert np.allclose(r, ref, rtol=0, atol=1e-13 * np.abs(g[name + tag]).max())
        pgd.used_numModes = 3


# ----------------------------------------------------------------------------- device
@pytest.mark.gpu
@pytest.mark.parametrize("kind,R", [("2d_p1", 3), ("2d_p2", 4), ("3d_p1", 5), ("3d_vec", 2), ("3d_vec", 1), ("1d_p2", 2)])
def test_sensor_on_device(kind, R):
    _sensor_checks(kind, R)


@pytest.mark.gpu
def test_injected_probe_cache_reproduces_reference_on_device():
    _golden_checks()


@pytest.mark.gpu
def test_locate_points_kernel_vs_oracle():
    """pgd_locate_points against the oracle: same winning cell (lowest index) and barycentric coordinates, incl. points on
    vertices / edges / outside, more points than one shared-memory chunk, and an empty point set."""
    import torch

    from oracle import meshes
    from oracle.evaluate import locate_points
    from pgdrome_b200 import _lib

    rng = np.random.default_rng(1)
    for co, ce in (meshes.box_mesh(0, 0, 0, 1, 2, 1, 9, 8, 7), meshes.rectangle_mesh(0, 0, 3, 1, 31, 17, "crossed"),
                   (np.linspace(-1, 1, 41).reshape(-1, 1), np.column_stack([np.arange(40), np.arange(1, 41)]))):
        co = np.asarray(co, dtype=np.float64).reshape(len(co), -1)
        P = _points(rng, co.min(axis=0), co.max(axis=0), 1300, co, ce)
        P[7] = co.max(axis=0) + 0.5  # outside
        P[8] = co.min(axis=0) - 1e-3
        cell, bary = _lib.locate_points(_lib.to_device(co), _lib.to_device(ce.astype(np.int32)), _lib.to_device(P))
        c_ref, b_ref = locate_points(co, ce, P)
        assert np.array_equal(cell.cpu().numpy(), c_ref)
        assert np.abs(bary.cpu().numpy() - b_ref).max() < 1e-12
    cell, bary = _lib.locate_points(_lib.to_device(co), _lib.to_device(ce.astype(np.int32)),
                                    torch.empty((0, 1), dtype=torch.float64, device=cell.device))
    assert cell.numel() == 0 and bary.numel() == 0
