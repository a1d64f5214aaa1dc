"""PXDMF I/O (pgdrome/model.py:202-575), Format="XML" variant.

Pinning: tests/golden/pxdmf/PGDsolution.pxdmf was written by the product writer and read back by the
UNMODIFIED reference loader (tests/golden/make_golden.py::pxdmf); pxdmf.npz holds what the reference stored
and what the reference's own evaluate() returned from it.  CPU tests check the product loader against that
bit for bit plus the size-independent round-trip / idempotence properties; the gpu test evaluates the loaded
model on the device."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURE = os.path.join(GOLD, "pxdmf", "PGDsolution.pxdmf")


def _gold():
    return np.load(os.path.join(GOLD, "pxdmf.npz"))


def _load(path=FIXTURE):
    from pgdrome_b200.model import PGD

    return PGD().load_pxdmf(path)


def test_loader_matches_reference_loader():
    g = _gold()
    pgd = _load()
    assert pgd.name == str(g["name"]) == "PGDsolution.pxdmf"  # the Domain name keeps the suffix (model.py:447)
    assert pgd.numModes == pgd.used_numModes == int(g["num_modes"]) == 3
    assert len(pgd.mesh) == 3
    for d, m in enumerate(pgd.mesh):
        assert m.name == str(g["g%d_name" % d])
        assert np.array_equal(np.array(m.info), g["g%d_info" % d])
        assert m.meshdim == int(g["g%d_meshdim" % d])
        assert m.numElements == int(g["g%d_num_elements" % d])
        assert m.typElements == str(g["g%d_typ_elements" % d])
        assert np.array_equal(m.topology, g["g%d_topology" % d])
        assert m.numNodes == int(g["g%d_num_nodes" % d])
        assert np.array_equal(m.dataX, g["g%d_x" % d]) and np.array_equal(m.dataY, g["g%d_y" % d])
        assert len(m.attributes) == 2
        for a, att in enumerate(m.attributes):
            assert [att.name, att._type, att.field] == list(g["g%d_a%d_meta" % (d, a)])
            assert np.array_equal(np.array(att.data), g["g%d_a%d_data" % (d, a)])  # bit-exact, 17 digits


def test_write_load_roundtrip_and_idempotence(tmp_path):
    pgd = _load()
    pgd.name = "again"
    p1 = pgd.write_pxdmf(str(tmp_path / "a"))
    back = _load(p1)
    assert back.name == "again.pxdmf" and back.numModes == pgd.numModes
    for m, n in zip(pgd.mesh, back.mesh):
        assert np.array_equal(m.topology, n.topology)
        for u, v in zip((m.dataX, m.dataY), (n.dataX, n.dataY)):
            assert np.array_equal(u, v)
        for a, b in zip(m.attributes, n.attributes):
            assert (a.name, a._type, a.field) == (b.name, b._type, b.field)
            assert all(np.array_equal(x, y) for x, y in zip(a.data, b.data))
    back.name = "again"
    p2 = back.write_pxdmf(str(tmp_path / "b"))
    assert open(p1).read() == open(p2).read()  # write(load(write(x))) == write(x)


def test_empty_modes_and_unknown_formats(tmp_path):
    from pgdrome_b200.model import PGD

    txt = open(FIXTURE).read()
    bad = tmp_path / "hdf.pxdmf"
    bad.write_text(txt.replace('Format = "XML">\n0 1 6', 'Format = "HDF">PGD1.h5:/Mesh/0/mesh/topology\n0 1 6', 1))
    try:
        import h5py  # noqa: F401

        expect = (OSError, FileNotFoundError)
    except ImportError:
        expect = ImportError
    with pytest.raises(expect):
        PGD().load_pxdmf(str(bad))
    odd = tmp_path / "odd.pxdmf"
    odd.write_text(txt.replace('Format = "XML"', 'Format = "Binary"', 1))
    with pytest.raises(ValueError):
        PGD().load_pxdmf(str(odd))


@pytest.mark.skipif(not os.path.isdir("/root/reference/pgdrome"), reason="reference sources not on this machine")
def test_reference_loader_reads_product_file(tmp_path):
    """Only where /root/reference exists (this container): the unmodified reference loader, under the NumPy dolfin
    stub, reads a file the product just wrote -- run in a subprocess so the stub never leaks into this session."""
    import subprocess
    import sys

    pgd = _load()
    pgd.name = "fresh"
    path = pgd.write_pxdmf(str(tmp_path))
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import make_golden as mg\n"
        "mg._install_stub()\n"
        "from pgdrome.model import PGD\n"
        "p = PGD().load_pxdmf(%r)\n"
        "g = np.load(%r)\n"
        "assert p.numModes == 3 and len(p.mesh) == 3\n"
        "for d, m in enumerate(p.mesh):\n"
        "    assert np.array_equal(m.dataX, g['g%%d_x' %% d])\n"
        "    for a, att in enumerate(m.attributes):\n"
        "        assert np.array_equal(np.array(att.data), g['g%%d_a%%d_data' %% (d, a)])\n"
        "print('ok')\n" % (GOLD, path, os.path.join(GOLD, "pxdmf.npz")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def _evaluate_checks(pgd):
    g = _gold()
    for at in (0, 1):
        for d in (1, 2):
            pgd.mesh[d].attributes[at].interpolationInfo = {"name": 0, "kind": "linear"}
        pgd.create_interpolation_fcts([1, 2], at)
        ref = g["eval_%d" % at]
        scale = np.abs(ref).max()
        for pt, r in zip(g["points"], ref):
            u = np.asarray(pgd.evaluate(0, [1, 2], list(pt), at))
            assert u.shape == r.shape
            assert np.abs(u - r).max() <= 1e-13 * scale
        U = pgd.evaluate_batch(0, [1, 2], g["points"], at)
        U = U.cpu().numpy() if hasattr(U, "cpu") else np.asarray(U)
        assert np.abs(U.reshape(ref.shape) - ref).max() <= 1e-13 * scale
    with pytest.raises(ValueError):
        pgd.evaluate(0, [1, 2], [2.5, 1.0], 0)  # outside PGD2's range


def test_loaded_model_evaluates_host_logic(monkeypatch):
    """Host logic of load -> create_interpolation_fcts -> evaluate against the reference's values, with the NumPy
    ABI stand-in in place of the device library (no GPU here)."""
    from tests import cpu_abi

    cpu_abi.install(monkeypatch)
    _evaluate_checks(_load())


@pytest.mark.gpu
def test_loaded_model_evaluates_on_device():
    _evaluate_checks(_load())


@pytest.mark.gpu
def test_solved_problem_written_loaded_evaluated(tmp_path):
    """solve_PGD on the device -> return_PGD -> write_pxdmf -> load_pxdmf -> evaluate: P1 modes written as vertex
    data and re-interpolated linearly reproduce the Lagrange-space evaluation (model.py:788-803 vs 822-842)."""
    from pgdrome_b200 import configs

    p = configs.heat2d_tk(n=12, nt=30, nk=8, PGD_nmax=3)
    p.solve_PGD(_problem="linear")
    pgd = p.return_PGD()
    path = pgd.write_pxdmf(str(tmp_path))
    back = _load(path)
    assert back.numModes == pgd.numModes and [m.numNodes for m in back.mesh] == [m.numNodes for m in pgd.mesh]
    for d in (1, 2):
        back.mesh[d].attributes[0].interpolationInfo = {"name": 0, "kind": "linear"}
    back.create_interpolation_fcts([1, 2], 0)
    t = pgd.mesh[1].dataX
    k = pgd.mesh[2].dataX
    for pt in ([float(t[3]), float(k[2])], [0.5 * (t.min() + t.max()), 0.3 * k.min() + 0.7 * k.max()], [t.max(), k.min()]):
        u_ref = pgd.evaluate(0, [1, 2], pt, 0).compute_vertex_values()
        u = np.asarray(back.evaluate(0, [1, 2], pt, 0))  # vertex data [numNodes, meshdim], values in column 0 (model.py:1510-1556)
        assert u.shape == (pgd.mesh[0].numNodes, 2) and not u[:, 1].any()
        u = u[:, 0]
        assert np.abs(u - u_ref).max() <= 1e-12 * max(np.abs(u_ref).max(), 1e-300)
