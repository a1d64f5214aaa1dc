"""Oracle (test infrastructure): PGD.evaluate and its callers, restated (pgdrome/model.py).

* ``evaluate_interp1d``  model.py:780-803 (interpolationInfo["name"] == 0): vertex-order data,
  scipy ``interp1d`` per mode (model.py:633-639) -> out-of-range coordinate raises ValueError
  (pinned by tests/unit/test_pgdclass.py:319-326).
* ``evaluate_dofs``      model.py:805-860 (default): dof-order vectors, each free-dim mode is a
  finite-element function evaluated at the coordinate (``Function.__call__``) [DOLFIN-knowledge].
* ``sampling_LHS`` / ``evaluate_error``  model.py:1704-1743, 1745-1825.
"""
import numpy as np
from scipy import interpolate
from scipy.stats import qmc

from .fem import tabulate


def evaluate_interp1d(fixed_data, free_x, free_data, coord, used_numModes=None, kind="linear"):
    """fixed_data[k]: [n, ncomp]; free_x[i]: [n_i]; free_data[i][k]: [n_i]; coord[i]: float."""
    R = len(fixed_data) if used_numModes is None else used_numModes
    out = np.zeros(np.asarray(fixed_data[0]).shape)
    for k in range(R):
        tmp = np.copy(fixed_data[k])
        for i in range(len(free_x)):
            fac = interpolate.interp1d(free_x[i], free_data[i][k], kind=kind)(coord[i])
            tmp = tmp * fac
        out += tmp
    return out


def point_eval(space, f, x):
    """Function.__call__ for a scalar/vector Lagrange function on a simplicial mesh."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    X = space.coords[space.cells]
    if space.tdim == 1:
        lo = np.minimum(X[:, 0, 0], X[:, 1, 0])
        hi = np.maximum(X[:, 0, 0], X[:, 1, 0])
        tol = 1e-12 * max(1.0, np.abs(space.coords).max())
        cand = np.nonzero((x[0] >= lo - tol) & (x[0] <= hi + tol))[0]
        if cand.size == 0:
            raise ValueError("point outside mesh")
        e = cand[0]
        xi = np.array([(x[0] - X[e, 0, 0]) / (X[e, 1, 0] - X[e, 0, 0])])
    else:
        J = np.swapaxes(X[:, 1:, :] - X[:, :1, :], 1, 2)
        xi_all = np.linalg.solve(J, (x[None, :] - X[:, 0, :])[:, :, None])[:, :, 0]
        lam = np.concatenate([1 - xi_all.sum(1, keepdims=True), xi_all], 1)
        ok = np.nonzero(lam.min(1) >= -1e-10)[0]
        if ok.size == 0:
            raise ValueError("point outside mesh")
        e = ok[0]
        xi = xi_all[e]
    phi, _ = tabulate(space.tdim, space.degree, xi[None, :])
    vals = np.asarray(f)[space.cell_dofs[e]].reshape(space.nd, space.bs)
    res = phi[0] @ vals
    return float(res[0]) if space.bs == 1 else res


def evaluate_dofs(fixed_modes, free_spaces, free_modes, coord, used_numModes=None):
    """fixed_modes[k]: dof vector; free_modes[i][k]: dof vector on free_spaces[i]."""
    R = len(fixed_modes) if used_numModes is None else used_numModes
    arr = np.zeros(len(fixed_modes[0]))
    for k in range(R):
        fac = 1.0
        for i in range(len(free_spaces)):
            fac *= point_eval(free_spaces[i], free_modes[i][k], coord[i])
        arr += np.asarray(fixed_modes[k]) * fac
    return arr


def sampling_LHS(min_bnd, max_bnd, n_samples):
    sampler = qmc.LatinHypercube(d=len(min_bnd), seed=3452)
    return qmc.scale(sampler.random(n=n_samples), min_bnd, max_bnd).tolist()


def sample_error(u_fom, u_pgd):
    r = np.asarray(u_pgd).reshape(-1) - np.asarray(u_fom).reshape(-1)
    return np.linalg.norm(r, 2) / np.linalg.norm(np.asarray(u_fom).reshape(-1), 2)


def evaluate_error(fom, pgd_eval, data_test):
    err = np.array([sample_error(fom(s), pgd_eval(s)) for s in data_test])
    return err, err.mean(), err.max()
