"""Separated representation of a MOVING heat source for the PGD load (north_star item 1: "moving-source load";
SURVEY.md 7.3: "the moving-source load is not separable in (x, t, v); needs a precomputed separated expansion
(`load` with several `ext` terms, as the API already allows: tests/integration/test_solver_problem.py:268-282)").

The reference has no moving source (its only source term is the static Goldak-type Expression of
tests/integration/test_heat1D.py:684-691), so there is nothing to be bit-compatible with: parity unpinned.  What a
PGDrome user needs is the list of separated load terms; this module computes it.

    q(x, t, v) = exp(-3 (x - x_s(t, v))^2 / a^2),      x_s(t, v) = x0 + v * speed * t

is sampled on the tensor grid of the problem's own 1-D node sets and approximated by a greedy sequence of rank-one terms

    q  ~  sum_m  G_m(x) H_m(t) W_m(v)

each obtained by alternating least squares on the current residual (the same fixed-point idea as the PGD itself, applied
a posteriori to a known tensor).  The approximation error is returned per number of terms, so the caller sees what a
given `n_terms` buys: a translating Gaussian is the classical slowly-separable case (the error decays algebraically,
roughly like the ratio source width / travelled distance).  Host NumPy, set-up only, deterministic.
"""
import numpy as np


def moving_gaussian(x, t, v, a=0.12, x0=0.2, speed=0.6):
    """q[i, j, k] = exp(-3 (x_i - x0 - speed * v_k * t_j)^2 / a^2) on the tensor grid of the three node sets."""
    x, t, v = (np.asarray(c, dtype=np.float64).ravel() for c in (x, t, v))
    xs = x0 + speed * t[:, None] * v[None, :]
    return np.exp(-3.0 * (x[:, None, None] - xs[None, :, :]) ** 2 / (a * a))


def separate(Q, n_terms, sweeps=60, tol=1e-9):
    """Greedy rank-one separation of a 3-way tensor.  Returns (G [m, n0], H [m, n1], W [m, n2], rel_err [m]) with
    rel_err[k] = ||Q - sum_{m<=k} G_m x H_m x W_m||_F / ||Q||_F.  Deterministic (fixed starting vectors)."""
    Q = np.asarray(Q, dtype=np.float64)
    R = Q.copy()
    n0, n1, n2 = Q.shape
    norm_q = np.linalg.norm(Q)
    G, H, W, errs = [], [], [], []
    for _ in range(n_terms):
        # start from the residual's largest fibre (deterministic, never orthogonal to the residual)
        i, j, k = np.unravel_index(np.argmax(np.abs(R)), R.shape)
        h, w = R[i, :, k].copy(), R[i, j, :].copy()
        if not np.any(h) or not np.any(w):
            break
        h /= np.linalg.norm(h)
        w /= np.linalg.norm(w)
        g = np.einsum("ijk,j,k->i", R, h, w)
        for _s in range(sweeps):
            g_old = g
            g = np.einsum("ijk,j,k->i", R, h, w)
            h = np.einsum("ijk,i,k->j", R, g, w) / max(g @ g, 1e-300)
            h /= np.linalg.norm(h)
            w = np.einsum("ijk,i,j->k", R, g, h) / max(g @ g, 1e-300)
            w /= np.linalg.norm(w)
            g = np.einsum("ijk,j,k->i", R, h, w)
            if np.linalg.norm(g - g_old) <= tol * np.linalg.norm(g):
                break
        R -= np.einsum("i,j,k->ijk", g, h, w)
        G.append(g)
        H.append(h)
        W.append(w)
        errs.append(np.linalg.norm(R) / norm_q)
    return np.array(G), np.array(H), np.array(W), np.array(errs)


def moving_gaussian_terms(x, t, v, n_terms=8, a=0.12, x0=0.2, speed=0.6):
    """Separated load terms of the moving Gaussian on the given node sets: dict with G, H, W (one row per term), the
    relative Frobenius error after every term and the parameters."""
    G, H, W, errs = separate(moving_gaussian(x, t, v, a, x0, speed), n_terms)
    return {"G": G, "H": H, "W": W, "rel_err": errs, "x": np.asarray(x, dtype=np.float64).ravel(),
            "t": np.asarray(t, dtype=np.float64).ravel(), "v": np.asarray(v, dtype=np.float64).ravel(), "a": a, "x0": x0,
            "speed": speed}
