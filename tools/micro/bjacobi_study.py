"""How much does a PETSc-style block-Jacobi (one block per CTA slice, solved EXACTLY -- the best case for any inner
IC(0)/Chebyshev sweeps) cut the PCG iteration count of the configs[2] operator?  CPU study with the oracle's matrices
(VERDICT r01 item 5).  Prints iterations to relres 1e-13 for
  node    3x3 node-block Jacobi (what the product runs),
  slices  296 contiguous node ranges of the mesh ordering (the rows a persistent CTA owns),
  cubes   the same number of blocks, but cube-shaped sub-domains (needs a renumbering of the mesh).
usage: python tools/micro/bjacobi_study.py [n=32] [blocks=296]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import fem as ofem
from oracle import meshes as omesh


def pcg(A, b, apply_minv, rtol=1e-13, maxit=20000):
    x = np.zeros_like(b)
    r = b.copy()
    z = apply_minv(r)
    p = z.copy()
    rz = r @ z
    bb = b @ b
    for it in range(1, maxit + 1):
        q = A @ p
        al = rz / (p @ q)
        x += al * p
        r -= al * q
        if r @ r <= rtol * rtol * bb:
            return x, it
        z = apply_minv(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxit


def block_solver(A, groups):
    lus = [(g, spla.splu(A[g][:, g].tocsc())) for g in groups]

    def apply(r):
        z = np.empty_like(r)
        for g, lu in lus:
            z[g] = lu.solve(r[g])
        return z

    return apply


def main(n=32, nblocks=296):
    X, C = omesh.box_mesh(0.0, 0.0, 0.0, 1.0, 1.0, 1.0, n, n, n)
    S = ofem.Space(X, C, 1, 3)
    nu = 0.3
    lam, mu = nu / ((1 + nu) * (1 - 2 * nu)), 1.0 / (2 * (1 + nu))
    Cm = ofem.isotropic_C(lam, mu, 3)
    T = ofem.T_voigt(Cm, 3)
    zones = [lambda x: (x[..., 0] < 1 / 3) * 1.0, lambda x: ((x[..., 0] >= 1 / 3) & (x[..., 0] < 2 / 3)) * 1.0,
             lambda x: (x[..., 0] >= 2 / 3) * 1.0]
    E = 1.3
    A = sum(c * ofem.assemble_bilinear(S, T, weight=w, weight_degree=0) for c, w in zip((1.0, E, E * E), zones)).tocsr()
    nodes = S.node_coords if hasattr(S, "node_coords") else X
    free_node = nodes[:, 0] > 1e-12
    free = np.repeat(free_node, 3)
    A = A[free][:, free].tocsr()
    nn = int(free_node.sum())
    rng = np.random.default_rng(0)
    b = A @ rng.uniform(-1, 1, A.shape[0])
    out = {"n": n, "dofs": A.shape[0], "blocks": nblocks}
    t0 = time.perf_counter()
    node_groups = [np.arange(3 * i, 3 * i + 3) for i in range(nn)]
    D = sp.block_diag([np.linalg.inv(A[g][:, g].toarray()) for g in node_groups[:0]]) if False else None
    # 3x3 node blocks: vectorised inverse
    Ad = A.tobsr(blocksize=(3, 3))
    diag = np.zeros((nn, 3, 3))
    for i in range(nn):
        s, e = Ad.indptr[i], Ad.indptr[i + 1]
        diag[i] = Ad.data[s + np.searchsorted(Ad.indices[s:e], i)]
    dinv = np.linalg.inv(diag)
    _, out["node"] = pcg(A, b, lambda r: np.einsum("nij,nj->ni", dinv, r.reshape(nn, 3)).ravel())
    bounds = np.linspace(0, nn, nblocks + 1).astype(int)
    slices = [np.arange(3 * bounds[k], 3 * bounds[k + 1]) for k in range(nblocks) if bounds[k + 1] > bounds[k]]
    _, out["slices"] = pcg(A, b, block_solver(A, slices))
    # cube-shaped sub-domains: m^3 boxes with m^3 ~ nblocks
    m = max(1, int(round(nblocks ** (1.0 / 3.0))))
    fx = nodes[free_node]
    cell = np.minimum((fx * m).astype(int), m - 1)
    cid = (cell[:, 0] * m + cell[:, 1]) * m + cell[:, 2]
    cubes = [np.repeat(3 * np.nonzero(cid == c)[0], 3) + np.tile(np.arange(3), int((cid == c).sum())) for c in range(m ** 3)]
    cubes = [g for g in cubes if len(g)]
    out["cube_blocks"] = len(cubes)
    _, out["cubes"] = pcg(A, b, block_solver(A, cubes))
    out["seconds"] = time.perf_counter() - t0
    print(out)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 296)
