"""Oracle (test infrastructure): DOLFIN 2019.1 built-in mesh generators, restated.

[DOLFIN-knowledge] vertex numbering and cell splitting of IntervalMesh, RectangleMesh
("right" / "left" / "crossed") and BoxMesh (6 tetrahedra per hexahedron).  The reference
only *calls* these (tests/integration/test_elastic.py:45, test_solver_problem.py:69-71).
Each generator returns ``(coords float64 [n_vertices, gdim], cells int32 [n_cells, gdim+1])``
with every cell's vertex indices ascending (DOLFIN ``mesh.order()``).
"""
import numpy as np


def interval_mesh(n, a, b):
    x = a + (b - a) * np.arange(n + 1, dtype=np.float64) / n
    cells = np.stack([np.arange(n), np.arange(1, n + 1)], axis=1).astype(np.int32)
    return x.reshape(-1, 1), cells


def rectangle_mesh(x0, y0, x1, y1, nx, ny, diagonal="right"):
    xs = x0 + (x1 - x0) * np.arange(nx + 1, dtype=np.float64) / nx
    ys = y0 + (y1 - y0) * np.arange(ny + 1, dtype=np.float64) / ny
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # row iy, col ix -> vertex iy*(nx+1)+ix
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v1 + (nx + 1)
    if diagonal == "right":
        cells = np.stack([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], axis=1)
    elif diagonal == "left":
        cells = np.stack([np.stack([v0, v1, v2], 1), np.stack([v1, v2, v3], 1)], axis=1)
    elif diagonal == "crossed":
        xm = 0.5 * (xs[:-1] + xs[1:])
        ym = 0.5 * (ys[:-1] + ys[1:])
        XM, YM = np.meshgrid(xm, ym, indexing="xy")
        coords = np.vstack([coords, np.stack([XM.ravel(), YM.ravel()], axis=1)])
        vm = (nx + 1) * (ny + 1) + (iy * nx + ix).ravel()
        cells = np.stack(
            [
                np.stack([v0, v1, vm], 1),
                np.stack([v0, v2, vm], 1),
                np.stack([v1, v3, vm], 1),
                np.stack([v2, v3, vm], 1),
            ],
            axis=1,
        )
    else:
        raise ValueError("diagonal must be right|left|crossed")
    cells = cells.reshape(-1, 3)
    return coords, np.sort(cells, axis=1).astype(np.int32)


def box_mesh(x0, y0, z0, x1, y1, z1, nx, ny, nz):
    xs = x0 + (x1 - x0) * np.arange(nx + 1, dtype=np.float64) / nx
    ys = y0 + (y1 - y0) * np.arange(ny + 1, dtype=np.float64) / ny
    zs = z0 + (z1 - z0) * np.arange(nz + 1, dtype=np.float64) / nz
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    v0 = (iz * sy + iy * sx + ix).ravel()
    v1, v2 = v0 + 1, v0 + sx
    v3 = v1 + sx
    v4, v5, v6, v7 = v0 + sy, v1 + sy, v2 + sy, v3 + sy
    tets = [
        (v0, v1, v3, v7),
        (v0, v1, v7, v5),
        (v0, v5, v7, v4),
        (v0, v3, v2, v7),
        (v0, v6, v4, v7),
        (v0, v2, v6, v7),
    ]
    cells = np.stack([np.stack(t, 1) for t in tets], axis=1).reshape(-1, 4)
    return coords, np.sort(cells, axis=1).astype(np.int32)
