#!/usr/bin/env python
"""Run this where a real FEniCS 2019.1 lives (it is NOT importable in the build container): dumps
what is needed to check the north-star parity targets against DOLFIN itself --

    sparsity patterns and dof maps bit-exact, assembled matrices to 1e-12, PGD modes to 1e-8.

    python tools/export_from_fenics.py baseline/_ref/fenics_export.npz

For a few small spaces (P1/P2 interval, P1 "right"/"crossed" rectangle, P1 box, vector P1 box) it
stores mesh coordinates and cells, the cell->dof table, dof coordinates, the boundary dofs of
`on_boundary`, and the PETSc CSR (indptr, indices, data) of the mass and stiffness matrices.
`tests/test_fenics_export.py` consumes the file when it is present (skipped otherwise) and feeds the
DOLFIN dofmap into `FunctionSpace.from_arrays`, so the pattern / matrix comparison is entry by entry.
"""
import sys

import numpy as np


def main(path):
    import dolfin as df

    out = {}
    cases = {
        "interval_p1": (df.IntervalMesh(17, 0.2, 2.0), "P", 1, 1),
        "interval_p2": (df.IntervalMesh(11, 0.0, 1.0), "P", 2, 1),
        "rect_right_p1": (df.RectangleMesh(df.Point(0, 0), df.Point(3, 1), 7, 4, "right"), "P", 1, 1),
        "rect_crossed_p1": (df.RectangleMesh(df.Point(0, 0), df.Point(3, 1), 5, 3, "crossed"), "P", 1, 1),
        "box_p1": (df.BoxMesh(df.Point(0, 0, 0), df.Point(1, 2, 1.5), 4, 3, 2), "P", 1, 1),
        "box_vec_p1": (df.BoxMesh(df.Point(0, 0, 0), df.Point(1, 1, 1), 3, 3, 3), "P", 1, 3),
    }
    for key, (mesh, fam, deg, bs) in cases.items():
        V = df.FunctionSpace(mesh, fam, deg) if bs == 1 else df.VectorFunctionSpace(mesh, fam, deg)
        u, v = df.TrialFunction(V), df.TestFunction(V)
        out[key + "_coords"] = mesh.coordinates().copy()
        out[key + "_cells"] = mesh.cells().copy()
        out[key + "_cell_dofs"] = np.array([V.dofmap().cell_dofs(c) for c in range(mesh.num_cells())])
        out[key + "_dof_coords"] = V.tabulate_dof_coordinates().copy()
        out[key + "_degree_bs"] = np.array([deg, bs])
        bc = df.DirichletBC(V, df.Constant(0.0) if bs == 1 else df.Constant((0.0,) * bs), lambda x, on_boundary: on_boundary)
        out[key + "_bc_dofs"] = np.array(sorted(bc.get_boundary_values().keys()))
        forms = {"mass": df.inner(u, v) * df.dx, "stiff": df.inner(df.grad(u), df.grad(v)) * df.dx}
        for name, a in forms.items():
            A = df.as_backend_type(df.assemble(a, keep_diagonal=True)).mat()
            indptr, indices, data = A.getValuesCSR()
            out["%s_%s_indptr" % (key, name)] = np.asarray(indptr)
            out["%s_%s_indices" % (key, name)] = np.asarray(indices)
            out["%s_%s_data" % (key, name)] = np.asarray(data)
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "baseline/_ref/fenics_export.npz")
