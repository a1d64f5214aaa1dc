"""CPU oracle for the PGD enrichment hot path of BAMresearch/PGDrome.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``pgdrome_b200/`` imports ``oracle``.

It restates, in plain NumPy/SciPy, the algorithm the reference executes through
FEniCS/DOLFIN 2019.1 + PETSc/MUMPS + SciPy:

* ``oracle.meshes``   DOLFIN built-in mesh generators (Interval/Rectangle/Box)       [DOLFIN-knowledge]
* ``oracle.fem``      Lagrange P1/P2 dofmaps, element tabulation, COO->CSR assembly,
                      symmetric Dirichlet elimination                                 [DOLFIN-knowledge]
* ``oracle.pgd``      ``get_Fsinit`` / ``solve_PGD`` / ``FP_solve`` / ``FD_matrices``
                      (pgdrome/solver.py:158-304, 306-506, 508-881, 947-988)
* ``oracle.evaluate`` ``PGD.evaluate`` both paths (pgdrome/model.py:724-860) and the
                      LHS sampling / error loop (pgdrome/model.py:1704-1825)
* ``oracle.problems`` matrix-form restatements of the reference's own callback sets
                      (tests/integration/*.py) and of the BASELINE.json configs.

PARITY STATUS: **partly pinned**.
  * PINNED to round-off by golden vectors generated from the UNMODIFIED reference code
    (tests/golden/make_golden.py runs pgdrome/solver.py + pgdrome/model.py + the reference's own
    test callbacks in this container through a NumPy stand-in for the few DOLFIN containers the
    finite-difference paths touch): ``FD_matrices``; the enrichment loop in FD mode (get_Fsinit,
    residual check, FP_solve with both stopping criteria, the three normalisations, the stopping
    test); ``PGD.evaluate`` interp1d path and mode point-evaluation loop; evaluate_min/max; LHS
    sampling; the error loop.  Checked in tests/test_oracle_golden.py.  Also pinned: the mode
    combination / return shapes of ``evaluate_sensor_response`` (tests/golden/sensor.npz) and the PXDMF
    XML layout (a product-written file read back by the unmodified reference loader, pxdmf.npz).
  * **parity unpinned** for everything that goes through DOLFIN's finite-element assembly
    (sparsity bit-exact, matrices 1e-12): fenics=2019.1.0 (environment.yml:8) cannot be imported
    here (Python 3.12, no dolfin/ufl/ffc/petsc4py) and the reference ships no stored matrices.
    That part is pinned only to the tolerances of the reference's own tests, restated in
    tests/test_oracle_kat.py: tests/unit/test_FD.py:166-169, tests/integration/test_elastic.py:353,380,
    test_laplace.py:970-971,1091-1092, test_heat1D.py:804-807,903-904, test_solver_problem.py:748-752.
"""
