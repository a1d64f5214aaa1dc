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

PARITY STATUS: **parity unpinned** at the north-star tolerances (sparsity bit-exact,
matrices 1e-12, modes/reconstruction 1e-8).  The arithmetic of the reference lives in
un-vendored third-party code (fenics=2019.1.0, environment.yml:8) that cannot be
imported in this image (Python 3.12, no dolfin/ufl/ffc/petsc4py), and the reference
ships no golden vectors.  What pins this oracle are the reference's own
tolerance-vs-analytic tests, restated in ``tests/test_oracle_kat.py``:
tests/unit/test_FD.py:166-169, tests/unit/test_pgdclass.py:298-326,
tests/integration/test_elastic.py:353,380, test_laplace.py:970-971,1091-1092,
test_heat1D.py:804-807,903-904.
"""
