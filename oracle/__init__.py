"""CPU oracle for the FEMuS assembly + geometric-multigrid hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``femus_b200/`` may import this package; it is used by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` as the checker, never as the thing measured or shipped.

Two halves:

* ``oracle/_ref/libfemus_fe_ref.so`` -- the reference's OWN FE kernel
  (``src/02_reference_geom_elements``) compiled where it lies by ``oracle/ref_capi/Makefile``;
  loaded through :mod:`oracle.ref`.  Pins :mod:`oracle.fe_hex` (tables, Jacobian, element matrices,
  local prolongators).
* a numpy / scipy restatement of everything that needs PETSc+MPI in the reference (box mesh,
  renumbering, refinement, dof maps, Dirichlet flags, sparsity, assembly, prolongators, Galerkin
  operators, V-cycle).  **Parity of that half is unpinned by the reference**: the reference ships
  no golden vectors for a 3-D hex Poisson problem and cannot be executed without PETSc
  (SURVEY.md section 4 and 8c); each function cites the file:line it restates.
"""
