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
  operators, V-cycle; systems of several variables, Stokes / Navier-Stokes element loops).  The reference ships no
  golden vectors for these and its build needs PETSc + MPI, but its OWN sources compile here on a single-process host
  backend (:mod:`oracle.ref_build`): the reference's unmodified 001_Poisson, and drivers that compile its Stokes callback
  and its Navier-Stokes library routine in place, generated tests/golden/ref_poisson_*.npz and ref_stokes_*.npz, which
  pin this half (tests/test_reference_pin.py, tests/test_reference_pin_stokes.py).  What stays unpinned is PETSc's own
  numerics (PCASM / ILU(0) / GMRES / MUMPS: :mod:`oracle.asm` restates their published algorithms); each function cites
  the file:line it restates.
"""
