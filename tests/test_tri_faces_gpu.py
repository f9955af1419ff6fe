"""GPU parity of the Neumann boundary integrals on the faces of ANY element type (b2_asm_neumann_faces:
triangular faces of tetrahedra and wedges with the 13-point rule, quadrilateral faces of wedges and of 20-node
hexahedra) against the CPU oracle (oracle/mesh_mixed.neumann_rhs, pinned through oracle/fe_face.py to the
compiled reference's elem_type_2D), through the C ABI.  Reference: applications/001_Poisson/main.cpp:495-548."""
import os
import numpy as np
import pytest

# never run on a GPU yet: a kernel that hangs must not hang the box (the thread method ends the process)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]
RTOL = 1e-12
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["cube_tet10", "cube_wedge18", "cube_mixed", "cube_hex27_2x2x2"])
@pytest.mark.parametrize("order", ["linear", "quadratic", "biquadratic"])
def test_neumann_vector_on_any_element_type(ctx, name, order):
    """With a zero source and a zero solution the residual IS the boundary vector: every (plan, face kind) group
    of b2_asm_neumann_faces against the oracle, and the discrete flux balance sum(F) = sum(flux x area)."""
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_mixed as mm
    path = os.path.join(GOLDEN, name + ".neu")
    neumann = {1: 0.2, 4: -1.5, 6: 0.7}
    H = hostapi.HostHierarchy.from_neu(path, 2)
    pb = PoissonMG(ctx, 0, 0, 0, 2, order, hier=H, dirichlet_faces=(2,), neumann=neumann, fsrc=0.0)
    if name != "cube_hex27_2x2x2" or order == "quadratic":
        kinds = {g[2][0].shape[0] for g in pb.nm_groups}
        assert kinds == {"cube_tet10": {13}, "cube_hex27_2x2x2": {16}}.get(name, {13, 16})
    pb.assemble()
    want = mm.neumann_rhs(mm.build_hierarchy(path, 2)[-1], order, neumann)
    got = pb.RES.get()
    assert np.abs(got - want).max() <= RTOL * np.abs(want).max()
    assert abs(got.sum() - (0.2 - 1.5 + 0.7)) <= 1e-12
    del pb


@pytest.mark.parametrize("name,order", [("cube_tet10", "quadratic"), ("cube_mixed", "biquadratic")])
def test_vcycle_trace_with_neumann_faces_on_unstructured_meshes(ctx, name, order):
    """The shipped boundary conditions of 001_Poisson's 3-D input (Dirichlet on one boundary set, a Neumann flux
    on another, natural elsewhere) on tetrahedra and on the mixed mesh: six V-cycles against the oracle."""
    from femus_b200 import hostapi
    from femus_b200.poisson import PoissonMG
    from oracle import mesh_mixed as mm, mg
    path = os.path.join(GOLDEN, name + ".neu")
    neumann, dirichlet = {3: 0.2}, (6,)
    H = hostapi.HostHierarchy.from_neu(path, 2)
    pb = PoissonMG(ctx, 0, 0, 0, 2, order, hier=H, dirichlet_faces=dirichlet, neumann=neumann, fsrc=0.0, coarse_rtol=1e-15,
                   omega=0.3)
    pb.assemble(); pb.galerkin(); pb.mg_set_levels()
    O = mg.Hierarchy(mm.build_hierarchy(path, 2), order, fsrc=0.0, dirichlet_faces=dirichlet, neumann=neumann, mesh=mm)
    trace_ref, eps_ref = O.mg_solve_trace(6, omega=0.3)
    trace = []
    for _ in range(6):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    for a, b in zip(trace, trace_ref):
        assert abs(a - b) <= RTOL * trace_ref[0], (trace, trace_ref)
    assert np.abs(pb.EPS.get() - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()
    del pb
