"""The Stokes / Navier-Stokes oracles (oracle/stokes.py, oracle/navier_stokes.py): internal consistency, since the
reference's routines cannot run here (PETSc, adept recording inside the library build) -- the analytic Newton Jacobian
equals the finite-difference derivative of the residual, the Navier-Stokes routine reduces to the Stokes one at zero
velocity, the Stokes residual is -B sol, constant pressure is in the kernel of the gradient block on a closed box."""
import numpy as np

from oracle import fe_hex, mesh_box as mb, navier_stokes as ons, stokes

TOF = lambda t, o: fe_hex.tables(o)      # noqa: E731


def _n(L):
    return 3 * L.dof_offset[2][-1] + L.dof_offset[0][-1]


def test_newton_jacobian_is_the_derivative_of_the_residual():
    L = mb.build_hierarchy(1, 2, 1, 1)[0]
    rng = np.random.default_rng(0)
    sol, v = 0.3 * rng.standard_normal(_n(L)), rng.standard_normal(_n(L))
    A, _ = ons.assemble(L, mb, "biquadratic", "linear", sol, 0.1, TOF)
    eps = 1e-6
    _, rp = ons.assemble(L, mb, "biquadratic", "linear", sol + eps * v, 0.1, TOF)
    _, rm = ons.assemble(L, mb, "biquadratic", "linear", sol - eps * v, 0.1, TOF)
    fd = -(rp - rm) / (2 * eps)               # RES = -aRes
    assert np.abs(fd - A @ v).max() <= 1e-8 * np.abs(A @ v).max()


def test_navier_stokes_reduces_to_stokes_and_stokes_residual_is_linear():
    L = mb.build_hierarchy(2, 1, 1, 1)[0]
    rng = np.random.default_rng(1)
    n, npr = _n(L), L.dof_offset[0][-1]
    sol0 = np.zeros(n)
    sol0[-npr:] = rng.standard_normal(npr)
    A0, r0 = ons.assemble(L, mb, "biquadratic", "linear", sol0, 0.1, TOF)
    As, rs = stokes.assemble(L, mb, "biquadratic", "linear", sol0, 0.1, TOF)
    assert np.abs(A0 - As).max() <= 1e-15 and np.abs(r0 - rs).max() <= 1e-15
    sol = rng.standard_normal(n)
    A, r = stokes.assemble(L, mb, "biquadratic", "linear", sol, 0.7, TOF)
    assert np.abs(A @ sol + r).max() <= 1e-13 * np.abs(r).max()
    assert np.abs(A - A.T).max() <= 1e-15
    # the gradient block annihilates constant pressure against velocities that vanish on the boundary
    interior = mb.bdc_flags(L, "biquadratic") > 1.5
    p1 = np.zeros(n)
    p1[-npr:] = 1.0
    g = (A @ p1)[:L.dof_offset[2][-1]]
    assert np.abs(g[interior]).max() <= 1e-15 if interior.any() else True
