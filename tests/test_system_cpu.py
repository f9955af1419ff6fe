"""Systems of several Lagrange variables on the host (first step): row numbering [rank][variable][dof]
(LinearEquation::InitPde / GetSystemDof), element dof lists, sparsity pattern with and without a coupling table
(GetSparsityPatternSize), system prolongator (BuildProlongatorMatrix, variable by variable) and Dirichlet flags --
bit-exact against the oracle (oracle/system.py) on boxes split over 1-2 ranks and on the mixed mesh."""
import os
import numpy as np
import pytest

from femus_b200 import hostapi
from oracle import mesh_box as mb, mesh_mixed as mm, system as osys

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TAYLOR_HOOD = ["biquadratic", "biquadratic", "biquadratic", "linear"]


@pytest.mark.parametrize("nprocs", [1, 2])
def test_taylor_hood_system_on_a_box(nprocs):
    lv = mb.build_hierarchy(2, 2, 3, 2, nprocs=nprocs)
    H = hostapi.HostHierarchy(2, 2, 3, 2, nprocs=nprocs)
    for l in (0, 1):
        S = hostapi.SystemOnLevel(H.levels[l], TAYLOR_HOOD)
        assert S.n == 3 * lv[l].dof_offset[2][-1] + lv[l].dof_offset[0][-1]
        d, od = S.elem_dofs(), osys.elem_system_dofs(lv[l], mb, TAYLOR_HOOD)
        for k in range(4):
            nve = 27 if k < 3 else 8
            assert np.array_equal(d[:, k, :nve], np.array(od[k])) and (d[:, k, nve:] == -1).all()
        for pat in (None, [[1, 0, 0, 1], [0, 1, 0, 1], [0, 0, 1, 1], [1, 1, 1, 0]]):       # all pairs; Stokes coupling table
            rp, ci = S.sparsity(pat)
            orp, oci = osys.sparsity(lv[l], mb, TAYLOR_HOOD, pat)
            assert np.array_equal(rp, orp) and np.array_equal(ci, oci)
        walls = [(1, 2, 3, 4, 5, 6)] * 3 + [()]                                              # velocity on every wall, free pressure
        assert np.array_equal(S.bdc(walls), osys.bdc(lv[l], mb, TAYLOR_HOOD, walls))
    rp, ci, v, shape = hostapi.SystemOnLevel(H.levels[1], TAYLOR_HOOD).prolongator()
    P = osys.prolongator(lv[0], lv[1], mb, TAYLOR_HOOD)
    assert shape == P.shape and np.array_equal(rp, P.indptr) and np.array_equal(ci, P.indices) and np.array_equal(v, P.data)
    if nprocs == 2:       # rows of rank 0 come first, variable by variable
        off = hostapi.system_offsets(H.levels[1], TAYLOR_HOOD)
        assert off[0, 1] == off[4, 0] and (np.diff(off, axis=0) >= 0).all()


def test_system_on_the_mixed_mesh():
    path = os.path.join(GOLDEN, "cube_mixed.neu")
    lv, H = mm.build_hierarchy(path, 2), hostapi.HostHierarchy.from_neu(path, 2)
    fams = ["quadratic", "quadratic", "linear"]
    for l in (0, 1):
        rp, ci = hostapi.SystemOnLevel(H.levels[l], fams).sparsity()
        orp, oci = osys.sparsity(lv[l], mm, fams)
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci)
    rp, ci, v, shape = hostapi.SystemOnLevel(H.levels[1], fams).prolongator()
    P = osys.prolongator(lv[0], lv[1], mm, fams)
    assert shape == P.shape and np.array_equal(rp, P.indptr) and np.array_equal(ci, P.indices) and np.abs(v - P.data).max() < 1e-15


def test_one_variable_system_is_the_scalar_case():
    H = hostapi.HostHierarchy(2, 2, 2, 2)
    S = hostapi.SystemOnLevel(H.levels[1], ["biquadratic"])
    rp, ci = S.sparsity()
    rp1, ci1 = H.levels[1].sparsity("biquadratic")
    assert np.array_equal(rp, rp1) and np.array_equal(ci, ci1)
    assert np.array_equal(S.elem_dofs()[:, 0, :], H.levels[1].system_dofs("biquadratic"))
    with pytest.raises(ValueError):
        hostapi.SystemOnLevel(H.levels[0], ["linear"]).prolongator()       # no level below the coarsest
