"""N>1 host logic on CPU: per-rank local hierarchies against the complete mesh, interface layout and
ownership against the reference's dof offsets, and a world_size-2 `gloo` run of the sharded V-cycle
(numpy mirror of the device sequence) against the serial oracle."""
import os
import socket
import tempfile

import numpy as np
import pytest

from femus_b200 import hostapi, dist as distlayout
from oracle import mesh_box as mb, mg


@pytest.mark.parametrize("box,nl,P", [((2, 3, 4), 3, 2), ((2, 2, 4), 2, 4), ((1, 1, 3), 3, 3)])
def test_local_hierarchy_matches_complete_mesh(box, nl, P):
    G = hostapi.HostHierarchy(*box, nl, nprocs=P)
    for r in range(P):
        Lh = hostapi.HostHierarchy(*box, nl, nprocs=P, local_rank=r)
        for l in range(nl):
            g, loc = G.levels[l], Lh.levels[l]
            e0, e1 = g.elem_offset[r], g.elem_offset[r + 1]
            assert loc.nel == e1 - e0
            # same elements in the same order, same nodes (by lattice name), same boundary faces
            assert np.array_equal(g.lattice_key()[g.conn[e0:e1]], loc.lattice_key()[loc.conn])
            assert np.array_equal(g.face[e0:e1], loc.face)
            assert np.abs(g.xyz[:, g.conn[e0:e1]] - loc.xyz[:, loc.conn]).max() <= 4e-16
            assert len(G.levels[l].interface_nodes()) == 0


@pytest.mark.parametrize("order", ["linear", "biquadratic"])
def test_layout_ownership_matches_reference_offsets(order):
    box, nl, P = (2, 2, 4), 3, 4
    G = hostapi.HostHierarchy(*box, nl, nprocs=P)
    H = [hostapi.HostHierarchy(*box, nl, nprocs=P, local_rank=r) for r in range(P)]
    for l in range(nl):
        nd = [H[r].levels[l].ndofs(order) for r in range(P)]
        allk = []
        for r in range(P):
            nodes = H[r].levels[l].interface_nodes()
            allk.append(H[r].levels[l].lattice_key(nodes[nodes < nd[r]]))
        lays = [distlayout.level_layout(H[r].levels[l], nd[r], r, lambda k: allk) for r in range(P)]
        # owned dofs per rank == _dofOffset differences of the complete mesh (Mesh.cpp:706-853)
        assert [L.n_owned for L in lays] == list(np.diff(G.levels[l].dof_offset[hostapi.FAMILY[order]]))
        assert all(L.n_packed == lays[0].n_packed for L in lays)
        assert max(L.mult.max() for L in lays) == 2        # slabs: an interface dof is on 2 ranks
        # each packed position is claimed by exactly `mult` ranks and owned by one
        claims = np.zeros(lays[0].n_packed, dtype=int)
        owners = np.zeros(lays[0].n_packed, dtype=int)
        for L in lays:
            claims[L.pos] += 1
            owners[L.pos] += L.owned[L.idx]
        assert np.all(claims == 2) and np.all(owners == 1)


def _synthetic_layouts(P, n_local, seed):
    """Ranks holding random subsets of a common key space (up to 4 holders per key): the exchange lists do not assume slabs."""
    rng = np.random.default_rng(seed)
    nkeys = 40
    holders = [rng.choice(P, size=rng.integers(2, min(P, 4) + 1), replace=False) for _ in range(nkeys)]
    nodes, keys = [], []
    for r in range(P):
        k = np.array([K for K in range(nkeys) if r in holders[K]], dtype=np.int64)
        k = k[rng.permutation(k.shape[0])]                 # every rank lists its interface entries in its own order
        keys.append(k * 7 + 3)
        nodes.append(rng.choice(n_local, size=k.shape[0], replace=False).astype(np.int64))
    return nodes, keys


@pytest.mark.parametrize("case", ["slabs", "synthetic"])
def test_peer_exchange_lists_reproduce_the_packed_interface_sum(case):
    """b2_halo_set_exchange's inputs (femus_b200.dist.exchange_lists): emulate the messages rank by rank and sum every
    interface entry over its holders in ascending rank order -- the result must be the packed all-reduce's, and equal
    on all holders of a dof bit for bit."""
    if case == "slabs":
        box, nl, P, order = (2, 2, 4), 2, 4, "biquadratic"
        H = [hostapi.HostHierarchy(*box, nl, nprocs=P, local_rank=r) for r in range(P)]
        l = nl - 1
        nd = [H[r].levels[l].ndofs(order) for r in range(P)]
        nodes = [H[r].levels[l].interface_nodes() for r in range(P)]
        nodes = [nodes[r][nodes[r] < nd[r]] for r in range(P)]
        keys = [H[r].levels[l].lattice_key(nodes[r]) for r in range(P)]
    else:
        P = 5
        nd = [60] * P
        nodes, keys = _synthetic_layouts(P, 60, 3)
    ex = [distlayout.exchange_lists(nodes[r], keys, r) for r in range(P)]
    rng = np.random.default_rng(1)
    v = [rng.standard_normal(nd[r]) for r in range(P)]
    # messages: every rank writes its value of entry k into cell hold_spos of its message to each other holder
    msg = [{} for _ in range(P)]
    for r in range(P):
        hptr, hrank, hpos, hspos, longest = ex[r]
        for k in range(nodes[r].shape[0]):
            for q, sp_ in zip(hrank[hptr[k]:hptr[k + 1]], hspos[hptr[k]:hptr[k + 1]]):
                if q != r:
                    cell = msg[q].setdefault(r, {})
                    assert sp_ not in cell and 0 <= sp_ < longest               # one value per cell, inside the slot
                    cell[int(sp_)] = v[r][nodes[r][k]]
    # reference: sum over holders through the union of keys
    union = np.unique(np.concatenate(keys))
    total = np.zeros(union.shape[0])
    for r in range(P):
        np.add.at(total, np.searchsorted(union, keys[r]), v[r][nodes[r]])
    got_by_key = {}
    for r in range(P):
        hptr, hrank, hpos = ex[r][0], ex[r][1], ex[r][2]
        assert hptr.shape[0] == nodes[r].shape[0] + 1
        for k in range(nodes[r].shape[0]):
            hs = hrank[hptr[k]:hptr[k + 1]]
            assert np.all(np.diff(hs) > 0) and r in hs                       # ascending, this rank included
            s = None
            for q, pos in zip(hs, hpos[hptr[k]:hptr[k + 1]]):
                a = v[r][nodes[r][k]] if q == r else msg[r][q][int(pos)]
                s = a if s is None else s + a
            K = int(keys[r][k])
            assert abs(s - total[np.searchsorted(union, K)]) <= 1e-14 * max(1.0, abs(s))
            assert got_by_key.setdefault(K, s) == s                           # bit-identical on every holder
    assert len(got_by_key) == union.shape[0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_rank(rank, world, port, box, nl, order, ncyc, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.dist_numpy import NumpyRank

    def allgather(obj):
        o = [None] * world
        dist.all_gather_object(o, obj)
        return o

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()
    R = NumpyRank(box, nl, order, rank, world, allgather, allreduce)
    trace, eps = R.mg_solve_trace(ncyc)
    top = R.H.levels[-1]
    own = R.lay[-1].owned.astype(bool)
    res = allgather((top.lattice_key(np.arange(R.nd[-1]))[own], eps[own], trace))
    if rank == 0:
        np.save(out, np.array(res, dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("order,box,nl", [("linear", (2, 2, 4), 3), ("biquadratic", (1, 2, 2), 3)])
def test_gloo_world2_sharded_vcycle_matches_serial_oracle(order, box, nl):
    import torch.multiprocessing as mp
    world, ncyc = 2, 3
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "r.npy")
        mp.spawn(_gloo_rank, args=(world, _free_port(), box, nl, order, ncyc, out), nprocs=world, join=True)
        res = np.load(out, allow_pickle=True)
    H = mg.Hierarchy(mb.build_hierarchy(*box, nl), order)
    trace_ref, eps_ref = H.mg_solve_trace(ncyc)
    G = hostapi.HostHierarchy(*box, nl)
    n = G.levels[-1].ndofs(order)
    gk = G.levels[-1].lattice_key(np.arange(n))
    srt = np.argsort(gk)
    eps = np.full(n, np.nan)
    for r in range(world):
        eps[srt[np.searchsorted(gk[srt], res[r][0])]] = res[r][1]
        for a, b in zip(res[r][2], trace_ref):
            assert abs(a - b) <= 1e-12 * trace_ref[0]
    assert not np.isnan(eps).any()
    assert np.abs(eps - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()
