"""Host-side layout of a distributed run: one rank per GPU, the mesh sharded by element (z-slabs of
the coarsest level, children inherit -- reference MeshMetisPartitioning.cpp:143-155), every rank
holding ALL dofs of its own elements.  Dofs on the partition interface are held by several ranks;
the lowest rank owns them (reference Mesh.cpp:530-553).  This module computes, per level, what
b2_halo_create needs: the interface dofs, their rank-independent position in the packed interface
vector, the ownership mask and the multiplicity.  Interface nodes of different ranks are matched
through their integer lattice coordinates (hostapi.HostLevel.lattice_key), gathered once at setup.

No arithmetic on field data happens here; `allgather(obj) -> [obj of rank 0, ..., obj of rank P-1]`
is supplied by the launcher (torch.distributed.all_gather_object under torchrun)."""
import numpy as np


class LevelLayout:
    """Interface layout of one level for one FE family on one rank."""

    def __init__(self, n_local, idx, pos, n_packed, owned, mult, exchange=None):
        # exchange: (hold_ptr, hold_rank, hold_pos, hold_spos, longest message) of the peer-memory form of the sum
        self.exchange = exchange
        self.n_local = n_local
        self.idx = idx                  # int32 [n_if] local dofs on the interface
        self.pos = pos                  # int32 [n_if] position in the packed interface vector
        self.n_packed = n_packed
        self.owned = owned              # uint8 [n_local]
        self.mult = mult                # uint8 [n_local]

    @property
    def n_owned(self):
        return int(self.owned.sum())


def level_layout(level, ndofs, rank, allgather):
    """level: hostapi.HostLevel of a rank-local hierarchy; ndofs: dofs of the family on it (local
    dof == local node id because the local mesh is numbered as a single rank: vertices first)."""
    nodes = level.interface_nodes()
    nodes = nodes[nodes < ndofs]
    keys = level.lattice_key(nodes)
    all_keys = allgather(keys)
    union = np.unique(np.concatenate(all_keys)) if len(all_keys) else np.zeros(0, dtype=np.int64)
    count = np.zeros(union.shape[0], dtype=np.int32)
    owner = np.full(union.shape[0], len(all_keys), dtype=np.int32)
    for r, k in enumerate(all_keys):
        p = np.searchsorted(union, k)
        count[p] += 1
        owner[p] = np.minimum(owner[p], r)
    pos = np.searchsorted(union, keys).astype(np.int32)
    owned = np.ones(ndofs, dtype=np.uint8)
    mult = np.ones(ndofs, dtype=np.uint8)
    owned[nodes] = (owner[pos] == rank)
    mult[nodes] = count[pos]
    return LevelLayout(ndofs, nodes.astype(np.int32), pos, int(union.shape[0]), owned, mult, exchange_lists(nodes, all_keys, rank))


def exchange_lists(nodes, all_keys, rank):
    """Peer-memory form of the interface sum (b2_halo_set_exchange).  all_keys[q] = lattice keys of rank q's interface
    entries in ITS entry order.  The message q -> r carries, in q's entry order, the values of q's entries whose key r
    also holds; so both sides know every position without talking to each other.  Returns
      hold_ptr [n_if+1], hold_rank, hold_pos, hold_spos: per entry its holders in ascending rank order (this rank
      included, positions -1), where each holder's value sits in that holder's message to this rank and where this
      rank's value sits in its message to the holder;  longest: the longest message this rank sends or receives."""
    mine = all_keys[rank]
    n_if = mine.shape[0]
    hk = [np.arange(n_if, dtype=np.int64)]
    hr = [np.full(n_if, rank, dtype=np.int64)]
    hp = [np.full(n_if, -1, dtype=np.int64)]
    hs = [np.full(n_if, -1, dtype=np.int64)]
    longest = 0
    for q, kq in enumerate(all_keys):
        if q == rank or kq.shape[0] == 0 or n_if == 0:
            continue
        to_q = np.isin(mine, kq)                    # my entries that q holds: my message to q, in my entry order
        if not to_q.any():
            continue
        from_q = kq[np.isin(kq, mine)]              # q's message to me, in q's entry order
        order = np.argsort(from_q, kind="stable")
        where = order[np.searchsorted(from_q[order], mine[to_q])]
        hk.append(np.nonzero(to_q)[0].astype(np.int64))
        hr.append(np.full(where.shape[0], q, dtype=np.int64))
        hp.append(where.astype(np.int64))
        hs.append(np.arange(where.shape[0], dtype=np.int64))
        longest = max(longest, int(where.shape[0]))
    hk, hr, hp, hs = np.concatenate(hk), np.concatenate(hr), np.concatenate(hp), np.concatenate(hs)
    o = np.lexsort((hr, hk))                        # by entry, then by holder rank
    hk, hr, hp, hs = hk[o], hr[o], hp[o], hs[o]
    hold_ptr = np.zeros(n_if + 1, dtype=np.int64)
    np.add.at(hold_ptr, hk + 1, 1)
    hold_ptr = np.cumsum(hold_ptr)
    return hold_ptr, hr.astype(np.int32), hp.astype(np.int32), hs.astype(np.int32), longest


def torch_allgather():
    """all_gather_object over the default torch.distributed group (any backend)."""
    import torch.distributed as dist

    def gather(obj):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, obj)
        return out
    return gather
