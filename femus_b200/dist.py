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

    def __init__(self, n_local, idx, pos, n_packed, owned, mult, keys_owned_order=None):
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
    return LevelLayout(ndofs, nodes.astype(np.int32), pos, int(union.shape[0]), owned, mult)


def torch_allgather():
    """all_gather_object over the default torch.distributed group (any backend)."""
    import torch.distributed as dist

    def gather(obj):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, obj)
        return out
    return gather
