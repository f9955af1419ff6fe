"""Worker of the multi-rank tests: one process per rank (spawned by the test), rendezvous over
127.0.0.1.  `run_gpu_rank` drives the real CUDA path (one GPU per rank, NCCL inside the library);
results travel back to rank 0 through gloo."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def init_gloo(rank, world, port):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    return dist


def run_gpu_rank(rank, world, port, box, nlevels, order, ncycles, out_path, peer=True):
    import torch
    dist = init_gloo(rank, world, port)
    from femus_b200 import capi
    from femus_b200.dist import torch_allgather
    from femus_b200.poisson import PoissonMG
    torch.cuda.set_device(rank)
    ctx = capi.Context(rank)
    uid = [capi.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    gather = torch_allgather()
    pb = PoissonMG(ctx, *box, nlevels, order, dist=(rank, world, gather), coarse_rtol=1e-15, peer=peer)
    pb.step()
    trace = [pb.residual_norm()]
    for _ in range(ncycles - 1):
        pb.mg_solve()
        trace.append(pb.residual_norm())
    top = pb.hier.levels[-1]
    own = pb.layout[-1].owned.astype(bool)
    keys = top.lattice_key(np.arange(pb.n))[own]
    eps = pb.EPS.get()[own]
    # level operators: every rank's partial matrices, by lattice key, for the summed-operator check
    mats = []
    for l in range(nlevels):
        A = pb.KK[l].to_scipy().tocoo()
        k = pb.hier.levels[l].lattice_key(np.arange(pb.ndofs[l]))
        mats.append((k[A.row], k[A.col], A.data))
    assert not ctx.peer_error(), "a wait of the peer-memory exchange timed out"
    allr = gather((keys, eps, trace, mats, pb.mg.coarse_iterations(), ctx.launches()))
    if rank == 0:
        np.save(out_path, np.array(allr, dtype=object), allow_pickle=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
