"""Multi-GPU parity: the z-slab sharded run (one rank per GPU, partial matrices, interface sums through peer memory
over NVLink, or by ncclAllReduce, inside the library) against the serial CPU oracle on the same mesh.  Needs >= 2 GPUs."""
import os
import socket
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("peer", [True, False], ids=["peer_memory", "nccl_allreduce"])
@pytest.mark.parametrize("order,box,nl,world", [("biquadratic", (2, 2, 4), 3, 2), ("linear", (2, 3, 4), 3, 2),
                                                 ("biquadratic", (2, 2, 4), 2, 4)])
def test_sharded_vcycle_matches_serial_oracle(order, box, nl, world, peer):
    if ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import scipy.sparse as sp
    from femus_b200 import hostapi
    from oracle import mesh_box as mb, mg
    from tests import dist_worker
    ncyc = 4
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "res.npy")
        mp.spawn(dist_worker.run_gpu_rank, args=(world, free_port(), box, nl, order, ncyc, out, peer), nprocs=world, join=True)
        res = np.load(out, allow_pickle=True)
    # serial oracle; global dof <-> lattice key through the host layer (bit-identical numbering)
    lv = mb.build_hierarchy(*box, nl)
    H = mg.Hierarchy(lv, order)
    trace_ref, eps_ref = H.mg_solve_trace(ncyc)
    G = hostapi.HostHierarchy(*box, nl)
    for l in range(nl):
        n = G.levels[l].ndofs(order)
        gk = G.levels[l].lattice_key(np.arange(n))
        srt = np.argsort(gk)
        # summed partial operators (before penalty rows are compared: after MGSetLevel) == oracle level operators
        rows, cols, vals = [], [], []
        for r in range(world):
            kr, kc, v = res[r][3][l]
            rows.append(srt[np.searchsorted(gk[srt], kr)])
            cols.append(srt[np.searchsorted(gk[srt], kc)])
            vals.append(v)
        S = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
        D = (S - H.A[l]).tocsr()
        assert np.abs(D.data).max() <= 1e-12 * np.abs(H.A[l].data).max(), f"level {l} operator"
    for r in range(world):
        for a, b in zip(res[r][2], trace_ref):
            assert abs(a - b) <= 1e-12 * trace_ref[0], (res[r][2], trace_ref)
    n = G.levels[-1].ndofs(order)
    gk = G.levels[-1].lattice_key(np.arange(n))
    srt = np.argsort(gk)
    eps = np.full(n, np.nan)
    for r in range(world):
        keys, e = res[r][0], res[r][1]
        eps[srt[np.searchsorted(gk[srt], keys)]] = e
    assert not np.isnan(eps).any(), "every dof is owned by exactly one rank"
    assert np.abs(eps - eps_ref).max() <= 1e-11 * np.abs(eps_ref).max()
    assert sum(len(res[r][0]) for r in range(world)) == n
