"""CPU mirror (numpy/scipy + gloo) of the sharded multigrid run, used by the world_size-2 CPU test:
the same host-side layout (femus_b200.dist, femus_b200.hostapi local hierarchies) and the same
sequence of operations as the device path (partial matrices per rank, interface sums, ownership
masks), with torch.distributed `gloo` all_reduce in place of the library's ncclAllReduce.
Test infrastructure only."""
import numpy as np
import scipy.sparse as sp

from femus_b200 import hostapi
from femus_b200 import dist as distlayout
from oracle import fe_hex, mg as omg


class NumpyRank:
    def __init__(self, box, nlevels, order, rank, world, allgather, allreduce, fsrc=1.0, omega=0.5):
        self.rank, self.world, self.allreduce = rank, world, allreduce
        self.order, self.nl, self.omega = order, nlevels, omega
        H = hostapi.HostHierarchy(*box, nlevels, nprocs=world, local_rank=rank)
        self.H = H
        lv = H.levels
        self.nd = [L.ndofs(order) for L in lv]
        self.lay = [distlayout.level_layout(lv[l], self.nd[l], rank, allgather) for l in range(nlevels)]
        self.bdc = [L.bdc(order) for L in lv]
        # prolongators (local), Dirichlet rows/cols zeroed
        self.P = [None] * nlevels
        for l in range(1, nlevels):
            rp, ci, v, shp = H.prolongator(l, order)
            P = sp.csr_matrix((v, ci, rp), shape=shp)
            Df = sp.diags((self.bdc[l] > 1.5).astype(float))
            Dc = sp.diags((self.bdc[l - 1] > 1.5).astype(float))
            self.P[l] = (Df @ P @ Dc).tocsr()
        # partial finest matrix + rhs from this rank's elements
        top = lv[-1]
        nve = 27 if order == "biquadratic" else 8
        d = top.system_dofs(order)
        X = top.xyz[:, top.conn[:, :nve]].transpose(1, 0, 2)
        F, B = fe_hex.poisson_elements(order, X, np.zeros((top.nel, nve)), fsrc)
        n = self.nd[-1]
        rows = np.repeat(d, nve, axis=1).ravel()
        cols = np.tile(d, (1, nve)).ravel()
        A = sp.csr_matrix((B.ravel(), (rows, cols)), shape=(n, n))
        rhs = np.zeros(n)
        np.add.at(rhs, d.ravel(), F.ravel())
        self.rhs = self.halo_sum(nlevels - 1, rhs)
        # Galerkin chain on the partial matrices (no communication), then penalty with ownership
        self.A = [None] * nlevels
        self.A[-1] = A
        for l in range(nlevels - 1, 0, -1):
            self.A[l - 1] = (self.P[l].T @ self.A[l] @ self.P[l]).tocsr()
        self.dinv, self.R = [None] * nlevels, [None] * nlevels
        for l in range(nlevels):
            idx = np.nonzero(self.bdc[l] < 1.5)[0]
            A = self.A[l].tolil()
            for r in idx:
                A.rows[r], A.data[r] = [r], [1.0 if self.lay[l].owned[r] else 0.0]
            self.A[l] = A.tocsr()
            self.dinv[l] = 1.0 / self.halo_sum(l, self.A[l].diagonal())
            if l > 0:
                self.R[l] = (sp.diags(self.lay[l].owned.astype(float)) @ self.P[l]).T.tocsr()

    def halo_sum(self, l, v):
        lay = self.lay[l]
        buf = np.zeros(lay.n_packed)
        buf[lay.pos] = v[lay.idx]
        buf = self.allreduce(buf)
        v = v.copy()
        v[lay.idx] = buf[lay.pos]
        return v

    def dot(self, l, x, y):
        own = self.lay[l].owned.astype(bool)
        return float(self.allreduce(np.array([x[own] @ y[own]]))[0])

    def resid(self, l, b, x):
        w = 1.0 / self.lay[l].mult
        return self.halo_sum(l, w * b - self.A[l] @ x)

    def smooth(self, l, x, b, n, zero_guess):
        for k in range(n):
            if zero_guess and k == 0:
                x = self.omega * self.dinv[l] * b
            else:
                x = x + self.omega * self.dinv[l] * self.resid(l, b, x)
        return x

    def halo_sum_scalars(self, l, v, scal):
        """b2_halo_sum_scalars: interface values and the scalars in ONE all_reduce."""
        lay = self.lay[l]
        buf = np.zeros(lay.n_packed + len(scal))
        buf[lay.pos] = v[lay.idx]
        buf[lay.n_packed:] = scal
        buf = self.allreduce(buf)
        v = v.copy()
        v[lay.idx] = buf[lay.pos]
        return v, buf[lay.n_packed:].copy()

    def coarse(self, b, rtol=1e-15, maxit=5000):
        """Single-reduction (Chronopoulos-Gear) Jacobi-PCG as b2_mg.cu coarse_solve: per iteration one
        partial product w = A u, the local sums (r.u over owned, w.u over ALL local entries -- u is
        complete on every holder, so the partial w can be used --, r.r over owned) and one collective
        that completes w on the interface and reduces the three scalars."""
        idx = np.nonzero(self.bdc[0] < 1.5)[0]
        own = self.lay[0].owned.astype(bool)
        x = np.zeros_like(b)
        x[idx] = b[idx]
        r = self.resid(0, b, x)
        u = self.dinv[0] * r
        p = np.zeros_like(b)
        s = np.zeros_like(b)
        gamma = alpha = bb = 0.0
        for it in range(maxit + 1):
            w = self.A[0] @ u
            scal = [r[own] @ u[own], w @ u, r[own] @ r[own]] + ([b[own] @ b[own]] if it == 0 else [])
            w, red = self.halo_sum_scalars(0, w, scal)
            gn, delta, rr = red[0], red[1], red[2]
            if it == 0:
                bb = red[3]
            if bb == 0.0 or not (rr > rtol * rtol * bb):
                break
            if it == 0:
                beta, alpha = 0.0, gn / delta
            else:
                beta = gn / gamma
                alpha = gn / (delta - beta * gn / alpha)
            gamma = gn
            p = u + beta * p
            s = w + beta * s
            x += alpha * p
            r -= alpha * s
            u = self.dinv[0] * r
        return x

    def vcycle(self, l, b):
        if l == 0:
            return self.coarse(b)
        x = self.smooth(l, None, b, 1, True)
        r = self.resid(l, b, x)
        bc = self.halo_sum(l - 1, self.R[l] @ r)
        xc = self.vcycle(l - 1, bc)
        x = x + self.P[l] @ xc
        return self.smooth(l, x, b, 1, False)

    def mg_solve_trace(self, ncyc):
        top = self.nl - 1
        res = self.rhs.copy()
        eps = np.zeros_like(res)
        idx = np.nonzero(self.bdc[top] < 1.5)[0]
        free = self.bdc[top] > 1.1
        trace = []
        for _ in range(ncyc):
            res[idx] = 0.0
            e = self.vcycle(top, res)
            res = self.resid(top, res, e)
            eps += e
            m = np.where(free, res, 0.0)
            trace.append(np.sqrt(self.dot(top, m, m)))
        return trace, eps
