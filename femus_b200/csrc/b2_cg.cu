// The coarse-level Jacobi-PCG of the V-cycle as ONE persistent cooperative kernel.
// (Reference: PREONLY + LU on level 0, PetscPreconditioner.cpp:147-160; here the exact coarse solve is a conjugate
// gradient run to round-off, b2_mg.cu.)  The host-driven loop costs 5 launches per iteration (SpMV, dot products,
// interface sum, update) on a problem of 10^4 - 10^5 dofs: ~25 us per iteration of pure launch latency on one GPU,
// ~45 us with the interface exchange, 80 - 150 iterations per V-cycle.  Here the whole iteration -- CSR product, the
// three dot products, the interface sum with the other ranks through peer memory, the vector update and the
// convergence test -- runs inside one kernel; iterations are separated by grid-wide barriers (2 on one rank, 3 with
// the exchange, which itself needs none), the scalars never leave the chip.
//
// Same recurrence as the host loop (single-reduction Chronopoulos-Gear PCG, b2_mg.cu): w = A u; gamma = (r,u),
// delta = (w,u), rr = (r,r) in one reduction; beta = gamma / gamma_old; alpha = gamma / (delta - beta gamma / alpha_old);
// p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s; u = D^-1 r.  Every reduction has a fixed order
// (thread -> warp shuffle -> block -> all blocks, ranks in ascending order): the result is run-to-run bit-identical and
// equal on all ranks.  The convergence test runs every iteration (the host loop looks every 8th).
#include <cooperative_groups.h>
#include "b2_common.cuh"
#include "b2_peer.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kCgBlock = 1024;     // few, large blocks: the grid-wide barrier costs per block, the product wants every thread slot of the SM

struct CgArgs {
  int64_t n;
  const int64_t* rowptr;
  const int32_t* col;
  const double* val;
  const double* dinv;
  const double* b;
  const uint8_t* owned;       // null: every entry is owned
  double *x, *r, *u, *p, *s, *w;
  double* partial;            // [gridDim][4]
  double* out;                // [4]: iterations, rr, bb
  double* trace;              // null, or [5]: cycles of the phases of iteration 5 seen by thread 0 (product, barrier, totals / exchange, update, barrier)
  double rtol2;
  int maxit;
  // interface exchange (nranks == 1: none)
  int nranks, me;
  void* const* bases;         // device array [nranks]: every rank's inbox
  void* base;                 // this rank's inbox
  int64_t slot;
  unsigned long long epoch0;  // exchanges done before this solve
  int64_t n_if;
  const int32_t* idx;
  const int64_t* hold_ptr;
  const int32_t* hold_rank;
  const int32_t* hold_pos;
  const int32_t* hold_spos;
  int* err;
};

__global__ void __launch_bounds__(kCgBlock, 1) cg_persistent_kernel(const CgArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[kCgBlock / 32][4];
  __shared__ double tot[4];
  __shared__ int s_abort;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (int64_t)gridDim.x * blockDim.x;
  const int sub = lane & 3;                                   // 4 lanes per row
  const int64_t grow = gtid >> 2, nrows_step = gthreads >> 2;
  double gamma_old = 0.0, alpha_old = 0.0, bb = 0.0;
  int it = 0;
  int aborted = 0;
  for (;; it++) {
    // ---- w = A u on this rank's rows; partial sums of (r,u) | owned, (w,u) | all, (r,r) | owned (, (b,b) | owned)
    long long tk[6];
    const bool tr = a.trace && it == 5 && gtid == 0;
    if (tr) tk[0] = clock64();
    double acc[4] = {0., 0., 0., 0.};
    // (the eight rows of a warp leave the loop together: the shuffles below name all 32 lanes)
    for (int64_t i0 = grow - (lane >> 2); i0 < a.n; i0 += nrows_step) {
      const int64_t i = i0 + (lane >> 2);
      double t = 0.0;
      if (i < a.n) {
        // latency-bound (a few 10^4 rows, everything in L2): eight independent (value, column) loads and eight gathers
        // in flight per lane instead of one dependent chain
        const int64_t q1 = a.rowptr[i + 1];
        int64_t q = a.rowptr[i] + sub;
        for (; q + 28 < q1; q += 32) {
          int32_t cc[8];
          double vv[8], uu[8];
#pragma unroll
          for (int k = 0; k < 8; k++) { cc[k] = a.col[q + 4 * k]; vv[k] = a.val[q + 4 * k]; }
#pragma unroll
          for (int k = 0; k < 8; k++) uu[k] = a.u[cc[k]];
#pragma unroll
          for (int k = 0; k < 8; k++) t = fma(vv[k], uu[k], t);
        }
        for (; q < q1; q += 4) t = fma(a.val[q], a.u[a.col[q]], t);
      }
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      if (sub == 0 && i < a.n) {
        a.w[i] = t;
        const double ui = a.u[i];
        acc[1] = fma(t, ui, acc[1]);
        if (!a.owned || a.owned[i]) {
          const double ri = a.r[i];
          acc[0] = fma(ri, ui, acc[0]);
          acc[2] = fma(ri, ri, acc[2]);
          if (it == 0) acc[3] = fma(a.b[i], a.b[i], acc[3]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
    if (lane == 0)
      for (int k = 0; k < 4; k++) sh[wib][k] = acc[k];
    __syncthreads();
    if (threadIdx.x < 4) {
      double t = 0.0;
      for (int ww = 0; ww < kCgBlock / 32; ww++) t += sh[ww][threadIdx.x];
      a.partial[4 * blockIdx.x + threadIdx.x] = t;
    }
    if (tr) tk[1] = clock64();
    grid.sync();
    if (tr) tk[2] = clock64();
    // ---- this rank's totals: every block sums the partials of all blocks in the same order
    if (wib == 0) {
      double t[4] = {0., 0., 0., 0.};
      for (int k = lane; k < (int)gridDim.x; k += 32)
        for (int j = 0; j < 4; j++) t[j] += __ldcg(&a.partial[4 * k + j]);
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t[j] += __shfl_xor_sync(0xffffffffu, t[j], o);
      if (lane == 0)
        for (int j = 0; j < 4; j++) tot[j] = t[j];
    }
    __syncthreads();
    if (a.nranks > 1) {
      // ---- interface sum of w and global sums of the scalars through peer memory (cells and protocol: b2_halo.cu): the
      // thread that owns an interface entry sends its value to the other holders, waits for theirs, sums in rank order
      const unsigned long long epoch = a.epoch0 + (unsigned long long)it + 1ull;
      const int parity = (int)(epoch & 1ull);
      const unsigned flag = b2_peer_flag(epoch);
      if (blockIdx.x == 0 && threadIdx.x < 4)
        for (int r = 0; r < a.nranks; r++)
          if (r != a.me) b2_peer_store(b2_peer_cell(a.bases[r], a.slot, a.nranks, parity, a.me, threadIdx.x), tot[threadIdx.x], flag);
      for (int64_t k = gtid; k < a.n_if; k += gthreads) {
        const int32_t d = a.idx[k];
        const double own = a.w[d];
        const int64_t h0 = a.hold_ptr[k], h1 = a.hold_ptr[k + 1];
        for (int64_t q = h0; q < h1; q++)
          if (a.hold_rank[q] != a.me)
            b2_peer_store(b2_peer_cell(a.bases[a.hold_rank[q]], a.slot, a.nranks, parity, a.me, kPeerScal + a.hold_spos[q]), own, flag);
        double t = 0.0;
        for (int64_t q = h0; q < h1; q++) {
          const int r = a.hold_rank[q];
          const double v = r == a.me ? own : b2_peer_wait(b2_peer_cell(a.base, a.slot, a.nranks, parity, r, kPeerScal + a.hold_pos[q]), flag, a.err);
          t = q == h0 ? v : t + v;
        }
        a.w[d] = t;
      }
      __syncthreads();      // (block 0: the local totals were read for sending before they are replaced)
      if (threadIdx.x < 4) {
        const double mine = tot[threadIdx.x];
        double t = 0.0;
        for (int r = 0; r < a.nranks; r++) {
          const double v = r == a.me ? mine : b2_peer_wait(b2_peer_cell(a.base, a.slot, a.nranks, parity, r, threadIdx.x), flag, a.err);
          t = r == 0 ? v : t + v;
        }
        tot[threadIdx.x] = t;
      }
      grid.sync();
      if (threadIdx.x == 0) s_abort = *reinterpret_cast<volatile int*>(a.err) != 0;
      __syncthreads();
      aborted = s_abort;
    } else {
      __syncthreads();
    }
    if (tr) tk[3] = clock64();
    const double gn = tot[0], delta = tot[1], rr = tot[2];
    if (it == 0) bb = tot[3];
    if (aborted || bb == 0.0 || !(rr > a.rtol2 * bb) || it == a.maxit) {
      if (gtid == 0) {
        a.out[0] = (double)it;
        a.out[1] = rr;
        a.out[2] = bb;
      }
      break;
    }
    // ---- update
    double beta = 0.0, alpha;
    if (it == 0) {
      alpha = delta != 0.0 ? gn / delta : 0.0;
    } else {
      beta = gamma_old != 0.0 ? gn / gamma_old : 0.0;
      const double den = delta - (alpha_old != 0.0 ? beta * gn / alpha_old : 0.0);
      alpha = den != 0.0 ? gn / den : 0.0;
    }
    gamma_old = gn;
    alpha_old = alpha;
    for (int64_t i = gtid; i < a.n; i += gthreads) {
      const double pi = fma(beta, a.p[i], a.u[i]);
      const double si = fma(beta, a.s[i], a.w[i]);
      a.p[i] = pi;
      a.s[i] = si;
      a.x[i] = fma(alpha, pi, a.x[i]);
      const double ri = fma(-alpha, si, a.r[i]);
      a.r[i] = ri;
      a.u[i] = a.dinv[i] * ri;
    }
    if (tr) tk[4] = clock64();
    grid.sync();
    if (tr) {
      tk[5] = clock64();
      for (int k = 0; k < 5; k++) a.trace[k] = (double)(tk[k + 1] - tk[k]);
    }
  }
}

}  // namespace

/* Runs the iteration loop of the coarse PCG in one cooperative launch.  On entry x holds the start vector, r = b - A x
 * (interface entries complete), u = D^-1 r, p = s = 0.  *ran = 0 (nothing done) when the device cannot co-schedule the
 * grid or a sharded level has no peer-memory exchange; the caller then runs its host-driven loop. */
int b2_cg_persistent(b2_ctx* c, const b2_csr* A, const double* dinv, const double* b, const uint8_t* owned, const b2_halo* halo,
                     double* x, double* r, double* u, double* p, double* s, double* w, double* partial, int partial_blocks,
                     double* out4, double rtol, int maxit, int* its, int* ran) {
  *ran = 0;
  *its = 0;
  if (c->nranks > 1 && !(halo && c->halo_peer && c->peer_base && halo->hold_ptr)) return 0;
  int coop = 0;
  B2_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
  if (!coop) return 0;
  int per_sm = 0;
  B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_persistent_kernel, kCgBlock, 0));
  if (per_sm < 1) return 0;
  int grid = c->sm_count * per_sm;
  const int64_t want = (A->nrows * 4 + kCgBlock - 1) / kCgBlock;      // 4 lanes per row
  if (grid > want) grid = (int)(want > 0 ? want : 1);
  if (grid > partial_blocks) grid = partial_blocks;
  CgArgs a = {};
  a.n = A->nrows;
  a.rowptr = A->rowptr;
  a.col = A->col;
  a.val = A->val;
  a.dinv = dinv;
  a.b = b;
  a.owned = owned;
  a.x = x; a.r = r; a.u = u; a.p = p; a.s = s; a.w = w;
  a.partial = partial;
  a.out = out4;
  a.trace = getenv("B2_CG_TRACE") ? out4 + 8 : nullptr;
  a.rtol2 = rtol * rtol;
  a.maxit = maxit;
  a.nranks = c->nranks;
  a.me = c->rank;
  if (c->nranks > 1) {
    a.bases = (void* const*)c->d_peer_base;
    a.base = c->peer_local;
    a.slot = c->peer_slot;
    a.epoch0 = c->peer_epoch;
    a.n_if = halo->n_if;
    a.idx = halo->idx;
    a.hold_ptr = halo->hold_ptr;
    a.hold_rank = halo->hold_rank;
    a.hold_pos = halo->hold_pos;
    a.hold_spos = halo->hold_spos;
    a.err = c->peer_err;
  }
  B2_CUDA(cudaMemsetAsync(out4, 0, 4 * sizeof(double), c->stream));
  void* params[] = {(void*)&a};
  B2_CUDA(cudaLaunchCooperativeKernel((void*)cg_persistent_kernel, dim3(grid), dim3(kCgBlock), params, 0, c->stream));
  c->launches++;
  double h[4];
  B2_TRY(b2_download(c, h, out4, 4));      // synchronises: the exchange counter of the host must follow the kernel's
  *its = (int)h[0];
  if (a.trace) {
    double t[5];
    B2_TRY(b2_download(c, t, a.trace, 5));
    fprintf(stderr, "b2_cg_persistent: n %lld nnz %lld grid %d x %d, %d iterations; cycles of iteration 5: product %.0f barrier %.0f totals/exchange %.0f update %.0f barrier %.0f\n",
            (long long)A->nrows, (long long)A->nnz, grid, kCgBlock, *its, t[0], t[1], t[2], t[3], t[4]);
  }
  if (c->nranks > 1) c->peer_epoch += (unsigned long long)*its + 1ull;
  if (c->nranks > 1) {
    int perr = 0;
    B2_TRY(b2_download(c, &perr, c->peer_err, 1));
    B2_CHECK(!perr, "coarse PCG: a wait of the peer-memory exchange timed out (a rank stopped taking part)");
  }
  *ran = 1;
  return 0;
}
