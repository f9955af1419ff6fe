// Geometric multigrid V-cycle on device CSR operators.  Replaces what
// LinearEquationSolverPetsc::{MGInit, MGSetLevel, MGSolve, MGClear} configure in PETSc's PCMG
// (reference src/08_algebra.../03_solvers_with_preconditioner/LinearEquationSolverPetsc.cpp:185-353):
// multiplicative V, zero initial guess on every level, npre/npost sweeps of
// KSPRICHARDSON(scale omega)+PCJACOBI (:516-519, PetscPreconditioner.cpp:209-212), restriction
// with P^T (PCMGSetRestriction(..., PP), :277), interpolation with P (:276), level operators with
// Dirichlet rows set to identity (SetPenalty, :428-436).  The coarse solve (reference: PREONLY +
// MUMPS LU, PetscPreconditioner.cpp:147-160) is a Jacobi-preconditioned CG run to a tight
// relative residual.
#include "b2_common.cuh"

struct b2_mg_level {
  b2_csr* A = nullptr;      // borrowed, penalised in place
  b2_csr* P = nullptr;      // borrowed
  b2_csr* R = nullptr;      // owned: explicit transpose of P
  b2_vec *dinv = nullptr, *x = nullptr, *t = nullptr, *b = nullptr, *r = nullptr;
  int32_t* bdc = nullptr;   // device copy of the Dirichlet row list
  int64_t nbdc = 0;
  b2_halo* halo = nullptr;  // borrowed: distributed layout of this level's vectors (null: single rank)
  int npre = 1, npost = 1;
  double omega = 0.5;
};

struct b2_mg {
  b2_ctx* ctx;
  int nlevels;
  std::vector<b2_mg_level> L;
  double coarse_rtol = 1e-14;
  int coarse_maxit = 5000;
  int coarse_its = 0;
  // PCG work vectors on level 0 and device scalars
  b2_vec *p = nullptr, *q = nullptr, *z = nullptr;
  double* scal = nullptr;   // [8]: 0 rz, 1 pq, 2 rz_new, 3 rr, 4 bb
};

namespace {

constexpr int kBlock = 256;

__global__ void recip_kernel(int64_t n, const double* __restrict__ d, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = 1.0 / d[i];
}
// first Richardson-Jacobi sweep from a zero guess: x = omega * dinv * b
__global__ void jacobi_first_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ b,
                                    double* __restrict__ x, double omega) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = (omega * dinv[i]) * b[i];
}
__global__ void fill_idx_kernel(double* __restrict__ v, const int32_t* __restrict__ idx, int64_t n, double a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[idx[i]] = a;
}
__global__ void copy_idx_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                const int32_t* __restrict__ idx, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[idx[i]] = src[idx[i]];
}

// ---- PCG kernels with device-resident scalars -------------------------------------------------
// z = dinv .* r ; p = z            (start)
__global__ void pcg_start_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ r,
                                 double* __restrict__ z, double* __restrict__ p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double zi = dinv[i] * r[i];
    z[i] = zi;
    p[i] = zi;
  }
}
// alpha = rz/pq ; x += alpha p ; r -= alpha q ; z = dinv .* r
__global__ void pcg_update_kernel(int64_t n, const double* __restrict__ scal, int cur, const double* __restrict__ dinv,
                                  const double* __restrict__ p, const double* __restrict__ q, double* __restrict__ x,
                                  double* __restrict__ r, double* __restrict__ z) {
  const double pq = scal[1];
  const double alpha = pq != 0.0 ? scal[cur] / pq : 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    z[i] = dinv[i] * ri;
  }
}
// beta = rz_new/rz ; p = z + beta p.  r.z is double buffered in slots 0 and 2 (`cur` = current).
__global__ void pcg_dir_kernel(int64_t n, const double* __restrict__ scal, int cur, const double* __restrict__ z,
                               double* __restrict__ p) {
  const double rz = scal[cur], rzn = scal[cur ^ 2];
  const double beta = rz != 0.0 ? rzn / rz : 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = fma(beta, p[i], z[i]);
}

// x += omega * dinv .* r   (second half of a distributed Richardson-Jacobi sweep)
__global__ void jacobi_update_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ r,
                                     double* __restrict__ x, double omega) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = fma(omega * dinv[i], r[i], x[i]);
}

int vec_grid(b2_ctx* c, int64_t n) { return b2_grid_for(c, n, kBlock, 8); }

// r = b - A x with every entry complete on every rank that holds it
int level_resid(b2_mg_level& L, const b2_vec* b, const b2_vec* x, b2_vec* r) {
  if (!L.halo) return b2_csr_resid(L.A, b, x, r);
  B2_TRY(b2_csr_resid_w(L.A, b->d, L.halo->invmult, x->d, r->d));
  return b2_halo_sum(L.halo, r);
}
// y = A x, complete
int level_spmv(b2_mg_level& L, const b2_vec* x, b2_vec* y) {
  B2_TRY(b2_csr_spmv(L.A, x, y));
  return L.halo ? b2_halo_sum(L.halo, y) : 0;
}
const uint8_t* owned(const b2_mg_level& L) { return L.halo ? L.halo->owned : nullptr; }

int smooth(b2_mg* mg, int l, int nsweeps, bool zero_guess) {
  b2_mg_level& L = mg->L[l];
  b2_ctx* c = mg->ctx;
  const int64_t n = L.A->nrows;
  int done = 0;
  if (zero_guess && nsweeps > 0) {
    B2_LAUNCH(c, jacobi_first_kernel, vec_grid(c, n), kBlock, 0, n, L.dinv->d, L.b->d, L.x->d, L.omega);
    done = 1;
  } else if (zero_guess) {
    B2_TRY(b2_vec_zero(L.x));
  }
  for (; done < nsweeps; done++) {
    if (L.halo) {      // interface rows need the other ranks' part of A x before the update
      B2_TRY(level_resid(L, L.b, L.x, L.t));
      B2_LAUNCH(c, jacobi_update_kernel, vec_grid(c, n), kBlock, 0, n, L.dinv->d, L.t->d, L.x->d, L.omega);
    } else {
      B2_TRY(b2_csr_jacobi_sweep(L.A, L.dinv, L.b, L.x, L.t, L.omega));
      std::swap(L.x, L.t);
    }
  }
  return 0;
}

int coarse_solve(b2_mg* mg) {
  // Jacobi-PCG on level 0: solves A x = b with x, b the level-0 work vectors.
  b2_mg_level& L = mg->L[0];
  b2_ctx* c = mg->ctx;
  const int64_t n = L.A->nrows;
  const int g = vec_grid(c, n);
  double* s = mg->scal;
  // x0 = b on Dirichlet rows (identity rows), 0 elsewhere; r = b - A x0
  B2_TRY(b2_vec_zero(L.x));
  if (L.nbdc) B2_LAUNCH(c, copy_idx_kernel, vec_grid(c, L.nbdc), kBlock, 0, L.x->d, L.b->d, L.bdc, L.nbdc);
  B2_TRY(level_resid(L, L.b, L.x, L.r));
  const uint8_t* own = owned(L);
  B2_TRY(b2_dev_dot(c, L.b->d, L.b->d, n, s + 4, own));
  B2_LAUNCH(c, pcg_start_kernel, g, kBlock, 0, n, L.dinv->d, L.r->d, mg->z->d, mg->p->d);
  B2_TRY(b2_dev_dot(c, L.r->d, mg->z->d, n, s + 0, own));
  B2_TRY(b2_dev_dot(c, L.r->d, L.r->d, n, s + 3, own));
  B2_TRY(b2_allreduce_sum(c, s + 3, 2));    // slots 3 (rr) and 4 (bb) together
  B2_TRY(b2_allreduce_sum(c, s + 0, 1));
  double h[8];
  B2_TRY(b2_download(c, h, s, 8));
  const double bb = h[4];
  mg->coarse_its = 0;
  if (bb == 0.0 || h[3] <= mg->coarse_rtol * mg->coarse_rtol * bb) return 0;
  int cur = 0;
  const int check_every = 8;
  for (int it = 1; it <= mg->coarse_maxit; it++) {
    B2_TRY(level_spmv(L, mg->p, mg->q));
    B2_TRY(b2_dev_dot(c, mg->p->d, mg->q->d, n, s + 1, own));
    B2_TRY(b2_allreduce_sum(c, s + 1, 1));
    B2_LAUNCH(c, pcg_update_kernel, g, kBlock, 0, n, s, cur, L.dinv->d, mg->p->d, mg->q->d, L.x->d, L.r->d, mg->z->d);
    B2_TRY(b2_dev_dot(c, L.r->d, mg->z->d, n, s + (cur ^ 2), own));
    B2_TRY(b2_allreduce_sum(c, s + (cur ^ 2), 1));
    B2_LAUNCH(c, pcg_dir_kernel, g, kBlock, 0, n, s, cur, mg->z->d, mg->p->d);
    cur ^= 2;
    mg->coarse_its = it;
    if (it % check_every == 0) {
      B2_TRY(b2_dev_dot(c, L.r->d, L.r->d, n, s + 3, own));
      B2_TRY(b2_allreduce_sum(c, s + 3, 1));
      B2_TRY(b2_download(c, h, s, 8));
      if (!(h[3] > mg->coarse_rtol * mg->coarse_rtol * bb)) break;   // also leaves on NaN
    }
  }
  return 0;
}

int vcycle(b2_mg* mg, int l) {
  // solves (approximately) A_l x_l = b_l, zero initial guess; operands are the level work vectors
  b2_mg_level& L = mg->L[l];
  if (l == 0) return coarse_solve(mg);
  B2_TRY(smooth(mg, l, L.npre, true));
  B2_TRY(level_resid(L, L.b, L.x, L.r));
  b2_mg_level& C = mg->L[l - 1];
  B2_TRY(b2_csr_spmv(L.R, L.r, C.b));        // R = P^T restricted to the fine rows this rank owns
  if (C.halo) B2_TRY(b2_halo_sum(C.halo, C.b));
  B2_TRY(vcycle(mg, l - 1));
  B2_TRY(b2_csr_spmv_add(L.P, C.x, L.x));
  B2_TRY(smooth(mg, l, L.npost, false));
  return 0;
}

}  // namespace

extern "C" {

int b2_mg_create(b2_ctx* c, int nlevels, b2_mg** out) {
  *out = nullptr;
  B2_CHECK(c && nlevels >= 1, "b2_mg_create: bad arguments");
  b2_mg* mg = new b2_mg();
  mg->ctx = c;
  mg->nlevels = nlevels;
  mg->L.resize(nlevels);
  B2_TRY(b2_malloc(c, &mg->scal, 8));
  B2_CUDA(cudaMemsetAsync(mg->scal, 0, 8 * sizeof(double), c->stream));
  *out = mg;
  return 0;
}

int b2_mg_set_level(b2_mg* mg, int level, b2_csr* A, b2_csr* P, const int32_t* bdc_idx, int64_t nbdc, int npre,
                    int npost, double omega) {
  B2_CHECK(level >= 0 && level < mg->nlevels && A, "b2_mg_set_level: bad level %d", level);
  B2_CHECK(A->nrows == A->ncols, "b2_mg_set_level: operator must be square");
  B2_CHECK(level == 0 || (P && P->nrows == A->nrows), "b2_mg_set_level: prolongator shape mismatch");
  b2_ctx* c = mg->ctx;
  b2_mg_level& L = mg->L[level];
  const bool newP = (level > 0) && (L.P != P || !L.R);
  L.A = A;
  L.P = level ? P : nullptr;
  L.npre = npre;
  L.npost = npost;
  L.omega = omega;
  const int64_t n = A->nrows;
  if (!L.dinv) {
    B2_TRY(b2_vec_create(c, n, &L.dinv));
    B2_TRY(b2_vec_create(c, n, &L.x));
    B2_TRY(b2_vec_create(c, n, &L.t));
    B2_TRY(b2_vec_create(c, n, &L.b));
    B2_TRY(b2_vec_create(c, n, &L.r));
    if (level == 0) {
      B2_TRY(b2_vec_create(c, n, &mg->p));
      B2_TRY(b2_vec_create(c, n, &mg->q));
      B2_TRY(b2_vec_create(c, n, &mg->z));
    }
  }
  if (L.bdc) { b2_free(c, L.bdc, (size_t)L.nbdc); L.bdc = nullptr; }
  L.nbdc = nbdc;
  B2_CHECK(!L.halo || L.halo->n_local == n, "b2_mg_set_level: layout of level %d has %lld dofs, operator %lld", level,
           (long long)(L.halo ? L.halo->n_local : 0), (long long)n);
  if (nbdc) {
    B2_TRY(b2_malloc(c, &L.bdc, (size_t)nbdc));
    B2_TRY(b2_upload(c, L.bdc, bdc_idx, (size_t)nbdc));
    B2_TRY(b2_csr_zero_rows_dev(A, L.bdc, nbdc, 1.0, owned(L)));     // SetPenalty
  }
  B2_TRY(b2_csr_diag(A, L.dinv));
  if (L.halo) B2_TRY(b2_halo_sum(L.halo, L.dinv));       // diagonal of the summed operator
  B2_LAUNCH(c, recip_kernel, vec_grid(c, n), kBlock, 0, n, L.dinv->d, L.dinv->d);
  if (newP) {   // the explicit restriction R = P^T is rebuilt only when P changes
    if (L.R) { b2_csr_destroy(L.R); L.R = nullptr; }
    B2_TRY(b2_csr_transpose(P, &L.R));
    if (L.halo) B2_TRY(b2_csr_zero_cols_notowned(L.R, L.halo->owned));   // shared fine rows restrict once
  }
  return 0;
}

int b2_mg_set_level_halo(b2_mg* mg, int level, b2_halo* halo) {
  B2_CHECK(level >= 0 && level < mg->nlevels, "b2_mg_set_level_halo: bad level %d", level);
  B2_CHECK(!mg->L[level].R, "b2_mg_set_level_halo: call it before b2_mg_set_level");
  mg->L[level].halo = halo;
  return 0;
}

int b2_mg_set_coarse(b2_mg* mg, double rtol, int maxit) {
  mg->coarse_rtol = rtol;
  mg->coarse_maxit = maxit;
  return 0;
}

int b2_mg_vcycle(b2_mg* mg, const b2_vec* rhs, b2_vec* x) {
  const int top = mg->nlevels - 1;
  b2_mg_level& L = mg->L[top];
  B2_CHECK(L.A && rhs->n >= L.A->nrows && x->n >= L.A->nrows, "b2_mg_vcycle: level %d not set or vectors too short", top);
  for (int l = 0; l <= top; l++) B2_CHECK(mg->L[l].A, "b2_mg_vcycle: level %d not set", l);
  B2_CUDA(cudaMemcpyAsync(L.b->d, rhs->d, (size_t)L.A->nrows * sizeof(double), cudaMemcpyDeviceToDevice, mg->ctx->stream));
  B2_TRY(vcycle(mg, top));
  B2_CUDA(cudaMemcpyAsync(x->d, L.x->d, (size_t)L.A->nrows * sizeof(double), cudaMemcpyDeviceToDevice, mg->ctx->stream));
  return 0;
}

int b2_mg_solve(b2_mg* mg, b2_vec* res, b2_vec* eps) {
  const int top = mg->nlevels - 1;
  b2_mg_level& L = mg->L[top];
  b2_ctx* c = mg->ctx;
  B2_CHECK(L.A, "b2_mg_solve: top level not set");
  // ZerosBoundaryResiduals
  if (L.nbdc) B2_LAUNCH(c, fill_idx_kernel, vec_grid(c, L.nbdc), kBlock, 0, res->d, L.bdc, L.nbdc, 0.0);
  // EPSC = Vcycle(RES): the cycle runs on the level work vectors, EPSC = L.x
  B2_CUDA(cudaMemcpyAsync(L.b->d, res->d, (size_t)L.A->nrows * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  B2_TRY(vcycle(mg, top));
  // RES -= A EPSC ; EPS += EPSC
  B2_TRY(level_resid(L, res, L.x, res));
  B2_TRY(b2_vec_axpy(eps, 1.0, L.x));
  return 0;
}

int b2_mg_coarse_iterations(const b2_mg* mg) { return mg->coarse_its; }

int b2_mg_destroy(b2_mg* mg) {
  if (!mg) return 0;
  b2_ctx* c = mg->ctx;
  cudaStreamSynchronize(c->stream);
  for (auto& L : mg->L) {
    b2_vec_destroy(L.dinv);
    b2_vec_destroy(L.x);
    b2_vec_destroy(L.t);
    b2_vec_destroy(L.b);
    b2_vec_destroy(L.r);
    if (L.R) b2_csr_destroy(L.R);
    if (L.bdc) b2_free(c, L.bdc, (size_t)L.nbdc);
  }
  b2_vec_destroy(mg->p);
  b2_vec_destroy(mg->q);
  b2_vec_destroy(mg->z);
  b2_free(c, mg->scal, 8);
  delete mg;
  return 0;
}

}  // extern "C"
