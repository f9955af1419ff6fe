// Geometric multigrid V-cycle on device CSR operators.  Replaces what
// LinearEquationSolverPetsc::{MGInit, MGSetLevel, MGSolve, MGClear} configure in PETSc's PCMG
// (reference src/08_algebra.../03_solvers_with_preconditioner/LinearEquationSolverPetsc.cpp:185-353):
// multiplicative V, zero initial guess on every level, npre/npost sweeps of
// KSPRICHARDSON(scale omega)+PCJACOBI (:516-519, PetscPreconditioner.cpp:209-212), restriction
// with P^T (PCMGSetRestriction(..., PP), :277), interpolation with P (:276), level operators with
// Dirichlet rows set to identity (SetPenalty, :428-436).  The coarse solve (reference: PREONLY +
// MUMPS LU, PetscPreconditioner.cpp:147-160) is a Jacobi-preconditioned CG run to a tight
// relative residual, in the single-reduction form of Chronopoulos and Gear: one operator
// application and ONE reduction per iteration, and in the sharded run the three scalars of that
// reduction travel in the same ncclAllReduce as the interface values of the product.
#include "b2_common.cuh"
#include "b2_gmres.hpp"

struct b2_mg_level {
  b2_csr* A = nullptr;      // borrowed, penalised in place
  b2_csr* P = nullptr;      // borrowed
  b2_csr* R = nullptr;      // owned: explicit transpose of P
  uint64_t Pversion = 0;    // P->version R was built from (values of P changed in place => R is rebuilt)
  b2_vec *dinv = nullptr, *x = nullptr, *t = nullptr, *b = nullptr, *r = nullptr;
  int32_t* bdc = nullptr;   // device copy of the Dirichlet row list
  std::vector<int32_t> bdc_host;      // what it holds
  int64_t nbdc = 0;
  b2_halo* halo = nullptr;  // borrowed: distributed layout of this level's vectors (null: single rank)
  int npre = 1, npost = 1;
  double omega = 0.5;
  // smoother: 0 = Richardson(omega) + Jacobi, 1 = Chebyshev + Jacobi on [emin, emax] of D^-1 A,
  // 2 = Richardson(omega) + element-block multiplicative Schwarz (b2_schwarz.cu)
  int smoother = 0;
  b2_schwarz* schwarz = nullptr;   // borrowed
  // level solver around the preconditioner of kinds 0 / 2: 0 = Richardson(omega), 1 = GMRES (left-preconditioned,
  // npre / npost iterations, no restart inside a smoothing call) -- KSPGMRES, the reference's default level solver
  int ksp = 0;
  std::vector<b2_vec*> krylov;     // GMRES basis (max(npre, npost) vectors) and one work vector
  double emin = 0., emax = 0.;     // bounds in use
  double emin_user = 0., emax_user = 0.;   // emax_user <= 0: estimated at MGSetLevel (power iteration)
  b2_vec* d = nullptr;             // Chebyshev direction
  // null space of the level operator (MatSetNullSpace / MatSetTransposeNullSpace, LinearEquationSolverPetsc.cpp:357-414):
  // one normalised vector (the constant pressure of an enclosed flow), owned; bproj = the smoother's projected right-hand side
  b2_vec *nullvec = nullptr, *bproj = nullptr;
};

struct b2_mg {
  b2_ctx* ctx;
  int nlevels;
  std::vector<b2_mg_level> L;
  b2_schwarz* coarse_schwarz = nullptr;   // borrowed: exact coarse solve (one block holding every dof) instead of the PCG
  double coarse_rtol = 1e-14;
  int coarse_maxit = 5000;
  int coarse_its = 0;
  // PCG work vectors on level 0 and device scalars
  b2_vec *p = nullptr, *q = nullptr, *z = nullptr, *w = nullptr;
  double* scal = nullptr;   // [16]: 0 gamma_new (r.u), 1 delta (w.u), 2 rr, 3 bb | 8 gamma, 9 alpha, 10 beta
  double* dot_partial = nullptr;      // [kRedBlocks][4]
  unsigned int* dot_counter = nullptr;
  // optional per-phase device timing of the cycles (b2_mg_set_timing): events around every phase of every level
  bool timing = false;
  struct Stamp { int level, phase; cudaEvent_t e0, e1; };
  std::vector<Stamp> stamps;
};

// brackets one phase of one level with events when timing is on (phases: 0 pre-smoothing, 1 residual, 2 restriction,
// 3 coarse solve, 4 prolongation, 5 post-smoothing)
struct b2_mg_phase {
  b2_mg* mg; int level, phase; cudaEvent_t e0 = nullptr, e1 = nullptr;
  b2_mg_phase(b2_mg* m, int l, int p) : mg(m), level(l), phase(p) {
    if (mg->timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, mg->ctx->stream); }
  }
  ~b2_mg_phase() {
    if (e0) { cudaEventRecord(e1, mg->ctx->stream); mg->stamps.push_back({level, phase, e0, e1}); }
  }
};

int b2_cg_persistent(b2_ctx* c, const b2_csr* A, const double* dinv, const double* b, const uint8_t* owned, const b2_halo* halo,
                     double* x, double* r, double* u, double* p, double* s, double* w, double* partial, int partial_blocks,
                     double* out4, double rtol, int maxit, int* its, int* ran);      // b2_cg.cu

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

// ---- single-reduction PCG (Chronopoulos-Gear) kernels, scalars resident on the device ------------
// out[0] = sum_owned r.u, out[1] = sum_all w.u, out[2] = sum_owned r.r (, out[3] = sum_owned b.b), one pass,
// deterministic two-level reduction with a ticket counter.  w is this rank's PARTIAL product: since u
// is complete on every rank that holds a dof, summing w.u over ALL local entries and over the ranks
// gives the dot product with the completed w -- so it can be formed before the interface sum.
__global__ void __launch_bounds__(kBlock) cg_dots_kernel(int64_t n, const double* __restrict__ r, const double* __restrict__ u,
                                                         const double* __restrict__ w, const double* __restrict__ b,
                                                         const uint8_t* __restrict__ owned, double* __restrict__ partial,
                                                         double* __restrict__ out, unsigned int* counter) {
  double a[4] = {0., 0., 0., 0.};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double ui = u[i];
    a[1] = fma(w[i], ui, a[1]);
    if (!owned || owned[i]) {
      const double ri = r[i];
      a[0] = fma(ri, ui, a[0]);
      a[2] = fma(ri, ri, a[2]);
      if (b) a[3] = fma(b[i], b[i], a[3]);
    }
  }
  __shared__ double sh[kBlock / 32][4];
  __shared__ bool last;
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_down_sync(0xffffffffu, a[k], o);
  const int wi = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0)
    for (int k = 0; k < 4; k++) sh[wi][k] = a[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; k++) {
      double t = 0.;
      for (int ww = 0; ww < kBlock / 32; ww++) t += sh[ww][k];
      partial[4 * blockIdx.x + k] = t;
    }
    __threadfence();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && wi == 0) {
    double t[4] = {0., 0., 0., 0.};
    for (int k = l; k < (int)gridDim.x; k += 32)
      for (int j = 0; j < 4; j++) t[j] += __ldcg(&partial[4 * k + j]);
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t[j] += __shfl_down_sync(0xffffffffu, t[j], o);
    if (l == 0) {
      out[0] = t[0]; out[1] = t[1]; out[2] = t[2];
      if (b) out[3] = t[3];
      *counter = 0;
    }
  }
}
// u = dinv .* r
__global__ void cg_precond_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ u) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) u[i] = dinv[i] * r[i];
}
// scalars of one iteration, by every thread from the reduced values (first: beta = 0, alpha = gamma/delta):
//   beta = gamma_new / gamma ; alpha = gamma_new / (delta - beta * gamma_new / alpha_old)
// then  p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = dinv .* r
// (thread 0 of block 0 publishes gamma, alpha for the next iteration AFTER a grid-wide read: the
//  scalars live in two banks selected by `bank`, so no thread can see the new values too early)
__global__ void cg_step_kernel(int64_t n, double* __restrict__ sc, int bank, int first, const double* __restrict__ dinv,
                               const double* __restrict__ w, double* __restrict__ u, double* __restrict__ p,
                               double* __restrict__ s, double* __restrict__ x, double* __restrict__ r) {
  const double gn = sc[0], delta = sc[1];
  const double* old = sc + 8 + 4 * bank;          // gamma, alpha of the previous iteration
  double beta = 0.0, alpha;
  if (first) {
    alpha = delta != 0.0 ? gn / delta : 0.0;
  } else {
    const double g = old[0], al = old[1];
    beta = g != 0.0 ? gn / g : 0.0;
    const double den = delta - (al != 0.0 ? beta * gn / al : 0.0);
    alpha = den != 0.0 ? gn / den : 0.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double* nw = sc + 8 + 4 * (bank ^ 1);
    nw[0] = gn;
    nw[1] = alpha;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double pi = fma(beta, p[i], u[i]);
    const double si = fma(beta, s[i], w[i]);
    p[i] = pi;
    s[i] = si;
    x[i] = fma(alpha, pi, x[i]);
    const double ri = fma(-alpha, si, r[i]);
    r[i] = ri;
    u[i] = dinv[i] * ri;
  }
}

// x += omega * dinv .* r   (second half of a distributed Richardson-Jacobi sweep)
__global__ void jacobi_update_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ r,
                                     double* __restrict__ x, double omega) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = fma(omega * dinv[i], r[i], x[i]);
}

// Chebyshev step: d = cd * d + cz * dinv .* r ; x = (zero ? 0 : x) + d
__global__ void cheb_update_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ r,
                                   double* __restrict__ d, double* __restrict__ x, double cd, double cz, int zero) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double di = fma(cz * dinv[i], r[i], cd * d[i]);
    d[i] = di;
    x[i] = zero ? di : x[i] + di;
  }
}
// deterministic start vector of the eigenvalue estimate
__global__ void power_start_kernel(int64_t n, double* __restrict__ v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = 1.0 + 0.5 * sin((double)(i % 1000));
}
// v = scale * dinv .* w
__global__ void scaled_precond_kernel(int64_t n, const double* __restrict__ dinv, const double* __restrict__ w,
                                      double* __restrict__ v, double scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = scale * (dinv[i] * w[i]);
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

// largest eigenvalue of D^-1 A by `its` power iterations from a fixed start vector (our own stated
// bound: PETSc's KSPCHEBYSHEV estimates it with GMRES on a random vector, which is not reproducible)
int estimate_emax(b2_mg* mg, b2_mg_level& L, int its, double* out) {
  b2_ctx* c = mg->ctx;
  const int64_t n = L.A->nrows;
  const int g = vec_grid(c, n);
  double* s = mg->scal + 4;
  B2_LAUNCH(c, power_start_kernel, g, kBlock, 0, n, L.x->d);
  if (L.nbdc) B2_LAUNCH(c, fill_idx_kernel, vec_grid(c, L.nbdc), kBlock, 0, L.x->d, L.bdc, L.nbdc, 0.0);
  if (L.halo) {        // shared dofs carry different local indices on different ranks: average the start values
    B2_TRY(b2_halo_sum(L.halo, L.x));
    B2_LAUNCH(c, scaled_precond_kernel, g, kBlock, 0, n, L.halo->invmult, L.x->d, L.t->d, 1.0);
    std::swap(L.x, L.t);
  }
  double lam = 0.0, h = 0.0;
  for (int it = 0; it <= its; it++) {
    B2_TRY(b2_dev_dot(c, L.x->d, L.x->d, n, s, owned(L)));
    B2_TRY(b2_allreduce_sum(c, s, 1));
    B2_TRY(b2_download(c, &h, s, 1));
    const double nrm = sqrt(h);
    if (it > 0) lam = nrm;            // ||D^-1 A v|| with ||v|| = 1
    if (it == its || nrm == 0.0) break;
    B2_TRY(b2_vec_scale(L.x, 1.0 / nrm));
    B2_TRY(level_spmv(L, L.x, L.t));
    B2_LAUNCH(c, scaled_precond_kernel, g, kBlock, 0, n, L.dinv->d, L.t->d, L.x->d, 1.0);
  }
  *out = lam;
  return 0;
}

// Chebyshev + Jacobi sweeps on [emin, emax] (Saad, Iterative Methods, alg. 12.1)
int smooth_chebyshev(b2_mg* mg, b2_mg_level& L, int nsweeps, bool zero_guess) {
  b2_ctx* c = mg->ctx;
  const int64_t n = L.A->nrows;
  if (nsweeps <= 0) return zero_guess ? b2_vec_zero(L.x) : 0;
  const double theta = 0.5 * (L.emax + L.emin), delta = 0.5 * (L.emax - L.emin), sigma1 = theta / delta;
  double rho = 1.0 / sigma1;
  for (int k = 0; k < nsweeps; k++) {
    const double* r = L.b->d;                       // zero guess: r = b
    if (!(zero_guess && k == 0)) {
      B2_TRY(level_resid(L, L.b, L.x, L.t));
      r = L.t->d;
    }
    double cd = 0.0, cz = 1.0 / theta;
    if (k > 0) {
      const double rho_new = 1.0 / (2.0 * sigma1 - rho);
      cd = rho_new * rho;
      cz = 2.0 * rho_new / delta;
      rho = rho_new;
    }
    B2_LAUNCH(c, cheb_update_kernel, vec_grid(c, n), kBlock, 0, n, L.dinv->d, r, L.d->d, L.x->d, cd, cz,
              (zero_guess && k == 0) ? 1 : 0);
  }
  return 0;
}

// z = M^-1 r with the level's preconditioner: Jacobi (kind 0) or the element-block sweep (kind 2)
int pc_apply(b2_mg* mg, b2_mg_level& L, const b2_vec* r, b2_vec* z) {
  if (L.smoother == 2) return b2_schwarz_apply(L.schwarz, r, z);
  B2_LAUNCH(mg->ctx, cg_precond_kernel, vec_grid(mg->ctx, L.A->nrows), kBlock, 0, L.A->nrows, L.dinv->d, r->d, z->d);
  return 0;
}

// KSPGMRES as level solver (LinearEquationSolverPetsc.cpp:452-536 with _levelSolverType = GMRES; PETSc defaults:
// left preconditioning, so the Krylov space is built from M^-1 A and the PRECONDITIONED residual is minimised):
// k iterations from the current iterate, x <- x + V y.  Modified Gram-Schmidt + Givens rotations; the iterate is the
// unique minimiser over x0 + K_k(M^-1 A, M^-1 r0), so it equals PETSc's (classical Gram-Schmidt) up to rounding.
// Scalars come back to the host (k <= a few iterations per call; one device synchronisation per dot product).
int smooth_gmres(b2_mg* mg, b2_mg_level& L, int k, bool zero_guess);

// Richardson(omega) around the element-block preconditioner: x <- x + omega M^-1 (b - A x)
int smooth_schwarz(b2_mg* mg, b2_mg_level& L, int nsweeps, bool zero_guess) {
  if (zero_guess) B2_TRY(b2_vec_zero(L.x));
  for (int k = 0; k < nsweeps; k++) {
    const b2_vec* r = L.b;                          // zero guess: r = b
    if (!(zero_guess && k == 0)) {
      B2_TRY(level_resid(L, L.b, L.x, L.t));
      r = L.t;
    }
    B2_TRY(b2_schwarz_apply(L.schwarz, r, L.d));
    B2_TRY(b2_vec_axpy(L.x, L.omega, L.d));
  }
  return 0;
}

// device-vector operations of the GMRES cycle (b2_gmres.hpp): basis in L.krylov[0..k-1], w = L.krylov[k]
struct gmres_level_ops {
  b2_mg* mg;
  b2_mg_level& L;
  bool zero_guess;
  b2_vec* w;
  int start(double* beta) {
    const b2_vec* r = L.b;                  // zero guess: r = b
    if (!zero_guess) {
      B2_TRY(level_resid(L, L.b, L.x, L.t));
      r = L.t;
    }
    B2_TRY(pc_apply(mg, L, r, L.krylov[0]));
    return b2_vec_norm(L.krylov[0], 2, beta);
  }
  int scale(int j, double a) { return b2_vec_scale(L.krylov[j], a); }
  int apply(int j) {
    B2_TRY(level_spmv(L, L.krylov[j], L.t));
    return pc_apply(mg, L, L.t, w);
  }
  int dot_w(int i, double* h) { return b2_vec_dot(w, L.krylov[i], h); }
  int axpy_w(double a, int i) { return b2_vec_axpy(w, a, L.krylov[i]); }
  int norm_w(double* n) { return b2_vec_norm(w, 2, n); }
  int store(int j) { return b2_vec_copy(L.krylov[j], w); }
  int update_x(double a, int i) { return b2_vec_axpy(L.x, a, L.krylov[i]); }
};

int smooth_gmres(b2_mg* mg, b2_mg_level& L, int k, bool zero_guess) {
  b2_ctx* c = mg->ctx;
  const int64_t n = L.A->nrows;
  B2_CHECK(!L.halo, "b2_mg: the GMRES level solver runs on one rank only");
  if (zero_guess) B2_TRY(b2_vec_zero(L.x));
  if (k <= 0) return 0;
  while ((int)L.krylov.size() < k + 1) {
    b2_vec* v = nullptr;
    B2_TRY(b2_vec_create(c, n, &v));
    L.krylov.push_back(v);
  }
  gmres_level_ops ops{mg, L, zero_guess, L.krylov[k]};
  return b2_gmres_cycle(ops, k);
}

// Richardson(omega) around the level's preconditioner on an operator with a null space, as KSPSolve does it once
// MatSetNullSpace / MatSetTransposeNullSpace are set: the right-hand side is projected (a copy: the cycle's residual keeps
// the original), and so is every preconditioned residual: b' = b - (n.b) n; x <- x + omega (I - n n^T) M^-1 (b' - A x)
int smooth_nullspace(b2_mg* mg, b2_mg_level& L, int nsweeps, bool zero_guess) {
  B2_CHECK(L.ksp == 0 && L.smoother != 1, "a level with a null space is smoothed by Richardson around Jacobi or the element blocks");
  if (zero_guess) B2_TRY(b2_vec_zero(L.x));
  if (nsweeps <= 0) return 0;
  double s = 0.0;
  B2_TRY(b2_vec_copy(L.bproj, L.b));
  B2_TRY(b2_vec_dot(L.nullvec, L.bproj, &s));
  B2_TRY(b2_vec_axpy(L.bproj, -s, L.nullvec));
  for (int k = 0; k < nsweeps; k++) {
    const b2_vec* r = L.bproj;
    if (!(zero_guess && k == 0)) {
      B2_TRY(level_resid(L, L.bproj, L.x, L.t));
      r = L.t;
    }
    B2_TRY(pc_apply(mg, L, r, L.d));
    B2_TRY(b2_vec_dot(L.nullvec, L.d, &s));
    B2_TRY(b2_vec_axpy(L.d, -s, L.nullvec));
    B2_TRY(b2_vec_axpy(L.x, L.omega, L.d));
  }
  return 0;
}

int smooth(b2_mg* mg, int l, int nsweeps, bool zero_guess) {
  b2_mg_level& L = mg->L[l];
  b2_ctx* c = mg->ctx;
  const int64_t n = L.A->nrows;
  if (L.nullvec) return smooth_nullspace(mg, L, nsweeps, zero_guess);
  if (L.smoother == 1) return smooth_chebyshev(mg, L, nsweeps, zero_guess);
  if (L.ksp == 1) return smooth_gmres(mg, L, nsweeps, zero_guess);
  if (L.smoother == 2) return smooth_schwarz(mg, L, nsweeps, zero_guess);
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
  // r, u = D^-1 r, w = A u, p, s = A p as in Chronopoulos & Gear; mg->z = u, mg->w = w, mg->p = p, mg->q = s.
  b2_mg_level& L = mg->L[0];
  b2_ctx* c = mg->ctx;
  if (mg->coarse_schwarz) {      // PREONLY + LU of the reference (PetscPreconditioner.cpp:147-160): x = A^-1 b, also for indefinite systems
    mg->coarse_its = 0;
    return b2_schwarz_apply(mg->coarse_schwarz, L.b, L.x);
  }
  const int64_t n = L.A->nrows;
  const int g = vec_grid(c, n);
  int gd = b2_grid_for(c, n, kBlock * 4, 8);
  if (gd > kRedBlocks) gd = kRedBlocks;
  double* sc = mg->scal;
  const uint8_t* own = owned(L);
  // x0 = b on Dirichlet rows (identity rows), 0 elsewhere; r = b - A x0
  B2_TRY(b2_vec_zero(L.x));
  if (L.nbdc) B2_LAUNCH(c, copy_idx_kernel, vec_grid(c, L.nbdc), kBlock, 0, L.x->d, L.b->d, L.bdc, L.nbdc);
  B2_TRY(level_resid(L, L.b, L.x, L.r));
  B2_TRY(b2_vec_zero(mg->p));
  B2_TRY(b2_vec_zero(mg->q));
  B2_LAUNCH(c, cg_precond_kernel, g, kBlock, 0, n, L.dinv->d, L.r->d, mg->z->d);
  double h[4];
  double bb = 0.0;
  mg->coarse_its = 0;
  if (c->coarse_persistent) {      // the whole iteration loop in one cooperative kernel (b2_cg.cu)
    int ran = 0, its = 0;
    B2_TRY(b2_cg_persistent(c, L.A, L.dinv->d, L.b->d, own, L.halo, L.x->d, L.r->d, mg->z->d, mg->p->d, mg->q->d, mg->w->d, mg->dot_partial,
                            kRedBlocks, sc, mg->coarse_rtol, mg->coarse_maxit, &its, &ran));
    if (ran) {
      mg->coarse_its = its;
      return 0;
    }
  }
  const int check_every = 8;
  int bank = 0;
  for (int it = 0; it <= mg->coarse_maxit; it++) {
    // w = A u (this rank's part), the three dot products, then ONE collective for interface + scalars
    B2_TRY(b2_csr_spmv(L.A, mg->z, mg->w));
    B2_LAUNCH(c, cg_dots_kernel, gd, kBlock, 0, n, L.r->d, mg->z->d, mg->w->d, it == 0 ? L.b->d : (const double*)nullptr, own,
              mg->dot_partial, sc, mg->dot_counter);
    if (L.halo) B2_TRY(b2_halo_sum_scalars(L.halo, mg->w, sc, it == 0 ? 4 : 3));
    if (it == 0 || it % check_every == 0) {      // convergence on ||r||^2 <= rtol^2 ||b||^2 (one small D2H)
      B2_TRY(b2_download(c, h, sc, 4));
      if (it == 0) bb = h[3];
      if (bb == 0.0 || !(h[2] > mg->coarse_rtol * mg->coarse_rtol * bb)) break;      // also leaves on NaN
    }
    if (it == mg->coarse_maxit) break;
    B2_LAUNCH(c, cg_step_kernel, g, kBlock, 0, n, sc, bank, it == 0 ? 1 : 0, L.dinv->d, mg->w->d, mg->z->d, mg->p->d, mg->q->d,
              L.x->d, L.r->d);
    bank ^= 1;
    mg->coarse_its = it + 1;
  }
  return 0;
}

int vcycle(b2_mg* mg, int l) {
  // solves (approximately) A_l x_l = b_l, zero initial guess; operands are the level work vectors
  b2_mg_level& L = mg->L[l];
  if (l == 0) {
    b2_mg_phase ph(mg, 0, 3);
    return coarse_solve(mg);
  }
  {
    b2_mg_phase ph(mg, l, 0);
    B2_TRY(smooth(mg, l, L.npre, true));
  }
  {
    b2_mg_phase ph(mg, l, 1);
    B2_TRY(level_resid(L, L.b, L.x, L.r));
  }
  b2_mg_level& C = mg->L[l - 1];
  {
    b2_mg_phase ph(mg, l, 2);
    B2_TRY(b2_csr_spmv(L.R, L.r, C.b));        // R = P^T restricted to the fine rows this rank owns
    if (C.halo) B2_TRY(b2_halo_sum(C.halo, C.b));
  }
  B2_TRY(vcycle(mg, l - 1));
  {
    b2_mg_phase ph(mg, l, 4);
    B2_TRY(b2_csr_spmv_add(L.P, C.x, L.x));
  }
  {
    b2_mg_phase ph(mg, l, 5);
    B2_TRY(smooth(mg, l, L.npost, false));
  }
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
  B2_TRY(b2_malloc(c, &mg->scal, 16));
  B2_CUDA(cudaMemsetAsync(mg->scal, 0, 16 * sizeof(double), c->stream));
  B2_TRY(b2_malloc(c, &mg->dot_partial, (size_t)kRedBlocks * 4));
  B2_TRY(b2_malloc(c, &mg->dot_counter, 1));
  B2_CUDA(cudaMemsetAsync(mg->dot_counter, 0, sizeof(unsigned int), c->stream));
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
  // PCMGSetRestriction(..., PP) is called at every MGSetLevel (LinearEquationSolverPetsc.cpp:277): the restriction
  // always matches the prolongator's current values
  const bool newP = (level > 0) && (L.P != P || !L.R || L.Pversion != P->version);
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
      B2_TRY(b2_vec_create(c, n, &mg->w));
    }
  }
  B2_CHECK(!L.halo || L.halo->n_local == n, "b2_mg_set_level: layout of level %d has %lld dofs, operator %lld", level,
           (long long)(L.halo ? L.halo->n_local : 0), (long long)n);
  // the Dirichlet row list rarely changes between two MGSetLevel calls: keep the device copy (a cudaFree here would wait
  // for every stream of the device, e.g. for the next step's inputs still uploading on the copy stream)
  const bool same_bdc = L.bdc_host.size() == (size_t)nbdc && (nbdc == 0 || memcmp(L.bdc_host.data(), bdc_idx, (size_t)nbdc * sizeof(int32_t)) == 0);
  if (!same_bdc) {
    if (L.bdc) { b2_free(c, L.bdc, (size_t)L.nbdc); L.bdc = nullptr; }
    L.bdc_host.assign(bdc_idx, bdc_idx + nbdc);
  }
  L.nbdc = nbdc;
  if (nbdc && !same_bdc) {
    B2_TRY(b2_malloc(c, &L.bdc, (size_t)nbdc));
    B2_TRY(b2_upload(c, L.bdc, bdc_idx, (size_t)nbdc));
  }
  if (nbdc) B2_TRY(b2_csr_zero_rows_dev(A, L.bdc, nbdc, 1.0, owned(L)));     // SetPenalty
  B2_TRY(b2_csr_diag(A, L.dinv));
  if (L.halo) B2_TRY(b2_halo_sum(L.halo, L.dinv));       // diagonal of the summed operator
  B2_LAUNCH(c, recip_kernel, vec_grid(c, n), kBlock, 0, n, L.dinv->d, L.dinv->d);
  if (level == 0 && mg->coarse_schwarz) {      // numeric phase of the direct coarse solve on the penalised operator
    B2_CHECK(!L.halo, "b2_mg_set_level: the direct coarse solve runs on one rank only");
    B2_CHECK(b2_schwarz_operator(mg->coarse_schwarz) == A, "b2_mg_set_level: the coarse solver was created on another operator");
    B2_TRY(b2_schwarz_setup(mg->coarse_schwarz));
  }
  if (L.smoother == 2 && level > 0) {          // numeric phase of the block smoother on the penalised operator
    B2_CHECK(!L.halo, "b2_mg_set_level: the element-block smoother runs on one rank only");
    B2_CHECK(L.schwarz && b2_schwarz_operator(L.schwarz) == A, "b2_mg_set_level: level %d: the block smoother was created on another operator", level);
    if (!L.d) B2_TRY(b2_vec_create(c, n, &L.d));
    B2_TRY(b2_schwarz_setup(L.schwarz));
  }
  if (L.nullvec && level > 0) {
    B2_CHECK(L.nullvec->n == n, "b2_mg_set_level: the null-space vector of level %d has %lld entries, the operator %lld rows", level,
             (long long)L.nullvec->n, (long long)n);
    if (!L.d) B2_TRY(b2_vec_create(c, n, &L.d));
  }
  if (L.smoother == 1 && level > 0) {
    if (!L.d) B2_TRY(b2_vec_create(c, n, &L.d));
    if (L.emax_user <= 0.0) {                       // our own stated bounds: [0.1, 1.1] x power-iteration estimate
      double lam = 0.0;
      B2_TRY(estimate_emax(mg, L, 10, &lam));
      L.emax = 1.1 * lam;
      L.emin = 0.1 * lam;
    } else {
      L.emax = L.emax_user;
      L.emin = L.emin_user;
    }
  }
  if (newP) {   // the explicit restriction R = P^T is rebuilt only when P changes
    if (L.R) { b2_csr_destroy(L.R); L.R = nullptr; }
    B2_TRY(b2_csr_transpose(P, &L.R));
    L.Pversion = P->version;
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

/* smoother of one level: kind 0 = Richardson(omega)+Jacobi (KSPRICHARDSON+PCJACOBI), 1 = Chebyshev+Jacobi
 * (KSPCHEBYSHEV+PCJACOBI, LinearEquationSolverPetsc.cpp:452-536) on the interval [emin, emax] of D^-1 A;
 * emax <= 0: estimated at b2_mg_set_level by 10 power iterations from a fixed start vector, interval
 * [0.1, 1.1] x estimate.  Call before b2_mg_set_level. */
int b2_mg_set_smoother(b2_mg* mg, int level, int kind, double emin, double emax) {
  B2_CHECK(level >= 0 && level < mg->nlevels && (kind == 0 || kind == 1), "b2_mg_set_smoother: bad arguments");
  B2_CHECK(kind == 0 || emax <= 0.0 || (emin > 0.0 && emin < emax), "b2_mg_set_smoother: need 0 < emin < emax");
  mg->L[level].smoother = kind;
  mg->L[level].emin_user = emin;
  mg->L[level].emax_user = emax;
  return 0;
}
int b2_mg_set_level_schwarz(b2_mg* mg, int level, b2_schwarz* s) {
  B2_CHECK(mg && level >= 1 && level < mg->nlevels, "b2_mg_set_level_schwarz: bad level %d (the coarsest level has no smoother)", level);
  mg->L[level].schwarz = s;
  mg->L[level].smoother = s ? 2 : 0;
  return 0;
}

/* Null space of the operator of one level >= 1 (RemoveNullSpace, LinearEquationSolverPetsc.cpp:357-414: levels above the
 * coarsest; the base vector of GetNullSpaceBase is 1 on the free dofs of the flagged variable, normalised): the level's
 * smoother projects its right-hand side and every preconditioned residual onto the complement.  nvec is copied and
 * normalised; NULL removes the setting. */
int b2_mg_set_level_nullspace(b2_mg* mg, int level, const b2_vec* nvec) {
  B2_CHECK(mg && level >= 1 && level < mg->nlevels, "b2_mg_set_level_nullspace: bad level %d (levels above the coarsest)", level);
  b2_mg_level& L = mg->L[level];
  if (L.nullvec) { b2_vec_destroy(L.nullvec); L.nullvec = nullptr; }
  if (L.bproj) { b2_vec_destroy(L.bproj); L.bproj = nullptr; }
  if (!nvec) return 0;
  double nn = 0.0;
  B2_TRY(b2_vec_dot(nvec, nvec, &nn));
  B2_CHECK(nn > 0.0, "b2_mg_set_level_nullspace: zero vector");
  B2_TRY(b2_vec_create(mg->ctx, nvec->n, &L.nullvec));
  B2_TRY(b2_vec_create(mg->ctx, nvec->n, &L.bproj));
  B2_TRY(b2_vec_copy(L.nullvec, nvec));
  B2_TRY(b2_vec_scale(L.nullvec, 1.0 / sqrt(nn)));
  return 0;
}

int b2_mg_set_level_ksp(b2_mg* mg, int level, int kind) {
  B2_CHECK(mg && level >= 1 && level < mg->nlevels && (kind == 0 || kind == 1), "b2_mg_set_level_ksp: level %d, kind %d (0 Richardson, 1 GMRES)", level, kind);
  mg->L[level].ksp = kind;
  return 0;
}

int b2_mg_set_coarse_schwarz(b2_mg* mg, b2_schwarz* s) {
  B2_CHECK(mg, "b2_mg_set_coarse_schwarz: null handle");
  mg->coarse_schwarz = s;
  return 0;
}

int b2_mg_level_bounds(const b2_mg* mg, int level, double* emin, double* emax) {
  B2_CHECK(level >= 0 && level < mg->nlevels, "b2_mg_level_bounds: bad level");
  *emin = mg->L[level].emin;
  *emax = mg->L[level].emax;
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

/* per-phase device timing of the cycles run from now on (on = 1); b2_mg_get_timing sums what was recorded since the last
 * call into ms[nlevels][6] (phases: pre-smoothing, residual, restriction, coarse solve, prolongation, post-smoothing) */
int b2_mg_set_timing(b2_mg* mg, int on) {
  mg->timing = on != 0;
  return 0;
}
int b2_mg_get_timing(b2_mg* mg, double* ms) {
  B2_CUDA(cudaStreamSynchronize(mg->ctx->stream));
  for (int k = 0; k < mg->nlevels * 6; k++) ms[k] = 0.0;
  for (auto& st : mg->stamps) {
    float t = 0.f;
    cudaEventElapsedTime(&t, st.e0, st.e1);
    ms[st.level * 6 + st.phase] += (double)t;
    cudaEventDestroy(st.e0);
    cudaEventDestroy(st.e1);
  }
  mg->stamps.clear();
  return 0;
}

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
    b2_vec_destroy(L.d);
    b2_vec_destroy(L.nullvec);
    b2_vec_destroy(L.bproj);
    for (b2_vec* v : L.krylov) b2_vec_destroy(v);
    if (L.R) b2_csr_destroy(L.R);
    if (L.bdc) b2_free(c, L.bdc, (size_t)L.nbdc);
  }
  b2_vec_destroy(mg->p);
  b2_vec_destroy(mg->q);
  b2_vec_destroy(mg->z);
  b2_vec_destroy(mg->w);
  b2_free(c, mg->scal, 16);
  b2_free(c, mg->dot_partial, (size_t)kRedBlocks * 4);
  b2_free(c, mg->dot_counter, 1);
  delete mg;
  return 0;
}

}  // extern "C"
