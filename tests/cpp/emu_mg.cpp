// TEST INFRASTRUCTURE: the library's multigrid orchestration (b2_mg.cu), vectors (b2_vec.cu) and element-block smoother
// (b2_schwarz.cu) compiled with g++ for the CPU thread emulator (emu_prefix.hpp, emu_rt/cuda_runtime.h) and linked with
// the host stand-ins below for what they call in the TMA-based SpMV family and in b2_csr.cu (plain CSR loops).  One C
// entry point runs MGSetLevel on every level and a number of MGSolve cycles -- the code path of LinearEquationSolverB200
// -- so that the V-cycle with the block smoother, the GMRES level solver and the direct coarse solve is checked against
// the oracle without a GPU (tests/test_kernel_emulation.py).
#include <cstdarg>
#include <cstdio>
#include <string>
#include "../../femus_b200/csrc/b2_common.cuh"

static thread_local std::string g_err;
void b2_set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}
int b2_allreduce_op(b2_ctx*, double*, int64_t, int) { return 0; }
int b2_allreduce_sum(b2_ctx*, double*, int64_t) { return 0; }
int b2_halo_sum_scalars(b2_halo*, b2_vec*, double*, int) { return 0; }
// the cooperative coarse-PCG kernel (b2_cg.cu) has no emulated form: the host-driven loop of b2_mg.cu runs instead
int b2_cg_persistent(b2_ctx*, const b2_csr*, const double*, const double*, const uint8_t*, const b2_halo*, double*, double*, double*, double*,
                     double*, double*, double*, int, double*, double, int, int* its, int* ran) {
  *its = 0;
  *ran = 0;
  return 0;
}
int b2_csr_zero_cols_notowned(b2_csr*, const uint8_t*) { return 0; }
int b2_csr_resid_w(const b2_csr* A, const double* b, const double* w, const double* x, double* r) {
  for (int64_t i = 0; i < A->nrows; i++) {
    double s = 0.0;
    for (int64_t q = A->rowptr[i]; q < A->rowptr[i + 1]; q++) s += A->val[q] * x[A->col[q]];
    r[i] = w[i] * b[i] - s;
  }
  return 0;
}
int b2_csr_zero_rows_dev(b2_csr* A, const int32_t* rows, int64_t n, double diag, const uint8_t*) {
  for (int64_t k = 0; k < n; k++)
    for (int64_t q = A->rowptr[rows[k]]; q < A->rowptr[rows[k] + 1]; q++) A->val[q] = A->col[q] == rows[k] ? diag : 0.0;
  return 0;
}

// the mesh object of b2_assemble.cu, as far as b2_stokes.cu looks at it
struct b2_mesh {
  b2_ctx* ctx;
  int64_t nnode, nel;
  const double* xyz;
  const int32_t* conn;
};
void b2_mesh_view(const b2_mesh* m, b2_ctx** ctx, int64_t* nnode, int64_t* nel, const double** xyz, const int32_t** conn) {
  *ctx = m->ctx;
  *nnode = m->nnode;
  *nel = m->nel;
  *xyz = m->xyz;
  *conn = m->conn;
}

extern "C" {
const char* b2_last_error(void) { return g_err.c_str(); }
int b2_halo_sum(b2_halo*, b2_vec*) { return 0; }
int b2_csr_destroy(b2_csr* A) {
  if (!A) return 0;
  std::free(A->rowptr); std::free(A->col); std::free(A->val);
  delete A;
  return 0;
}
int b2_csr_spmv(const b2_csr* A, const b2_vec* x, b2_vec* y) {
  for (int64_t i = 0; i < A->nrows; i++) {
    double s = 0.0;
    for (int64_t q = A->rowptr[i]; q < A->rowptr[i + 1]; q++) s += A->val[q] * x->d[A->col[q]];
    y->d[i] = s;
  }
  return 0;
}
int b2_csr_spmv_add(const b2_csr* A, const b2_vec* x, b2_vec* y) {
  for (int64_t i = 0; i < A->nrows; i++) {
    double s = 0.0;
    for (int64_t q = A->rowptr[i]; q < A->rowptr[i + 1]; q++) s += A->val[q] * x->d[A->col[q]];
    y->d[i] += s;
  }
  return 0;
}
int b2_csr_resid(const b2_csr* A, const b2_vec* b, const b2_vec* x, b2_vec* r) {
  for (int64_t i = 0; i < A->nrows; i++) {
    double s = 0.0;
    for (int64_t q = A->rowptr[i]; q < A->rowptr[i + 1]; q++) s += A->val[q] * x->d[A->col[q]];
    r->d[i] = b->d[i] - s;
  }
  return 0;
}
int b2_csr_jacobi_sweep(const b2_csr* A, const b2_vec* dinv, const b2_vec* b, const b2_vec* x, b2_vec* xnew, double omega) {
  for (int64_t i = 0; i < A->nrows; i++) {
    double s = 0.0;
    for (int64_t q = A->rowptr[i]; q < A->rowptr[i + 1]; q++) s += A->val[q] * x->d[A->col[q]];
    xnew->d[i] = x->d[i] + omega * dinv->d[i] * (b->d[i] - s);
  }
  return 0;
}
int b2_csr_diag(const b2_csr* A, b2_vec* d) {
  for (int64_t i = 0; i < A->nrows; i++) {
    d->d[i] = 0.0;
    for (int64_t q = A->rowptr[i]; q < A->rowptr[i + 1]; q++)
      if (A->col[q] == i) d->d[i] = A->val[q];
  }
  return 0;
}
static b2_csr* make_csr(b2_ctx* c, int64_t nr, int64_t nc, const int64_t* rp, const int32_t* col, const double* val) {
  b2_csr* A = new b2_csr();
  std::memset(A, 0, sizeof *A);
  A->ctx = c; A->nrows = nr; A->ncols = nc; A->nnz = rp[nr];
  A->rowptr = (int64_t*)std::malloc((size_t)(nr + 1) * 8);
  A->col = (int32_t*)std::malloc((size_t)A->nnz * 4 + 4);
  A->val = (double*)std::malloc((size_t)A->nnz * 8 + 8);
  std::memcpy(A->rowptr, rp, (size_t)(nr + 1) * 8);
  std::memcpy(A->col, col, (size_t)A->nnz * 4);
  std::memcpy(A->val, val, (size_t)A->nnz * 8);
  return A;
}
int b2_csr_transpose(const b2_csr* A, b2_csr** out) {
  std::vector<int64_t> rp((size_t)A->ncols + 1, 0);
  for (int64_t q = 0; q < A->nnz; q++) rp[A->col[q] + 1]++;
  for (int64_t i = 0; i < A->ncols; i++) rp[i + 1] += rp[i];
  std::vector<int32_t> col((size_t)A->nnz);
  std::vector<double> val((size_t)A->nnz);
  std::vector<int64_t> fill(rp.begin(), rp.end() - 1);
  for (int64_t i = 0; i < A->nrows; i++)
    for (int64_t q = A->rowptr[i]; q < A->rowptr[i + 1]; q++) {
      const int64_t p = fill[A->col[q]]++;
      col[p] = (int32_t)i;
      val[p] = A->val[q];
    }
  *out = make_csr(A->ctx, A->ncols, A->nrows, rp.data(), col.data(), val.data());
  return 0;
}

// b2_stokes_create / b2_ns_create + assemble (+ the boundary pressure faces) of the real b2_stokes.cu on host arrays:
// the system matrix values on the pattern (rp, col) and the residual.  ns != 0: the Navier-Stokes routine.
int emu_stokes_plan(int64_t nnode, int64_t nel, const double* xyz, const int32_t* conn, int64_t n, const int64_t* rp, const int32_t* col,
                    const int32_t* edof, int nv, int np, int ng, const double* phi_v, const double* dxi, const double* deta, const double* dzeta,
                    const double* w, const double* phi_p, const double* sol, double coef, int ns, int64_t nfaces, const int32_t* felem,
                    const int32_t* flocal, const double* fvalue, int nvf, int ngf, const double* fphi, const double* fdxi, const double* fdeta,
                    const double* fw, const int32_t* fnodes, double* val_out, double* rhs_out) {
  b2_ctx ctx;
  ctx.sm_count = 1;
  b2_mesh mesh{&ctx, nnode, nel, xyz, conn};
  std::vector<double> zeros((size_t)rp[n], 0.0);
  b2_csr* A = make_csr(&ctx, n, n, rp, col, zeros.data());
  auto run = [&]() -> int {
    b2_stokes* plan = nullptr;
    if (ns) B2_TRY(b2_ns_create(&mesh, A, edof, nv, np, ng, phi_v, dxi, deta, dzeta, w, phi_p, &plan));
    else B2_TRY(b2_stokes_create(&mesh, A, edof, nv, np, ng, dxi, deta, dzeta, w, phi_p, &plan));
    b2_vec *S = nullptr, *R = nullptr;
    B2_TRY(b2_vec_create(&ctx, n, &S));
    B2_TRY(b2_vec_create(&ctx, n, &R));
    B2_TRY(b2_vec_put(S, sol, n));
    if (ns) B2_TRY(b2_ns_assemble(plan, S, R, coef));
    else B2_TRY(b2_stokes_assemble(plan, S, R, coef));
    if (nfaces) B2_TRY(b2_ns_pressure_faces(plan, nfaces, felem, flocal, fvalue, nvf, ngf, fphi, fdxi, fdeta, fw, fnodes, R));
    B2_TRY(b2_vec_get(R, rhs_out, n));
    std::memcpy(val_out, A->val, (size_t)rp[n] * sizeof(double));
    b2_vec_destroy(S);
    b2_vec_destroy(R);
    b2_stokes_destroy(plan);
    return 0;
  };
  const int rc = run();
  b2_csr_destroy(A);
  return rc;
}

// One multigrid solve sequence on nlevels levels.  Level l: operator (rp, col, val)[l] (un-penalised), prolongator from
// level l-1 (l >= 1), Dirichlet rows bdc[l].  smoother: 0 Richardson + Jacobi, 2 Richardson / GMRES + element blocks
// (blocks and schedule of level l >= 1 in blk_* / grp_*; sub = block solve); ksp: 0 Richardson, 1 GMRES; coarse_direct:
// one exact block on level 0.  res (finest) holds the right-hand side on entry; trace[c] = ||res||_2 over the rows with
// free[i] != 0 after cycle c; eps = the accumulated correction.  Returns 0, or 1 with b2_last_error set.
int emu_mg_run(int nlevels, const int64_t* n, const int64_t* const* rp, const int32_t* const* col, const double* const* val,
               const int64_t* const* prp, const int32_t* const* pcol, const double* const* pval, const int64_t* nbdc, const int32_t* const* bdc,
               int smoother, int sub, int ksp, int coarse_direct, int row_levels, const int64_t* nblk, const int64_t* const* blk_ptr,
               const int32_t* const* blk_dofs, const int64_t* ngrp, const int64_t* const* grp_ptr, const int32_t* const* grp_blocks, int npre,
               int npost, double omega, int ncycles, double* res, double* eps, const uint8_t* free_rows, double* trace) {
  b2_ctx ctx;
  ctx.sm_count = 1;
  ctx.red_partial = (double*)std::calloc(kRedBlocks * 2, 8);
  ctx.red_result = (double*)std::calloc(8, 8);
  ctx.red_counter = (unsigned*)std::calloc(1, 4);
  ctx.h_result = (double*)std::calloc(8, 8);
  b2_ctx* c = &ctx;
  auto run = [&]() -> int {
    b2_mg* mg = nullptr;
    B2_TRY(b2_mg_create(c, nlevels, &mg));
    B2_TRY(b2_mg_set_coarse(mg, 1e-15, 10000));
    std::vector<b2_csr*> A(nlevels, nullptr), P(nlevels, nullptr);
    std::vector<b2_schwarz*> S(nlevels, nullptr);
    for (int l = 0; l < nlevels; l++) {
      A[l] = make_csr(c, n[l], n[l], rp[l], col[l], val[l]);
      if (l) P[l] = make_csr(c, n[l], n[l - 1], prp[l], pcol[l], pval[l]);
    }
    if (coarse_direct) {
      const int64_t bp[2] = {0, n[0]}, gp[2] = {0, 1};
      std::vector<int32_t> all((size_t)n[0]);
      for (int64_t i = 0; i < n[0]; i++) all[i] = (int32_t)i;
      const int32_t gb[1] = {0};
      B2_TRY(b2_schwarz_create(c, A[0], 1, bp, all.data(), 1, gp, gb, &S[0]));
      B2_TRY(b2_mg_set_coarse_schwarz(mg, S[0]));
    }
    for (int l = 1; l < nlevels; l++) {
      if (smoother == 2) {
        B2_TRY(b2_schwarz_create(c, A[l], nblk[l], blk_ptr[l], blk_dofs[l], ngrp[l], grp_ptr[l], grp_blocks[l], &S[l]));
        B2_TRY(b2_schwarz_set_subsolver(S[l], sub));
        B2_TRY(b2_schwarz_set_row_levels(S[l], row_levels));
        B2_TRY(b2_mg_set_level_schwarz(mg, l, S[l]));
      }
      B2_TRY(b2_mg_set_level_ksp(mg, l, ksp));
    }
    for (int l = 0; l < nlevels; l++) B2_TRY(b2_mg_set_level(mg, l, A[l], P[l], bdc[l], nbdc[l], npre, npost, omega));
    b2_vec *R = nullptr, *E = nullptr;
    B2_TRY(b2_vec_create(c, n[nlevels - 1], &R));
    B2_TRY(b2_vec_create(c, n[nlevels - 1], &E));
    B2_TRY(b2_vec_put(R, res, n[nlevels - 1]));
    for (int cyc = 0; cyc < ncycles; cyc++) {
      B2_TRY(b2_mg_solve(mg, R, E));
      B2_TRY(b2_vec_get(R, res, n[nlevels - 1]));
      double s = 0.0;
      for (int64_t i = 0; i < n[nlevels - 1]; i++)
        if (free_rows[i]) s += res[i] * res[i];
      trace[cyc] = std::sqrt(s);
    }
    B2_TRY(b2_vec_get(E, eps, n[nlevels - 1]));
    b2_vec_destroy(R);
    b2_vec_destroy(E);
    b2_mg_destroy(mg);
    for (int l = 0; l < nlevels; l++) {
      b2_schwarz_destroy(S[l]);
      b2_csr_destroy(A[l]);
      b2_csr_destroy(P[l]);
    }
    return 0;
  };
  const int rc = run();
  std::free(ctx.red_partial); std::free(ctx.red_result); std::free(ctx.red_counter); std::free(ctx.h_result);
  return rc;
}

}  // extern "C"
