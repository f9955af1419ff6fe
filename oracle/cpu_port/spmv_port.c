/* Oracle / CPU baseline (TEST INFRASTRUCTURE ONLY): OpenMP C restatement of the CSR kernels the
 * V-cycle spends its time in -- what PETSc's MatMult / MatMultAdd (PetscVector.cpp:193-247 in the
 * reference), KSPRICHARDSON+PCJACOBI (LinearEquationSolverPetsc.cpp:516-519) and VecAXPY/VecDot do
 * on AIJ matrices.  Used by bench.py's cpu_baseline / --impl reference legs and by tests as a
 * second checker; never linked by the product. */
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* mode 0: y = A x ; 1: y += A x ; 2: y = b - A x ; 3: y = x + omega * dinv * (b - A x) */
void port_spmv(int mode, int64_t nrows, const int64_t* rowptr, const int32_t* col, const double* val,
               const double* x, const double* b, const double* dinv, double* y, double omega, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t r = 0; r < nrows; r++) {
    double s = 0.0;
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) s += val[k] * x[col[k]];
    if (mode == 0) y[r] = s;
    else if (mode == 1) y[r] += s;
    else if (mode == 2) y[r] = b[r] - s;
    else y[r] = x[r] + omega * dinv[r] * (b[r] - s);
  }
}

double port_dot(int64_t n, const double* x, const double* y, int nthreads) {
  double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s) num_threads(nthreads)
  for (int64_t i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}

/* y = a x + b y */
void port_axpby(int64_t n, double a, const double* x, double b, double* y, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t i = 0; i < n; i++) y[i] = a * x[i] + b * y[i];
}

/* z = d .* r */
void port_pmult(int64_t n, const double* d, const double* r, double* z, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t i = 0; i < n; i++) z[i] = d[i] * r[i];
}

int port_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* C = P^T A P onto the known pattern (Cp, Ci) of C -- what MatPtAP does for the Galerkin chain
 * (PetscMatrix.cpp:733-751, LinearImplicitSystem.cpp:347-370), OpenMP over rows in two Gustavson
 * sweeps: T = A P (rows of T kept in per-row buffers of capacity tcap), then C[I,:] = sum_i R[I,i] T[i,:]
 * with R = P^T given in CSR.  Entries of the product outside C's pattern are an error (returns 1). */
#include <stdlib.h>
#include <string.h>
int port_ptap(int64_t nf, int64_t nc, const int64_t* Ap, const int32_t* Ai, const double* Ax, const int64_t* Pp,
              const int32_t* Pi, const double* Px, const int64_t* Rp, const int32_t* Ri, const double* Rx,
              const int64_t* Cp, const int32_t* Ci, double* Cx, int nthreads) {
  /* pass 1: size of every row of T = A P (upper bound: sum of the P row lengths, at most nc) */
  int64_t* Tp = (int64_t*)malloc((size_t)(nf + 1) * sizeof(int64_t));
  if (!Tp) return 2;
  Tp[0] = 0;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t i = 0; i < nf; i++) {
    int64_t cap = 0;
    for (int64_t k = Ap[i]; k < Ap[i + 1]; k++) cap += Pp[Ai[k] + 1] - Pp[Ai[k]];
    Tp[i + 1] = cap < nc ? cap : nc;
  }
  for (int64_t i = 0; i < nf; i++) Tp[i + 1] += Tp[i];
  int32_t* Ti = (int32_t*)malloc((size_t)(Tp[nf] + 1) * sizeof(int32_t));
  double* Tx = (double*)malloc((size_t)(Tp[nf] + 1) * sizeof(double));
  int64_t* Tn = (int64_t*)malloc((size_t)(nf + 1) * sizeof(int64_t));
  if (!Ti || !Tx || !Tn) { free(Tp); free(Ti); free(Tx); free(Tn); return 2; }
  int err = 0;
#pragma omp parallel num_threads(nthreads)
  {
    double* acc = (double*)calloc((size_t)nc, sizeof(double));
    int32_t* mark = (int32_t*)malloc((size_t)nc * sizeof(int32_t));
    for (int64_t j = 0; j < nc; j++) mark[j] = -1;
    /* T = A P */
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < nf; i++) {
      int64_t n = 0;
      int32_t* ti = Ti + Tp[i];
      for (int64_t k = Ap[i]; k < Ap[i + 1]; k++) {
        const double a = Ax[k];
        const int32_t r = Ai[k];
        for (int64_t q = Pp[r]; q < Pp[r + 1]; q++) {
          const int32_t J = Pi[q];
          if (mark[J] != (int32_t)i) { mark[J] = (int32_t)i; ti[n++] = J; acc[J] = a * Px[q]; }
          else acc[J] += a * Px[q];
        }
      }
      for (int64_t t = 0; t < n; t++) { Tx[Tp[i] + t] = acc[ti[t]]; }
      Tn[i] = n;
    }
    /* C = R T, gathered onto C's pattern */
    for (int64_t j = 0; j < nc; j++) mark[j] = -1;
#pragma omp for schedule(dynamic, 64)
    for (int64_t I = 0; I < nc; I++) {
      for (int64_t k = Cp[I]; k < Cp[I + 1]; k++) { mark[Ci[k]] = (int32_t)I; acc[Ci[k]] = 0.0; }
      for (int64_t k = Rp[I]; k < Rp[I + 1]; k++) {
        const double rv = Rx[k];
        const int64_t i = Ri[k];
        for (int64_t t = 0; t < Tn[i]; t++) {
          const int32_t J = Ti[Tp[i] + t];
          if (mark[J] != (int32_t)I) { if (rv * Tx[Tp[i] + t] != 0.0) err = 1; }
          else acc[J] += rv * Tx[Tp[i] + t];
        }
      }
      for (int64_t k = Cp[I]; k < Cp[I + 1]; k++) Cx[k] = acc[Ci[k]];
    }
    free(acc);
    free(mark);
  }
  free(Tp); free(Ti); free(Tx); free(Tn);
  return err;
}
