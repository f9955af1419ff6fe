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
