// Galerkin triple product C = P^T A P (numeric phase onto a known pattern) and y = A^T x.
// Replaces SparseMatrix::matrix_PtAP -> MatPtAP (reference src/03_algebra/01_matrices/
// PetscMatrix.cpp:733-751) as called down the hierarchy by LinearImplicitSystem::MGsolve
// (src/08_equations/00_stationary/LinearImplicitSystem.cpp:347-370), and MatMultTranspose
// (PetscVector.cpp:219-228).
//
// One warp per FINE row i:
//   1. t = (A P)[i,:] accumulated in a per-warp shared-memory hash table keyed by coarse column
//      (<= 125 distinct columns for triquadratic elements, table of 512);
//   2. for every coarse row I with P[i,I] != 0:  C[I,J] += P[i,I] * t[J]  (binary search in row I
//      of C, fp64 atomicAdd in L2).
// No temporary A*P matrix is materialised (it would be 18 GB at 128^3 Hex27).
#include "b2_common.cuh"

namespace {

constexpr int kHash = 512;
constexpr int kWarps = 4;

__device__ __forceinline__ int64_t lower_find(const int32_t* __restrict__ col, int64_t s, int64_t e, int32_t c) {
  int64_t lo = s, hi = e;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1;
    else hi = mid;
  }
  return (lo < e && col[lo] == c) ? lo : -1;
}

__global__ void __launch_bounds__(kWarps * 32) ptap_kernel(
    int64_t nrows_f, const int64_t* __restrict__ Ap, const int32_t* __restrict__ Ac, const double* __restrict__ Av,
    const int64_t* __restrict__ Pp, const int32_t* __restrict__ Pc, const double* __restrict__ Pv,
    const int64_t* __restrict__ Cp, const int32_t* __restrict__ Cc, double* __restrict__ Cv, int* err) {
  __shared__ int32_t hkey[kWarps][kHash];
  __shared__ double hval[kWarps][kHash];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int32_t* key = hkey[wib];
  double* acc = hval[wib];
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w; i < nrows_f; i += nw) {
    const int64_t ps = Pp[i], pe = Pp[i + 1];
    // rows of P that are entirely zero (fine Dirichlet dofs) contribute nothing
    bool any = false;
    for (int64_t q = ps + lane; q < pe; q += 32) any |= (Pv[q] != 0.0);
    if (!__any_sync(0xffffffffu, any)) continue;
    for (int t = lane; t < kHash; t += 32) { key[t] = -1; acc[t] = 0.0; }
    __syncwarp();
    // 1. t = A[i,:] * P
    for (int64_t k = Ap[i] + lane; k < Ap[i + 1]; k += 32) {
      const double a = Av[k];
      if (a == 0.0) continue;
      const int32_t j = Ac[k];
      for (int64_t q = Pp[j]; q < Pp[j + 1]; q++) {
        const double pv = Pv[q];
        if (pv == 0.0) continue;
        const int32_t J = Pc[q];
        unsigned h = ((unsigned)J * 2654435761u) & (kHash - 1);
        int probes = 0;
        while (true) {
          const int32_t old = atomicCAS(&key[h], -1, J);
          if (old == -1 || old == J) break;
          h = (h + 1) & (kHash - 1);
          if (++probes >= kHash) { atomicExch(err, 1); break; }
        }
        if (probes < kHash) atomicAdd(&acc[h], a * pv);
      }
    }
    __syncwarp();
    // 2. scatter P[i,I] * t into C
    for (int64_t q = ps; q < pe; q++) {
      const double pv = Pv[q];
      if (pv == 0.0) continue;
      const int32_t I = Pc[q];
      const int64_t cs = Cp[I], ce = Cp[I + 1];
      for (int t = lane; t < kHash; t += 32) {
        const int32_t J = key[t];
        if (J < 0) continue;
        const double v = acc[t];
        if (v == 0.0) continue;
        const int64_t pos = lower_find(Cc, cs, ce, J);
        if (pos < 0) atomicExch(err, 2);
        else atomicAdd(&Cv[pos], pv * v);
      }
    }
    __syncwarp();
  }
}

__global__ void spmv_t_kernel(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                              const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w; r < nrows; r += nw) {
    const double xr = x[r];
    if (xr == 0.0) continue;
    for (int64_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) atomicAdd(&y[col[k]], val[k] * xr);
  }
}

}  // namespace

extern "C" {

int b2_csr_ptap(const b2_csr* P, const b2_csr* A, b2_csr* C) {
  B2_CHECK(P->nrows == A->nrows && A->nrows == A->ncols && C->nrows == P->ncols && C->ncols == P->ncols,
           "b2_csr_ptap: shape mismatch P %lldx%lld A %lldx%lld C %lldx%lld", (long long)P->nrows,
           (long long)P->ncols, (long long)A->nrows, (long long)A->ncols, (long long)C->nrows, (long long)C->ncols);
  b2_ctx* c = A->ctx;
  int* d_err = nullptr;
  B2_TRY(b2_malloc(c, &d_err, 1));
  B2_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  B2_CUDA(cudaMemsetAsync(C->val, 0, (size_t)C->nnz * sizeof(double), c->stream));
  const int grid = b2_grid_for(c, A->nrows * 32, kWarps * 32, 8);
  B2_LAUNCH(c, ptap_kernel, grid, kWarps * 32, 0, A->nrows, A->rowptr, A->col, A->val, P->rowptr, P->col, P->val,
            C->rowptr, C->col, C->val, d_err);
  int err = 0;
  B2_TRY(b2_download(c, &err, d_err, 1));
  b2_free(c, d_err, 1);
  B2_CHECK(err != 1, "b2_csr_ptap: a row of A*P has more than %d distinct columns", kHash);
  B2_CHECK(err != 2, "b2_csr_ptap: product entry outside the pattern of C");
  return 0;
}

int b2_csr_spmv_t(const b2_csr* A, const b2_vec* x, b2_vec* y) {
  B2_CHECK(x->n >= A->nrows && y->n >= A->ncols && x != y, "b2_csr_spmv_t: bad operands");
  b2_ctx* c = A->ctx;
  B2_CUDA(cudaMemsetAsync(y->d, 0, (size_t)A->ncols * sizeof(double), c->stream));
  if (A->nrows == 0) return 0;
  const int grid = b2_grid_for(c, A->nrows * 32, 256, 8);
  B2_LAUNCH(c, spmv_t_kernel, grid, 256, 0, A->nrows, A->rowptr, A->col, A->val, x->d, y->d);
  return 0;
}

}  // extern "C"
