// General sparse products and sums for the plugin surface: what the reference's AMR path asks of SparseMatrix beyond the
// Galerkin product --
//   matrix_RightMatMult / matrix_LeftMatMult (MatMatMult, PetscMatrix.cpp:766-790): _PP[ig] <- _PP[ig] * _PPamr[ig-1]
//                                            (LinearImplicitSystem.cpp:255-258),
//   matrix_ABC (MatMatMatMult, :755-764):    KK <- RRamr * KKamr * PPamr (:336-341),
//   matrix_add / add (MatAXPY, :793-812).
// C = A B by expand - sort - compress, all on the device: one (row, column) key and one product per pair (A_ik, B_kj),
// a STABLE radix sort by key, then the products of equal keys are summed by one thread in their sorted order (k ascending) -- the result does not
// depend on thread scheduling (run to run bit-identical), and structural zeros are kept, as MatMatMult keeps them.
// Memory: 32 B per product; the products of one call are bounded by what fits next to the operands.
#include <cub/cub.cuh>
#include "b2_common.cuh"

namespace {

constexpr int kBlock = 256;

// products per row of A (one thread per row)
__global__ void mm_count_kernel(int64_t nrows, const int64_t* __restrict__ Ap, const int32_t* __restrict__ Ac,
                                const int64_t* __restrict__ Bp, unsigned long long* __restrict__ cnt) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride) {
    unsigned long long s = 0;
    for (int64_t k = Ap[i]; k < Ap[i + 1]; k++) s += (unsigned long long)(Bp[Ac[k] + 1] - Bp[Ac[k]]);
    cnt[i + 1] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) cnt[0] = 0;
}

// one warp per row of A: the products of row i in the order (k ascending, j ascending)
__global__ void mm_expand_kernel(int64_t nrows, const int64_t* __restrict__ Ap, const int32_t* __restrict__ Ac,
                                 const double* __restrict__ Av, const int64_t* __restrict__ Bp, const int32_t* __restrict__ Bc,
                                 const double* __restrict__ Bv, const unsigned long long* __restrict__ off,
                                 unsigned long long* __restrict__ key, double* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w; i < nrows; i += nw) {
    unsigned long long o = off[i];
    for (int64_t k = Ap[i]; k < Ap[i + 1]; k++) {
      const int32_t r = Ac[k];
      const double a = Av[k];
      const int64_t b0 = Bp[r], b1 = Bp[r + 1];
      for (int64_t q = b0 + lane; q < b1; q += 32) {
        key[o + (unsigned long long)(q - b0)] = ((unsigned long long)i << 32) | (unsigned int)Bc[q];
        val[o + (unsigned long long)(q - b0)] = a * Bv[q];
      }
      o += (unsigned long long)(b1 - b0);
    }
  }
}

// head flag of every run of equal keys (an inclusive scan turns it into 1 + run index)
__global__ void mm_heads_kernel(int64_t np, const unsigned long long* __restrict__ key, unsigned long long* __restrict__ head) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < np; t += stride) head[t] = (t == 0 || key[t] != key[t - 1]) ? 1ull : 0ull;
}
// the thread at the head of a run sums it, first product to last
__global__ void mm_compress_kernel(int64_t np, const unsigned long long* __restrict__ key, const double* __restrict__ val,
                                   const unsigned long long* __restrict__ run, unsigned long long* __restrict__ ukey,
                                   double* __restrict__ uval) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < np; t += stride) {
    if (t > 0 && key[t] == key[t - 1]) continue;
    const unsigned long long kk = key[t];
    double s = val[t];
    for (int64_t q = t + 1; q < np && key[q] == kk; q++) s += val[q];
    ukey[run[t] - 1] = kk;
    uval[run[t] - 1] = s;
  }
}

// unique keys -> CSR: entries per row (keys are sorted, so a row's entries are a contiguous run), columns
__global__ void mm_rows_kernel(int64_t nuniq, const unsigned long long* __restrict__ key, int32_t* __restrict__ col,
                               unsigned long long* __restrict__ rowcnt) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nuniq; t += stride) {
    const unsigned long long kk = key[t];
    col[t] = (int32_t)(kk & 0xffffffffull);
    atomicAdd(&rowcnt[(kk >> 32) + 1], 1ull);
  }
}
__global__ void mm_u64_to_i64_kernel(int64_t n, const unsigned long long* __restrict__ a, int64_t* __restrict__ b) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = (int64_t)a[i];
}

// Y += a X for a pattern of X contained in the pattern of Y: one warp per row, bisection in the row of Y
__global__ void axpy_pattern_kernel(int64_t nrows, const int64_t* __restrict__ Xp, const int32_t* __restrict__ Xc,
                                    const double* __restrict__ Xv, const int64_t* __restrict__ Yp, const int32_t* __restrict__ Yc,
                                    double* __restrict__ Yv, double a, int* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w; i < nrows; i += nw) {
    const int64_t y0 = Yp[i], y1 = Yp[i + 1];
    for (int64_t q = Xp[i] + lane; q < Xp[i + 1]; q += 32) {
      const int32_t cq = Xc[q];
      int64_t lo = y0, hi = y1;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (Yc[mid] < cq) lo = mid + 1;
        else hi = mid;
      }
      if (lo < y1 && Yc[lo] == cq) Yv[lo] = fma(a, Xv[q], Yv[lo]);
      else *err = 1;
    }
  }
}
__global__ void count_missing_kernel(int64_t nrows, const int64_t* __restrict__ Xp, const int32_t* __restrict__ Xc,
                                     const int64_t* __restrict__ Yp, const int32_t* __restrict__ Yc, int* __restrict__ missing) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w; i < nrows; i += nw) {
    const int64_t y0 = Yp[i], y1 = Yp[i + 1];
    for (int64_t q = Xp[i] + lane; q < Xp[i + 1]; q += 32) {
      const int32_t cq = Xc[q];
      int64_t lo = y0, hi = y1;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (Yc[mid] < cq) lo = mid + 1;
        else hi = mid;
      }
      if (!(lo < y1 && Yc[lo] == cq)) *missing = 1;
    }
  }
}

int scan_u64(b2_ctx* c, unsigned long long* d, int64_t n) {
  size_t tmp_bytes = 0;
  B2_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, d, d, n, c->stream));
  void* tmp = nullptr;
  B2_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
  cudaError_t e = cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, d, d, n, c->stream);
  c->launches += 2;
  cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  B2_CUDA(e);
  return 0;
}

}  // namespace

extern "C" {

/* C = A B (new matrix; the caller owns it) */
int b2_csr_matmat(const b2_csr* A, const b2_csr* B, b2_csr** out) {
  B2_CHECK(A && B && out, "b2_csr_matmat: null argument");
  *out = nullptr;
  B2_CHECK(A->ncols == B->nrows, "b2_csr_matmat: %lld x %lld times %lld x %lld", (long long)A->nrows, (long long)A->ncols,
           (long long)B->nrows, (long long)B->ncols);
  B2_CHECK(A->nrows < ((int64_t)1 << 31) && B->ncols < ((int64_t)1 << 31), "b2_csr_matmat: dimensions beyond 2^31");
  b2_ctx* c = A->ctx;
  const int64_t m = A->nrows;
  unsigned long long* off = nullptr;
  B2_TRY(b2_malloc(c, &off, (size_t)m + 1));
  B2_LAUNCH(c, mm_count_kernel, b2_grid_for(c, m, kBlock, 8), kBlock, 0, m, A->rowptr, A->col, B->rowptr, off);
  B2_TRY(scan_u64(c, off, m + 1));
  unsigned long long total = 0;
  B2_TRY(b2_download(c, &total, off + m, 1));
  const int64_t np = (int64_t)total;
  b2_csr* C = nullptr;
  if (np == 0) {
    B2_TRY(b2_csr_alloc(c, m, B->ncols, 0, &C));
    B2_CUDA(cudaMemsetAsync(C->rowptr, 0, ((size_t)m + 1) * sizeof(int64_t), c->stream));
    b2_free(c, off, (size_t)m + 1);
    B2_TRY(b2_csr_finalize(C));
    *out = C;
    return 0;
  }
  unsigned long long *k0 = nullptr, *k1 = nullptr, *rowcnt = nullptr;
  double *v0 = nullptr, *v1 = nullptr;
  B2_TRY(b2_malloc(c, &k0, (size_t)np));
  B2_TRY(b2_malloc(c, &k1, (size_t)np));
  B2_TRY(b2_malloc(c, &v0, (size_t)np));
  B2_TRY(b2_malloc(c, &v1, (size_t)np));
  B2_LAUNCH(c, mm_expand_kernel, b2_grid_for(c, m * 32, kBlock, 8), kBlock, 0, m, A->rowptr, A->col, A->val, B->rowptr, B->col, B->val, off,
            k0, v0);
  int cbits = 1;
  while (cbits < 32 && ((int64_t)1 << cbits) < B->ncols) cbits++;
  int rbits = 1;
  while (rbits < 32 && ((int64_t)1 << rbits) < m) rbits++;
  {
    // the row lives in bits 32.., the column in bits 0..cbits: two stable passes (column, then row) skip the dead bits
    size_t tb1 = 0, tb2 = 0;
    B2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb1, k0, k1, v0, v1, np, 0, cbits, c->stream));
    B2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, k1, k0, v1, v0, np, 32, 32 + rbits, c->stream));
    void* tmp = nullptr;
    const size_t tb = tb1 > tb2 ? tb1 : tb2;
    B2_CUDA(cudaMalloc(&tmp, tb ? tb : 16));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tb1, k0, k1, v0, v1, np, 0, cbits, c->stream);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tb2, k1, k0, v1, v0, np, 32, 32 + rbits, c->stream);
    c->launches += 16;
    cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    B2_CUDA(e);
  }
  // compress: run heads -> run index (integer scan) -> every run summed by ONE thread in sorted order (deterministic;
  // a parallel reduce-by-key would associate the floating-point sums differently from run to run)
  unsigned long long* pos = nullptr;
  B2_TRY(b2_malloc(c, &pos, (size_t)np + 1));
  B2_LAUNCH(c, mm_heads_kernel, b2_grid_for(c, np, kBlock, 8), kBlock, 0, np, k0, pos);
  B2_TRY(scan_u64(c, pos, np));
  unsigned long long nu = 0;
  B2_TRY(b2_download(c, &nu, pos + (np - 1), 1));
  const int64_t nuniq = (int64_t)nu;
  B2_LAUNCH(c, mm_compress_kernel, b2_grid_for(c, np, kBlock, 8), kBlock, 0, np, k0, v0, pos, k1, v1);
  B2_CUDA(cudaStreamSynchronize(c->stream));
  b2_free(c, pos, (size_t)np + 1);
  B2_TRY(b2_csr_alloc(c, m, B->ncols, nuniq, &C));
  B2_TRY(b2_malloc(c, &rowcnt, (size_t)m + 1));
  B2_CUDA(cudaMemsetAsync(rowcnt, 0, ((size_t)m + 1) * 8, c->stream));
  B2_LAUNCH(c, mm_rows_kernel, b2_grid_for(c, nuniq, kBlock, 8), kBlock, 0, nuniq, k1, C->col, rowcnt);
  B2_TRY(scan_u64(c, rowcnt, m + 1));
  B2_LAUNCH(c, mm_u64_to_i64_kernel, b2_grid_for(c, m + 1, kBlock, 8), kBlock, 0, m + 1, rowcnt, C->rowptr);
  B2_CUDA(cudaMemcpyAsync(C->val, v1, (size_t)nuniq * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  b2_free(c, off, (size_t)m + 1);
  b2_free(c, k0, (size_t)np);
  b2_free(c, k1, (size_t)np);
  b2_free(c, v0, (size_t)np);
  b2_free(c, v1, (size_t)np);
  b2_free(c, rowcnt, (size_t)m + 1);
  B2_TRY(b2_csr_finalize(C));
  *out = C;
  return 0;
}

/* *contained = 1 when every entry of X's pattern is in Y's pattern (same shape) */
int b2_csr_pattern_contains(const b2_csr* Y, const b2_csr* X, int* contained) {
  B2_CHECK(Y && X && contained, "b2_csr_pattern_contains: null argument");
  B2_CHECK(Y->nrows == X->nrows && Y->ncols == X->ncols, "b2_csr_pattern_contains: shapes differ");
  b2_ctx* c = Y->ctx;
  int* d = nullptr;
  B2_TRY(b2_malloc(c, &d, 1));
  B2_CUDA(cudaMemsetAsync(d, 0, sizeof(int), c->stream));
  if (X->nnz) B2_LAUNCH(c, count_missing_kernel, b2_grid_for(c, X->nrows * 32, kBlock, 8), kBlock, 0, X->nrows, X->rowptr, X->col, Y->rowptr, Y->col, d);
  int miss = 0;
  B2_TRY(b2_download(c, &miss, d, 1));
  b2_free(c, d, 1);
  *contained = !miss;
  return 0;
}

/* Y += a X (MatAXPY with SUBSET_NONZERO_PATTERN); fails if X has an entry outside Y's pattern */
int b2_csr_axpy(b2_csr* Y, double a, const b2_csr* X) {
  B2_CHECK(Y && X, "b2_csr_axpy: null argument");
  B2_CHECK(Y->nrows == X->nrows && Y->ncols == X->ncols, "b2_csr_axpy: shapes differ");
  b2_ctx* c = Y->ctx;
  int* d = nullptr;
  B2_TRY(b2_malloc(c, &d, 1));
  B2_CUDA(cudaMemsetAsync(d, 0, sizeof(int), c->stream));
  if (X->nnz) B2_LAUNCH(c, axpy_pattern_kernel, b2_grid_for(c, X->nrows * 32, kBlock, 8), kBlock, 0, X->nrows, X->rowptr, X->col, X->val, Y->rowptr,
                        Y->col, Y->val, a, d);
  int err = 0;
  B2_TRY(b2_download(c, &err, d, 1));
  b2_free(c, d, 1);
  Y->version++;
  B2_CHECK(!err, "b2_csr_axpy: X has an entry outside the pattern of Y");
  return 0;
}

}  // extern "C"
