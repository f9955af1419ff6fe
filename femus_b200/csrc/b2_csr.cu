// Device CSR matrix (fp64 values, int32 columns, int64 row pointers): storage, setup kernels
// (pattern from element lists, transpose, row/column zeroing, diagonal, staged block adds) and
// the row chunking consumed by the SpMV family (b2_spmv.cu).  Replaces PetscMatrix
// (reference src/03_algebra/01_matrices/PetscMatrix.cpp).
#include "b2_common.cuh"
#include <cub/cub.cuh>

namespace {

constexpr int kBlock = 256;

// ------------------------------------------------------------------------------------------
__global__ void row_stats_kernel(int64_t nrows, const int64_t* __restrict__ rowptr, int* max_row) {
  int m = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const int len = (int)(rowptr[r + 1] - rowptr[r]);
    m = max(m, len);
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(max_row, m);
}

__device__ __forceinline__ int64_t find_in_row(const int32_t* __restrict__ col, int64_t s, int64_t e, int32_t c) {
  // lower_bound over the sorted columns of one row; returns -1 if absent
  int64_t lo = s, hi = e;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1;
    else hi = mid;
  }
  return (lo < e && col[lo] == c) ? lo : -1;
}

__global__ void zero_rows_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                 double* __restrict__ val, const int32_t* __restrict__ rows, int64_t n, double diag) {
  // one warp per listed row
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w; i < n; i += nw) {
    const int32_t r = rows[i];
    const int64_t s = rowptr[r], e = rowptr[r + 1];
    for (int64_t k = s + lane; k < e; k += 32) val[k] = (col[k] == r) ? diag : 0.0;
  }
}

__global__ void mark_kernel(unsigned char* __restrict__ mask, const int32_t* __restrict__ idx, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) mask[idx[i]] = 1;
}
__global__ void zero_cols_kernel(int64_t nnz, const int32_t* __restrict__ col, double* __restrict__ val,
                                 const unsigned char* __restrict__ mask) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride)
    if (mask[col[k]]) val[k] = 0.0;
}

// position of the diagonal inside every row (-1: structurally absent): found once per pattern, then
// MatGetDiagonal is a plain gather
__global__ void diag_pos_kernel(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                int32_t* __restrict__ pos) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const int64_t p = find_in_row(col, rowptr[r], rowptr[r + 1], (int32_t)r);
    pos[r] = p >= 0 ? (int32_t)(p - rowptr[r]) : -1;
  }
}
__global__ void diag_gather_kernel(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ pos,
                                   const double* __restrict__ val, double* __restrict__ d) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const int32_t p = pos[r];
    d[r] = p >= 0 ? val[rowptr[r] + p] : 0.0;
  }
}

// staged host blocks -> CSR (compat path of add_matrix_blocked / insert_row)
__global__ void add_blocks_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                  double* __restrict__ val, int64_t nblk, int nrow, int ncol,
                                  const int32_t* __restrict__ rows, const int32_t* __restrict__ cols,
                                  const double* __restrict__ v, int* err) {
  const int64_t total = nblk * nrow * ncol;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t bI = t / (nrow * ncol);
    const int ij = (int)(t - bI * nrow * ncol);
    const int i = ij / ncol, j = ij - i * ncol;
    const int32_t r = rows[bI * nrow + i], cc = cols[bI * ncol + j];
    const int64_t p = find_in_row(col, rowptr[r], rowptr[r + 1], cc);
    if (p < 0) atomicExch(err, 1);
    else atomicAdd(&val[p], v[t]);
  }
}
__global__ void set_rows_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                double* __restrict__ val, int64_t nset, const int32_t* __restrict__ rows,
                                const int64_t* __restrict__ ptr, const int32_t* __restrict__ cols,
                                const double* __restrict__ v, int* err) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w; i < nset; i += nw) {
    const int32_t r = rows[i];
    for (int64_t k = ptr[i] + lane; k < ptr[i + 1]; k += 32) {
      const int64_t p = find_in_row(col, rowptr[r], rowptr[r + 1], cols[k]);
      if (p < 0) atomicExch(err, 1);
      else val[p] = v[k];
    }
  }
}

// ---- transpose through a key sort: key = (col << 32) | row, payload = source position -------
__global__ void transpose_keys_kernel(int64_t nrows, const int64_t* __restrict__ rowptr,
                                      const int32_t* __restrict__ col, unsigned long long* __restrict__ keys,
                                      unsigned int* __restrict__ pos) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w; r < nrows; r += nw) {
    for (int64_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) {
      keys[k] = ((unsigned long long)(unsigned int)col[k] << 32) | (unsigned long long)(unsigned int)r;
      pos[k] = (unsigned int)k;
    }
  }
}
__global__ void transpose_fill_kernel(int64_t nnz, const unsigned long long* __restrict__ keys,
                                      const unsigned int* __restrict__ pos, const double* __restrict__ val,
                                      int32_t* __restrict__ tcol, double* __restrict__ tval,
                                      unsigned long long* __restrict__ count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    const unsigned long long key = keys[k];
    tcol[k] = (int32_t)(key & 0xffffffffull);
    tval[k] = val[pos[k]];
    atomicAdd(&count[(key >> 32) + 1], 1ull);
  }
}

// ---- pattern from element->dof lists ------------------------------------------------------
__global__ void adj_count_kernel(int64_t total, const int32_t* __restrict__ dof, unsigned long long* __restrict__ cnt) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
    atomicAdd(&cnt[dof[t] + 1], 1ull);
}
__global__ void adj_fill_kernel(int64_t nel, int nve, const int32_t* __restrict__ dof,
                                const unsigned long long* __restrict__ adjptr, unsigned int* __restrict__ cursor,
                                int32_t* __restrict__ adj) {
  const int64_t total = nel * nve;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int32_t r = dof[t];
    const unsigned int k = atomicAdd(&cursor[r], 1u);
    adj[adjptr[r] + k] = (int32_t)(t / nve);
  }
}

constexpr int kCand = 1024;   // candidate columns per row held in shared memory (per warp)

// One warp per row: gather the dofs of all adjacent elements, bitonic-sort them in shared
// memory, drop duplicates.  FILL=false writes the row length, FILL=true the sorted columns.
template <bool FILL>
__global__ void __launch_bounds__(128) pattern_rows_kernel(int64_t nrows, int nve, const int32_t* __restrict__ dof,
                                                           const unsigned long long* __restrict__ adjptr,
                                                           const int32_t* __restrict__ adj,
                                                           unsigned long long* __restrict__ rowlen_or_ptr,
                                                           int32_t* __restrict__ col, int* err) {
  __shared__ int32_t buf[4][kCand];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int32_t* s = buf[wib];
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w; r < nrows; r += nw) {
    const unsigned long long a0 = adjptr[r], a1 = adjptr[r + 1];
    const int ncand = (int)(a1 - a0) * nve;
    if (ncand > kCand) {
      if (lane == 0) atomicExch(err, 2);
      continue;
    }
    int n2 = 32;
    while (n2 < ncand) n2 <<= 1;
    for (int t = lane; t < n2; t += 32) {
      int32_t v = 0x7fffffff;
      if (t < ncand) {
        const int32_t el = adj[a0 + t / nve];
        v = dof[(int64_t)el * nve + (t % nve)];
      }
      s[t] = v;
    }
    __syncwarp();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < n2; t += 32) {
          const int p = t ^ j;
          if (p > t) {
            const int32_t x = s[t], y = s[p];
            const bool up = ((t & k) == 0);
            if ((x > y) == up) { s[t] = y; s[p] = x; }
          }
        }
        __syncwarp();
      }
    }
    // unique count / compaction
    int base = 0;
    const int64_t out0 = FILL ? (int64_t)rowlen_or_ptr[r] : 0;
    for (int t0 = 0; t0 < n2; t0 += 32) {
      const int t = t0 + lane;
      const int32_t v = s[t];
      const bool keep = (v != 0x7fffffff) && (t == 0 || s[t - 1] != v);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (FILL && keep) col[out0 + base + __popc(m & ((1u << lane) - 1u))] = v;
      base += __popc(m);
    }
    if (!FILL && lane == 0) rowlen_or_ptr[r + 1] = (unsigned long long)base;
    __syncwarp();
  }
}

__global__ void u64_to_i64_kernel(int64_t n, const unsigned long long* __restrict__ a, int64_t* __restrict__ b) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = (int64_t)a[i];
}

int inclusive_scan_u64(b2_ctx* c, unsigned long long* d, int64_t n) {
  size_t tmp_bytes = 0;
  B2_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, d, d, (int)0, c->stream));
  B2_CHECK(n < (int64_t)1 << 31, "scan length too large");
  B2_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, d, d, (int)n, c->stream));
  void* tmp = nullptr;
  B2_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
  cudaError_t e = cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, d, d, (int)n, c->stream);
  c->launches += 2;
  cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  B2_CUDA(e);
  return 0;
}

}  // namespace

__global__ void zero_cols_notowned_kernel(int64_t nnz, const int32_t* __restrict__ col, double* __restrict__ val,
                                          const uint8_t* __restrict__ owned) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride)
    if (!owned[col[k]]) val[k] = 0.0;
}
int b2_csr_zero_cols_notowned(b2_csr* A, const uint8_t* d_owned) {
  if (A->nnz == 0) return 0;
  b2_ctx* c = A->ctx;
  B2_LAUNCH(c, zero_cols_notowned_kernel, b2_grid_for(c, A->nnz, kBlock, 8), kBlock, 0, A->nnz, A->col, A->val, d_owned);
  return 0;
}
// rows[i] -> zero, diagonal = diag where the row is owned by this rank, 0 otherwise (the owner's
// 1 completes the identity row in the interface sum)
__global__ void zero_rows_owned_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                       double* __restrict__ val, const int32_t* __restrict__ rows, int64_t n, double diag,
                                       const uint8_t* __restrict__ owned) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w; i < n; i += nw) {
    const int32_t r = rows[i];
    const double d = owned[r] ? diag : 0.0;
    for (int64_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) val[k] = (col[k] == r) ? d : 0.0;
  }
}
int b2_csr_zero_rows_dev(b2_csr* A, const int32_t* d_rows, int64_t n, double diag, const uint8_t* d_owned) {
  if (n == 0) return 0;
  A->version++;
  b2_ctx* c = A->ctx;
  const int grid = b2_grid_for(c, n * 32, kBlock, 8);
  if (d_owned) B2_LAUNCH(c, zero_rows_owned_kernel, grid, kBlock, 0, A->rowptr, A->col, A->val, d_rows, n, diag, d_owned);
  else B2_LAUNCH(c, zero_rows_kernel, grid, kBlock, 0, A->rowptr, A->col, A->val, d_rows, n, diag);
  return 0;
}

int b2_csr_alloc(b2_ctx* c, int64_t nrows, int64_t ncols, int64_t nnz, b2_csr** out) {
  b2_csr* A = new b2_csr();
  A->ctx = c;
  A->nrows = nrows;
  A->ncols = ncols;
  A->nnz = nnz;
  A->tpr = 8;
  A->max_row = 0;
  A->last_ms = 0.;
  A->version = 0;
  A->diag_pos = nullptr;
  A->chunk_row = nullptr;
  A->dict_ptr = nullptr;
  A->cdesc = nullptr;
  A->dict = nullptr;
  A->lidx = nullptr;
  A->nchunks = 0;
  A->dict_total = 0;
  A->dict_cap = 0;
  B2_TRY(b2_malloc(c, &A->rowptr, (size_t)nrows + 3));
  B2_TRY(b2_malloc(c, &A->col, (size_t)nnz + 16));
  B2_TRY(b2_malloc(c, &A->val, (size_t)nnz + 16));
  *out = A;
  return 0;
}

int b2_csr_finalize(b2_csr* A) {
  b2_ctx* c = A->ctx;
  int* d_max = nullptr;
  B2_TRY(b2_malloc(c, &d_max, 1));
  B2_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), c->stream));
  if (A->nrows > 0) {
    const int grid = b2_grid_for(c, A->nrows, kBlock, 8);
    B2_LAUNCH(c, row_stats_kernel, grid, kBlock, 0, A->nrows, A->rowptr, d_max);
  }
  B2_TRY(b2_download(c, &A->max_row, d_max, 1));
  b2_free(c, d_max, 1);
  const double mean = A->nrows ? (double)A->nnz / (double)A->nrows : 0.;
  int tpr = 1;
  while (tpr < 32 && tpr * 3 < mean) tpr <<= 1;   // ~3+ entries per lane
  A->tpr = tpr;
  B2_TRY(b2_csr_build_chunks(A));
  return 0;
}

extern "C" {

int b2_csr_create(b2_ctx* c, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int32_t* col,
                  const double* vals, b2_csr** out) {
  *out = nullptr;
  B2_CHECK(c && nrows >= 0 && ncols >= 0 && rowptr, "b2_csr_create: bad arguments");
  const int64_t nnz = rowptr[nrows];
  b2_csr* A = nullptr;
  B2_TRY(b2_csr_alloc(c, nrows, ncols, nnz, &A));
  B2_TRY(b2_upload(c, A->rowptr, rowptr, (size_t)nrows + 1));
  B2_TRY(b2_upload(c, A->col, col, (size_t)nnz));
  if (vals) B2_TRY(b2_upload(c, A->val, vals, (size_t)nnz));
  else B2_CUDA(cudaMemsetAsync(A->val, 0, (size_t)nnz * sizeof(double), c->stream));
  B2_TRY(b2_csr_finalize(A));
  *out = A;
  return 0;
}

int b2_csr_create_from_elements(b2_ctx* c, int64_t nrows, int64_t nel, int nve, const int32_t* dof, b2_csr** out) {
  *out = nullptr;
  B2_CHECK(c && nrows > 0 && nel > 0 && nve > 0 && dof, "b2_csr_create_from_elements: bad arguments");
  const int64_t total = nel * nve;
  int32_t* d_dof = nullptr;
  unsigned long long* adjptr = nullptr;
  unsigned int* cursor = nullptr;
  int32_t* adj = nullptr;
  unsigned long long* rp = nullptr;
  int* d_err = nullptr;
  B2_TRY(b2_malloc(c, &d_dof, (size_t)total));
  B2_TRY(b2_upload(c, d_dof, dof, (size_t)total));
  B2_TRY(b2_malloc(c, &adjptr, (size_t)nrows + 1));
  B2_TRY(b2_malloc(c, &cursor, (size_t)nrows));
  B2_TRY(b2_malloc(c, &adj, (size_t)total));
  B2_TRY(b2_malloc(c, &rp, (size_t)nrows + 1));
  B2_TRY(b2_malloc(c, &d_err, 1));
  B2_CUDA(cudaMemsetAsync(adjptr, 0, ((size_t)nrows + 1) * 8, c->stream));
  B2_CUDA(cudaMemsetAsync(cursor, 0, (size_t)nrows * 4, c->stream));
  B2_CUDA(cudaMemsetAsync(rp, 0, ((size_t)nrows + 1) * 8, c->stream));
  B2_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  int grid = b2_grid_for(c, total, kBlock, 8);
  B2_LAUNCH(c, adj_count_kernel, grid, kBlock, 0, total, d_dof, adjptr);
  B2_TRY(inclusive_scan_u64(c, adjptr, nrows + 1));
  B2_LAUNCH(c, adj_fill_kernel, grid, kBlock, 0, nel, nve, d_dof, adjptr, cursor, adj);
  const int gridw = b2_grid_for(c, nrows * 32, 128, 16);
  B2_LAUNCH(c, pattern_rows_kernel<false>, gridw, 128, 0, nrows, nve, d_dof, adjptr, adj, rp, (int32_t*)nullptr, d_err);
  B2_TRY(inclusive_scan_u64(c, rp, nrows + 1));
  unsigned long long nnz_u = 0;
  B2_TRY(b2_download(c, &nnz_u, rp + nrows, 1));
  int err = 0;
  B2_TRY(b2_download(c, &err, d_err, 1));
  B2_CHECK(err == 0, "pattern build: a node touches more than %d candidate columns", kCand);
  b2_csr* A = nullptr;
  B2_TRY(b2_csr_alloc(c, nrows, nrows, (int64_t)nnz_u, &A));
  B2_LAUNCH(c, pattern_rows_kernel<true>, gridw, 128, 0, nrows, nve, d_dof, adjptr, adj, rp, A->col, d_err);
  grid = b2_grid_for(c, nrows + 1, kBlock, 8);
  B2_LAUNCH(c, u64_to_i64_kernel, grid, kBlock, 0, nrows + 1, rp, A->rowptr);
  B2_CUDA(cudaMemsetAsync(A->val, 0, (size_t)A->nnz * sizeof(double), c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  b2_free(c, d_dof, (size_t)total);
  b2_free(c, adjptr, (size_t)nrows + 1);
  b2_free(c, cursor, (size_t)nrows);
  b2_free(c, adj, (size_t)total);
  b2_free(c, rp, (size_t)nrows + 1);
  b2_free(c, d_err, 1);
  B2_TRY(b2_csr_finalize(A));
  *out = A;
  return 0;
}

int b2_csr_destroy(b2_csr* A) {
  if (!A) return 0;
  cudaStreamSynchronize(A->ctx->stream);
  b2_free(A->ctx, A->rowptr, (size_t)A->nrows + 3);
  b2_csr_free_plan(A);
  if (A->diag_pos) b2_free(A->ctx, A->diag_pos, (size_t)A->nrows);
  b2_free(A->ctx, A->col, (size_t)A->nnz + 16);
  b2_free(A->ctx, A->val, (size_t)A->nnz + 16);
  delete A;
  return 0;
}
int64_t b2_csr_nrows(const b2_csr* A) { return A->nrows; }
int64_t b2_csr_ncols(const b2_csr* A) { return A->ncols; }
int64_t b2_csr_nnz(const b2_csr* A) { return A->nnz; }
double b2_csr_last_kernel_ms(const b2_csr* A) { return A->last_ms; }

int b2_csr_get(const b2_csr* A, int64_t* rowptr, int32_t* col, double* vals) {
  if (rowptr) B2_TRY(b2_download(A->ctx, rowptr, A->rowptr, (size_t)A->nrows + 1));
  if (col) B2_TRY(b2_download(A->ctx, col, A->col, (size_t)A->nnz));
  if (vals) B2_TRY(b2_download(A->ctx, vals, A->val, (size_t)A->nnz));
  return 0;
}
int b2_csr_put_vals(b2_csr* A, const double* vals) {
  A->version++;
  return b2_upload(A->ctx, A->val, vals, (size_t)A->nnz);
}
int b2_csr_zero(b2_csr* A) {
  A->version++;
  B2_CUDA(cudaMemsetAsync(A->val, 0, (size_t)A->nnz * sizeof(double), A->ctx->stream));
  return 0;
}
int b2_csr_copy_vals(b2_csr* dst, const b2_csr* src) {
  B2_CHECK(dst->nnz == src->nnz && dst->nrows == src->nrows, "b2_csr_copy_vals: pattern mismatch");
  dst->version++;
  B2_CUDA(cudaMemcpyAsync(dst->val, src->val, (size_t)src->nnz * sizeof(double), cudaMemcpyDeviceToDevice,
                          dst->ctx->stream));
  return 0;
}

int b2_csr_add_blocks(b2_csr* A, int64_t nblk, int nrow, int ncol, const int32_t* rows, const int32_t* cols,
                      const double* vals) {
  if (nblk == 0) return 0;
  A->version++;
  b2_ctx* c = A->ctx;
  int32_t *d_r = nullptr, *d_c = nullptr;
  double* d_v = nullptr;
  int* d_err = nullptr;
  const size_t nr = (size_t)nblk * nrow, nc = (size_t)nblk * ncol, nv = (size_t)nblk * nrow * ncol;
  B2_TRY(b2_malloc(c, &d_r, nr));
  B2_TRY(b2_malloc(c, &d_c, nc));
  B2_TRY(b2_malloc(c, &d_v, nv));
  B2_TRY(b2_malloc(c, &d_err, 1));
  B2_TRY(b2_upload(c, d_r, rows, nr));
  B2_TRY(b2_upload(c, d_c, cols, nc));
  B2_TRY(b2_upload(c, d_v, vals, nv));
  B2_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  const int grid = b2_grid_for(c, (int64_t)nv, kBlock, 8);
  B2_LAUNCH(c, add_blocks_kernel, grid, kBlock, 0, A->rowptr, A->col, A->val, nblk, nrow, ncol, d_r, d_c, d_v, d_err);
  int err = 0;
  B2_TRY(b2_download(c, &err, d_err, 1));
  b2_free(c, d_r, nr);
  b2_free(c, d_c, nc);
  b2_free(c, d_v, nv);
  b2_free(c, d_err, 1);
  B2_CHECK(err == 0, "b2_csr_add_blocks: entry outside the preallocated pattern");
  return 0;
}

int b2_csr_set_rows(b2_csr* A, int64_t nset, const int32_t* rows, const int64_t* ptr, const int32_t* cols,
                    const double* vals) {
  if (nset == 0) return 0;
  A->version++;
  b2_ctx* c = A->ctx;
  const size_t nv = (size_t)ptr[nset];
  int32_t *d_r = nullptr, *d_c = nullptr;
  int64_t* d_p = nullptr;
  double* d_v = nullptr;
  int* d_err = nullptr;
  B2_TRY(b2_malloc(c, &d_r, (size_t)nset));
  B2_TRY(b2_malloc(c, &d_p, (size_t)nset + 1));
  B2_TRY(b2_malloc(c, &d_c, nv));
  B2_TRY(b2_malloc(c, &d_v, nv));
  B2_TRY(b2_malloc(c, &d_err, 1));
  B2_TRY(b2_upload(c, d_r, rows, (size_t)nset));
  B2_TRY(b2_upload(c, d_p, ptr, (size_t)nset + 1));
  B2_TRY(b2_upload(c, d_c, cols, nv));
  B2_TRY(b2_upload(c, d_v, vals, nv));
  B2_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  const int grid = b2_grid_for(c, nset * 32, kBlock, 8);
  B2_LAUNCH(c, set_rows_kernel, grid, kBlock, 0, A->rowptr, A->col, A->val, nset, d_r, d_p, d_c, d_v, d_err);
  int err = 0;
  B2_TRY(b2_download(c, &err, d_err, 1));
  b2_free(c, d_r, (size_t)nset);
  b2_free(c, d_p, (size_t)nset + 1);
  b2_free(c, d_c, nv);
  b2_free(c, d_v, nv);
  b2_free(c, d_err, 1);
  B2_CHECK(err == 0, "b2_csr_set_rows: entry outside the preallocated pattern");
  return 0;
}

int b2_csr_zero_rows(b2_csr* A, const int32_t* rows, int64_t n, double diag) {
  if (n == 0) return 0;
  A->version++;
  b2_ctx* c = A->ctx;
  int32_t* d_r = nullptr;
  B2_TRY(b2_malloc(c, &d_r, (size_t)n));
  B2_TRY(b2_upload(c, d_r, rows, (size_t)n));
  const int grid = b2_grid_for(c, n * 32, kBlock, 8);
  B2_LAUNCH(c, zero_rows_kernel, grid, kBlock, 0, A->rowptr, A->col, A->val, d_r, n, diag);
  B2_CUDA(cudaStreamSynchronize(c->stream));
  b2_free(c, d_r, (size_t)n);
  return 0;
}

int b2_csr_zero_cols(b2_csr* A, const int32_t* cols, int64_t n) {
  if (n == 0 || A->nnz == 0) return 0;
  A->version++;
  b2_ctx* c = A->ctx;
  int32_t* d_c = nullptr;
  unsigned char* mask = nullptr;
  B2_TRY(b2_malloc(c, &d_c, (size_t)n));
  B2_TRY(b2_malloc(c, &mask, (size_t)A->ncols));
  B2_TRY(b2_upload(c, d_c, cols, (size_t)n));
  B2_CUDA(cudaMemsetAsync(mask, 0, (size_t)A->ncols, c->stream));
  int grid = b2_grid_for(c, n, kBlock, 8);
  B2_LAUNCH(c, mark_kernel, grid, kBlock, 0, mask, d_c, n);
  grid = b2_grid_for(c, A->nnz, kBlock, 8);
  B2_LAUNCH(c, zero_cols_kernel, grid, kBlock, 0, A->nnz, A->col, A->val, mask);
  B2_CUDA(cudaStreamSynchronize(c->stream));
  b2_free(c, d_c, (size_t)n);
  b2_free(c, mask, (size_t)A->ncols);
  return 0;
}

int b2_csr_diag(const b2_csr* A, b2_vec* d) {
  B2_CHECK(d->n >= A->nrows, "b2_csr_diag: vector too short");
  b2_ctx* c = A->ctx;
  const int grid = b2_grid_for(c, A->nrows, kBlock, 8);
  if (A->nrows == 0) return 0;
  if (!A->diag_pos) {       // first call on this pattern (the pattern of a b2_csr never changes)
    b2_csr* M = const_cast<b2_csr*>(A);
    B2_TRY(b2_malloc(c, &M->diag_pos, (size_t)A->nrows));
    B2_LAUNCH(c, diag_pos_kernel, grid, kBlock, 0, A->nrows, A->rowptr, A->col, M->diag_pos);
  }
  B2_LAUNCH(c, diag_gather_kernel, grid, kBlock, 0, A->nrows, A->rowptr, A->diag_pos, A->val, d->d);
  return 0;
}

int b2_csr_transpose(const b2_csr* A, b2_csr** out) {
  *out = nullptr;
  b2_ctx* c = A->ctx;
  B2_CHECK(A->nnz < ((int64_t)1 << 31), "b2_csr_transpose: nnz >= 2^31 not supported");
  b2_csr* T = nullptr;
  B2_TRY(b2_csr_alloc(c, A->ncols, A->nrows, A->nnz, &T));
  const size_t nnz = (size_t)A->nnz;
  unsigned long long *k0 = nullptr, *k1 = nullptr, *cnt = nullptr;
  unsigned int *p0 = nullptr, *p1 = nullptr;
  B2_TRY(b2_malloc(c, &k0, nnz));
  B2_TRY(b2_malloc(c, &k1, nnz));
  B2_TRY(b2_malloc(c, &p0, nnz));
  B2_TRY(b2_malloc(c, &p1, nnz));
  B2_TRY(b2_malloc(c, &cnt, (size_t)A->ncols + 1));
  B2_CUDA(cudaMemsetAsync(cnt, 0, ((size_t)A->ncols + 1) * 8, c->stream));
  if (nnz) {
    int grid = b2_grid_for(c, A->nrows * 32, kBlock, 8);
    B2_LAUNCH(c, transpose_keys_kernel, grid, kBlock, 0, A->nrows, A->rowptr, A->col, k0, p0);
    size_t tmp_bytes = 0;
    B2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, p0, p1, (int)nnz, 0, 64, c->stream));
    void* tmp = nullptr;
    B2_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, p0, p1, (int)nnz, 0, 64, c->stream);
    c->launches += 8;
    cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    B2_CUDA(e);
    grid = b2_grid_for(c, (int64_t)nnz, kBlock, 8);
    B2_LAUNCH(c, transpose_fill_kernel, grid, kBlock, 0, (int64_t)nnz, k1, p1, A->val, T->col, T->val, cnt);
  }
  B2_TRY(inclusive_scan_u64(c, cnt, A->ncols + 1));
  const int grid = b2_grid_for(c, A->ncols + 1, kBlock, 8);
  B2_LAUNCH(c, u64_to_i64_kernel, grid, kBlock, 0, A->ncols + 1, cnt, T->rowptr);
  B2_CUDA(cudaStreamSynchronize(c->stream));
  b2_free(c, k0, nnz);
  b2_free(c, k1, nnz);
  b2_free(c, p0, nnz);
  b2_free(c, p1, nnz);
  b2_free(c, cnt, (size_t)A->ncols + 1);
  B2_TRY(b2_csr_finalize(T));
  *out = T;
  return 0;
}

}  // extern "C"
