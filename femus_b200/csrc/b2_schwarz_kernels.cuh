// Device code of the element-block smoother (b2_schwarz.cu), kept free of host / runtime calls so that the SAME
// source also compiles for the CPU thread emulator of tests/cpp/cuda_emu.hpp (tests/test_kernel_emulation.py runs
// these kernels on host threads against the oracle when no GPU is present).  Included inside an anonymous namespace.
#pragma once
#ifndef B2_DYN_SHARED
#define B2_DYN_SHARED(type, name) extern __shared__ type name[]
#endif

constexpr int kApplyThreads = 256;
constexpr int kRowBatch = 8;         // row entries per lane in flight in the row-walking block solves
constexpr int kInvertThreads = 512;

__global__ void schwarz_extract_kernel(int64_t nblocks, const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_dofs,
                                       const int64_t* __restrict__ inv_ptr, const int64_t* __restrict__ rowptr,
                                       const int32_t* __restrict__ col, const double* __restrict__ val, double* __restrict__ inv) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const int32_t* D = blk_dofs + blk_ptr[b];
    const int m = (int)(blk_ptr[b + 1] - blk_ptr[b]);
    double* M = inv + inv_ptr[b];
    for (int i = warp; i < m; i += nwarps) {
      double* row = M + (int64_t)i * m;
      for (int j = lane; j < m; j += 32) row[j] = 0.0;
      __syncwarp();
      const int64_t k0 = rowptr[D[i]], k1 = rowptr[D[i] + 1];
      for (int64_t k = k0 + lane; k < k1; k += 32) {
        const int32_t c = col[k];
        int lo = 0, hi = m;                 // first position with D[pos] >= c
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (D[mid] < c) lo = mid + 1; else hi = mid;
        }
        if (lo < m && D[lo] == c) row[lo] = val[k];
      }
    }
  }
}

// In-place Gauss-Jordan with partial (row) pivoting: at step k the row with the largest entry of column k among the rows
// not used yet is swapped into place (the reference's exact block solve is an LU with pivoting, MLU_PRECOND; a
// velocity-pressure block can be regular and still meet an exactly zero Schur pivot); row k is scaled by 1/pivot (its k-th
// entry becomes 1/pivot, the image of the identity column) and every other row i gets
// M[i][j] = (j == k ? 0 : M[i][j]) - M[i][k] * rowk[j]; the row swaps are undone at the end as column swaps in reverse
// order.  Ties go to the lowest row, so the result does not depend on the number of threads.
// Shared memory: 2 * max_m doubles + max_m ints (dynamic) + the reduction scratch below.
constexpr int kInvertMaxThreads = 1024;
__global__ void __launch_bounds__(kInvertThreads) schwarz_invert_kernel(int64_t nblocks, const int64_t* __restrict__ blk_ptr,
                                                                         const int64_t* __restrict__ inv_ptr, double* __restrict__ inv,
                                                                         int max_m, int* __restrict__ err) {
  B2_DYN_SHARED(double, sh);
  __shared__ double red_v[kInvertMaxThreads];
  __shared__ int red_i[kInvertMaxThreads];
  double* rowk = sh;
  double* colk = sh + max_m;
  int* perm = reinterpret_cast<int*>(sh + 2 * max_m);
  for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const int m = (int)(blk_ptr[b + 1] - blk_ptr[b]);
    double* M = inv + inv_ptr[b];
    __syncthreads();                       // the extract launch finished; the buffers of the previous block are free
    for (int k = 0; k < m; k++) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) colk[i] = M[(int64_t)i * m + k];
      __syncthreads();
      // pivot row: largest |M[i][k]|, i >= k, lowest i on ties
      double bv = -1.0;
      int bi = k;
      for (int i = k + (int)threadIdx.x; i < m; i += blockDim.x) {
        const double a = fabs(colk[i]);
        if (a > bv) { bv = a; bi = i; }
      }
      red_v[threadIdx.x] = bv;
      red_i[threadIdx.x] = bi;
      __syncthreads();
      for (int off = 1; off < (int)blockDim.x; off <<= 1) {
        if ((threadIdx.x & (2 * off - 1)) == 0 && threadIdx.x + off < blockDim.x) {
          const double ov = red_v[threadIdx.x + off];
          const int oi = red_i[threadIdx.x + off];
          if (ov > red_v[threadIdx.x] || (ov == red_v[threadIdx.x] && oi < red_i[threadIdx.x])) {
            red_v[threadIdx.x] = ov;
            red_i[threadIdx.x] = oi;
          }
        }
        __syncthreads();
      }
      const int r = red_i[0];
      if (threadIdx.x == 0) perm[k] = r;
      if (r != k) {
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
          const double t = M[(int64_t)k * m + j];
          M[(int64_t)k * m + j] = M[(int64_t)r * m + j];
          M[(int64_t)r * m + j] = t;
        }
        if (threadIdx.x == 0) {
          const double t = colk[k];
          colk[k] = colk[r];
          colk[r] = t;
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < m; i += blockDim.x) rowk[i] = M[(int64_t)k * m + i];
      __syncthreads();
      const double piv = colk[k];
      if (threadIdx.x == 0 && !(fabs(piv) > 0.0)) atomicCAS(err, 0, (int)(b < 0x7ffffffe ? b + 1 : 0x7fffffff));
      const double p = 1.0 / piv;
      __syncthreads();                     // every thread has read the pivot before row k is rewritten
      for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const double v = (j == k ? 1.0 : rowk[j]) * p;
        rowk[j] = v;
        M[(int64_t)k * m + j] = v;
      }
      __syncthreads();
      const int64_t total = (int64_t)m * m;
      for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
        const int i = (int)(e / m), j = (int)(e - (int64_t)i * m);
        if (i == k) continue;
        const double old = (j == k) ? 0.0 : M[e];
        M[e] = fma(-colk[i], rowk[j], old);
      }
      __syncthreads();
    }
    for (int k = m - 1; k >= 0; k--) {     // (P A)^-1 P: the row swaps become column swaps, last one first
      const int r = perm[k];
      if (r != k)
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
          const double t = M[(int64_t)i * m + k];
          M[(int64_t)i * m + k] = M[(int64_t)i * m + r];
          M[(int64_t)i * m + r] = t;
        }
      __syncthreads();
    }
  }
}

// One CTA per block of the group.  y is read (columns of the block's rows) and written (the block's own dofs) through
// the same plain pointer: no other CTA of this launch touches those entries (the schedule's guarantee).
__global__ void __launch_bounds__(kApplyThreads) schwarz_apply_kernel(int64_t g0, int64_t g1, const int32_t* __restrict__ group_blocks,
                                                                       const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_dofs,
                                                                       const int64_t* __restrict__ inv_ptr, const double* __restrict__ inv,
                                                                       const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                       const double* __restrict__ val, const double* __restrict__ r,
                                                                       double* y, int max_m) {
  B2_DYN_SHARED(double, sh);
  double* t = sh;
  double* z = sh + max_m;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t q = g0 + blockIdx.x; q < g1; q += gridDim.x) {
    const int64_t b = group_blocks[q];
    const int32_t* D = blk_dofs + blk_ptr[b];
    const int m = (int)(blk_ptr[b + 1] - blk_ptr[b]);
    const double* M = inv + inv_ptr[b];
    __syncthreads();                       // t / z of the previous block are free
    for (int i = warp; i < m; i += nwarps) {
      const int64_t row = D[i];
      double acc = 0.0;
      for (int64_t k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) acc = fma(val[k], y[col[k]], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) t[i] = r[row] - acc;
    }
    __syncthreads();
    for (int i = warp; i < m; i += nwarps) {
      const double* Mi = M + (int64_t)i * m;
      double acc = 0.0;
      for (int j = lane; j < m; j += 32) acc = fma(Mi[j], t[j], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) z[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) y[D[i]] += z[i];
  }
}


// ---- block solves that walk the block's rows of A: one SSOR iteration, ILU(0) ------------------------------------
// Gauss-Seidel and the triangular solves are sequential in the block's (sorted) dofs.  Two ways to run them:
//   LEV = false  one warp walks the rows in order, lanes over a row's non-zeros; the parallelism is across the
//                blocks of a group (right for the small blocks of a few elements);
//   LEV = true   the rows of a block arrive sorted into DEPENDENCY LEVELS of its lower (forward sweeps, ILU
//                elimination) and upper (backward sweeps) triangular pattern: the rows of a level are independent, so
//                every warp of the CTA takes rows of the level and the CTA synchronises between levels -- same
//                arithmetic per row, hence the same result bit for bit (the reference's applications use blocks of
//                8^4 elements, or one block per level, where the one-warp walk is correct but serial).
// Per dof scratch in HBM (a block's dofs belong to no other block of its group): mark = the block that last claimed the
// dof (membership test of a column), zg = the block solution, tg / dg = right-hand side and diagonal (SSOR), foff = where
// the claiming block keeps the dof's factor row (ILU).
template <bool LEV, class Body>
__device__ __forceinline__ void schwarz_rows(bool reverse, int m, const int32_t* __restrict__ rows, const int32_t* __restrict__ off, int nlev,
                                             Body body) {
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (!LEV) {
    if (warp == 0)
      for (int ii = 0; ii < m; ii++) {
        body(reverse ? m - 1 - ii : ii);
        __syncwarp();
      }
    __syncthreads();
  } else {
    for (int l = 0; l < nlev; l++) {
      for (int t = off[l] + warp; t < off[l + 1]; t += nwarps) body(rows[t]);
      __syncthreads();
    }
  }
}

struct schwarz_ctx {          // what every row body needs
  int32_t b;
  const int32_t* D;
  const int64_t* rowptr;
  const int32_t* col;
  const double* val;
  int32_t* mark;
  double* zg;
};

// SSOR (PCSOR's default on the sub-block: local symmetric sweep, omega 1, zero initial guess -- what 001_Poisson's
// SetPreconditionerFineGrids(SOR_PRECOND) puts on the ASM blocks, LinearEquationSolverPetscAsm.cpp:300-317)
struct ssor_row {
  schwarz_ctx c;
  const double* tg;
  const double* dg;
  bool forward;
  // Entry k of the row belongs to lane (k - row start) % 32 and is added in ascending k: the arithmetic of the plain
  // strided loop.  Eight entries per lane are in flight at a time and nothing branches: column, membership mark and
  // block solution are three dependent loads per BATCH instead of three per entry (the sweep is a chain of such rows).
  __device__ void operator()(int i) const {
    const int lane = threadIdx.x & 31;
    const int64_t row = c.D[i];
    const int64_t k1 = c.rowptr[row + 1];
    double s = 0.0;
    for (int64_t kb = c.rowptr[row] + lane; kb < k1; kb += 32 * kRowBatch) {
      int32_t cc[kRowBatch];
      double v[kRowBatch], z[kRowBatch];
      bool use[kRowBatch];
#pragma unroll
      for (int j = 0; j < kRowBatch; j++) {
        const int64_t k = kb + 32 * j;
        const bool ok = k < k1;
        cc[j] = ok ? c.col[k] : (int32_t)row;
        v[j] = ok ? c.val[k] : 0.0;
        use[j] = forward ? cc[j] < row : cc[j] != row;
      }
#pragma unroll
      for (int j = 0; j < kRowBatch; j++) use[j] = use[j] && c.mark[use[j] ? cc[j] : (int32_t)row] == c.b;
#pragma unroll
      for (int j = 0; j < kRowBatch; j++) z[j] = use[j] ? c.zg[cc[j]] : 0.0;      // only members are read: no stray reads of other blocks' dofs
#pragma unroll
      for (int j = 0; j < kRowBatch; j++)
        if (use[j]) s = fma(v[j], z[j], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) c.zg[row] = (tg[row] - s) / dg[row];
  }
};

template <bool LEV>
__global__ void __launch_bounds__(kApplyThreads) schwarz_apply_ssor_kernel(int64_t g0, int64_t g1, const int32_t* __restrict__ group_blocks,
                                                                            const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_dofs,
                                                                            const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                            const double* __restrict__ val, const double* __restrict__ r, double* y,
                                                                            double* tg, double* dg, double* zg, int32_t* mark,
                                                                            const int64_t* __restrict__ lvptr_f, const int32_t* __restrict__ lvoff_f,
                                                                            const int32_t* __restrict__ lvrows_f, const int64_t* __restrict__ lvptr_b,
                                                                            const int32_t* __restrict__ lvoff_b, const int32_t* __restrict__ lvrows_b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t q = g0 + blockIdx.x; q < g1; q += gridDim.x) {
    const int32_t b = group_blocks[q];
    const int32_t* D = blk_dofs + blk_ptr[b];
    const int m = (int)(blk_ptr[b + 1] - blk_ptr[b]);
    for (int i = warp; i < m; i += nwarps) {
      const int64_t row = D[i];
      double acc = 0.0, diag = 0.0;
      for (int64_t k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) {
        const int32_t c = col[k];
        acc = fma(val[k], y[c], acc);
        if (c == row) diag = val[k];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        diag += __shfl_xor_sync(0xffffffffu, diag, o);
      }
      if (lane == 0) {
        tg[row] = r[row] - acc;
        dg[row] = diag;
        zg[row] = 0.0;
        mark[row] = b;
      }
    }
    __syncthreads();
    const schwarz_ctx c{b, D, rowptr, col, val, mark, zg};
    schwarz_rows<LEV>(false, m, LEV ? lvrows_f + blk_ptr[b] : nullptr, LEV ? lvoff_f + lvptr_f[b] : nullptr,
                      LEV ? (int)(lvptr_f[b + 1] - lvptr_f[b]) - 1 : 0, ssor_row{c, tg, dg, true});      // z = (D + L)^-1 t
    schwarz_rows<LEV>(true, m, LEV ? lvrows_b + blk_ptr[b] : nullptr, LEV ? lvoff_b + lvptr_b[b] : nullptr,
                      LEV ? (int)(lvptr_b[b + 1] - lvptr_b[b]) - 1 : 0, ssor_row{c, tg, dg, false});     // backward sweep from that iterate
    for (int i = threadIdx.x; i < m; i += blockDim.x) y[D[i]] += zg[D[i]];
    __syncthreads();
  }
}

// ---- ILU(0) block solves: ILU_PRECOND on the blocks, by far the most common fine-grid preconditioner of the
// reference's applications (PCILU: levels 0, natural ordering = the block's sorted dofs, PetscPreconditioner.cpp;
// LinearEquationSolverPetscAsm.cpp:300-317).  The factor of block b lives on the pattern of the block's rows of A:
// row i of the block (global row r = D[i]) owns len(r) slots at fac[frow[blk_ptr[b] + i] ..), slot q belonging to
// column col[rowptr[r] + q]; slots of columns outside the block are unused.  IKJ elimination of a row by one warp: for
// every earlier column k of the row, l_ik = a_ik / u_kk, then a_ij -= l_ik u_kj wherever (k, j) is in the pattern of
// row k (found by bisection); lanes work across the row's entries.
__device__ __forceinline__ int64_t schwarz_bsearch(const int32_t* __restrict__ col, int64_t lo, int64_t hi, int32_t c) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct ilu_factor_row {
  schwarz_ctx c;
  const int64_t* F;
  double* fac;
  const int64_t* foff;
  int* err;
  __device__ void operator()(int i) const {
    const int lane = threadIdx.x & 31;
    const int32_t r = c.D[i];
    const int64_t rp = c.rowptr[r], len = c.rowptr[r + 1] - rp, fi = F[i];
    for (int64_t q = 0; q < len; q++) {             // the row's entries in column order: earlier columns are pivots
      const int32_t k = c.col[rp + q];
      if (k >= r) break;
      if (c.mark[k] != c.b) continue;
      const int64_t kp = c.rowptr[k], klen = c.rowptr[k + 1] - kp, fk = foff[k];
      const int64_t dk = schwarz_bsearch(c.col, kp, kp + klen, k) - kp;          // the pivot row's diagonal slot
      const double lik = fac[fi + q] / fac[fk + dk];
      __syncwarp();
      if (lane == 0) fac[fi + q] = lik;
      for (int64_t q2 = q + 1 + lane; q2 < len; q2 += 32) {
        const int32_t c2 = c.col[rp + q2];
        if (c.mark[c2] != c.b) continue;
        const int64_t p = schwarz_bsearch(c.col, kp, kp + klen, c2);
        if (p < kp + klen && c.col[p] == c2) fac[fi + q2] = fma(-lik, fac[fk + (p - kp)], fac[fi + q2]);
      }
      __syncwarp();
    }
    const int64_t di = schwarz_bsearch(c.col, rp, rp + len, r) - rp;
    if (lane == 0 && !(fabs(fac[fi + di]) > 0.0)) atomicCAS(err, 0, c.b + 1);
  }
};

template <bool LEV>
__global__ void __launch_bounds__(kApplyThreads) schwarz_ilu_factor_kernel(int64_t g0, int64_t g1, const int32_t* __restrict__ group_blocks,
                                                                            const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_dofs,
                                                                            const int64_t* __restrict__ frow, const int64_t* __restrict__ rowptr,
                                                                            const int32_t* __restrict__ col, const double* __restrict__ val, double* fac,
                                                                            int32_t* mark, int64_t* foff, int* err, const int64_t* __restrict__ lvptr_f,
                                                                            const int32_t* __restrict__ lvoff_f, const int32_t* __restrict__ lvrows_f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t q0 = g0 + blockIdx.x; q0 < g1; q0 += gridDim.x) {
    const int32_t b = group_blocks[q0];
    const int32_t* D = blk_dofs + blk_ptr[b];
    const int64_t* F = frow + blk_ptr[b];
    const int m = (int)(blk_ptr[b + 1] - blk_ptr[b]);
    for (int i = warp; i < m; i += nwarps) {            // copy the rows, claim the dofs
      const int64_t r = D[i], rp = rowptr[r], len = rowptr[r + 1] - rp;
      for (int64_t q = lane; q < len; q += 32) fac[F[i] + q] = val[rp + q];
      if (lane == 0) { mark[r] = b; foff[r] = F[i]; }
    }
    __syncthreads();
    const schwarz_ctx c{b, D, rowptr, col, val, mark, nullptr};
    schwarz_rows<LEV>(false, m, LEV ? lvrows_f + blk_ptr[b] : nullptr, LEV ? lvoff_f + lvptr_f[b] : nullptr,
                      LEV ? (int)(lvptr_f[b + 1] - lvptr_f[b]) - 1 : 0, ilu_factor_row{c, F, fac, foff, err});
  }
}

struct ilu_solve_row {
  schwarz_ctx c;
  const int64_t* F;
  const double* fac;
  bool forward;
  // same batching as ssor_row: entry q of the row belongs to lane q % 32, added in ascending q, nothing branches
  __device__ void operator()(int i) const {
    const int lane = threadIdx.x & 31;
    const int32_t row = c.D[i];
    const int64_t rp = c.rowptr[row], len = c.rowptr[row + 1] - rp;
    const double* f = fac + F[i];
    double s = 0.0, d = 0.0;
    for (int64_t qb = lane; qb < len; qb += 32 * kRowBatch) {
      int32_t cc[kRowBatch];
      double v[kRowBatch], z[kRowBatch];
      bool use[kRowBatch];
#pragma unroll
      for (int j = 0; j < kRowBatch; j++) {
        const int64_t q = qb + 32 * j;
        const bool ok = q < len;
        cc[j] = ok ? c.col[rp + q] : row;
        v[j] = ok ? f[q] : 0.0;
        use[j] = forward ? cc[j] < row : cc[j] > row;
        if (!forward && ok && cc[j] == row) d = v[j];
      }
#pragma unroll
      for (int j = 0; j < kRowBatch; j++) use[j] = use[j] && c.mark[use[j] ? cc[j] : row] == c.b;
#pragma unroll
      for (int j = 0; j < kRowBatch; j++) z[j] = use[j] ? c.zg[cc[j]] : 0.0;
#pragma unroll
      for (int j = 0; j < kRowBatch; j++)
        if (use[j]) s = fma(v[j], z[j], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if (lane == 0) c.zg[row] = forward ? c.zg[row] - s : (c.zg[row] - s) / d;      // L z = t (unit diagonal), then U z = z
  }
};

template <bool LEV>
__global__ void __launch_bounds__(kApplyThreads) schwarz_apply_ilu_kernel(int64_t g0, int64_t g1, const int32_t* __restrict__ group_blocks,
                                                                           const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_dofs,
                                                                           const int64_t* __restrict__ frow, const int64_t* __restrict__ rowptr,
                                                                           const int32_t* __restrict__ col, const double* __restrict__ val,
                                                                           const double* __restrict__ fac, const double* __restrict__ r, double* y,
                                                                           double* zg, int32_t* mark, const int64_t* __restrict__ lvptr_f,
                                                                           const int32_t* __restrict__ lvoff_f, const int32_t* __restrict__ lvrows_f,
                                                                           const int64_t* __restrict__ lvptr_b, const int32_t* __restrict__ lvoff_b,
                                                                           const int32_t* __restrict__ lvrows_b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t q0 = g0 + blockIdx.x; q0 < g1; q0 += gridDim.x) {
    const int32_t b = group_blocks[q0];
    const int32_t* D = blk_dofs + blk_ptr[b];
    const int64_t* F = frow + blk_ptr[b];
    const int m = (int)(blk_ptr[b + 1] - blk_ptr[b]);
    for (int i = warp; i < m; i += nwarps) {            // t = (r - A y)[B] into zg, claim the dofs
      const int64_t row = D[i];
      double acc = 0.0;
      for (int64_t k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) acc = fma(val[k], y[col[k]], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) { zg[row] = r[row] - acc; mark[row] = b; }
    }
    __syncthreads();
    const schwarz_ctx c{b, D, rowptr, col, val, mark, zg};
    schwarz_rows<LEV>(false, m, LEV ? lvrows_f + blk_ptr[b] : nullptr, LEV ? lvoff_f + lvptr_f[b] : nullptr,
                      LEV ? (int)(lvptr_f[b + 1] - lvptr_f[b]) - 1 : 0, ilu_solve_row{c, F, fac, true});
    schwarz_rows<LEV>(true, m, LEV ? lvrows_b + blk_ptr[b] : nullptr, LEV ? lvoff_b + lvptr_b[b] : nullptr,
                      LEV ? (int)(lvptr_b[b + 1] - lvptr_b[b]) - 1 : 0, ilu_solve_row{c, F, fac, false});
    for (int i = threadIdx.x; i < m; i += blockDim.x) y[D[i]] += zg[D[i]];
    __syncthreads();
  }
}
