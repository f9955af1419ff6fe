// Device code of the element-block smoother (b2_schwarz.cu), kept free of host / runtime calls so that the SAME
// source also compiles for the CPU thread emulator of tests/cpp/cuda_emu.hpp (tests/test_kernel_emulation.py runs
// these kernels on host threads against the oracle when no GPU is present).  Included inside an anonymous namespace.
#pragma once
#ifndef B2_DYN_SHARED
#define B2_DYN_SHARED(type, name) extern __shared__ type name[]
#endif

constexpr int kApplyThreads = 256;
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

// In-place Gauss-Jordan: for every pivot k, row k is scaled by 1/pivot (its k-th entry becomes 1/pivot, the image of
// the identity column) and every other row i gets  M[i][j] = (j == k ? 0 : M[i][j]) - M[i][k] * rowk[j].
__global__ void __launch_bounds__(kInvertThreads) schwarz_invert_kernel(int64_t nblocks, const int64_t* __restrict__ blk_ptr,
                                                                         const int64_t* __restrict__ inv_ptr, double* __restrict__ inv,
                                                                         int max_m, int* __restrict__ err) {
  B2_DYN_SHARED(double, sh);
  double* rowk = sh;
  double* colk = sh + max_m;
  for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const int m = (int)(blk_ptr[b + 1] - blk_ptr[b]);
    double* M = inv + inv_ptr[b];
    __syncthreads();                       // the extract launch finished; rowk / colk of the previous block are free
    for (int k = 0; k < m; k++) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        colk[i] = M[(int64_t)i * m + k];
        rowk[i] = M[(int64_t)k * m + i];
      }
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


// The same sweep with ONE SSOR ITERATION as the block solve (PCSOR's default on the sub-block: local symmetric sweep,
// omega 1, zero initial guess -- what 001_Poisson's SetPreconditionerFineGrids(SOR_PRECOND) puts on the ASM blocks,
// LinearEquationSolverPetscAsm.cpp:300-317, PetscPreconditioner.cpp SOR_PRECOND).  No factor storage: the block's
// rows of A are used as they are.  Gauss-Seidel is sequential in the block's (sorted) dofs, so one warp walks the rows
// -- lanes over a row's non-zeros -- while the other warps only help with t = (r - A y)[B]; the parallelism is across
// the blocks of a group.  Scratch per dof in HBM (a block's dofs belong to no other block of its group): tg = t,
// dg = diagonal, zg = the block solution, mark = the block that last claimed the dof (membership test of a column).
__global__ void __launch_bounds__(kApplyThreads) schwarz_apply_ssor_kernel(int64_t g0, int64_t g1, const int32_t* __restrict__ group_blocks,
                                                                            const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_dofs,
                                                                            const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                            const double* __restrict__ val, const double* __restrict__ r, double* y,
                                                                            double* tg, double* dg, double* zg, int32_t* mark) {
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
    if (warp == 0) {
      for (int i = 0; i < m; i++) {                 // forward sweep: z = (D + L)^-1 t
        const int64_t row = D[i];
        double s = 0.0;
        for (int64_t k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) {
          const int32_t c = col[k];
          if (c < row && mark[c] == b) s = fma(val[k], zg[c], s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) zg[row] = (tg[row] - s) / dg[row];
        __syncwarp();
      }
      for (int i = m - 1; i >= 0; i--) {            // backward sweep from that iterate
        const int64_t row = D[i];
        double s = 0.0;
        for (int64_t k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) {
          const int32_t c = col[k];
          if (c != row && mark[c] == b) s = fma(val[k], zg[c], s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) zg[row] = (tg[row] - s) / dg[row];
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) y[D[i]] += zg[D[i]];
    __syncthreads();
  }
}

// ---- ILU(0) block solves: ILU_PRECOND on the blocks, by far the most common fine-grid preconditioner of the
// reference's applications (PCILU: levels 0, natural ordering = the block's sorted dofs, PetscPreconditioner.cpp;
// LinearEquationSolverPetscAsm.cpp:300-317).  The factor of block b lives on the pattern of the block's rows of A:
// row i of the block (global row r = D[i]) owns len(r) slots at fac[frow[blk_ptr[b] + i] ..), slot q belonging to
// column col[rowptr[r] + q]; slots of columns outside the block are unused.  One warp per block walks the rows in order
// (IKJ elimination: for every earlier column k of the row, l_ik = a_ik / u_kk, then a_ij -= l_ik u_kj wherever (k, j) is
// in the pattern of row k); lanes work across a row's entries.  Per dof scratch as in the SSOR kernel: mark = the
// block that claimed the dof, foff = where that block keeps the dof's factor row.
__device__ __forceinline__ int64_t schwarz_bsearch(const int32_t* __restrict__ col, int64_t lo, int64_t hi, int32_t c) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kApplyThreads) schwarz_ilu_factor_kernel(int64_t g0, int64_t g1, const int32_t* __restrict__ group_blocks,
                                                                            const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_dofs,
                                                                            const int64_t* __restrict__ frow, const int64_t* __restrict__ rowptr,
                                                                            const int32_t* __restrict__ col, const double* __restrict__ val, double* fac,
                                                                            int32_t* mark, int64_t* foff, int* __restrict__ err) {
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
    if (warp == 0) {
      for (int i = 0; i < m; i++) {
        const int32_t r = D[i];
        const int64_t rp = rowptr[r], len = rowptr[r + 1] - rp, fi = F[i];
        for (int64_t q = 0; q < len; q++) {             // the row's entries in column order: earlier columns are pivots
          const int32_t k = col[rp + q];
          if (k >= r) break;
          if (mark[k] != b) continue;
          const int64_t kp = rowptr[k], klen = rowptr[k + 1] - kp, fk = foff[k];
          const int64_t dk = schwarz_bsearch(col, kp, kp + klen, k) - kp;          // the pivot row's diagonal slot
          const double lik = fac[fi + q] / fac[fk + dk];
          __syncwarp();
          if (lane == 0) fac[fi + q] = lik;
          for (int64_t q2 = q + 1 + lane; q2 < len; q2 += 32) {
            const int32_t c2 = col[rp + q2];
            if (mark[c2] != b) continue;
            const int64_t p = schwarz_bsearch(col, kp, kp + klen, c2);
            if (p < kp + klen && col[p] == c2) fac[fi + q2] = fma(-lik, fac[fk + (p - kp)], fac[fi + q2]);
          }
          __syncwarp();
        }
        const int64_t di = schwarz_bsearch(col, rp, rp + len, r) - rp;
        if (lane == 0 && !(fabs(fac[fi + di]) > 0.0)) atomicCAS(err, 0, b + 1);
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kApplyThreads) schwarz_apply_ilu_kernel(int64_t g0, int64_t g1, const int32_t* __restrict__ group_blocks,
                                                                           const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_dofs,
                                                                           const int64_t* __restrict__ frow, const int64_t* __restrict__ rowptr,
                                                                           const int32_t* __restrict__ col, const double* __restrict__ val,
                                                                           const double* __restrict__ fac, const double* __restrict__ r, double* y,
                                                                           double* zg, int32_t* mark) {
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
    if (warp == 0) {
      for (int i = 0; i < m; i++) {                     // L z = t, unit lower triangle
        const int32_t row = D[i];
        const int64_t rp = rowptr[row], len = rowptr[row + 1] - rp;
        double s = 0.0;
        for (int64_t q = lane; q < len; q += 32) {
          const int32_t c = col[rp + q];
          if (c < row && mark[c] == b) s = fma(fac[F[i] + q], zg[c], s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) zg[row] -= s;
        __syncwarp();
      }
      for (int i = m - 1; i >= 0; i--) {                // U z = z
        const int32_t row = D[i];
        const int64_t rp = rowptr[row], len = rowptr[row + 1] - rp;
        double s = 0.0, d = 0.0;
        for (int64_t q = lane; q < len; q += 32) {
          const int32_t c = col[rp + q];
          if (c == row) d = fac[F[i] + q];
          else if (c > row && mark[c] == b) s = fma(fac[F[i] + q], zg[c], s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o);
          d += __shfl_xor_sync(0xffffffffu, d, o);
        }
        if (lane == 0) zg[row] = (zg[row] - s) / d;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) y[D[i]] += zg[D[i]];
    __syncthreads();
  }
}
