// Row-walking block solves of the element-block smoother WITHOUT row levels, staged: the variant for the many small
// blocks of a coloured sweep (b2_schwarz.cu; included after b2_schwarz_kernels.cuh inside the same anonymous
// namespace, free of host / runtime calls so that the same source compiles for the CPU thread emulator).
//
// The plain walk (ssor_row / ilu_solve_row with LEV = false) is a chain of rows, each row a chain of three dependent
// global loads (column -> membership mark -> block solution) before its shuffle reduction: latency, not bandwidth.
// Here the only dependence between two rows goes through SHARED memory:
//   * once per pattern, every (block, row) gets its row of LOCAL indices lidx[frow + q] = position of the column of
//     entry q in the block's sorted dof list, 0xffff outside the block (schwarz_lidx_kernel): membership and position in
//     one 16-bit load that does not depend on the sweep;
//   * the block's right-hand side, diagonal and solution live in the CTA's shared memory;
//   * the first batch of entries (local index, value) of the NEXT row is requested before the current row is reduced,
//     so the global latency of a row hides behind the row before it.
// The arithmetic of a row is that of the plain walk -- entry q belongs to lane q % 32, members are added in ascending
// q, butterfly reduction, (t - s) / d by lane 0 -- so both walks and the level-scheduled rows agree bit for bit.
#pragma once

constexpr int kWalkMaxM = 4096;          // largest block of the staged walk (5 m + 2 doubles of shared memory)
constexpr unsigned short kNotInBlock = 0xffff;

__host__ __device__ __forceinline__ size_t walk_smem_bytes(int max_m) { return (size_t)(5 * max_m + 2) * sizeof(double); }

// lidx of every (block, row): one warp per block row
__global__ void schwarz_lidx_kernel(int64_t nblocks, const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_dofs,
                                    const int64_t* __restrict__ frow, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                    unsigned short* __restrict__ lidx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const int32_t* D = blk_dofs + blk_ptr[b];
    const int m = (int)(blk_ptr[b + 1] - blk_ptr[b]);
    for (int i = warp; i < m; i += nwarps) {
      const int64_t rp = rowptr[D[i]], len = rowptr[D[i] + 1] - rp, f = frow[blk_ptr[b] + i];
      for (int64_t q = lane; q < len; q += 32) {
        const int32_t c = col[rp + q];
        int lo = 0, hi = m;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (D[mid] < c) lo = mid + 1; else hi = mid;
        }
        lidx[f + q] = (lo < m && D[lo] == c) ? (unsigned short)lo : kNotInBlock;
      }
    }
  }
}

// a batch of a row: entry lane + 32 j, j < NB.  NB is a template parameter of the walk: the smallest of 1, 2, 4, 8 whose
// batch holds the longest row of the operator (27 entries for trilinear, 125 for triquadratic hexahedra), so that the usual
// row is ONE batch without empty slots -- the slots are the instructions of a row visit
template <int NB>
struct walk_batch {
  int l[NB];
  double v[NB];
};
template <int NB>
__device__ __forceinline__ void walk_load(walk_batch<NB>& B, const unsigned short* __restrict__ lx, const double* __restrict__ v, int len, int lane) {
#pragma unroll
  for (int j = 0; j < NB; j++) {
    const int q = lane + 32 * j;
    const bool ok = q < len;
    B.l[j] = ok ? (int)lx[q] : (int)kNotInBlock;
    B.v[j] = ok ? v[q] : 0.0;
  }
}

// One sweep over the rows of a block by ONE warp.  KIND 0: SSOR forward (members l < i), 1: SSOR backward (l != i),
// 2: ILU forward (l < i, unit diagonal), 3: ILU backward (l > i, diagonal from the row).  lx / vals: the block's rows
// of local indices / values, row i at F[i] - F[0] resp. at vrow(i); z, t, d in shared memory.
template <int KIND, int NB, class VRow>
__device__ __forceinline__ void walk_sweep(int m, const int64_t* F, const unsigned short* __restrict__ lidx, VRow vrow, double* z,
                                           const double* t, const double* d, int lane) {
  constexpr bool reverse = KIND == 1 || KIND == 3;
  walk_batch<NB> cur, nxt;
  {
    const int i0 = reverse ? m - 1 : 0;
    walk_load(cur, lidx + F[i0], vrow(i0), (int)(F[i0 + 1] - F[i0]), lane);
    nxt = cur;
  }
  for (int ii = 0; ii < m; ii++) {
    const int i = reverse ? m - 1 - ii : ii;
    const int len = (int)(F[i + 1] - F[i]);
    if (ii + 1 < m) {                      // the next row's first batch: independent of this row's result
      const int in = reverse ? i - 1 : i + 1;
      walk_load(nxt, lidx + F[in], vrow(in), (int)(F[in + 1] - F[in]), lane);
    }
    double s = 0.0, dd = 0.0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const int l = cur.l[j];
      const bool use = l != (int)kNotInBlock && (KIND == 0 || KIND == 2 ? l < i : (KIND == 1 ? l != i : l > i));
      if (KIND == 3 && l == i) dd = cur.v[j];
      if (use) s = fma(cur.v[j], z[l], s);
    }
    if (len > 32 * NB) {                   // rows longer than one batch (only when the longest row exceeds 256 entries)
      const unsigned short* lx = lidx + F[i];
      const double* v = vrow(i);
      for (int qb = 32 * NB; qb < len; qb += 32 * NB) {
        walk_batch<NB> more;
        walk_load(more, lx + qb, v + qb, len - qb, lane);
#pragma unroll
        for (int j = 0; j < NB; j++) {
          const int l = more.l[j];
          const bool use = l != (int)kNotInBlock && (KIND == 0 || KIND == 2 ? l < i : (KIND == 1 ? l != i : l > i));
          if (KIND == 3 && l == i) dd = more.v[j];
          if (use) s = fma(more.v[j], z[l], s);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      if (KIND == 3) dd += __shfl_xor_sync(0xffffffffu, dd, o);
    }
    if (lane == 0) {
      if (KIND <= 1) z[i] = (t[i] - s) / d[i];
      else if (KIND == 2) z[i] = z[i] - s;
      else z[i] = (z[i] - s) / dd;
    }
    __syncwarp();
    cur = nxt;
  }
}

// shared memory of a CTA: z[m], t[m], d[m], F[m + 1] (int64), rs[m] (int64: first entry of the row in A)
struct walk_smem {
  double *z, *t, *d;
  int64_t *F, *rs;
  __device__ walk_smem(double* base, int max_m) : z(base), t(base + max_m), d(base + 2 * max_m),
                                                  F(reinterpret_cast<int64_t*>(base + 3 * max_m)),
                                                  rs(reinterpret_cast<int64_t*>(base + 4 * max_m + 2)) {}
};

// y[B] += (one SSOR iteration | the ILU(0) solve) of (r - A y)[B], one CTA per block of the group
template <bool ILU, int NB>
__global__ void __launch_bounds__(kApplyThreads) schwarz_walk_apply_kernel(int64_t g0, int64_t g1, const int32_t* __restrict__ group_blocks,
                                                                            const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_dofs,
                                                                            const int64_t* __restrict__ frow, const unsigned short* __restrict__ lidx,
                                                                            const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                            const double* __restrict__ val, const double* __restrict__ fac,
                                                                            const double* __restrict__ r, double* y, int max_m) {
  B2_DYN_SHARED(double, sh);
  walk_smem S(sh, max_m);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t q = g0 + blockIdx.x; q < g1; q += gridDim.x) {
    const int32_t b = group_blocks[q];
    const int32_t* D = blk_dofs + blk_ptr[b];
    const int64_t* Fg = frow + blk_ptr[b];
    const int m = (int)(blk_ptr[b + 1] - blk_ptr[b]);
    __syncthreads();                       // the previous block's shared arrays are free
    for (int i = threadIdx.x; i <= m; i += blockDim.x) S.F[i] = Fg[i];
    for (int i = warp; i < m; i += nwarps) {            // t = (r - A y)[B], the diagonal
      const int64_t row = D[i], k0 = rowptr[row];
      double acc = 0.0, diag = 0.0;
      for (int64_t k = k0 + lane; k < rowptr[row + 1]; k += 32) {
        const int32_t c = col[k];
        acc = fma(val[k], y[c], acc);
        if (!ILU && c == row) diag = val[k];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (!ILU) diag += __shfl_xor_sync(0xffffffffu, diag, o);
      }
      if (lane == 0) {
        S.rs[i] = k0;
        if (ILU) S.z[i] = r[row] - acc;
        else { S.t[i] = r[row] - acc; S.d[i] = diag; S.z[i] = 0.0; }
      }
    }
    __syncthreads();
    if (warp == 0) {
      if (ILU) {
        auto vrow = [&](int i) { return fac + S.F[i]; };
        walk_sweep<2, NB>(m, S.F, lidx, vrow, S.z, S.t, S.d, lane);      // L z = t (unit diagonal)
        walk_sweep<3, NB>(m, S.F, lidx, vrow, S.z, S.t, S.d, lane);      // U z = z
      } else {
        auto vrow = [&](int i) { return val + S.rs[i]; };
        walk_sweep<0, NB>(m, S.F, lidx, vrow, S.z, S.t, S.d, lane);      // z = (D + L)^-1 t
        walk_sweep<1, NB>(m, S.F, lidx, vrow, S.z, S.t, S.d, lane);      // backward sweep from that iterate
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) y[D[i]] += S.z[i];
  }
}

// the instantiation for an operator whose longest row has max_row entries
typedef void (*walk_apply_kernel_t)(int64_t, int64_t, const int32_t*, const int64_t*, const int32_t*, const int64_t*, const unsigned short*,
                                    const int64_t*, const int32_t*, const double*, const double*, const double*, double*, int);
template <bool ILU>
inline walk_apply_kernel_t schwarz_walk_apply_kernel_for(int max_row) {
  if (max_row <= 32) return schwarz_walk_apply_kernel<ILU, 1>;
  if (max_row <= 64) return schwarz_walk_apply_kernel<ILU, 2>;
  if (max_row <= 128) return schwarz_walk_apply_kernel<ILU, 4>;
  return schwarz_walk_apply_kernel<ILU, 8>;
}

// ILU(0) of every block of the group on the pattern of its rows of A, IKJ order by one warp: row i's slots are
// published in a shared map local index -> slot, every pivot row l < i of the row is scanned ONCE (lanes over its
// entries beyond its diagonal) and subtracted wherever row i has the column -- the updates of ilu_factor_row, each
// found by one shared-memory lookup instead of a bisection in global memory.
__global__ void __launch_bounds__(kApplyThreads) schwarz_walk_ilu_factor_kernel(int64_t g0, int64_t g1, const int32_t* __restrict__ group_blocks,
                                                                                 const int64_t* __restrict__ blk_ptr, const int32_t* __restrict__ blk_dofs,
                                                                                 const int64_t* __restrict__ frow, const unsigned short* __restrict__ lidx,
                                                                                 const int64_t* __restrict__ rowptr, const double* __restrict__ val,
                                                                                 double* fac, int* err, int max_m) {
  B2_DYN_SHARED(double, sh);
  walk_smem S(sh, max_m);
  unsigned short* map = reinterpret_cast<unsigned short*>(S.z);       // [m] slot of local column l in the current row
  unsigned short* dpos = reinterpret_cast<unsigned short*>(S.t);      // [m] diagonal slot of every finished row
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t q0 = g0 + blockIdx.x; q0 < g1; q0 += gridDim.x) {
    const int32_t b = group_blocks[q0];
    const int32_t* D = blk_dofs + blk_ptr[b];
    const int64_t* Fg = frow + blk_ptr[b];
    const int m = (int)(blk_ptr[b + 1] - blk_ptr[b]);
    __syncthreads();
    for (int i = threadIdx.x; i <= m; i += blockDim.x) S.F[i] = Fg[i];
    for (int i = threadIdx.x; i < m; i += blockDim.x) map[i] = kNotInBlock;
    for (int i = warp; i < m; i += nwarps) {            // copy the rows
      const int64_t rp = rowptr[D[i]], len = rowptr[D[i] + 1] - rp, f = Fg[i];
      for (int64_t q = lane; q < len; q += 32) fac[f + q] = val[rp + q];
    }
    __syncthreads();
    if (warp == 0) {
      for (int i = 0; i < m; i++) {
        const int64_t fi = S.F[i];
        const int len = (int)(S.F[i + 1] - fi);
        const unsigned short* lx = lidx + fi;
        for (int q = lane; q < len; q += 32) {
          const int l = lx[q];
          if (l != (int)kNotInBlock) map[l] = (unsigned short)q;
        }
        __syncwarp();
        for (int q = 0; q < len; q++) {              // the row's entries in column order: earlier in-block columns are pivots
          const int l = lx[q];
          if (l == (int)kNotInBlock) continue;
          if (l >= i) break;
          const int64_t fl = S.F[l];
          const int llen = (int)(S.F[l + 1] - fl);
          const double lik = fac[fi + q] / fac[fl + dpos[l]];
          __syncwarp();
          if (lane == 0) fac[fi + q] = lik;
          const unsigned short* lxl = lidx + fl;
          for (int q2 = (int)dpos[l] + 1 + lane; q2 < llen; q2 += 32) {      // the pivot row beyond its diagonal
            const int l2 = lxl[q2];
            if (l2 == (int)kNotInBlock) continue;
            const int p = map[l2];
            if (p != (int)kNotInBlock) fac[fi + p] = fma(-lik, fac[fl + q2], fac[fi + p]);
          }
          __syncwarp();
        }
        const int di = map[i];
        if (lane == 0) {
          dpos[i] = (unsigned short)di;
          if (di == (int)kNotInBlock || !(fabs(fac[fi + di]) > 0.0)) atomicCAS(err, 0, b + 1);
        }
        __syncwarp();
        for (int q = lane; q < len; q += 32) {
          const int l = lx[q];
          if (l != (int)kNotInBlock) map[l] = kNotInBlock;
        }
        __syncwarp();
      }
    }
  }
}
