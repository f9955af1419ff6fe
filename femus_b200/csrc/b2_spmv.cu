// fp64 CSR SpMV family with TMA-staged row blocks and a per-block column dictionary (sm_100a).
// Replaces the MatMult / MatMultAdd calls behind PetscVector::{matrix_mult, resid} (reference
// src/03_algebra/00_vectors/PetscVector.cpp:193-247) and the KSPRICHARDSON+PCJACOBI sweep PETSc
// runs on every level (LinearEquationSolverPetsc.cpp:516-519, PetscPreconditioner.cpp:209-212).
//
// PLAN (b2_csr_build_chunks, once per sparsity pattern -- the role of MatAssemblyEnd's
// MatSetUpMultiply in PETSc).  The matrix is cut into CHUNKS of whole consecutive rows: chunk c
// holds the rows whose key  rowptr[r] + kRowWeight * r  lies in [c*T, (c+1)*T), so a chunk never
// has more than kCap nonzeros nor more than kMaxRows rows.  Neighbouring rows of an FE matrix share
// most of their columns, so for every chunk the plan stores the sorted list of its DISTINCT columns
// (the dictionary, ~nnz/4 entries for Q2 hexahedra) and, for every nonzero, the 16-bit position of
// its column in that list.  The CSR arrays themselves are untouched (assembly, Galerkin products
// and the host see plain CSR); the plan costs 2 B per nonzero + 4 B per dictionary entry.
//
// KERNEL.  A producer warp streams a chunk's values, local indices, dictionary, row pointers and
// the epilogue operands of its rows (all contiguous byte ranges) into a shared-memory ring with
// 1-D bulk async copies (cp.async.bulk, the TMA engine) that complete on an mbarrier; the matrix
// stream carries an L2 evict-first policy so that it does not push x out of L2.  Eight consumer
// warps wait on the stage's "full" barrier, then
//   A. gather x once per DISTINCT column of the chunk into shared memory (sorted addresses, a few
//      cache lines per warp load instead of one per lane),
//   B. give every row to a sub-warp of TPR lanes (TPR from the chunk's mean row length) that
//      multiplies values by the staged x entirely out of shared memory, reduces with shuffles and
//      applies the epilogue selected at compile time:
//   Y_AX    y = A x                                 (MatMult)
//   Y_ADD   y += A x                                (MatMultAdd)
//   RESID   y = b - A x                             (resid)
//   JACOBI  y = x + omega * dinv .* (b - A x)       (one Richardson+Jacobi sweep, x != y)
//   RESID_W y = w .* b - A x                        (distributed residual, w = 1/multiplicity)
// Consumers synchronise among themselves with a named barrier (the producer warp never joins it)
// and hand the stage back through an "empty" mbarrier.
// HBM traffic: 8 B value + 2 B index per nonzero + 4 B per dictionary entry + rowptr/y per row;
// the roofline in bench.py is still quoted on the plain-CSR algorithmic bytes (12 B per nonzero).
// Matrices whose longest row does not fit a chunk use the register-streaming kernel at the bottom.
#include "b2_common.cuh"
#include <cub/cub.cuh>

namespace {

enum SpmvMode { Y_AX, Y_ADD, RESID, JACOBI, RESID_W };

// ---- chunk geometry --------------------------------------------------------------------------
constexpr int kCap = 2048;                    // nonzeros per chunk (dictionary entries likewise)
constexpr int kRowWeight = 8;                 // a row counts as this many nonzeros when cutting chunks
constexpr int kMaxRows = kCap / kRowWeight;   // 256
constexpr int kValPad = kCap + 16;            // staged entries incl. the 16-byte alignment of both ends
constexpr int kRpPad = kMaxRows + 4;

// Shared-memory layout, computed on the host per (matrix, epilogue): the dictionary area is sized
// by the matrix's largest chunk dictionary, the epilogue area by the number of staged row operands.
struct SpmvLayout {
  int off_idx, off_dict, off_rp, off_epi, stage_bytes;   // inside a stage
  int off_bar, off_desc, off_xs, xs_doubles, total;      // after the ring
};
static SpmvLayout make_layout(int stages, int dict_cap, int nepi, int xs_buffers) {
  SpmvLayout L;
  L.off_idx = kValPad * 8;
  L.off_dict = L.off_idx + kValPad * 2;
  L.off_rp = L.off_dict + dict_cap * 4;
  L.off_epi = L.off_rp + kRpPad * 8;
  L.stage_bytes = L.off_epi + nepi * kRpPad * 8;
  L.off_bar = stages * L.stage_bytes;
  L.off_desc = L.off_bar + stages * 16;
  L.off_xs = (L.off_desc + stages * 32 + 15) / 16 * 16;
  L.xs_doubles = dict_cap;
  L.total = L.off_xs + xs_buffers * dict_cap * 8;
  return L;
}

struct ChunkDesc {       // written by the producer into shared memory before it arms the barrier
  int r0, r1;            // rows [r0, r1)
  long long ka;          // first staged nonzero (k0 rounded down to a multiple of 8)
  int ra, nd;            // first staged row (r0 rounded down to even), dictionary entries
  long long pad;
};
static_assert(sizeof(ChunkDesc) == 32, "ChunkDesc layout");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// 1-D bulk async copy global -> shared (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
template <int NCON>
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCON) : "memory"); }

// ---- consumer, step B: rows of one staged chunk, TPR lanes per row, all operands in shared memory.
// Four predicated entries per lane and batch, four independent accumulators.
template <int TPR, int MODE, int NCON>
__device__ __forceinline__ void reduce_rows(int r0, int r1, int ra, long long ka, const double* __restrict__ sval,
                                            const unsigned short* __restrict__ sidx, const double* __restrict__ xs,
                                            const long long* __restrict__ srp, const double* __restrict__ sepi,
                                            double* y, double omega, int ctid) {
  constexpr int NSUB = NCON / TPR;
  const int lane = ctid & (TPR - 1);
  const int sub = ctid / TPR;
  // the lanes of one sub-warp leave the row loop together, other sub-warps of the warp may not:
  // shuffles name exactly the sub-warp
  const unsigned submask = TPR == 32 ? 0xffffffffu : (((1u << TPR) - 1u) << ((ctid & 31) & ~(TPR - 1)));
  for (int row = r0 + sub; row < r1; row += NSUB) {
    const int s = (int)(srp[row - ra] - ka), e = (int)(srp[row - ra + 1] - ka);
    double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
#pragma unroll 2
    for (int k = s + lane; k < e; k += 4 * TPR) {
      int ix[4];
      double v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int kk = k + j * TPR;
        const bool ok = kk < e;
        ix[j] = ok ? (int)sidx[kk] : 0;
        v[j] = ok ? sval[kk] : 0.0;
      }
      a0 = fma(v[0], xs[ix[0]], a0);
      a1 = fma(v[1], xs[ix[1]], a1);
      a2 = fma(v[2], xs[ix[2]], a2);
      a3 = fma(v[3], xs[ix[3]], a3);
    }
    double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(submask, acc, o, TPR);
    if (lane == 0) {
      const int t = row - ra;
      if (MODE == Y_AX) y[row] = acc;
      else if (MODE == Y_ADD) y[row] = sepi[t] + acc;
      else if (MODE == RESID) y[row] = sepi[t] - acc;
      else if (MODE == RESID_W) y[row] = fma(sepi[kRpPad + t], sepi[t], -acc);
      else y[row] = fma(omega * sepi[kRpPad + t], sepi[t] - acc, sepi[2 * kRpPad + t]);
    }
  }
}

template <int MODE, int NCON>
__device__ __forceinline__ void reduce_chunk(int r0, int r1, int ra, long long ka, const unsigned char* st,
                                             const SpmvLayout& L, const double* xs, double* y, double omega, int ctid) {
  const double* sval = reinterpret_cast<const double*>(st);
  const unsigned short* sidx = reinterpret_cast<const unsigned short*>(st + L.off_idx);
  const long long* srp = reinterpret_cast<const long long*>(st + L.off_rp);
  const double* sepi = reinterpret_cast<const double*>(st + L.off_epi);
  // The reduction is latency-bound (dependent fp64 operations and shuffles; measured ~2300 cycles
  // per pass whatever the lane count), so rows in flight matter more than lanes per row: as many
  // lanes per row as still give every row of the chunk its own sub-warp in ONE pass, at most half
  // the mean row length (uniform over the CTA).
  const int nr = r1 - r0;
  const int nz = (int)(srp[r1 - ra] - srp[r0 - ra]);
  const int mean = nr > 0 ? nz / nr : 0;
  int tpr = 32;
  while (tpr > 2 && (tpr * nr > NCON || tpr * 2 > mean)) tpr >>= 1;
  if (tpr == 32) reduce_rows<32, MODE, NCON>(r0, r1, ra, ka, sval, sidx, xs, srp, sepi, y, omega, ctid);
  else if (tpr == 16) reduce_rows<16, MODE, NCON>(r0, r1, ra, ka, sval, sidx, xs, srp, sepi, y, omega, ctid);
  else if (tpr == 8) reduce_rows<8, MODE, NCON>(r0, r1, ra, ka, sval, sidx, xs, srp, sepi, y, omega, ctid);
  else if (tpr == 4) reduce_rows<4, MODE, NCON>(r0, r1, ra, ka, sval, sidx, xs, srp, sepi, y, omega, ctid);
  else reduce_rows<2, MODE, NCON>(r0, r1, ra, ka, sval, sidx, xs, srp, sepi, y, omega, ctid);
}

// CW consumer warps + 1 producer warp.  PIPE: the x gathers of chunk i+1 are in flight (registers)
// while the rows of chunk i are reduced; needs STAGES >= 3 and two xs buffers.
template <int CW, int STAGES, int MODE, bool PIPE, bool TIMING>
__global__ void __launch_bounds__((CW + 1) * 32, 4) spmv_tma_kernel(int64_t nchunks, const longlong4* __restrict__ cdesc,
                                                                 const int64_t* __restrict__ rowptr,
                                                                 const unsigned short* __restrict__ lidx,
                                                                 const int32_t* __restrict__ dict,
                                                                 const double* __restrict__ val, const double* __restrict__ x,
                                                                 const double* b, const double* __restrict__ dinv, double* y,
                                                                 double omega, const SpmvLayout L,
                                                                 long long* __restrict__ timing) {
  constexpr int NCON = CW * 32;
  constexpr int NG = kCap / NCON;         // gathers per thread for the largest dictionary
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* empty = full + STAGES;
  ChunkDesc* desc = reinterpret_cast<ChunkDesc*>(smem + L.off_desc);
  double* xs0 = reinterpret_cast<double*>(smem + L.off_xs);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == CW) {
    // ---------------- producer: one lane streams this CTA's chunks into the ring
    if ((threadIdx.x & 31) != 0) return;
    const uint64_t pol = policy_evict_first();
    const uint64_t keep = policy_evict_last();
    // chunk descriptors {k0, d0, r0} are one 32-byte record per chunk; they are requested two
    // chunks ahead so that the producer never waits on their latency
    const int64_t G = gridDim.x;
    int64_t c = blockIdx.x;
    longlong4 a0 = make_longlong4(0, 0, 0, 0), b0 = a0, a1 = a0, b1 = a0;
    if (c < nchunks) { a0 = cdesc[c]; b0 = cdesc[c + 1]; }
    if (c + G < nchunks) { a1 = cdesc[c + G]; b1 = cdesc[c + G + 1]; }
    for (int i = 0; c < nchunks; i++, c += G) {
      const int s = i % STAGES;
      longlong4 a2 = make_longlong4(0, 0, 0, 0), b2 = a2;
      if (c + 2 * G < nchunks) { a2 = cdesc[c + 2 * G]; b2 = cdesc[c + 2 * G + 1]; }
      const int64_t k0 = a0.x, k1 = b0.x, d0 = a0.y, d1 = b0.y;
      const int r0 = (int)a0.z, r1 = (int)b0.z;
      if (i >= STAGES) mbar_wait(empty + s, ((i / STAGES) - 1) & 1);
      unsigned char* st = smem + (size_t)s * L.stage_bytes;
      const int64_t ka = k0 & ~(int64_t)7, kb = (k1 + 7) & ~(int64_t)7;
      const int ra = r0 & ~1;
      const uint32_t nrp = (uint32_t)(((r1 - ra + 1) + 1) & ~1);
      const uint32_t n = (uint32_t)(kb - ka);
      const uint32_t nd = (uint32_t)(d1 - d0);          // multiple of 4
      desc[s].r0 = r0;
      desc[s].r1 = r1;
      desc[s].ka = ka;
      desc[s].ra = ra;
      desc[s].nd = (int)nd;
      constexpr uint32_t nepi = MODE == Y_AX ? 0u : (MODE == JACOBI ? 3u : (MODE == RESID_W ? 2u : 1u));
      mbar_expect_tx(full + s, n * 10u + nd * 4u + nrp * 8u * (1u + nepi));
      if (n) {
        bulk_g2s(st, val + ka, n * 8u, full + s, pol);
        bulk_g2s(st + L.off_idx, lidx + ka, n * 2u, full + s, pol);
      }
      if (nd) bulk_g2s(st + L.off_dict, dict + d0, nd * 4u, full + s, pol);
      bulk_g2s(st + L.off_rp, rowptr + ra, nrp * 8u, full + s, pol);
      if (MODE == Y_ADD) bulk_g2s(st + L.off_epi, y + ra, nrp * 8u, full + s, pol);
      if (MODE == RESID || MODE == JACOBI || MODE == RESID_W) bulk_g2s(st + L.off_epi, b + ra, nrp * 8u, full + s, pol);
      if (MODE == JACOBI || MODE == RESID_W) bulk_g2s(st + L.off_epi + kRpPad * 8, dinv + ra, nrp * 8u, full + s, pol);
      if (MODE == JACOBI) bulk_g2s(st + L.off_epi + 2 * kRpPad * 8, x + ra, nrp * 8u, full + s, keep);
      a0 = a1; b0 = b1; a1 = a2; b1 = b2;
    }
    return;
  }

  // ---------------- consumers
  const int ctid = threadIdx.x;
  const int64_t G = gridDim.x;
  long long t_wait = 0, t_a = 0, t_b = 0, t_rel = 0, t0 = 0, t1 = 0;     // TIMING only
  int i = 0;
  if (!PIPE) {
    for (int64_t c = blockIdx.x; c < nchunks; c += G, i++) {
      const int s = i % STAGES;
      if (TIMING) t0 = clock64();
      mbar_wait(full + s, (i / STAGES) & 1);
      if (TIMING) { t1 = clock64(); t_wait += t1 - t0; t0 = t1; }
      unsigned char* st = smem + (size_t)s * L.stage_bytes;
      const int* sdict = reinterpret_cast<const int*>(st + L.off_dict);
      const int r0 = desc[s].r0, r1 = desc[s].r1, ra = desc[s].ra, nd = desc[s].nd;
      const long long ka = desc[s].ka;
      // A. x at the chunk's distinct columns -> shared memory (independent gathers, sorted addresses)
      for (int d = ctid; d < nd; d += 4 * NCON) {
        int cc[4];
        double xv[4];
#pragma unroll
        for (int j = 0; j < 4; j++) cc[j] = d + j * NCON < nd ? sdict[d + j * NCON] : 0;
#pragma unroll
        for (int j = 0; j < 4; j++) xv[j] = (j == 0 || d + j * NCON < nd) ? __ldg(x + cc[j]) : 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (d + j * NCON < nd) xs0[d + j * NCON] = xv[j];
      }
      consumer_sync<NCON>();
      if (TIMING) { t1 = clock64(); t_a += t1 - t0; t0 = t1; }
      // B. rows
      reduce_chunk<MODE, NCON>(r0, r1, ra, ka, st, L, xs0, y, omega, ctid);
      if (TIMING) { t1 = clock64(); t_b += t1 - t0; t0 = t1; }
      // the stage goes back to the producer, and xs may be rewritten, only when every consumer is done
      consumer_sync<NCON>();
      if (ctid == 0) mbar_arrive(empty + s);
      if (TIMING) { t1 = clock64(); t_rel += t1 - t0; }
    }
  } else {
    int64_t c = blockIdx.x;
    if (c < nchunks) {          // prologue: dictionary of the first chunk
      mbar_wait(full + 0, 0);
      const int* sdict = reinterpret_cast<const int*>(smem + L.off_dict);
      const int nd = desc[0].nd;
      for (int d = ctid; d < nd; d += NCON) xs0[d] = __ldg(x + sdict[d]);
      consumer_sync<NCON>();
    }
    for (; c < nchunks; c += G, i++) {
      const int s = i % STAGES;
      unsigned char* st = smem + (size_t)s * L.stage_bytes;
      const int r0 = desc[s].r0, r1 = desc[s].r1, ra = desc[s].ra;
      const long long ka = desc[s].ka;
      const bool has_next = c + G < nchunks;
      double xv[NG];
      int ndn = 0;
      if (TIMING) t0 = clock64();
      if (has_next) {           // gathers of the next chunk fly while this chunk's rows are reduced
        const int sn = (i + 1) % STAGES;
        mbar_wait(full + sn, ((i + 1) / STAGES) & 1);
        if (TIMING) { t1 = clock64(); t_wait += t1 - t0; t0 = t1; }
        const int* sdict = reinterpret_cast<const int*>(smem + (size_t)sn * L.stage_bytes + L.off_dict);
        ndn = desc[sn].nd;
        int cc[NG];
#pragma unroll
        for (int j = 0; j < NG; j++) cc[j] = ctid + j * NCON < ndn ? sdict[ctid + j * NCON] : 0;
#pragma unroll
        for (int j = 0; j < NG; j++) xv[j] = (j == 0 || ctid + j * NCON < ndn) ? __ldg(x + cc[j]) : 0.0;
      }
      if (TIMING) { t1 = clock64(); t_a += t1 - t0; t0 = t1; }
      reduce_chunk<MODE, NCON>(r0, r1, ra, ka, st, L, xs0 + (size_t)(i & 1) * L.xs_doubles, y, omega, ctid);
      if (TIMING) { t1 = clock64(); t_b += t1 - t0; t0 = t1; }
      if (has_next) {
        double* xn = xs0 + (size_t)((i + 1) & 1) * L.xs_doubles;
#pragma unroll
        for (int j = 0; j < NG; j++)
          if (ctid + j * NCON < ndn) xn[ctid + j * NCON] = xv[j];
      }
      consumer_sync<NCON>();
      if (ctid == 0) mbar_arrive(empty + s);
      if (TIMING) { t1 = clock64(); t_rel += t1 - t0; }
    }
  }
  if (TIMING && blockIdx.x == 0 && (ctid & 31) == 0 && (ctid >> 5) < 8) {
    long long* o = timing + (ctid >> 5) * 5;
    o[0] = t_wait; o[1] = t_a; o[2] = t_b; o[3] = t_rel; o[4] = i;
  }
}

// chunk_row[c] = first row whose key rowptr[r] + kRowWeight * r is >= c * T   (c = 0 .. nchunks)
__global__ void chunk_rows_kernel(int64_t nrows, const int64_t* __restrict__ rowptr, int64_t T, int64_t nchunks,
                                  int32_t* __restrict__ chunk_row) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= nchunks; c += stride) {
    const int64_t target = c * T;
    int64_t lo = 0, hi = nrows;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (rowptr[mid] + (int64_t)kRowWeight * mid < target) lo = mid + 1;
      else hi = mid;
    }
    chunk_row[c] = (int32_t)(c == nchunks ? nrows : lo);
  }
}

// packed producer descriptors: cdesc[c] = {rowptr[chunk_row[c]], dict_ptr[c], chunk_row[c], 0}, c = 0..nchunks
__global__ void chunk_desc_kernel(int64_t nchunks, const int32_t* __restrict__ chunk_row,
                                  const int64_t* __restrict__ rowptr, const int64_t* __restrict__ dict_ptr,
                                  longlong4* __restrict__ cdesc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= nchunks; c += stride) {
    const int r = chunk_row[c];
    cdesc[c] = make_longlong4(rowptr[r], dict_ptr[c], r, 0);
  }
}

// Column dictionary of every chunk: one CTA per chunk sorts the chunk's columns in shared memory
// (bitonic), compacts the distinct ones and (FILL) writes them to dict[dict_ptr[c]..] and the
// position of every nonzero's column to lidx.  FILL=false only counts (rounded up to 4 entries
// so that every dictionary starts 16-byte aligned).
constexpr int kDictThreads = 256;
template <bool FILL>
__global__ void __launch_bounds__(kDictThreads) chunk_dict_kernel(int64_t nchunks, const int32_t* __restrict__ chunk_row,
                                                                  const int64_t* __restrict__ rowptr,
                                                                  const int32_t* __restrict__ col,
                                                                  int64_t* __restrict__ count_or_ptr,
                                                                  int32_t* __restrict__ dict,
                                                                  unsigned short* __restrict__ lidx) {
  __shared__ int32_t skey[kCap];
  __shared__ int32_t sdict[kCap];
  __shared__ int s_nd;
  __shared__ int s_warp[kDictThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int64_t k0 = rowptr[chunk_row[c]], k1 = rowptr[chunk_row[c + 1]];
    const int n = (int)(k1 - k0);
    int n2 = 64;
    while (n2 < n) n2 <<= 1;
    for (int t = tid; t < n2; t += kDictThreads) skey[t] = t < n ? col[k0 + t] : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < n2; t += kDictThreads) {
          const int p = t ^ j;
          if (p > t) {
            const int32_t a = skey[t], bb = skey[p];
            const bool up = ((t & k) == 0);
            if ((a > bb) == up) { skey[t] = bb; skey[p] = a; }
          }
        }
        __syncthreads();
      }
    }
    // compaction of the distinct keys: every thread owns a contiguous run of n2/256 sorted keys
    const int per = n2 / kDictThreads > 0 ? n2 / kDictThreads : 1;
    const int t0 = tid * per;
    int mine = 0;
    for (int t = t0; t < t0 + per && t < n; t++) mine += (t == 0 || skey[t] != skey[t - 1]) ? 1 : 0;
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wib] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < wib; w++) base += s_warp[w];
    if (tid == kDictThreads - 1) s_nd = base + incl;
    int pos = base + incl - mine;
    if (FILL)
      for (int t = t0; t < t0 + per && t < n; t++)
        if (t == 0 || skey[t] != skey[t - 1]) sdict[pos++] = skey[t];
    __syncthreads();
    const int nd = s_nd;
    const int ndp = (nd + 3) & ~3;
    if (!FILL) {
      if (tid == 0) count_or_ptr[c] = ndp;
    } else {
      const int64_t d0 = count_or_ptr[c];
      for (int t = tid; t < ndp; t += kDictThreads) dict[d0 + t] = sdict[t < nd ? t : nd - 1];
      for (int t = tid; t < n; t += kDictThreads) {
        const int32_t key = col[k0 + t];
        int lo = 0, hi = nd;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (sdict[mid] < key) lo = mid + 1;
          else hi = mid;
        }
        lidx[k0 + t] = (unsigned short)lo;
      }
    }
    __syncthreads();
  }
}

// ---- register-streaming kernel: any row length, no staging (small or irregular matrices) -------
constexpr int kBlock = 256;
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <int TPR, int MODE>
__global__ void __launch_bounds__(kBlock) spmv_kernel(int64_t nrows, const int64_t* __restrict__ rowptr,
                                                      const int32_t* __restrict__ col,
                                                      const double* __restrict__ val, const double* __restrict__ x,
                                                      const double* __restrict__ b, const double* __restrict__ dinv,
                                                      double* __restrict__ y, double omega) {
  const int lane = threadIdx.x & (TPR - 1);
  const int64_t sub = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / TPR;
  const int64_t nsub = ((int64_t)gridDim.x * blockDim.x) / TPR;
  const unsigned submask = TPR == 32 ? 0xffffffffu : (((1u << TPR) - 1u) << ((threadIdx.x & 31) & ~(TPR - 1)));
  for (int64_t row = sub; row < nrows; row += nsub) {
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    double acc0 = 0., acc1 = 0.;
    int64_t k = s + lane;
    for (; k + TPR < e; k += 2 * TPR) {
      const double v0 = ld_stream(val + k), v1 = ld_stream(val + k + TPR);
      const int c0 = ld_stream(col + k), c1 = ld_stream(col + k + TPR);
      acc0 = fma(v0, x[c0], acc0);
      acc1 = fma(v1, x[c1], acc1);
    }
    if (k < e) acc0 = fma(ld_stream(val + k), x[ld_stream(col + k)], acc0);
    double acc = acc0 + acc1;
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(submask, acc, o, TPR);
    if (lane == 0) {
      if (MODE == Y_AX) y[row] = acc;
      else if (MODE == Y_ADD) y[row] += acc;
      else if (MODE == RESID) y[row] = b[row] - acc;
      else if (MODE == RESID_W) y[row] = fma(dinv[row], b[row], -acc);
      else y[row] = fma(omega * dinv[row], b[row] - acc, x[row]);
    }
  }
}

template <int MODE>
int launch_stream(const b2_csr* A, const double* x, const double* b, const double* dinv, double* y, double omega) {
  b2_ctx* c = A->ctx;
  const int tpr = A->tpr;
  const int64_t threads = A->nrows * tpr;
  const int grid = b2_grid_for(c, threads, kBlock, 8 * 4);
#define B2_SPMV_CASE(T)                                                                                    \
  case T:                                                                                                  \
    B2_LAUNCH(c, (spmv_kernel<T, MODE>), grid, kBlock, 0, A->nrows, A->rowptr, A->col, A->val, x, b, dinv, \
              y, omega);                                                                                   \
    break;
  switch (tpr) {
    B2_SPMV_CASE(1)
    B2_SPMV_CASE(2)
    B2_SPMV_CASE(4)
    B2_SPMV_CASE(8)
    B2_SPMV_CASE(16)
    B2_SPMV_CASE(32)
    default:
      B2_CHECK(false, "bad tpr %d", tpr);
  }
#undef B2_SPMV_CASE
  return 0;
}

template <int CW, int STAGES, int MODE, bool PIPE>
int launch_tma(const b2_csr* A, const double* x, const double* b, const double* dinv, double* y, double omega) {
  b2_ctx* c = A->ctx;
  constexpr int nepi = MODE == Y_AX ? 0 : (MODE == JACOBI ? 3 : (MODE == RESID_W ? 2 : 1));
  constexpr int threads = (CW + 1) * 32;
  const SpmvLayout L = make_layout(STAGES, (int)A->dict_cap, nepi, PIPE ? 2 : 1);
  B2_CHECK(L.total <= 227 * 1024, "spmv: %d bytes of shared memory needed", L.total);
  auto kern = spmv_tma_kernel<CW, STAGES, MODE, PIPE, false>;
  static int configured_bytes = 0, ctas_per_sm = 1;      // per instantiation
  if (configured_bytes < L.total) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    configured_bytes = L.total;
  }
  B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, threads, (size_t)L.total));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  int64_t grid = (int64_t)c->sm_count * ctas_per_sm;
  if (grid > A->nchunks) grid = A->nchunks;
  if (MODE == Y_AX && c->spmv_timing) {      // diagnostic: cycles per phase of the consumer warps of CTA 0
    auto tk = spmv_tma_kernel<CW, STAGES, MODE, PIPE, true>;
    B2_CUDA(cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    long long* d_t = nullptr;
    B2_CUDA(cudaMalloc(&d_t, 8 * 5 * sizeof(long long)));
    tk<<<(int)grid, threads, L.total, c->stream>>>(A->nchunks, (const longlong4*)A->cdesc, A->rowptr, A->lidx, A->dict,
                                                   A->val, x, b, dinv, y, omega, L, d_t);
    long long h[40];
    B2_CUDA(cudaMemcpyAsync(h, d_t, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_t);
    fprintf(stderr, "[spmv timing] CW %d stages %d pipe %d: %d CTAs/SM, %d B smem, dict cap %d\n", CW, STAGES, (int)PIPE,
            ctas_per_sm, L.total, (int)A->dict_cap);
    for (int w = 0; w < 8; w += 3)
      fprintf(stderr, "[spmv timing]   warp %d: chunks %lld  cycles/chunk: wait %.0f  stepA %.0f  stepB %.0f  release %.0f\n",
              w, h[w * 5 + 4], (double)h[w * 5] / h[w * 5 + 4], (double)h[w * 5 + 1] / h[w * 5 + 4],
              (double)h[w * 5 + 2] / h[w * 5 + 4], (double)h[w * 5 + 3] / h[w * 5 + 4]);
    return 0;
  }
  B2_LAUNCH(c, kern, (int)grid, threads, (size_t)L.total, A->nchunks, (const longlong4*)A->cdesc, A->rowptr, A->lidx,
            A->dict, A->val, x, b, dinv, y, omega, L, (long long*)nullptr);
  return 0;
}

template <int MODE>
int launch_spmv(const b2_csr* A, const double* x, const double* b, const double* dinv, double* y, double omega) {
  if (A->nrows == 0) return 0;
  b2_prof_scope prof(A->ctx, A);
  if (!A->chunk_row || A->ctx->spmv_variant == 0) return launch_stream<MODE>(A, x, b, dinv, y, omega);
  if (A->ctx->spmv_variant == 2) return launch_tma<8, 3, MODE, true>(A, x, b, dinv, y, omega);
  return launch_tma<8, 2, MODE, false>(A, x, b, dinv, y, omega);
}

}  // namespace

// SpMV plan: row chunks, per-chunk column dictionaries and 16-bit local column indices
// (called from b2_csr_finalize; the pattern of a b2_csr never changes afterwards)
void b2_csr_free_plan(b2_csr* A) {
  b2_ctx* c = A->ctx;
  if (A->chunk_row) b2_free(c, A->chunk_row, (size_t)A->nchunks + 1);
  if (A->dict_ptr) b2_free(c, A->dict_ptr, (size_t)A->nchunks + 1);
  if (A->dict) b2_free(c, A->dict, (size_t)A->dict_total + 4);
  if (A->cdesc) b2_free(c, A->cdesc, ((size_t)A->nchunks + 1) * 4);
  A->cdesc = nullptr;
  if (A->lidx) b2_free(c, A->lidx, (size_t)A->nnz + 16);
  A->chunk_row = nullptr;
  A->dict_ptr = nullptr;
  A->dict = nullptr;
  A->lidx = nullptr;
  A->nchunks = 0;
  A->dict_total = 0;
  A->dict_cap = 0;
}

int b2_csr_build_chunks(b2_csr* A) {
  b2_ctx* c = A->ctx;
  b2_csr_free_plan(A);
  if (A->nrows == 0 || A->nnz == 0) return 0;
  if (A->max_row + kRowWeight + 16 > kCap / 2) return 0;      // rows too long to stage: streaming kernel
  if (A->nnz < 16 * A->nrows) return 0;     // short rows (prolongators): chunks would be half empty, streaming kernel
  const int64_t T = kCap - A->max_row - kRowWeight - 8;
  const int64_t total = A->nnz + (int64_t)kRowWeight * A->nrows;
  A->nchunks = total / T + 1;
  B2_TRY(b2_malloc(c, &A->chunk_row, (size_t)A->nchunks + 1));
  B2_TRY(b2_malloc(c, &A->dict_ptr, (size_t)A->nchunks + 1));
  B2_TRY(b2_malloc(c, &A->lidx, (size_t)A->nnz + 16));
  B2_CUDA(cudaMemsetAsync(A->lidx, 0, ((size_t)A->nnz + 16) * sizeof(unsigned short), c->stream));
  B2_LAUNCH(c, chunk_rows_kernel, b2_grid_for(c, A->nchunks + 1, 256, 8), 256, 0, A->nrows, A->rowptr, T, A->nchunks,
            A->chunk_row);
  const int grid = (int)(A->nchunks < (int64_t)c->sm_count * 8 ? A->nchunks : (int64_t)c->sm_count * 8);
  B2_LAUNCH(c, chunk_dict_kernel<false>, grid, kDictThreads, 0, A->nchunks, A->chunk_row, A->rowptr, A->col, A->dict_ptr,
            (int32_t*)nullptr, (unsigned short*)nullptr);
  // largest dictionary (sizes the shared-memory layout), then
  // exclusive scan of the padded dictionary sizes -> dict_ptr[0..nchunks]
  {
    int64_t* d_max = nullptr;
    B2_TRY(b2_malloc(c, &d_max, 1));
    size_t rb = 0;
    B2_CUDA(cub::DeviceReduce::Max(nullptr, rb, A->dict_ptr, d_max, (int)A->nchunks, c->stream));
    void* rtmp = nullptr;
    B2_CUDA(cudaMalloc(&rtmp, rb ? rb : 16));
    cudaError_t re = cub::DeviceReduce::Max(rtmp, rb, A->dict_ptr, d_max, (int)A->nchunks, c->stream);
    c->launches += 1;
    int64_t mx = 0;
    if (re == cudaSuccess) re = cudaMemcpyAsync(&mx, d_max, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(rtmp);
    b2_free(c, d_max, 1);
    B2_CUDA(re);
    A->dict_cap = (mx + 15) / 16 * 16;
    if (A->dict_cap < 16) A->dict_cap = 16;
  }
  B2_CUDA(cudaMemsetAsync(A->dict_ptr + A->nchunks, 0, sizeof(int64_t), c->stream));
  B2_CHECK(A->nchunks + 1 < ((int64_t)1 << 31), "too many chunks");
  size_t tmp_bytes = 0;
  B2_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, A->dict_ptr, A->dict_ptr, (int)(A->nchunks + 1), c->stream));
  void* tmp = nullptr;
  B2_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, A->dict_ptr, A->dict_ptr, (int)(A->nchunks + 1), c->stream);
  c->launches += 2;
  int64_t dict_total = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&dict_total, A->dict_ptr + A->nchunks, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream);
  cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  B2_CUDA(e);
  A->dict_total = dict_total;
  B2_TRY(b2_malloc(c, &A->dict, (size_t)dict_total + 4));
  B2_CUDA(cudaMemsetAsync(A->dict, 0, ((size_t)dict_total + 4) * sizeof(int32_t), c->stream));
  B2_LAUNCH(c, chunk_dict_kernel<true>, grid, kDictThreads, 0, A->nchunks, A->chunk_row, A->rowptr, A->col, A->dict_ptr,
            A->dict, A->lidx);
  B2_TRY(b2_malloc(c, &A->cdesc, ((size_t)A->nchunks + 1) * 4));
  B2_LAUNCH(c, chunk_desc_kernel, b2_grid_for(c, A->nchunks + 1, 256, 8), 256, 0, A->nchunks, A->chunk_row, A->rowptr,
            A->dict_ptr, (longlong4*)A->cdesc);
  return 0;
}

int b2_csr_resid_w(const b2_csr* A, const double* b, const double* w, const double* x, double* r) {
  return launch_spmv<RESID_W>(A, x, b, w, r, 0.);
}

extern "C" {

int b2_csr_spmv(const b2_csr* A, const b2_vec* x, b2_vec* y) {
  B2_CHECK(x->n >= A->ncols && y->n >= A->nrows && x != y, "b2_csr_spmv: bad operands");
  return launch_spmv<Y_AX>(A, x->d, nullptr, nullptr, y->d, 0.);
}
int b2_csr_spmv_add(const b2_csr* A, const b2_vec* x, b2_vec* y) {
  B2_CHECK(x->n >= A->ncols && y->n >= A->nrows && x != y, "b2_csr_spmv_add: bad operands");
  return launch_spmv<Y_ADD>(A, x->d, nullptr, nullptr, y->d, 0.);
}
int b2_csr_resid(const b2_csr* A, const b2_vec* b, const b2_vec* x, b2_vec* r) {
  B2_CHECK(x->n >= A->ncols && r->n >= A->nrows && b->n >= A->nrows && x != r, "b2_csr_resid: bad operands");
  return launch_spmv<RESID>(A, x->d, b->d, nullptr, r->d, 0.);
}
int b2_csr_jacobi_sweep(const b2_csr* A, const b2_vec* dinv, const b2_vec* b, const b2_vec* xin, b2_vec* xout,
                        double omega) {
  B2_CHECK(A->nrows == A->ncols && xin != xout && xin->n >= A->nrows && xout->n >= A->nrows,
           "b2_csr_jacobi_sweep: bad operands");
  return launch_spmv<JACOBI>(A, xin->d, b->d, dinv->d, xout->d, omega);
}

}  // extern "C"
