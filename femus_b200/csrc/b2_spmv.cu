// fp64 CSR SpMV family with TMA-staged row blocks (sm_100a).  Replaces the MatMult / MatMultAdd
// calls behind PetscVector::{matrix_mult, resid} (reference src/03_algebra/00_vectors/
// PetscVector.cpp:193-247) and the KSPRICHARDSON+PCJACOBI sweep PETSc runs on every level
// (LinearEquationSolverPetsc.cpp:516-519, PetscPreconditioner.cpp:209-212).
//
// The matrix is cut into CHUNKS of whole consecutive rows (b2_csr_finalize): chunk c holds the rows
// whose key  rowptr[r] + kRowWeight * r  lies in [c*T, (c+1)*T), so that a chunk never has more
// than kCap nonzeros nor more than kMaxRows rows.  The values, columns and row pointers of a chunk
// are three contiguous byte ranges of the CSR arrays; a producer warp streams them into a
// kStages-deep shared-memory ring with 1-D bulk async copies (cp.async.bulk, the TMA engine) that
// complete on an mbarrier, with an L2 evict-first policy so the 12 B/nnz stream does not push the
// x vector out of L2.  Eight consumer warps wait on the stage's "full" barrier and work in two
// phases.  Phase 1 is FLAT over the chunk's nonzeros -- thread t takes entries t, t+256, ... so the
// load is balanced whatever the row lengths and every thread has eight independent x gathers in
// flight (read-only path); the products overwrite the staged values.  Phase 2 gives every row to a
// sub-warp of TPR lanes (TPR from the chunk's mean row length) that sums its products from shared
// memory and applies the epilogue, selected at compile time:
//   Y_AX    y = A x                                 (MatMult)
//   Y_ADD   y += A x                                (MatMultAdd)
//   RESID   y = b - A x                             (resid)
//   JACOBI  y = x + omega * dinv .* (b - A x)       (one Richardson+Jacobi sweep, x != y)
//   RESID_W y = w .* b - A x                        (distributed residual, w = 1/multiplicity)
// Consumers synchronise among themselves with a named barrier (the producer warp never joins it)
// and hand the stage back through an "empty" mbarrier.
// Algorithmic traffic: 12 B per nonzero + 8 B rowptr + 8 B y per row, x read once from HBM.
// Matrices whose longest row does not fit a chunk use the register-streaming kernel at the bottom.
#include "b2_common.cuh"

namespace {

enum SpmvMode { Y_AX, Y_ADD, RESID, JACOBI, RESID_W };

// ---- chunk geometry --------------------------------------------------------------------------
constexpr int kCap = 2048;          // staged nonzeros per chunk
constexpr int kRowWeight = 8;       // a row counts as this many nonzeros when cutting chunks
constexpr int kMaxRows = kCap / kRowWeight;   // 256
constexpr int kCapPad = kCap + 8;   // room for the 16-byte alignment of both ends
constexpr int kRpPad = kMaxRows + 4;
constexpr int kConsumerWarps = 8;
constexpr int kThreads = (kConsumerWarps + 1) * 32;

template <int STAGES>
struct SpmvSmem {
  static constexpr size_t stage_bytes = (size_t)kCapPad * 8 + (size_t)kCapPad * 4 + (size_t)kRpPad * 8;
  static constexpr size_t epi_offset = STAGES * stage_bytes + STAGES * 2 * 8 + STAGES * 32;
  static constexpr size_t bytes = epi_offset + 3 * kMaxRows * 8;
};

struct ChunkDesc {       // written by the producer into shared memory before it arms the barrier
  int r0, r1;            // rows [r0, r1)
  long long ka;          // first staged nonzero (k0 rounded down to a multiple of 4)
  int ra, pad;           // first staged row pointer (r0 rounded down to even)
};
static_assert(sizeof(ChunkDesc) == 24 || sizeof(ChunkDesc) == 32, "ChunkDesc layout");

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
// 1-D bulk async copy global -> shared (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// ---- consumer, phase 2: row sums of the products left in shared memory, TPR lanes per row ------
template <int TPR, int MODE>
__device__ __forceinline__ void reduce_rows(int r0, int r1, int ra, long long ka, const double* __restrict__ sval,
                                            const long long* __restrict__ srp, const double* __restrict__ sepi,
                                            double* y, double omega, int ctid) {
  constexpr int NSUB = kConsumerWarps * 32 / TPR;
  const int lane = ctid & (TPR - 1);
  const int sub = ctid / TPR;
  // the lanes of one sub-warp leave the row loop together, other sub-warps of the warp may not:
  // shuffles name exactly the sub-warp
  const unsigned submask = TPR == 32 ? 0xffffffffu : (((1u << TPR) - 1u) << ((ctid & 31) & ~(TPR - 1)));
  for (int row = r0 + sub; row < r1; row += NSUB) {
    const int s = (int)(srp[row - ra] - ka), e = (int)(srp[row - ra + 1] - ka);
    double a0 = 0., a1 = 0.;
    int k = s + lane;
    for (; k + TPR < e; k += 2 * TPR) {
      a0 += sval[k];
      a1 += sval[k + TPR];
    }
    if (k < e) a0 += sval[k];
    double acc = a0 + a1;
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(submask, acc, o, TPR);
    if (lane == 0) {
      const int t = row - r0;
      if (MODE == Y_AX) y[row] = acc;
      else if (MODE == Y_ADD) y[row] = sepi[t] + acc;
      else if (MODE == RESID) y[row] = sepi[t] - acc;
      else if (MODE == RESID_W) y[row] = fma(sepi[kMaxRows + t], sepi[t], -acc);
      else y[row] = fma(omega * sepi[kMaxRows + t], sepi[t] - acc, sepi[2 * kMaxRows + t]);
    }
  }
}

template <int STAGES, int MODE>
__global__ void __launch_bounds__(kThreads) spmv_tma_kernel(int64_t nchunks, const int32_t* __restrict__ chunk_row,
                                                            const int64_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ col, const double* __restrict__ val,
                                                            const double* __restrict__ x, const double* b,
                                                            const double* __restrict__ dinv, double* y, double omega) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr size_t SB = SpmvSmem<STAGES>::stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * SB);
  uint64_t* empty = full + STAGES;
  ChunkDesc* desc = reinterpret_cast<ChunkDesc*>(empty + STAGES);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ---------------- producer: one lane streams this CTA's chunks into the ring
    if ((threadIdx.x & 31) != 0) return;
    const uint64_t pol = policy_evict_first();
    int64_t c = blockIdx.x;
    int r0 = 0, r1 = 0;
    int64_t k0 = 0, k1 = 0;
    if (c < nchunks) {
      r0 = chunk_row[c];
      r1 = chunk_row[c + 1];
      k0 = rowptr[r0];
      k1 = rowptr[r1];
    }
    for (int i = 0; c < nchunks; i++, c += gridDim.x) {
      const int s = i % STAGES;
      // descriptors of the next chunk are requested before this one is issued (hidden latency)
      const int64_t cn = c + gridDim.x;
      int nr0 = 0, nr1 = 0;
      int64_t nk0 = 0, nk1 = 0;
      if (cn < nchunks) {
        nr0 = chunk_row[cn];
        nr1 = chunk_row[cn + 1];
        nk0 = rowptr[nr0];
        nk1 = rowptr[nr1];
      }
      if (i >= STAGES) mbar_wait(empty + s, ((i / STAGES) - 1) & 1);
      unsigned char* st = smem + (size_t)s * SB;
      const int64_t ka = k0 & ~(int64_t)3, kb = (k1 + 3) & ~(int64_t)3;
      const int ra = r0 & ~1;
      const int nrp = ((r1 - ra + 1) + 1) & ~1;
      const uint32_t n = (uint32_t)(kb - ka);
      desc[s].r0 = r0;
      desc[s].r1 = r1;
      desc[s].ka = ka;
      desc[s].ra = ra;
      mbar_expect_tx(full + s, n * 12u + (uint32_t)nrp * 8u);
      if (n) {
        bulk_g2s(st, val + ka, n * 8u, full + s, pol);
        bulk_g2s(st + (size_t)kCapPad * 8, col + ka, n * 4u, full + s, pol);
      }
      bulk_g2s(st + (size_t)kCapPad * 12, rowptr + ra, (uint32_t)nrp * 8u, full + s, pol);
      r0 = nr0; r1 = nr1; k0 = nk0; k1 = nk1;
    }
    return;
  }

  // ---------------- consumers
  // Phase 1 is flat over the chunk's nonzeros (perfect balance whatever the row lengths, eight
  // independent x gathers in flight per thread): products overwrite the staged values.  Phase 2
  // sums the rows from shared memory.  The epilogue operands of the chunk's rows (one row per
  // thread, coalesced) are requested before the gathers and parked in shared memory.
  const int ctid = threadIdx.x;
  double* sepi = reinterpret_cast<double*>(smem + SpmvSmem<STAGES>::epi_offset);   // [3][kMaxRows]
  int i = 0;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x, i++) {
    const int s = i % STAGES;
    mbar_wait(full + s, (i / STAGES) & 1);
    unsigned char* st = smem + (size_t)s * SB;
    double* sval = reinterpret_cast<double*>(st);
    const int* scol = reinterpret_cast<const int*>(st + (size_t)kCapPad * 8);
    const long long* srp = reinterpret_cast<const long long*>(st + (size_t)kCapPad * 12);
    const int r0 = desc[s].r0, r1 = desc[s].r1, ra = desc[s].ra;
    const long long ka = desc[s].ka;
    const int nr = r1 - r0;
    const int lo = (int)(srp[r0 - ra] - ka), hi = (int)(srp[r1 - ra] - ka);
    double e0 = 0., e1 = 0., e2 = 0.;
    if (MODE != Y_AX && ctid < nr) {
      const int row = r0 + ctid;
      if (MODE == Y_ADD) e0 = y[row];
      else e0 = b[row];
      if (MODE == JACOBI || MODE == RESID_W) e1 = dinv[row];
      if (MODE == JACOBI) e2 = x[row];
    }
    for (int k = lo + ctid; k < hi; k += kConsumerWarps * 32 * 8) {
      int cc[8];
      double xv[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int kk = k + j * kConsumerWarps * 32;
        cc[j] = kk < hi ? scol[kk] : 0;
      }
#pragma unroll
      for (int j = 0; j < 8; j++) xv[j] = __ldg(x + cc[j]);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int kk = k + j * kConsumerWarps * 32;
        if (kk < hi) sval[kk] *= xv[j];
      }
    }
    if (MODE != Y_AX && ctid < nr) {
      sepi[ctid] = e0;
      if (MODE == JACOBI || MODE == RESID_W) sepi[kMaxRows + ctid] = e1;
      if (MODE == JACOBI) sepi[2 * kMaxRows + ctid] = e2;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
    // lanes per row from the chunk's mean row length (uniform over the CTA)
    const int mean = nr > 0 ? (hi - lo) / nr : 0;
    if (mean >= 48) reduce_rows<32, MODE>(r0, r1, ra, ka, sval, srp, sepi, y, omega, ctid);
    else if (mean >= 24) reduce_rows<16, MODE>(r0, r1, ra, ka, sval, srp, sepi, y, omega, ctid);
    else if (mean >= 12) reduce_rows<8, MODE>(r0, r1, ra, ka, sval, srp, sepi, y, omega, ctid);
    else if (mean >= 6) reduce_rows<4, MODE>(r0, r1, ra, ka, sval, srp, sepi, y, omega, ctid);
    else reduce_rows<2, MODE>(r0, r1, ra, ka, sval, srp, sepi, y, omega, ctid);
    // the stage goes back to the producer, and the epilogue scratch may be rewritten, only when
    // every consumer warp is done with this chunk
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
    if (ctid == 0) mbar_arrive(empty + s);
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

template <int STAGES, int MODE>
int launch_tma(const b2_csr* A, const double* x, const double* b, const double* dinv, double* y, double omega,
               int ctas_per_sm) {
  b2_ctx* c = A->ctx;
  auto kern = spmv_tma_kernel<STAGES, MODE>;
  const size_t smem = SpmvSmem<STAGES>::bytes;
  static bool configured = false;      // per instantiation
  if (!configured) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int64_t grid = (int64_t)c->sm_count * ctas_per_sm;
  if (grid > A->nchunks) grid = A->nchunks;
  B2_LAUNCH(c, kern, (int)grid, kThreads, smem, A->nchunks, A->chunk_row, A->rowptr, A->col, A->val, x, b, dinv, y, omega);
  return 0;
}

template <int MODE>
int launch_spmv(const b2_csr* A, const double* x, const double* b, const double* dinv, double* y, double omega) {
  if (A->nrows == 0) return 0;
  b2_prof_scope prof(A->ctx, A);
  if (!A->chunk_row || A->ctx->spmv_variant == 0) return launch_stream<MODE>(A, x, b, dinv, y, omega);
  switch (A->ctx->spmv_variant) {
    case 2: return launch_tma<2, MODE>(A, x, b, dinv, y, omega, 3);
    case 3: return launch_tma<3, MODE>(A, x, b, dinv, y, omega, 2);
    case 4: return launch_tma<4, MODE>(A, x, b, dinv, y, omega, 2);
    case 6: return launch_tma<6, MODE>(A, x, b, dinv, y, omega, 1);
    default: return launch_tma<3, MODE>(A, x, b, dinv, y, omega, 2);
  }
}

}  // namespace

// cut A into row chunks for the staged kernel (called from b2_csr_finalize)
int b2_csr_build_chunks(b2_csr* A) {
  b2_ctx* c = A->ctx;
  if (A->chunk_row) { b2_free(c, A->chunk_row, (size_t)A->nchunks + 1); A->chunk_row = nullptr; }
  A->nchunks = 0;
  if (A->nrows == 0 || A->nnz == 0) return 0;
  if (A->max_row + kRowWeight + 16 > kCap / 2) return 0;      // rows too long to stage: streaming kernel
  const int64_t T = kCap - A->max_row - kRowWeight - 8;
  const int64_t total = A->nnz + (int64_t)kRowWeight * A->nrows;
  A->nchunks = total / T + 1;
  B2_TRY(b2_malloc(c, &A->chunk_row, (size_t)A->nchunks + 1));
  B2_LAUNCH(c, chunk_rows_kernel, b2_grid_for(c, A->nchunks + 1, 256, 8), 256, 0, A->nrows, A->rowptr, T, A->nchunks,
            A->chunk_row);
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
