// Element-block (ASM / Vanka) smoother: the preconditioner LinearEquationSolverPetscAsm configures in PETSc
// (reference src/08_algebra.../03_solvers_with_preconditioner/petsc_asm/LinearEquationSolverPetscAsm.cpp:266-340,
// 03_algebra/02_preconditioners/PetscPreconditioner.cpp:179-184): PCASM with the caller's overlapping index sets
// (BuildASMIndex, :91-262), PC_ASM_BASIC, local type PC_COMPOSITE_MULTIPLICATIVE, overlap 0, exact block solves
// (MLU_PRECOND on the blocks).  On one rank PCApply_ASM is the multiplicative sweep
//     y = 0;  for i = 0 .. nblocks-1:   y[B_i] += A[B_i,B_i]^-1 (r - A y)[B_i]
// which is sequential as written.  Here the blocks arrive with a SCHEDULE: ordered groups of mutually independent
// blocks (no block of a group reads or writes a dof another block of the group writes).  Blocks of one group
// commute, so one launch per group -- one CTA per block -- reproduces the sweep in any block order that respects
// the groups: the reference's own order when the groups are the dependency levels of that order, a coloured order
// (few groups, the GPU-sized choice) when they are colours (the host layer builds both, host/AsmPartition.hpp).
//
// Data: the inverse of every diagonal block A[B_i,B_i], dense fp64 [m_i][m_i] row-major in HBM (125 x 125 for a
// 2x2x2 block of HEX27 elements = 125 KB).  The apply kernel is HBM-bound on those inverses: per block m^2 x 8 B of
// inverse + the block's CSR rows (12 B per non-zero) in, m x 8 B out.
//   schwarz_extract_kernel   A[B,B] -> dense (binary search of every column in the block's sorted dof list)
//   schwarz_invert_kernel    in-place Gauss-Jordan with partial pivoting, one CTA per block, pivot row / column staged
//                            in shared memory
//   schwarz_apply_kernel     t = (r - A y)[B] (warp per row), z = inv . t (warp per row, coalesced), y[B] += z
#include <algorithm>
#include "b2_common.cuh"
#include "b2_schwarz_levels.hpp"

struct b2_schwarz {
  b2_ctx* ctx = nullptr;
  b2_csr* A = nullptr;            // borrowed
  int64_t nblocks = 0, ngroups = 0, ndofs_total = 0, inv_total = 0;
  int max_m = 0;
  int64_t* blk_ptr = nullptr;     // [nblocks+1] device
  int32_t* blk_dofs = nullptr;    // [ndofs_total] device, every block's dofs sorted
  int64_t* inv_ptr = nullptr;     // [nblocks+1] device: start of every block's inverse
  double* inv = nullptr;          // [inv_total]
  int32_t* group_blocks = nullptr;   // [nblocks] device: blocks in schedule order
  std::vector<int64_t> group_ptr;    // [ngroups+1] host
  int* err = nullptr;             // device: 1 + first block whose pivot vanished, or 0
  int64_t n = 0;                  // A->nrows at creation (the borrowed operator may be gone when this object is destroyed)
  int sub = 0;                    // block solve: 0 = exact (dense inverse), 1 = one SSOR iteration on the block's rows, 2 = ILU(0)
  int64_t* frow = nullptr;        // [ndofs_total+1] start of every (block, row)'s row of factor values / local indices
  unsigned short* lidx = nullptr; // [fac_total] staged walk: position of every row entry's column in its block's dof list (b2_schwarz_walk.cuh)
  double* fac = nullptr;          // [fac_total]   ILU factors on the pattern of the blocks' rows of A
  int64_t* foff = nullptr;        // [n]           ILU: factor row of the dof in the block that claimed it
  int64_t fac_total = 0;
  double *tg = nullptr, *dg = nullptr, *zg = nullptr;   // [n] scratch of the SSOR sweep
  int32_t* mark = nullptr;        // [n]
  // optional dependency levels of every block's rows (b2_schwarz_set_row_levels): forward (lower pattern) / backward
  bool row_levels = false;
  int64_t *lvptr_f = nullptr, *lvptr_b = nullptr;       // [nblocks+1] into lvoff_*
  int32_t *lvoff_f = nullptr, *lvoff_b = nullptr;       // per block: nlev+1 offsets into its row list
  int32_t *lvrows_f = nullptr, *lvrows_b = nullptr;     // [ndofs_total] local rows of every block sorted by level
  int64_t lvoff_f_n = 0, lvoff_b_n = 0, max_levels = 0;
  bool ready = false;
};

namespace {

#include "b2_schwarz_kernels.cuh"
#include "b2_schwarz_walk.cuh"

// the row-walking block solves run staged (b2_schwarz_walk.cuh) unless the rows are level-scheduled or a block is too large
bool staged_walk(const b2_schwarz* s) { return s->sub != 0 && !s->row_levels && s->max_m <= kWalkMaxM; }

// rows of the blocks laid end to end: frow[k] = first slot of block row k (k over blk_dofs), frow[ndofs_total] = total
int build_frow(b2_schwarz* s) {
  if (s->frow) return 0;
  b2_ctx* c = s->ctx;
  const size_t n = (size_t)s->A->nrows;
  std::vector<int64_t> rp(n + 1);
  std::vector<int32_t> bd((size_t)s->ndofs_total);
  B2_TRY(b2_download(c, rp.data(), s->A->rowptr, n + 1));
  B2_TRY(b2_download(c, bd.data(), s->blk_dofs, (size_t)s->ndofs_total));
  std::vector<int64_t> frow((size_t)s->ndofs_total + 1);
  int64_t tot = 0;
  for (int64_t k = 0; k < s->ndofs_total; k++) { frow[k] = tot; tot += rp[bd[k] + 1] - rp[bd[k]]; }
  frow[(size_t)s->ndofs_total] = tot;
  s->fac_total = tot;
  B2_TRY(b2_malloc(c, &s->frow, (size_t)s->ndofs_total + 1));
  B2_TRY(b2_upload(c, s->frow, frow.data(), (size_t)s->ndofs_total + 1));
  return 0;
}
// local indices of the staged walk (pattern only: built once)
int build_lidx(b2_schwarz* s) {
  if (s->lidx) return 0;
  b2_ctx* c = s->ctx;
  B2_TRY(build_frow(s));
  B2_TRY(b2_malloc(c, &s->lidx, (size_t)s->fac_total));
  B2_LAUNCH(c, schwarz_lidx_kernel, b2_grid_for(c, s->nblocks, 1, 8), 256, 0, s->nblocks, s->blk_ptr, s->blk_dofs, s->frow, s->A->rowptr,
            s->A->col, s->lidx);
  return 0;
}
// grid / CTA size / shared memory of a staged-walk launch over one group
template <class K>
int walk_launch_shape(b2_schwarz* s, K kern, int64_t nblk, int* grid, int* threads, size_t* smem) {
  b2_ctx* c = s->ctx;
  *smem = walk_smem_bytes(s->max_m);
  // the attribute belongs to the function: always the fixed maximum, so that no other object's launch can lower it
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)walk_smem_bytes(kWalkMaxM)));
  // one warp sweeps a block; the other warps of the CTA only help with the block's residual.  Many blocks: the smallest
  // CTAs put the most sweeping warps on an SM (registers allow ~25 one-warp CTAs, 12 of two warps)
  *threads = nblk >= 16 * (int64_t)c->sm_count ? 32 : (nblk >= 8 * (int64_t)c->sm_count ? 64 : (nblk >= 4 * (int64_t)c->sm_count ? 128 : kApplyThreads));
  int per_sm = 1;
  B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, *threads, *smem));
  *grid = b2_grid_for(c, nblk, 1, per_sm < 1 ? 1 : per_sm);
  return 0;
}

}  // namespace

const b2_csr* b2_schwarz_operator(const b2_schwarz* s) { return s ? s->A : nullptr; }

// Row-walking block solves without row levels: ONE warp sweeps a block (the other warps of the CTA only help with the
// block's residual), so when a group has many blocks small CTAs put more sweeping warps on an SM (32 resident CTAs of
// 64 threads instead of 8 of 256)
static inline int walk_threads(const b2_ctx* c, int64_t nblocks_in_group) {
  return nblocks_in_group >= 8 * (int64_t)c->sm_count ? 64 : (nblocks_in_group >= 4 * (int64_t)c->sm_count ? 128 : kApplyThreads);
}

extern "C" {

int b2_schwarz_create(b2_ctx* c, b2_csr* A, int64_t nblocks, const int64_t* blk_ptr, const int32_t* blk_dofs, int64_t ngroups,
                      const int64_t* group_ptr, const int32_t* group_blocks, b2_schwarz** out) {
  B2_CHECK(c && A && out && blk_ptr && blk_dofs && group_ptr && group_blocks, "b2_schwarz_create: null argument");
  B2_CHECK(A->nrows == A->ncols, "b2_schwarz_create: operator must be square");
  B2_CHECK(nblocks >= 1 && ngroups >= 1 && ngroups <= nblocks, "b2_schwarz_create: %lld blocks in %lld groups", (long long)nblocks,
           (long long)ngroups);
  B2_CHECK(blk_ptr[0] == 0 && group_ptr[0] == 0 && group_ptr[ngroups] == nblocks, "b2_schwarz_create: offsets must start at 0 and the groups hold every block");
  int max_m = 0;
  std::vector<int64_t> inv_ptr((size_t)nblocks + 1, 0);
  for (int64_t b = 0; b < nblocks; b++) {
    const int64_t m = blk_ptr[b + 1] - blk_ptr[b];
    B2_CHECK(m >= 1, "b2_schwarz_create: block %lld is empty", (long long)b);
    for (int64_t k = blk_ptr[b]; k < blk_ptr[b + 1]; k++) {
      B2_CHECK(blk_dofs[k] >= 0 && blk_dofs[k] < A->nrows, "b2_schwarz_create: block %lld: dof %d outside the operator", (long long)b, (int)blk_dofs[k]);
      B2_CHECK(k == blk_ptr[b] || blk_dofs[k] > blk_dofs[k - 1], "b2_schwarz_create: block %lld: dofs must be sorted and distinct", (long long)b);
    }
    if (m > max_m) max_m = (int)m;
    inv_ptr[b + 1] = inv_ptr[b] + m * m;
  }
  std::vector<uint8_t> seen((size_t)nblocks, 0);
  for (int64_t g = 0; g < ngroups; g++) B2_CHECK(group_ptr[g + 1] > group_ptr[g], "b2_schwarz_create: group %lld is empty", (long long)g);
  for (int64_t q = 0; q < nblocks; q++) {
    B2_CHECK(group_blocks[q] >= 0 && group_blocks[q] < nblocks && !seen[group_blocks[q]], "b2_schwarz_create: the schedule must list every block once");
    seen[group_blocks[q]] = 1;
  }
  b2_schwarz* s = new b2_schwarz();
  s->ctx = c;
  s->A = A;
  s->n = A->nrows;
  s->nblocks = nblocks;
  s->ngroups = ngroups;
  s->ndofs_total = blk_ptr[nblocks];
  s->inv_total = inv_ptr[nblocks];
  s->max_m = max_m;
  s->group_ptr.assign(group_ptr, group_ptr + ngroups + 1);
  *out = s;
  const int rc = [&]() -> int {
    B2_TRY(b2_malloc(c, &s->blk_ptr, (size_t)nblocks + 1));
    B2_TRY(b2_malloc(c, &s->blk_dofs, (size_t)s->ndofs_total));
    B2_TRY(b2_malloc(c, &s->inv_ptr, (size_t)nblocks + 1));
    B2_TRY(b2_malloc(c, &s->group_blocks, (size_t)nblocks));
    B2_TRY(b2_malloc(c, &s->err, 1));
    B2_TRY(b2_upload(c, s->blk_ptr, blk_ptr, (size_t)nblocks + 1));
    B2_TRY(b2_upload(c, s->blk_dofs, blk_dofs, (size_t)s->ndofs_total));
    B2_TRY(b2_upload(c, s->inv_ptr, inv_ptr.data(), (size_t)nblocks + 1));
    B2_TRY(b2_upload(c, s->group_blocks, group_blocks, (size_t)nblocks));
    return 0;
  }();
  if (rc) {               // nothing half-built is handed out
    b2_schwarz_destroy(s);
    *out = nullptr;
  }
  return rc;
}

/* block solve: 0 = exact (MLU_PRECOND on the blocks; dense inverses, blocks of at most 4096 dofs), 1 = one SSOR
 * iteration (SOR_PRECOND on the blocks, 001_Poisson's own choice; no storage, any block size) */
int b2_schwarz_set_subsolver(b2_schwarz* s, int kind) {
  B2_CHECK(s && kind >= 0 && kind <= 2, "b2_schwarz_set_subsolver: kind must be 0 (exact), 1 (SSOR) or 2 (ILU(0))");
  if (kind != s->sub) s->ready = false;
  s->sub = kind;
  return 0;
}

namespace {
// dependency levels of the rows of every block in its lower / upper triangular in-block pattern, rows sorted by level
int build_row_levels(b2_schwarz* s) {
  b2_ctx* c = s->ctx;
  const size_t n = (size_t)s->A->nrows;
  std::vector<int64_t> rp(n + 1), bp((size_t)s->nblocks + 1);
  std::vector<int32_t> col((size_t)s->A->nnz), bd((size_t)s->ndofs_total);
  B2_TRY(b2_download(c, rp.data(), s->A->rowptr, n + 1));
  B2_TRY(b2_download(c, col.data(), s->A->col, (size_t)s->A->nnz));
  B2_TRY(b2_download(c, bp.data(), s->blk_ptr, (size_t)s->nblocks + 1));
  B2_TRY(b2_download(c, bd.data(), s->blk_dofs, (size_t)s->ndofs_total));
  std::vector<int64_t> ptr[2];
  std::vector<int32_t> off[2], rows[2];
  s->max_levels = b2_schwarz_row_level_schedule(s->nblocks, bp.data(), bd.data(), rp.data(), col.data(), ptr, off, rows);
  s->lvoff_f_n = (int64_t)off[0].size();
  s->lvoff_b_n = (int64_t)off[1].size();
  B2_TRY(b2_malloc(c, &s->lvptr_f, (size_t)s->nblocks + 1));
  B2_TRY(b2_malloc(c, &s->lvptr_b, (size_t)s->nblocks + 1));
  B2_TRY(b2_malloc(c, &s->lvoff_f, off[0].size()));
  B2_TRY(b2_malloc(c, &s->lvoff_b, off[1].size()));
  B2_TRY(b2_malloc(c, &s->lvrows_f, (size_t)s->ndofs_total));
  B2_TRY(b2_malloc(c, &s->lvrows_b, (size_t)s->ndofs_total));
  B2_TRY(b2_upload(c, s->lvptr_f, ptr[0].data(), (size_t)s->nblocks + 1));
  B2_TRY(b2_upload(c, s->lvptr_b, ptr[1].data(), (size_t)s->nblocks + 1));
  B2_TRY(b2_upload(c, s->lvoff_f, off[0].data(), off[0].size()));
  B2_TRY(b2_upload(c, s->lvoff_b, off[1].data(), off[1].size()));
  B2_TRY(b2_upload(c, s->lvrows_f, rows[0].data(), (size_t)s->ndofs_total));
  B2_TRY(b2_upload(c, s->lvrows_b, rows[1].data(), (size_t)s->ndofs_total));
  return 0;
}
}  // namespace

/* the SSOR and ILU(0) block solves walk a block's rows in order (one warp per block).  on != 0: the rows are sorted into
 * dependency levels of the block's triangular patterns at the next b2_schwarz_setup and every warp of the CTA takes rows
 * of a level -- the same arithmetic per row, the same result bit for bit; meant for large blocks (the reference's
 * applications use 8^4 elements per block, or FEMuS_DEFAULT = one block per level).  b2_schwarz_row_levels: the
 * longest dependency chain found (0 before the setup). */
int b2_schwarz_set_row_levels(b2_schwarz* s, int on) {
  B2_CHECK(s, "b2_schwarz_set_row_levels: null handle");
  if ((on != 0) != s->row_levels) s->ready = false;
  s->row_levels = on != 0;
  return 0;
}
int64_t b2_schwarz_row_levels(const b2_schwarz* s) { return s ? s->max_levels : 0; }

/* numeric phase: A[B_i,B_i] of the operator's CURRENT values (call it after the penalty rows are set), inverted */
int b2_schwarz_setup(b2_schwarz* s) {
  B2_CHECK(s, "b2_schwarz_setup: null handle");
  b2_ctx* c = s->ctx;
  if (s->sub != 0 && s->row_levels && !s->lvrows_f) B2_TRY(build_row_levels(s));
  if (s->sub == 1 && staged_walk(s)) {      // SSOR works on A's rows: only the local indices are needed
    B2_TRY(build_lidx(s));
    s->ready = true;
    return 0;
  }
  if (s->sub == 1) {              // plain / level-scheduled walk: scratch vectors
    const size_t n = (size_t)s->A->nrows;
    if (!s->tg) B2_TRY(b2_malloc(c, &s->tg, n));
    if (!s->dg) B2_TRY(b2_malloc(c, &s->dg, n));
    if (!s->zg) B2_TRY(b2_malloc(c, &s->zg, n));
    if (!s->mark) {
      B2_TRY(b2_malloc(c, &s->mark, n));
      B2_CUDA(cudaMemsetAsync(s->mark, 0xff, n * sizeof(int32_t), c->stream));     // -1: claimed by no block
    }
    s->ready = true;
    return 0;
  }
  if (s->sub == 2) {              // ILU(0) of every block on the pattern of its rows of A, group by group
    const size_t n = (size_t)s->A->nrows;
    B2_TRY(build_frow(s));
    if (!s->fac) B2_TRY(b2_malloc(c, &s->fac, (size_t)s->fac_total));
    if (staged_walk(s)) {
      B2_TRY(build_lidx(s));
      B2_CUDA(cudaMemsetAsync(s->err, 0, sizeof(int), c->stream));
      for (int64_t g = 0; g < s->ngroups; g++) {
        const int64_t g0 = s->group_ptr[g], g1 = s->group_ptr[g + 1];
        int grid = 1, threads = 64;
        size_t smem = 0;
        B2_TRY(walk_launch_shape(s, schwarz_walk_ilu_factor_kernel, g1 - g0, &grid, &threads, &smem));
        B2_LAUNCH(c, schwarz_walk_ilu_factor_kernel, grid, threads, smem, g0, g1, s->group_blocks, s->blk_ptr, s->blk_dofs, s->frow, s->lidx,
                  s->A->rowptr, s->A->val, s->fac, s->err, s->max_m);
      }
      int err = 0;
      B2_TRY(b2_download(c, &err, s->err, 1));
      B2_CHECK(err == 0, "b2_schwarz_setup: ILU(0) of block %d met a zero pivot", err - 1);
      s->ready = true;
      return 0;
    }
    if (!s->foff) B2_TRY(b2_malloc(c, &s->foff, n));
    if (!s->zg) B2_TRY(b2_malloc(c, &s->zg, n));
    if (!s->mark) {
      B2_TRY(b2_malloc(c, &s->mark, n));
      B2_CUDA(cudaMemsetAsync(s->mark, 0xff, n * sizeof(int32_t), c->stream));
    }
    B2_CUDA(cudaMemsetAsync(s->err, 0, sizeof(int), c->stream));
    for (int64_t g = 0; g < s->ngroups; g++) {
      const int64_t g0 = s->group_ptr[g], g1 = s->group_ptr[g + 1];
      if (s->row_levels)
        B2_LAUNCH(c, schwarz_ilu_factor_kernel<true>, b2_grid_for(c, g1 - g0, 1, 16), kApplyThreads, 0, g0, g1, s->group_blocks, s->blk_ptr,
                  s->blk_dofs, s->frow, s->A->rowptr, s->A->col, s->A->val, s->fac, s->mark, s->foff, s->err, s->lvptr_f, s->lvoff_f, s->lvrows_f);
      else
        B2_LAUNCH(c, schwarz_ilu_factor_kernel<false>, b2_grid_for(c, g1 - g0, 1, 2048 / walk_threads(c, g1 - g0)), walk_threads(c, g1 - g0), 0, g0, g1, s->group_blocks, s->blk_ptr,
                  s->blk_dofs, s->frow, s->A->rowptr, s->A->col, s->A->val, s->fac, s->mark, s->foff, s->err, nullptr, nullptr, nullptr);
    }
    int err = 0;
    B2_TRY(b2_download(c, &err, s->err, 1));
    B2_CHECK(err == 0, "b2_schwarz_setup: ILU(0) of block %d met a zero pivot", err - 1);
    s->ready = true;
    return 0;
  }
  B2_CHECK(s->max_m <= 4096, "b2_schwarz_setup: a block has %d dofs; exact block solves support at most 4096 (use the SSOR block solve)", s->max_m);
  const int smem = 2 * s->max_m * (int)sizeof(double);
  if (!s->inv) B2_TRY(b2_malloc(c, &s->inv, (size_t)s->inv_total));
  // shared memory of the two kernels: 2 x max_m doubles.  The attribute belongs to the FUNCTION, not to this object: it is
  // always the fixed maximum (max_m <= 4096 => 64 KB), so that no other object's setup can lower it under this one
  B2_CUDA(cudaFuncSetAttribute(schwarz_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * (int)sizeof(double) + 4096 * (int)sizeof(int)));
  B2_CUDA(cudaFuncSetAttribute(schwarz_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * (int)sizeof(double)));
  B2_CUDA(cudaMemsetAsync(s->err, 0, sizeof(int), c->stream));
  B2_LAUNCH(c, schwarz_extract_kernel, b2_grid_for(c, s->nblocks, 1, 8), 256, 0, s->nblocks, s->blk_ptr, s->blk_dofs, s->inv_ptr,
            s->A->rowptr, s->A->col, s->A->val, s->inv);
  B2_LAUNCH(c, schwarz_invert_kernel, b2_grid_for(c, s->nblocks, 1, 4), kInvertThreads, smem + s->max_m * (int)sizeof(int), s->nblocks, s->blk_ptr,
            s->inv_ptr, s->inv, s->max_m, s->err);
  int err = 0;
  B2_TRY(b2_download(c, &err, s->err, 1));
  B2_CHECK(err == 0, "b2_schwarz_setup: block %d is singular (zero pivot column)", err - 1);
  s->ready = true;
  return 0;
}

/* y = M^-1 r: the multiplicative sweep over the schedule's groups, one launch per group */
int b2_schwarz_apply(b2_schwarz* s, const b2_vec* r, b2_vec* y) {
  B2_CHECK(s && r && y, "b2_schwarz_apply: null argument");
  B2_CHECK(s->ready, "b2_schwarz_apply: call b2_schwarz_setup first");
  B2_CHECK(r->n >= s->A->nrows && y->n >= s->A->nrows && r->d != y->d, "b2_schwarz_apply: vectors too short or aliased");
  b2_ctx* c = s->ctx;
  const int smem = 2 * s->max_m * (int)sizeof(double);
  B2_CUDA(cudaMemsetAsync(y->d, 0, (size_t)s->A->nrows * sizeof(double), c->stream));
  for (int64_t g = 0; g < s->ngroups; g++) {
    const int64_t g0 = s->group_ptr[g], g1 = s->group_ptr[g + 1];
    if (staged_walk(s)) {
      int grid = 1, threads = 64;
      size_t wsm = 0;
      const walk_apply_kernel_t kern = s->sub == 2 ? schwarz_walk_apply_kernel_for<true>(s->A->max_row) : schwarz_walk_apply_kernel_for<false>(s->A->max_row);
      B2_TRY(walk_launch_shape(s, kern, g1 - g0, &grid, &threads, &wsm));
      B2_LAUNCH(c, kern, grid, threads, wsm, g0, g1, s->group_blocks, s->blk_ptr, s->blk_dofs, s->frow, s->lidx, s->A->rowptr, s->A->col, s->A->val,
                s->sub == 2 ? (const double*)s->fac : (const double*)nullptr, r->d, y->d, s->max_m);
      continue;
    }
    if (s->sub == 2) {
      if (s->row_levels)
        B2_LAUNCH(c, schwarz_apply_ilu_kernel<true>, b2_grid_for(c, g1 - g0, 1, 16), kApplyThreads, 0, g0, g1, s->group_blocks, s->blk_ptr,
                  s->blk_dofs, s->frow, s->A->rowptr, s->A->col, s->A->val, s->fac, r->d, y->d, s->zg, s->mark, s->lvptr_f, s->lvoff_f, s->lvrows_f,
                  s->lvptr_b, s->lvoff_b, s->lvrows_b);
      else
        B2_LAUNCH(c, schwarz_apply_ilu_kernel<false>, b2_grid_for(c, g1 - g0, 1, 2048 / walk_threads(c, g1 - g0)), walk_threads(c, g1 - g0), 0, g0, g1, s->group_blocks, s->blk_ptr,
                  s->blk_dofs, s->frow, s->A->rowptr, s->A->col, s->A->val, s->fac, r->d, y->d, s->zg, s->mark, nullptr, nullptr, nullptr, nullptr,
                  nullptr, nullptr);
      continue;
    }
    if (s->sub == 1) {
      if (s->row_levels)
        B2_LAUNCH(c, schwarz_apply_ssor_kernel<true>, b2_grid_for(c, g1 - g0, 1, 16), kApplyThreads, 0, g0, g1, s->group_blocks, s->blk_ptr,
                  s->blk_dofs, s->A->rowptr, s->A->col, s->A->val, r->d, y->d, s->tg, s->dg, s->zg, s->mark, s->lvptr_f, s->lvoff_f, s->lvrows_f,
                  s->lvptr_b, s->lvoff_b, s->lvrows_b);
      else
        B2_LAUNCH(c, schwarz_apply_ssor_kernel<false>, b2_grid_for(c, g1 - g0, 1, 2048 / walk_threads(c, g1 - g0)), walk_threads(c, g1 - g0), 0, g0, g1, s->group_blocks, s->blk_ptr,
                  s->blk_dofs, s->A->rowptr, s->A->col, s->A->val, r->d, y->d, s->tg, s->dg, s->zg, s->mark, nullptr, nullptr, nullptr, nullptr,
                  nullptr, nullptr);
      continue;
    }
    B2_LAUNCH(c, schwarz_apply_kernel, b2_grid_for(c, g1 - g0, 1, 16), kApplyThreads, smem, g0, g1, s->group_blocks, s->blk_ptr,
              s->blk_dofs, s->inv_ptr, s->inv, s->A->rowptr, s->A->col, s->A->val, r->d, y->d, s->max_m);
  }
  return 0;
}

int64_t b2_schwarz_bytes(const b2_schwarz* s) {
  if (!s) return 0;
  return ((s->inv ? s->inv_total : 0) + (s->fac ? s->fac_total : 0)) * (int64_t)sizeof(double) + (s->lidx ? s->fac_total : 0) * (int64_t)sizeof(unsigned short);
}
int64_t b2_schwarz_groups(const b2_schwarz* s) { return s ? s->ngroups : 0; }

int b2_schwarz_destroy(b2_schwarz* s) {
  if (!s) return 0;
  b2_ctx* c = s->ctx;
  b2_free(c, s->blk_ptr, (size_t)s->nblocks + 1);
  b2_free(c, s->blk_dofs, (size_t)s->ndofs_total);
  b2_free(c, s->inv_ptr, (size_t)s->nblocks + 1);
  b2_free(c, s->inv, (size_t)s->inv_total);
  b2_free(c, s->lvptr_f, (size_t)s->nblocks + 1);
  b2_free(c, s->lvptr_b, (size_t)s->nblocks + 1);
  b2_free(c, s->lvoff_f, (size_t)s->lvoff_f_n);
  b2_free(c, s->lvoff_b, (size_t)s->lvoff_b_n);
  b2_free(c, s->lvrows_f, (size_t)s->ndofs_total);
  b2_free(c, s->lvrows_b, (size_t)s->ndofs_total);
  b2_free(c, s->frow, (size_t)s->ndofs_total + 1);
  b2_free(c, s->lidx, (size_t)s->fac_total);
  b2_free(c, s->fac, (size_t)s->fac_total);
  b2_free(c, s->foff, (size_t)s->n);
  b2_free(c, s->tg, (size_t)s->n);
  b2_free(c, s->dg, (size_t)s->n);
  b2_free(c, s->zg, (size_t)s->n);
  b2_free(c, s->mark, (size_t)s->n);
  b2_free(c, s->group_blocks, (size_t)s->nblocks);
  b2_free(c, s->err, 1);
  delete s;
  return 0;
}

}  // extern "C"
