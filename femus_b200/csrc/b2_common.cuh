// Internal declarations shared by the femus_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/femus_b200.h"

void b2_set_error(const char* fmt, ...);

#define B2_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      b2_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

#define B2_CHECK(cond, ...)              \
  do {                                   \
    if (!(cond)) {                       \
      b2_set_error(__VA_ARGS__);         \
      return 1;                          \
    }                                    \
  } while (0)

#define B2_TRY(call)        \
  do {                      \
    int s__ = (call);       \
    if (s__) return s__;    \
  } while (0)

// every kernel launch of the library goes through this macro: counts the launch and checks it
// (the CPU emulator build of the tests predefines it, tests/cpp/emu_prefix.hpp)
#ifndef B2_LAUNCH
#define B2_LAUNCH(ctx, kernel, grid, block, smem, ...)                         \
  do {                                                                         \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);           \
    (ctx)->launches++;                                                         \
    B2_CUDA(cudaGetLastError());                                               \
  } while (0)
#endif

struct b2_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // second stream for host->device prefetches that overlap the compute of the previous step
  // (b2_mesh_prefetch / b2_vec_prefetch / b2_ctx_join_copies)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied = nullptr, ev_free = nullptr, ev_marked = nullptr;
  int64_t bytes = 0;
  int64_t launches = 0;
  // scratch for reductions: partial sums + result (device) and a pinned host mirror
  double* red_partial = nullptr;    // [kRedBlocks * 2]
  double* red_result = nullptr;     // [8]
  unsigned int* red_counter = nullptr;
  double* h_result = nullptr;       // pinned [8]
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // SpMV kernel: 0 = register-streaming, 1 = TMA-staged ring (default), 2 = staged + software-pipelined
  // gathers (B2_SPMV_VARIANT)
  int spmv_variant = 1;
  int spmv_timing = 0;
  // assembly kernel for triquadratic elements: 3 = sum factorisation (default; falls back to 1 when the tables are not
  // tensor products), 1 = FP64 tensor cores (mma.sync m8n8k4), 0 = CUDA-core tiles, 2 = table-driven kernel
  int asm_variant = 3;
  int asm_warps = 12;      // warps per CTA of the sum-factorised kernel: 12 or 16
  // multi-GPU
  int nranks = 1, rank = 0;
  void* nccl_comm = nullptr;
  // peer-memory exchange over NVLink / NVSwitch (b2_halo.cu): one inbox per rank, opened by every other rank of the node
  // through CUDA IPC.  Layout: 2 (parity of the exchange number) x nranks (sender) x peer_slot cells of 16 bytes
  // {low word, flag, high word, flag}; peer_base[r] = rank r's inbox as seen from this process (own: local pointer).
  void* peer_local = nullptr;
  void** peer_base = nullptr;           // host array [nranks]
  void** d_peer_base = nullptr;         // the same on the device
  int64_t peer_slot = 0;                // cells per (parity, sender): 8 scalars + interface values
  unsigned long long peer_epoch = 0;    // exchanges done so far (identical on every rank: SPMD call sequence)
  int* peer_err = nullptr;              // device: set when a wait timed out
  int coarse_persistent = 1;            // 1: iteration loop of the coarse PCG as one cooperative kernel (b2_cg.cu), 0: host-driven loop
  int halo_peer = 1;                    // 1: interface sums through the peer-memory exchange when it is set up; 0: packed ncclAllReduce
  // optional per-launch event timing of the SpMV family / assembly (b2_ctx_profile)
  bool profiling = false;
  const void* prof_only = nullptr;   // when set, only launches tagged with this handle are timed
  struct ProfRec { const void* tag; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof;
};

// brackets one launch with events when profiling is on
struct b2_prof_scope {
  b2_ctx* c; const void* tag; cudaEvent_t e0 = nullptr, e1 = nullptr;
  b2_prof_scope(b2_ctx* c_, const void* tag_) : c(c_), tag(tag_) {
    if (c->profiling && (!c->prof_only || c->prof_only == tag)) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, c->stream); }
  }
  ~b2_prof_scope() {
    if (e0) { cudaEventRecord(e1, c->stream); c->prof.push_back({tag, e0, e1}); }
  }
};

static constexpr int kRedBlocks = 1184;   // 148 SMs x 8

struct b2_halo;
struct b2_vec {
  b2_ctx* ctx;
  int64_t n;
  double* d;
  const b2_halo* halo;   // distributed layout (owned mask for reductions) or null
};

// Layout of a rank-local vector whose entries on the partition interface are shared with other
// ranks (every rank holds ALL dofs of its own elements): interface sum through one packed
// ncclAllReduce, ownership mask (lowest rank owns, Mesh.cpp:530-553) for reductions.
struct b2_halo {
  b2_ctx* ctx;
  int64_t n_local, n_if, n_packed;
  int32_t* idx;        // [n_if] local dof of interface entry k
  int32_t* pos;        // [n_if] its position in the packed interface vector (same on every rank)
  double* send;        // [n_packed] zero except at own positions
  double* recv;        // [n_packed]
  uint8_t* owned;      // [n_local] 1 if this rank owns the dof
  double* invmult;     // [n_local] 1 / (number of ranks holding the dof)
  int64_t n_owned;
  // peer-memory exchange (b2_halo_set_exchange): for every interface entry its holders in ascending rank order (the sum
  // is taken in that order on every holder), where each holder's value arrives in this rank's inbox and where this
  // rank's value goes in the holder's
  int64_t* hold_ptr = nullptr;     // [n_if+1]
  int32_t* hold_rank = nullptr;    // [hold_ptr[n_if]] holder ranks, ascending, this rank included
  int32_t* hold_pos = nullptr;     // cell of the holder's value in its message to this rank (unused for this rank)
  int32_t* hold_spos = nullptr;    // cell of this rank's value in its message to the holder
  int64_t n_hold = 0;
};

struct b2_csr {
  b2_ctx* ctx;
  int64_t nrows, ncols, nnz;
  int64_t* rowptr;   // [nrows+1]
  int32_t* col;      // [nnz]
  double* val;       // [nnz]
  int32_t* diag_pos; // [nrows] position of the diagonal inside each row (built by the first b2_csr_diag), or null
  // SpMV plan (b2_spmv.cu), null: streaming kernel
  int32_t* chunk_row;        // [nchunks+1] row cuts
  int64_t* cdesc;            // [nchunks+1][4] packed {first nonzero, dictionary start, first row, 0} per chunk
  int64_t* dict_ptr;         // [nchunks+1] start of every chunk's column dictionary (multiples of 4)
  int32_t* dict;             // [dict_total] distinct columns of every chunk, sorted
  unsigned short* lidx;      // [nnz] position of each nonzero's column in its chunk's dictionary
  int64_t nchunks, dict_total, dict_cap;   // dict_cap: largest chunk dictionary (sizes the shared memory)
  int tpr;           // threads per row of the streaming SpMV kernel (power of two <= 32)
  int max_row;       // longest row
  double last_ms;
  // bumped by every C-ABI entry point that rewrites values of an EXISTING matrix in place (zero, zero rows / columns,
  // set rows, add blocks, put values, copy): consumers that cache something derived from the values (the explicit
  // restriction R = P^T of b2_mg) compare it.  Zero-initialised by b2_csr_alloc (value-initialised struct).
  uint64_t version;
};

template <class T>
int b2_malloc(b2_ctx* c, T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  B2_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
  c->bytes += (int64_t)(count * sizeof(T));
  return 0;
}
template <class T>
void b2_free(b2_ctx* c, T* p, size_t count) {
  if (!p) return;
  cudaFree((void*)p);
  if (count == 0) count = 1;
  c->bytes -= (int64_t)(count * sizeof(T));
}
template <class T>
int b2_upload(b2_ctx* c, T* dst, const T* src, size_t count) {
  if (count) B2_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
template <class T>
int b2_download(b2_ctx* c, T* dst, const T* src, size_t count) {
  if (count) B2_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

static inline int b2_grid_for(b2_ctx* c, int64_t work_items, int per_block, int blocks_per_sm) {
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)c->sm_count * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// internal cross-TU helpers
const b2_csr* b2_schwarz_operator(const b2_schwarz* s);
void b2_mesh_view(const b2_mesh* m, b2_ctx** ctx, int64_t* nnode, int64_t* nel, const double** xyz, const int32_t** conn);
int b2_csr_alloc(b2_ctx* c, int64_t nrows, int64_t ncols, int64_t nnz, b2_csr** out);
int b2_csr_finalize(b2_csr* A);   // row statistics -> tpr / max_row, SpMV row chunks
int b2_csr_build_chunks(b2_csr* A);
void b2_csr_free_plan(b2_csr* A);
int b2_csr_resid_w(const b2_csr* A, const double* b, const double* w, const double* x, double* r);   // r = w.*b - A x
int b2_csr_zero_cols_notowned(b2_csr* A, const uint8_t* d_owned);
int b2_csr_zero_rows_dev(b2_csr* A, const int32_t* d_rows, int64_t n, double diag, const uint8_t* d_owned);
// result stays on device; owned != null restricts the sum to entries with owned[i] != 0
int b2_dev_dot(b2_ctx* c, const double* x, const double* y, int64_t n, double* d_out, const uint8_t* owned = nullptr);
int b2_halo_sum_scalars(b2_halo* h, b2_vec* v, double* d_scal, int nscal);   // interface sum + scalars, one collective
int b2_allreduce_sum(b2_ctx* c, double* d_buf, int64_t n);
int b2_allreduce_op(b2_ctx* c, double* d_buf, int64_t n, int op);   // 2 = max, 3 = min
