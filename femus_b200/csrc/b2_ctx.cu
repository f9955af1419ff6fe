// Context: device, stream, timers, reduction scratch, NCCL communicator (one rank per GPU).
// Replaces FemusInit/PetscInitialize + MPI_COMM_WORLD of the reference
// (src/00_utils/00_application_initialization/FemusInit.cpp:46-73).
#include "b2_common.cuh"
#include <cstdarg>
#include <dlfcn.h>

static thread_local char g_err[1024] = "";

void b2_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- NCCL through dlopen: the launcher (torch.distributed) usually has libnccl.so.2 loaded
// already, in which case dlopen hands back that very copy.
namespace {
struct NcclApi {
  void* h = nullptr;
  void* get_uid = nullptr;
  void* init_rank = nullptr;
  void* all_reduce = nullptr;
  void* destroy = nullptr;
  void* err_string = nullptr;
  void* send = nullptr;
  void* recv = nullptr;
  void* group_start = nullptr;
  void* group_end = nullptr;
} g_nccl;
struct Uid { char b[128]; };

int load_nccl() {
  if (g_nccl.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  B2_CHECK(g_nccl.h, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.get_uid = dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.init_rank = dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.all_reduce = dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.destroy = dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.err_string = dlsym(g_nccl.h, "ncclGetErrorString");
  g_nccl.send = dlsym(g_nccl.h, "ncclSend");
  g_nccl.recv = dlsym(g_nccl.h, "ncclRecv");
  g_nccl.group_start = dlsym(g_nccl.h, "ncclGroupStart");
  g_nccl.group_end = dlsym(g_nccl.h, "ncclGroupEnd");
  B2_CHECK(g_nccl.get_uid && g_nccl.init_rank && g_nccl.all_reduce && g_nccl.send && g_nccl.recv,
           "libnccl lacks expected symbols");
  return 0;
}
const char* nccl_err(int r) {
  if (!g_nccl.err_string) return "?";
  return ((const char* (*)(int))g_nccl.err_string)(r);
}
}  // namespace

extern "C" {

const char* b2_last_error(void) { return g_err; }
int b2_version(void) { return 100; }

int b2_ctx_create(int device, b2_ctx** out) {
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  B2_CHECK(e == cudaSuccess && ndev > 0, "no CUDA device visible (%s): femus_b200 has no CPU fallback",
           cudaGetErrorString(e));
  B2_CHECK(device >= 0 && device < ndev, "device %d out of range (%d visible)", device, ndev);
  B2_CUDA(cudaSetDevice(device));
  b2_ctx* c = new b2_ctx();
  c->device = device;
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  if (const char* v = getenv("B2_SPMV_VARIANT")) c->spmv_variant = atoi(v);
  if (const char* v = getenv("B2_ASM_VARIANT")) c->asm_variant = atoi(v);
  if (const char* v = getenv("B2_SF_WARPS")) c->asm_warps = atoi(v) == 16 ? 16 : 12;
  B2_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  B2_CUDA(cudaEventCreate(&c->ev0));
  B2_CUDA(cudaEventCreate(&c->ev1));
  B2_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  B2_CUDA(cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming));
  B2_CUDA(cudaEventCreateWithFlags(&c->ev_free, cudaEventDisableTiming));
  B2_CUDA(cudaEventCreateWithFlags(&c->ev_marked, cudaEventDisableTiming));
  B2_TRY(b2_malloc(c, &c->red_partial, (size_t)kRedBlocks * 2));
  B2_TRY(b2_malloc(c, &c->red_result, 8));
  B2_TRY(b2_malloc(c, &c->red_counter, 1));
  B2_CUDA(cudaMemsetAsync(c->red_counter, 0, sizeof(unsigned int), c->stream));
  B2_CUDA(cudaMallocHost((void**)&c->h_result, 8 * sizeof(double)));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  *out = c;
  return 0;
}

int b2_ctx_destroy(b2_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->peer_base) {
    for (int r = 0; r < c->nranks; r++)
      if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
    delete[] c->peer_base;
    b2_free(c, c->d_peer_base, (size_t)c->nranks);
    b2_free(c, c->peer_err, 1);
  }
  if (c->peer_local) cudaFree(c->peer_local);
  if (c->nccl_comm && g_nccl.destroy) ((int (*)(void*))g_nccl.destroy)(c->nccl_comm);
  b2_free(c, c->red_partial, (size_t)kRedBlocks * 2);
  b2_free(c, c->red_result, 8);
  b2_free(c, c->red_counter, 1);
  if (c->flush_buf) cudaFree(c->flush_buf);
  cudaFreeHost(c->h_result);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ev_copied) cudaEventDestroy(c->ev_copied);
  if (c->ev_free) cudaEventDestroy(c->ev_free);
  if (c->ev_marked) cudaEventDestroy(c->ev_marked);
  delete c;
  return 0;
}

/* Peer-memory exchange, step 1: allocate this rank's inbox (double buffered, `slot_cells` 16-byte cells per sender: 8
 * for scalars + the longest message) and export its CUDA IPC handle (64 bytes) for the launcher to all-gather. */
int b2_ctx_peer_export(b2_ctx* c, int64_t slot_cells, void* handle64) {
  const int64_t slot_doubles = slot_cells;
  B2_CHECK(c && handle64 && slot_cells >= 8, "b2_ctx_peer_export: bad arguments");
  B2_CHECK(c->nranks > 1, "b2_ctx_peer_export: communicator not initialised (b2_ctx_comm_init first)");
  B2_CHECK(!c->peer_local, "b2_ctx_peer_export: already exported");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  const size_t bytes = (size_t)2 * c->nranks * (size_t)slot_doubles * 16;
  B2_CUDA(cudaMalloc(&c->peer_local, bytes));
  c->bytes += (int64_t)bytes;
  B2_CUDA(cudaMemset(c->peer_local, 0, bytes));
  c->peer_slot = slot_doubles;
  cudaIpcMemHandle_t h;
  B2_CUDA(cudaIpcGetMemHandle(&h, c->peer_local));
  memcpy(handle64, &h, 64);
  return 0;
}
/* step 2: open the blocks of all ranks (handles[nranks][64], gathered in rank order) */
int b2_ctx_peer_open(b2_ctx* c, const void* handles) {
  B2_CHECK(c && handles && c->peer_local && !c->peer_base, "b2_ctx_peer_open: export first, open once");
  c->peer_base = new void*[c->nranks];
  for (int r = 0; r < c->nranks; r++) {
    if (r == c->rank) { c->peer_base[r] = c->peer_local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * 64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&c->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess);
    B2_CHECK(e == cudaSuccess, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
  }
  B2_TRY(b2_malloc(c, &c->d_peer_base, (size_t)c->nranks));
  B2_TRY(b2_upload(c, c->d_peer_base, c->peer_base, (size_t)c->nranks));
  B2_TRY(b2_malloc(c, &c->peer_err, 1));
  B2_CUDA(cudaMemsetAsync(c->peer_err, 0, sizeof(int), c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
/* 1 when a wait of the peer-memory exchange timed out since the last call (a rank stopped taking part) */
int b2_ctx_peer_error(b2_ctx* c, int* err) {
  *err = 0;
  if (!c->peer_err) return 0;
  B2_TRY(b2_download(c, err, c->peer_err, 1));
  return 0;
}

int b2_ctx_sync(b2_ctx* c) {
  B2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
void* b2_ctx_stream(b2_ctx* c) { return (void*)c->stream; }
int b2_ctx_device(b2_ctx* c) { return c->device; }
int b2_ctx_nranks(b2_ctx* c) { return c->nranks; }
int b2_ctx_rank(b2_ctx* c) { return c->rank; }
int64_t b2_ctx_bytes_in_use(b2_ctx* c) { return c->bytes; }
int64_t b2_ctx_launch_count(b2_ctx* c, int reset) {
  int64_t n = c->launches;
  if (reset) c->launches = 0;
  return n;
}

int b2_timer_start(b2_ctx* c) {
  B2_CUDA(cudaEventRecord(c->ev0, c->stream));
  return 0;
}
int b2_timer_stop_ms(b2_ctx* c, double* ms) {
  B2_CUDA(cudaEventRecord(c->ev1, c->stream));
  B2_CUDA(cudaEventSynchronize(c->ev1));
  float f = 0.f;
  B2_CUDA(cudaEventElapsedTime(&f, c->ev0, c->ev1));
  *ms = (double)f;
  return 0;
}

int b2_ctx_flush_l2(b2_ctx* c) {
  if (!c->flush_buf) {
    c->flush_bytes = (size_t)256 << 20;   // 256 MiB > 126 MB L2
    B2_CUDA(cudaMalloc(&c->flush_buf, c->flush_bytes));
  }
  B2_CUDA(cudaMemsetAsync(c->flush_buf, 0, c->flush_bytes, c->stream));
  return 0;
}

// Prefetch protocol (double-buffered inputs): b2_ctx_open_copies makes the copy stream wait for
// everything enqueued so far on the compute stream (the buffers about to be overwritten may still be
// read by it); the b2_*_prefetch calls then copy on the copy stream; b2_ctx_join_copies makes the
// compute stream wait for those copies.
int b2_ctx_open_copies(b2_ctx* c) {
  B2_CUDA(cudaEventRecord(c->ev_free, c->stream));
  B2_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_free, 0));
  return 0;
}
int b2_ctx_join_copies(b2_ctx* c) {      // the compute stream waits for EVERYTHING issued on the copy stream so far
  B2_CUDA(cudaEventRecord(c->ev_copied, c->copy_stream));
  B2_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied, 0));
  return 0;
}
// mark the end of a batch of prefetches; b2_ctx_wait_marked makes the compute stream wait for that
// batch only (later copies, e.g. a result download of the previous step, keep overlapping)
int b2_ctx_mark_copies(b2_ctx* c) {
  B2_CUDA(cudaEventRecord(c->ev_marked, c->copy_stream));
  return 0;
}
int b2_ctx_wait_marked(b2_ctx* c) {
  B2_CUDA(cudaStreamWaitEvent(c->stream, c->ev_marked, 0));
  return 0;
}

int b2_ctx_set_option(b2_ctx* c, const char* name, int value) {
  if (!strcmp(name, "spmv_variant")) {
    B2_CHECK(value >= 0 && value <= 2, "spmv_variant %d (0, 1, 2)", value);
    c->spmv_variant = value;
    return 0;
  }
  if (!strcmp(name, "coarse_persistent")) {      // 1: coarse PCG loop in one cooperative kernel (default), 0: host-driven loop
    c->coarse_persistent = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(name, "halo_peer")) {      // 1: interface sums through peer memory (default once set up), 0: packed ncclAllReduce
    c->halo_peer = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(name, "asm_warps")) {
    B2_CHECK(value == 12 || value == 16, "asm_warps %d (12, 16)", value);
    c->asm_warps = value;
    return 0;
  }
  if (!strcmp(name, "asm_variant")) {
    B2_CHECK(value >= 0 && value <= 3, "asm_variant %d (0, 1, 2, 3)", value);
    c->asm_variant = value;
    return 0;
  }
  if (!strcmp(name, "spmv_timing")) {
    c->spmv_timing = value;
    return 0;
  }
  B2_CHECK(false, "b2_ctx_set_option: unknown option '%s'", name);
}

int b2_ctx_profile(b2_ctx* c, int on) {
  c->profiling = on != 0;
  c->prof_only = nullptr;
  return 0;
}
int b2_ctx_profile_only(b2_ctx* c, const void* handle) {
  c->profiling = handle != nullptr;
  c->prof_only = handle;
  return 0;
}
// total device time and launch count of the profiled launches tagged with `handle`
// (a b2_csr* or b2_asm*); the records are consumed.
int b2_ctx_profile_read(b2_ctx* c, const void* handle, int* count, double* total_ms) {
  B2_CUDA(cudaStreamSynchronize(c->stream));
  int n = 0;
  double t = 0.;
  std::vector<b2_ctx::ProfRec> keep;
  for (auto& r : c->prof) {
    if (r.tag == handle) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.e0, r.e1);
      t += ms;
      n++;
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    } else {
      keep.push_back(r);
    }
  }
  c->prof.swap(keep);
  *count = n;
  *total_ms = t;
  return 0;
}
int b2_ctx_profile_clear(b2_ctx* c) {
  B2_CUDA(cudaStreamSynchronize(c->stream));
  for (auto& r : c->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  c->prof.clear();
  return 0;
}

int b2_nccl_unique_id(void* id128) {
  B2_TRY(load_nccl());
  int r = ((int (*)(void*))g_nccl.get_uid)(id128);
  B2_CHECK(r == 0, "ncclGetUniqueId: %s", nccl_err(r));
  return 0;
}

int b2_ctx_comm_init(b2_ctx* c, int nranks, int rank, const void* id128) {
  B2_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d / %d", rank, nranks);
  c->nranks = nranks;
  c->rank = rank;
  if (nranks == 1) return 0;
  B2_TRY(load_nccl());
  B2_CUDA(cudaSetDevice(c->device));
  Uid uid;
  memcpy(uid.b, id128, 128);
  int r = ((int (*)(void**, int, Uid, int))g_nccl.init_rank)(&c->nccl_comm, nranks, uid, rank);
  B2_CHECK(r == 0, "ncclCommInitRank: %s", nccl_err(r));
  return 0;
}

}  // extern "C"

// sum-allreduce from a send buffer into a receive buffer
int b2_allreduce_into(b2_ctx* c, const double* d_send, double* d_recv, int64_t n) {
  if (n == 0) return 0;
  B2_CHECK(c->nranks > 1 && c->nccl_comm, "communicator not initialised");
  typedef int (*ar_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  int r = ((ar_t)g_nccl.all_reduce)(d_send, d_recv, (size_t)n, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->nccl_comm, c->stream);
  B2_CHECK(r == 0, "ncclAllReduce: %s", nccl_err(r));
  c->launches++;
  return 0;
}
// max (op 2) / min (op 3) allreduce in place
int b2_allreduce_op(b2_ctx* c, double* d_buf, int64_t n, int op) {
  if (c->nranks == 1 || n == 0) return 0;
  B2_CHECK(c->nccl_comm, "communicator not initialised");
  typedef int (*ar_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  int r = ((ar_t)g_nccl.all_reduce)(d_buf, d_buf, (size_t)n, /*ncclDouble*/ 8, op, c->nccl_comm, c->stream);
  B2_CHECK(r == 0, "ncclAllReduce: %s", nccl_err(r));
  return 0;
}

// sum-allreduce of n doubles in place on the library stream (ncclDouble = 8, ncclSum = 0)
int b2_allreduce_sum(b2_ctx* c, double* d_buf, int64_t n) {
  if (c->nranks == 1 || n == 0) return 0;
  B2_CHECK(c->nccl_comm, "communicator not initialised");
  typedef int (*ar_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  int r = ((ar_t)g_nccl.all_reduce)(d_buf, d_buf, (size_t)n, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->nccl_comm, c->stream);
  B2_CHECK(r == 0, "ncclAllReduce: %s", nccl_err(r));
  return 0;
}
