// Streaming vector kernels (HBM-bound): replaces the VecXxx calls behind PetscVector
// (reference src/03_algebra/00_vectors/PetscVector.cpp).  All fp64, 128-bit accesses where the
// alignment allows, grid sized as a multiple of the SM count.
#include "b2_common.cuh"

namespace {

constexpr int kBlock = 256;

enum VecOp { OP_FILL, OP_AXPY, OP_AYPX, OP_SCALE, OP_SHIFT, OP_PMULT, OP_COPY_MASKED, OP_ABS };

template <int OP>
__device__ __forceinline__ double vec_apply(double y, double x, double z, double a) {
  if (OP == OP_FILL) return a;
  if (OP == OP_AXPY) return fma(a, x, y);
  if (OP == OP_AYPX) return fma(a, y, x);
  if (OP == OP_SCALE) return a * y;
  if (OP == OP_SHIFT) return y + a;
  if (OP == OP_ABS) return fabs(y);
  if (OP == OP_PMULT) return x * z;
  if (OP == OP_COPY_MASKED) return z > a ? x : 0.0;
  return y;
}

// y[i] = f(y[i], x[i], z[i], a); two doubles per access (all buffers come from cudaMalloc, so
// 16-byte aligned), tail handled by the last thread.
template <int OP>
__global__ void __launch_bounds__(kBlock) vec_map_kernel(double* __restrict__ y, const double* __restrict__ x,
                                                         const double* __restrict__ z, double a, int64_t n) {
  const int64_t n2 = n >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 yv = make_double2(0., 0.), xv = make_double2(0., 0.), zv = make_double2(0., 0.);
    if (OP != OP_FILL && OP != OP_PMULT && OP != OP_COPY_MASKED) yv = reinterpret_cast<const double2*>(y)[i];
    if (OP == OP_AXPY || OP == OP_AYPX || OP == OP_PMULT || OP == OP_COPY_MASKED)
      xv = reinterpret_cast<const double2*>(x)[i];
    if (OP == OP_PMULT || OP == OP_COPY_MASKED) zv = reinterpret_cast<const double2*>(z)[i];
    yv.x = vec_apply<OP>(yv.x, xv.x, zv.x, a);
    yv.y = vec_apply<OP>(yv.y, xv.y, zv.y, a);
    reinterpret_cast<double2*>(y)[i] = yv;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t i = n - 1;
    double yv = (OP != OP_FILL && OP != OP_PMULT && OP != OP_COPY_MASKED) ? y[i] : 0.;
    double xv = (OP == OP_AXPY || OP == OP_AYPX || OP == OP_PMULT || OP == OP_COPY_MASKED) ? x[i] : 0.;
    double zv = (OP == OP_PMULT || OP == OP_COPY_MASKED) ? z[i] : 0.;
    y[i] = vec_apply<OP>(yv, xv, zv, a);
  }
}

enum RedOp { RED_DOT, RED_SUM, RED_L1, RED_LINF, RED_MINMAX };

template <int OP>
__device__ __forceinline__ void red_combine(double& a0, double& a1, double b0, double b1) {
  if (OP == RED_LINF) a0 = fmax(a0, b0);
  else if (OP == RED_MINMAX) { a0 = fmin(a0, b0); a1 = fmax(a1, b1); }
  else a0 += b0;
}

// Deterministic two-level reduction in one launch: every block writes its partial, the last
// block to finish (ticket counter) folds the partials in index order.
template <int OP>
__global__ void __launch_bounds__(kBlock) vec_reduce_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                            int64_t n, double* __restrict__ partial,
                                                            double* __restrict__ result, unsigned int* counter,
                                                            const uint8_t* __restrict__ owned) {
  double a0 = (OP == RED_MINMAX) ? INFINITY : 0.0, a1 = -INFINITY;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (owned && !owned[i]) continue;       // entries owned by another rank are counted there
    const double v = x[i];
    if (OP == RED_DOT) a0 = fma(v, y[i], a0);
    else if (OP == RED_SUM) a0 += v;
    else if (OP == RED_L1) a0 += fabs(v);
    else if (OP == RED_LINF) a0 = fmax(a0, fabs(v));
    else { a0 = fmin(a0, v); a1 = fmax(a1, v); }
  }
  __shared__ double s0[kBlock / 32], s1[kBlock / 32];
  __shared__ bool last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double b0 = __shfl_down_sync(0xffffffffu, a0, o);
    double b1 = __shfl_down_sync(0xffffffffu, a1, o);
    red_combine<OP>(a0, a1, b0, b1);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s0[w] = a0; s1[w] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kBlock / 32; k++) red_combine<OP>(a0, a1, s0[k], s1[k]);
    partial[2 * blockIdx.x] = a0;
    partial[2 * blockIdx.x + 1] = a1;
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    // fold partials in index order with one warp (fixed association => reproducible)
    if (w == 0) {
      double r0 = (OP == RED_MINMAX) ? INFINITY : 0.0, r1 = -INFINITY;
      for (int k = l; k < (int)gridDim.x; k += 32) {
        red_combine<OP>(r0, r1, __ldcg(&partial[2 * k]), __ldcg(&partial[2 * k + 1]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double b0 = __shfl_down_sync(0xffffffffu, r0, o);
        double b1 = __shfl_down_sync(0xffffffffu, r1, o);
        red_combine<OP>(r0, r1, b0, b1);
      }
      if (l == 0) {
        result[0] = r0;
        if (OP == RED_MINMAX) result[1] = r1;
        *counter = 0;
      }
    }
  }
}

__global__ void vec_scatter_kernel(double* __restrict__ v, const int32_t* __restrict__ idx,
                                   const double* __restrict__ vals, int64_t n, int mode, double a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t k = idx[i];
    if (mode == 0) v[k] = vals[i];
    else if (mode == 1) atomicAdd(&v[k], vals[i]);
    else v[k] = a;
  }
}
__global__ void vec_gather_kernel(const double* __restrict__ v, const int32_t* __restrict__ idx,
                                  double* __restrict__ vals, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) vals[i] = v[idx[i]];
}

template <int OP>
int launch_map(b2_vec* y, const b2_vec* x, const b2_vec* z, double a) {
  b2_ctx* c = y->ctx;
  if (y->n == 0) return 0;
  const int grid = b2_grid_for(c, (y->n + 1) / 2, kBlock, 8);
  B2_LAUNCH(c, vec_map_kernel<OP>, grid, kBlock, 0, y->d, x ? x->d : nullptr, z ? z->d : nullptr, a, y->n);
  return 0;
}

template <int OP>
int launch_reduce(b2_ctx* c, const double* x, const double* y, int64_t n, double* d_result,
                  const uint8_t* owned = nullptr) {
  int grid = b2_grid_for(c, n, kBlock * 4, 8);
  if (grid > kRedBlocks) grid = kRedBlocks;
  B2_LAUNCH(c, vec_reduce_kernel<OP>, grid, kBlock, 0, x, y, n, c->red_partial, d_result, c->red_counter, owned);
  return 0;
}
inline const uint8_t* owned_of(const b2_vec* v) { return v->halo ? v->halo->owned : nullptr; }

int fetch_result(b2_ctx* c, int count) {
  B2_CUDA(cudaMemcpyAsync(c->h_result, c->red_result, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

}  // namespace

int b2_dev_dot(b2_ctx* c, const double* x, const double* y, int64_t n, double* d_out, const uint8_t* owned) {
  return launch_reduce<RED_DOT>(c, x, y, n, d_out, owned);
}

extern "C" {

int b2_vec_create(b2_ctx* c, int64_t n, b2_vec** out) {
  *out = nullptr;
  B2_CHECK(c && n >= 0, "b2_vec_create: bad arguments");
  b2_vec* v = new b2_vec{c, n, nullptr, nullptr};
  B2_TRY(b2_malloc(c, &v->d, (size_t)n + 4));
  B2_CUDA(cudaMemsetAsync(v->d, 0, ((size_t)n + 4) * sizeof(double), c->stream));
  *out = v;
  return 0;
}
int b2_vec_destroy(b2_vec* v) {
  if (!v) return 0;
  cudaStreamSynchronize(v->ctx->stream);
  b2_free(v->ctx, v->d, (size_t)v->n + 4);
  delete v;
  return 0;
}
int64_t b2_vec_size(const b2_vec* v) { return v->n; }
void* b2_vec_device_ptr(b2_vec* v) { return v->d; }

int b2_vec_zero(b2_vec* v) {
  B2_CUDA(cudaMemsetAsync(v->d, 0, (size_t)v->n * sizeof(double), v->ctx->stream));
  return 0;
}
int b2_vec_fill(b2_vec* v, double a) { return launch_map<OP_FILL>(v, nullptr, nullptr, a); }
int b2_vec_put(b2_vec* v, const double* host, int64_t n) {
  B2_CHECK(n <= v->n, "b2_vec_put: %lld > size %lld", (long long)n, (long long)v->n);
  return b2_upload(v->ctx, v->d, host, (size_t)n);
}
int b2_vec_get(const b2_vec* v, double* host, int64_t n) {
  B2_CHECK(n <= v->n, "b2_vec_get: %lld > size %lld", (long long)n, (long long)v->n);
  return b2_download(v->ctx, host, v->d, (size_t)n);
}
int b2_vec_put_async(b2_vec* v, const double* host, int64_t n) {
  B2_CHECK(n <= v->n, "b2_vec_put_async: too long");
  B2_CUDA(cudaMemcpyAsync(v->d, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, v->ctx->stream));
  return 0;
}
// host -> device on the copy stream (see b2_ctx_open_copies / b2_ctx_join_copies)
int b2_vec_prefetch(b2_vec* v, const double* host, int64_t n) {
  B2_CHECK(n <= v->n, "b2_vec_prefetch: too long");
  B2_CUDA(cudaMemcpyAsync(v->d, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, v->ctx->copy_stream));
  return 0;
}
// device -> host on the copy stream, ordered after everything enqueued so far on the compute stream
// (the result of the step just issued); b2_ctx_join_copies / b2_ctx_sync make it visible to the host
int b2_vec_fetch(const b2_vec* v, double* host, int64_t n) {
  B2_CHECK(n <= v->n, "b2_vec_fetch: too long");
  b2_ctx* c = v->ctx;
  B2_CUDA(cudaEventRecord(c->ev_free, c->stream));
  B2_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_free, 0));
  B2_CUDA(cudaMemcpyAsync(host, v->d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
  return 0;
}
int b2_vec_get_async(const b2_vec* v, double* host, int64_t n) {
  B2_CHECK(n <= v->n, "b2_vec_get_async: too long");
  B2_CUDA(cudaMemcpyAsync(host, v->d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, v->ctx->stream));
  return 0;
}
int b2_vec_copy(b2_vec* dst, const b2_vec* src) {
  B2_CHECK(dst->n == src->n, "b2_vec_copy: size mismatch");
  B2_CUDA(cudaMemcpyAsync(dst->d, src->d, (size_t)src->n * sizeof(double), cudaMemcpyDeviceToDevice, dst->ctx->stream));
  return 0;
}
int b2_vec_axpy(b2_vec* y, double a, const b2_vec* x) {
  B2_CHECK(y->n == x->n, "b2_vec_axpy: size mismatch");
  return launch_map<OP_AXPY>(y, x, nullptr, a);
}
int b2_vec_aypx(b2_vec* y, double a, const b2_vec* x) {
  B2_CHECK(y->n == x->n, "b2_vec_aypx: size mismatch");
  return launch_map<OP_AYPX>(y, x, nullptr, a);
}
int b2_vec_scale(b2_vec* v, double a) { return launch_map<OP_SCALE>(v, nullptr, nullptr, a); }
int b2_vec_add_scalar(b2_vec* v, double a) { return launch_map<OP_SHIFT>(v, nullptr, nullptr, a); }
int b2_vec_abs(b2_vec* v) { return launch_map<OP_ABS>(v, nullptr, nullptr, 0.); }
int b2_vec_pointwise_mult(b2_vec* w, const b2_vec* x, const b2_vec* y) {
  B2_CHECK(w->n == x->n && w->n == y->n, "b2_vec_pointwise_mult: size mismatch");
  return launch_map<OP_PMULT>(w, x, y, 0.);
}
int b2_vec_copy_masked(b2_vec* dst, const b2_vec* src, const b2_vec* mask, double thr) {
  B2_CHECK(dst->n == src->n && dst->n == mask->n, "b2_vec_copy_masked: size mismatch");
  return launch_map<OP_COPY_MASKED>(dst, src, mask, thr);
}

int b2_vec_dot(const b2_vec* x, const b2_vec* y, double* out) {
  B2_CHECK(x->n == y->n, "b2_vec_dot: size mismatch");
  b2_ctx* c = x->ctx;
  B2_TRY(launch_reduce<RED_DOT>(c, x->d, y->d, x->n, c->red_result, owned_of(x)));
  B2_TRY(b2_allreduce_sum(c, c->red_result, 1));
  B2_TRY(fetch_result(c, 1));
  *out = c->h_result[0];
  return 0;
}
int b2_vec_norm(const b2_vec* x, int kind, double* out) {
  b2_ctx* c = x->ctx;
  if (kind == 2) {
    B2_TRY(launch_reduce<RED_DOT>(c, x->d, x->d, x->n, c->red_result, owned_of(x)));
    B2_TRY(b2_allreduce_sum(c, c->red_result, 1));
    B2_TRY(fetch_result(c, 1));
    *out = sqrt(c->h_result[0]);
  } else if (kind == 1) {
    B2_TRY(launch_reduce<RED_L1>(c, x->d, nullptr, x->n, c->red_result, owned_of(x)));
    B2_TRY(b2_allreduce_sum(c, c->red_result, 1));
    B2_TRY(fetch_result(c, 1));
    *out = c->h_result[0];
  } else if (kind == 0) {
    B2_TRY(launch_reduce<RED_LINF>(c, x->d, nullptr, x->n, c->red_result));
    B2_TRY(b2_allreduce_op(c, c->red_result, 1, 2));
    B2_TRY(fetch_result(c, 1));
    *out = c->h_result[0];
  } else {
    B2_CHECK(false, "b2_vec_norm: kind %d", kind);
  }
  return 0;
}
int b2_vec_sum(const b2_vec* x, double* out) {
  b2_ctx* c = x->ctx;
  B2_TRY(launch_reduce<RED_SUM>(c, x->d, nullptr, x->n, c->red_result, owned_of(x)));
  B2_TRY(b2_allreduce_sum(c, c->red_result, 1));
  B2_TRY(fetch_result(c, 1));
  *out = c->h_result[0];
  return 0;
}
int b2_vec_minmax(const b2_vec* x, double* mn, double* mx) {
  b2_ctx* c = x->ctx;
  B2_TRY(launch_reduce<RED_MINMAX>(c, x->d, nullptr, x->n, c->red_result));
  B2_TRY(b2_allreduce_op(c, c->red_result, 1, 3));
  B2_TRY(b2_allreduce_op(c, c->red_result + 1, 1, 2));
  B2_TRY(fetch_result(c, 2));
  if (mn) *mn = c->h_result[0];
  if (mx) *mx = c->h_result[1];
  return 0;
}

// Entries arrive in CALL ORDER, duplicates allowed, with the meaning of VecSetValues (PetscVector.hpp:595-612): for
// INSERT the last value given for an index stays, for ADD the contributions of an index are summed in the order
// given.  A device scatter has no order, so duplicates are combined here, on the host arrays, before the upload.
static int indexed_op(b2_vec* v, const int32_t* idx, const double* vals, int64_t n, int mode, double a) {
  if (n == 0) return 0;
  b2_ctx* c = v->ctx;
  std::vector<int32_t> u_idx;
  std::vector<double> u_val;
  if (vals) {
    bool dup = false;
    {
      std::vector<uint8_t> seen((size_t)v->n, 0);
      for (int64_t k = 0; k < n; k++) {
        B2_CHECK(idx[k] >= 0 && idx[k] < v->n, "indexed vector access: index %d outside [0, %lld)", idx[k], (long long)v->n);
        if (seen[idx[k]]) { dup = true; break; }
        seen[idx[k]] = 1;
      }
    }
    if (dup) {
      std::vector<int64_t> slot((size_t)v->n, -1);
      for (int64_t k = 0; k < n; k++) {
        int64_t& s = slot[idx[k]];
        if (s < 0) {
          s = (int64_t)u_idx.size();
          u_idx.push_back(idx[k]);
          u_val.push_back(vals[k]);
        } else if (mode == 0) {
          u_val[(size_t)s] = vals[k];
        } else {
          u_val[(size_t)s] += vals[k];
        }
      }
      idx = u_idx.data();
      vals = u_val.data();
      n = (int64_t)u_idx.size();
    }
  }
  int32_t* d_idx = nullptr;
  double* d_vals = nullptr;
  B2_TRY(b2_malloc(c, &d_idx, (size_t)n));
  B2_TRY(b2_upload(c, d_idx, idx, (size_t)n));
  if (vals) {
    B2_TRY(b2_malloc(c, &d_vals, (size_t)n));
    B2_TRY(b2_upload(c, d_vals, vals, (size_t)n));
  }
  const int grid = b2_grid_for(c, n, kBlock, 8);
  B2_LAUNCH(c, vec_scatter_kernel, grid, kBlock, 0, v->d, d_idx, d_vals, n, mode, a);
  B2_CUDA(cudaStreamSynchronize(c->stream));
  b2_free(c, d_idx, (size_t)n);
  if (d_vals) b2_free(c, d_vals, (size_t)n);
  return 0;
}
int b2_vec_set_indexed(b2_vec* v, const int32_t* idx, const double* vals, int64_t n) {
  return indexed_op(v, idx, vals, n, 0, 0.);
}
int b2_vec_add_indexed(b2_vec* v, const int32_t* idx, const double* vals, int64_t n) {
  return indexed_op(v, idx, vals, n, 1, 0.);
}
int b2_vec_fill_indexed(b2_vec* v, const int32_t* idx, int64_t n, double a) {
  return indexed_op(v, idx, nullptr, n, 2, a);
}
int b2_vec_get_indexed(const b2_vec* v, const int32_t* idx, double* vals, int64_t n) {
  if (n == 0) return 0;
  b2_ctx* c = v->ctx;
  int32_t* d_idx = nullptr;
  double* d_vals = nullptr;
  B2_TRY(b2_malloc(c, &d_idx, (size_t)n));
  B2_TRY(b2_malloc(c, &d_vals, (size_t)n));
  B2_TRY(b2_upload(c, d_idx, idx, (size_t)n));
  const int grid = b2_grid_for(c, n, kBlock, 8);
  B2_LAUNCH(c, vec_gather_kernel, grid, kBlock, 0, v->d, d_idx, d_vals, n);
  B2_TRY(b2_download(c, vals, d_vals, (size_t)n));
  b2_free(c, d_idx, (size_t)n);
  b2_free(c, d_vals, (size_t)n);
  return 0;
}

}  // extern "C"
