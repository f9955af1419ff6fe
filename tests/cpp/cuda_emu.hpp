// TEST INFRASTRUCTURE: a minimal CPU emulator of the CUDA execution model, enough to run the library's simple
// kernels (the .cuh device sources under femus_b200/csrc that avoid tensor-core / TMA instructions) on host threads
// when no GPU is present: one std::thread per CUDA thread of a CTA, CTAs one after the other, __syncthreads /
// __syncwarp as barriers, warp shuffles through a per-warp exchange buffer, atomics through std::atomic_ref.
// It checks the LOGIC of a kernel (indexing, synchronisation protocol, arithmetic) -- not performance, and not
// memory-model subtleties (host memory is sequentially consistent at the barriers).
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local emu_dim3 threadIdx, blockIdx;
inline emu_dim3 blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __constant__
#define __shared__ static            /* CTAs run one at a time: one static copy per kernel is the CTA's shared memory */
#define B2_DYN_SHARED(type, name) type* name = reinterpret_cast<type*>(emu::dyn_shared.data())

namespace emu {
inline std::vector<double> dyn_shared;                      // dynamic shared memory of the running CTA
inline std::unique_ptr<std::barrier<>> cta_barrier;
inline std::vector<std::unique_ptr<std::barrier<>>> warp_barrier;
inline std::vector<std::vector<double>> warp_buf;           // [warp][32] shuffle exchange
inline int warp_of() { return (int)(threadIdx.x >> 5); }

template <class K, class... Args>
void launch(K kernel, unsigned grid, unsigned block, size_t smem_bytes, Args... args) {
  gridDim.x = grid;
  blockDim.x = block;
  dyn_shared.assign(smem_bytes / sizeof(double) + 1, 0.0);
  const unsigned nwarps = (block + 31) / 32;
  cta_barrier.reset(new std::barrier<>(block));
  warp_barrier.clear();
  warp_buf.assign(nwarps, std::vector<double>(32, 0.0));
  for (unsigned w = 0; w < nwarps; w++) {
    const unsigned lanes = (w + 1) * 32 <= block ? 32 : block - w * 32;
    warp_barrier.emplace_back(new std::barrier<>(lanes));
  }
  // one host thread per CUDA thread of a CTA, reused for every CTA of the grid: the CTAs run one after the other (the
  // threads meet at the CTA barrier between two of them), so the static __shared__ storage always belongs to one CTA
  std::vector<std::thread> th;
  for (unsigned t = 0; t < block; t++)
    th.emplace_back([=] {
      threadIdx.x = t;
      for (unsigned b = 0; b < grid; b++) {
        blockIdx.x = b;
        kernel(args...);
        cta_barrier->arrive_and_wait();
      }
    });
  for (auto& x : th) x.join();
}
}  // namespace emu

inline void __syncthreads() { emu::cta_barrier->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier[emu::warp_of()]->arrive_and_wait(); }
inline double emu_shfl(double v, int src_lane) {
  const int w = emu::warp_of(), lane = (int)(threadIdx.x & 31);
  emu::warp_buf[w][lane] = v;
  emu::warp_barrier[w]->arrive_and_wait();
  const double r = (src_lane >= 0 && src_lane < 32) ? emu::warp_buf[w][src_lane] : v;
  emu::warp_barrier[w]->arrive_and_wait();
  return r;
}
inline double __shfl_xor_sync(unsigned, double v, int mask) { return emu_shfl(v, (int)(threadIdx.x & 31) ^ mask); }
inline double __shfl_down_sync(unsigned, double v, int delta) {
  const int src = (int)(threadIdx.x & 31) + delta;
  return emu_shfl(v, src < 32 ? src : (int)(threadIdx.x & 31));
}
inline unsigned __ballot_sync(unsigned, int pred) {
  const int w = emu::warp_of(), lane = (int)(threadIdx.x & 31);
  emu::warp_buf[w][lane] = pred ? 1.0 : 0.0;
  emu::warp_barrier[w]->arrive_and_wait();
  unsigned m = 0;
  for (int l = 0; l < 32; l++)
    if (emu::warp_buf[w][l] != 0.0) m |= 1u << l;
  emu::warp_barrier[w]->arrive_and_wait();
  return m;
}
inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }
inline int atomicCAS(int* p, int expected, int desired) {
  std::atomic_ref<int>(*p).compare_exchange_strong(expected, desired);
  return expected;
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <class T>
inline T __ldcg(const T* p) { return *p; }
template <class T>
inline T __ldg(const T* p) { return *p; }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
using std::fabs;
using std::fma;
using std::sqrt;
