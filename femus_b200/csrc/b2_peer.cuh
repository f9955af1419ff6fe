// Cells of the peer-memory exchange (b2_halo.cu, b2_cg.cu): a double travels as 16 bytes {low word, flag, high word, flag}.
// 8-byte stores are atomic, so a reader sees each half either stale or complete with its flag; it re-reads until both
// flags carry the number of the current exchange.  Inbox of a rank: [2 parity][nranks senders][slot cells].
#pragma once
#include <cstdint>

constexpr int kPeerScal = 8;      // cells at the start of every slot that carry scalars

__device__ __forceinline__ unsigned b2_peer_flag(unsigned long long epoch) { return (unsigned)(epoch % 0xfffffffeull) + 1u; }      // never 0 (fresh memory)

__device__ __forceinline__ uint4* b2_peer_cell(void* base, int64_t slot, int nranks, int parity, int sender, int64_t cell) {
  return reinterpret_cast<uint4*>(base) + ((int64_t)parity * nranks + sender) * slot + cell;
}
__device__ __forceinline__ void b2_peer_store(uint4* cell, double value, unsigned flag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(value);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"((unsigned)(bits & 0xffffffffull)), "r"(flag),
               "r"((unsigned)(bits >> 32)), "r"(flag)
               : "memory");
}
// waits for the cell to carry `flag`; gives up after ~10 s (a rank stopped taking part), sets *err and returns 0
__device__ __forceinline__ double b2_peer_wait(const uint4* cell, unsigned flag, int* err) {
  unsigned a, fa, b, fb;
  long long t_start = 0;
  for (unsigned spins = 0;; spins++) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(cell) : "memory");
    if (fa == flag && fb == flag) break;
    if ((spins & 1023u) == 1023u) {
      if (t_start == 0) t_start = clock64();
      else if (clock64() - t_start > 20000000000ll) {
        *err = 1;
        return 0.0;
      }
    }
  }
  return __longlong_as_double((long long)(((unsigned long long)b << 32) | (unsigned long long)a));
}
