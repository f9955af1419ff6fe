// TEST INFRASTRUCTURE: proves that the emulator + ThreadSanitizer pair detects a missing __syncthreads (run without
// arguments: racy; with one argument: synchronised).  Used by tests/test_kernel_emulation.py.
#include "cuda_emu.hpp"
#include <cstdio>
__global__ void racy(double* out, int sync) {
  __shared__ double s[64];
  s[threadIdx.x] = threadIdx.x;
  if (sync) __syncthreads();
  out[threadIdx.x] = s[(threadIdx.x + 1) % 64];
}
int main(int argc, char** argv) {
  double out[64];
  emu::launch(racy, 1u, 64u, 0, out, argc > 1);
  std::printf("done %g\n", out[0]);
}
