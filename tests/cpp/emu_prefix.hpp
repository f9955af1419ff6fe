// TEST INFRASTRUCTURE: force-included (-include) in front of the library's orchestration translation units when they
// are compiled for the CPU thread emulator: CUDA keywords and intrinsics (cuda_emu.hpp), and every kernel launch of
// the library -- all of them go through B2_LAUNCH -- mapped onto emu::launch.
#pragma once
#include "cuda_emu.hpp"
#define B2_LAUNCH(ctx, kernel, grid, block, smem, ...)                                    \
  do {                                                                                    \
    emu::launch(kernel, (unsigned)(grid), (unsigned)(block), (size_t)(smem), __VA_ARGS__); \
    (ctx)->launches++;                                                                    \
  } while (0)
