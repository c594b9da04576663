// Two spellings that differ between the product build (nvcc, sm_100a) and the kernel-logic emulator used by the CPU
// tests (tests/emu: the same sources compiled by g++ against a fake cuda_runtime.h, CTA threads run as fibers).
// The emulator is test infrastructure only -- nothing in genfft_b200 loads it and libgenfft_cuda.so never contains
// it; GENFFT_EMU is defined solely by tests/emu/build.sh.
#pragma once

#ifdef GENFFT_EMU
// kernel(args...) runs once per emulated thread; __syncthreads() yields to the emulator's scheduler
#define GENFFT_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ::genfft_emu::launch((grid), (block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#define GENFFT_DYN_SMEM(name) unsigned char* name = ::genfft_emu::dyn_smem()
#else
#define GENFFT_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define GENFFT_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
