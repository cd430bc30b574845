// CUDA headers + the two macros through which every kernel launch and every dynamic shared memory declaration of this
// library goes. For nvcc they expand to the usual syntax (the library's SASS is byte-identical with and without them);
// tests/simt defines AMB_SIMT_EMUL and supplies host versions, which is how the CPU test tier runs the kernel source
// and the host orchestration under a SIMT emulator.
#pragma once
#ifndef AMB_SIMT_EMUL
#include <cuda.h>
#include <cuda_runtime.h>
#define AMB_ID(...) __VA_ARGS__
#define AMB_LAUNCH(kernel, grid, block, smem, stream, ...) AMB_ID kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define AMB_DYN_SMEM(type, name, align) extern __shared__ __align__(align) type name[]
#endif
