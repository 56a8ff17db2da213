// Build-mode glue.  The product build is nvcc for sm_100a.  The same kernel sources can
// also be compiled by g++ against tests/sim/cusim.h (-DCPB_SIM), a test-only cooperative
// fiber model of a CUDA grid, so that kernel *logic* is checked against the oracle on the
// GPU-less build box.  Nothing in the product path loads the simulated library.
#pragma once

#include <stddef.h>
#include <stdint.h>

#ifdef CPB_SIM
#include "cusim.h"
#define CPB_KERNEL static void
#define CPB_DEVICE static inline
#define CPB_SHARED static thread_local
#define CPB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(cusim::dyn_smem())
#define CPB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cusim::launch(#kernel, (grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define CPB_RESTRICT
#define CPB_LAUNCH_BOUNDS(t, b)
#else
#include <cuda_runtime.h>
#define CPB_KERNEL __global__ void
#define CPB_DEVICE __device__ __forceinline__
#define CPB_SHARED __shared__
#define CPB_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; \
    type* name = reinterpret_cast<type*>(name##_raw_)
#define CPB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define CPB_RESTRICT __restrict__
#define CPB_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#endif
