// kernel_common.cuh -- small device helpers shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace anl {

#ifndef FULL
#define FULL 0xFFFFFFFFu
#endif

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// process-wide count of kernel launches issued by this library (bench.py reports it as gpu_launches)
void count_launch(unsigned n = 1);

}  // namespace anl
