// devmem.h -- tiny per-device cache of workspace blocks.
// cudaMalloc / cudaFree cost 10-100+ ms each on a device that holds tens of GB of mappings (measured: engine teardown
// 0.48 s, almost all of it in ~20 cudaFree calls), which is as long as 100 sweeps of the headline configuration.
// Engines are created and destroyed once per nmf() call with identical buffer sizes, so freed blocks are kept and
// handed out again by exact size.  Blocks larger than RRI_CACHE_BLOCK_MAX or beyond RRI_CACHE_TOTAL_MAX are really freed.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace rri {
cudaError_t cached_malloc(void** p, size_t bytes);     // current device; contents undefined
void cached_free(void* p);                             // current device must be the allocating one
void cache_trim();                                     // really free everything cached on the current device
}  // namespace rri
