// calico_b200 — platform glue.
//
// The product is CUDA for sm_100a and nothing else: libcalico_b200.so is built by nvcc and every entry point that
// touches the device fails with CB2_INTERNAL when no GPU is present. There is NO CPU fallback in the product.
//
// CB2_EMUL is a developer/test harness only (tests/emul/): the same kernel sources are compiled by g++ against a
// tiny SIMT emulator (one OS thread per CUDA thread, std::barrier for __syncthreads) so that indexing and
// synchronisation logic can be debugged in a container without a GPU. That build produces
// tests/emul/libcalico_b200_emul.so, which the package never loads.
#pragma once

#if defined(CB2_EMUL)
#include "cuda_emul.h"
#else
#include <cuda_runtime.h>
#endif

#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define CB2_HD __host__ __device__ __forceinline__
#define CB2_D __device__ __forceinline__
#else
#define CB2_HD inline
#define CB2_D inline
#endif

#if defined(CB2_EMUL)
#define CB2_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ::cb2emul::launch(dim3(grid), dim3(block), size_t(smem), [&]() { kernel(__VA_ARGS__); })
#else
#define CB2_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<dim3(grid), dim3(block), size_t(smem), (stream)>>>(__VA_ARGS__)
#endif

namespace cb2 {

// Dynamic shared memory base pointer (16-byte aligned).
template <typename T>
CB2_D T* dyn_smem() {
#if defined(CB2_EMUL)
  return reinterpret_cast<T*>(::cb2emul::dyn_smem_base());
#else
  extern __shared__ __align__(16) unsigned char cb2_dyn_smem_[];
  return reinterpret_cast<T*>(cb2_dyn_smem_);
#endif
}

}  // namespace cb2
