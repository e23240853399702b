// Kernel launch helper with optional programmatic dependent launch (see pdl_wait() in common.cuh).
//
// The training step is a chain of ~300 short kernels inside one CUDA graph; a tiny kernel of ours costs
// 4.5-4.8 us there against 1.75 us for a trivial elementwise kernel (profiles/launch_overhead.txt) -- the
// difference is prologue (mbarrier init, TMEM allocation, tensor-map fetch, ring fill) and drain.  With
// EDL_PDL=1 (or set_pdl(true)) the hot kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization
// so that the prologue of kernel i+1 overlaps the tail of kernel i; stream capture turns these into programmatic
// graph edges.  Off by default until measured on hardware.
#pragma once
#include <cuda_runtime.h>

namespace edl {

bool pdl_enabled();
void set_pdl(bool on);

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  if (!pdl_enabled()) {
    kern<<<grid, block, smem, stream>>>(static_cast<KArgs>(args)...);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif  // __CUDACC__

}  // namespace edl
