// Shared by the translation units of libalignn_b200.so: launch bookkeeping.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>
namespace alignn {
// "Has this call site configured the CURRENT device yet?"  cudaFuncSetAttribute (opt-in dynamic shared memory) is per
// device: a process that uses two GPUs has to set it on each.  One static instance per call site.
struct DeviceOnce {
  std::atomic<uint64_t> mask{0};
  bool needed(int* dev) {
    *dev = 0;
    cudaGetDevice(dev);
    return !((mask.load(std::memory_order_acquire) >> (*dev & 63)) & 1ull);
  }
  void done(int dev) { mask.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};
int check_launch();             // counts the launch, maps cudaGetLastError() to an alignn status
int record_cuda_error(int e);   // remembers a failed runtime call, returns ALIGNN_ERR_CUDA
// Grid of a grid-stride row kernel: never more blocks than fit on the device at once (SMs x resident blocks per SM of
// THIS kernel, from the occupancy calculator, cached per kernel) -- 592 blocks of a kernel with three resident blocks
// per SM would run as 1.33 waves, the last third of the time at a third of the occupancy.
int one_wave_grid(const void* kernel, int threads, size_t dyn_smem, int wanted_blocks);
}
