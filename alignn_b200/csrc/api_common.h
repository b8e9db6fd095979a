// Shared by the translation units of libalignn_b200.so: launch bookkeeping.
#pragma once
namespace alignn {
int check_launch();             // counts the launch, maps cudaGetLastError() to an alignn status
int record_cuda_error(int e);   // remembers a failed runtime call, returns ALIGNN_ERR_CUDA
// Grid of a grid-stride row kernel: never more blocks than fit on the device at once (SMs x resident blocks per SM of
// THIS kernel, from the occupancy calculator, cached per kernel) -- 592 blocks of a kernel with three resident blocks
// per SM would run as 1.33 waves, the last third of the time at a third of the occupancy.
int one_wave_grid(const void* kernel, int threads, size_t dyn_smem, int wanted_blocks);
}
