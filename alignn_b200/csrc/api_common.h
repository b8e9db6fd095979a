// Shared by the translation units of libalignn_b200.so: launch bookkeeping.
#pragma once
namespace alignn {
int check_launch();             // counts the launch, maps cudaGetLastError() to an alignn status
int record_cuda_error(int e);   // remembers a failed runtime call, returns ALIGNN_ERR_CUDA
}
