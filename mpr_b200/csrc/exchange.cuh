// Multi-GPU exchange (exchange.cu): bytes one rank contributes, and the pack / unpack launches.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace mprb {
size_t exchange_rank_bytes(int size, int world, int dim);
void launch_exchange(bool pack, int size, int world, int rank, int col_step, int dim, int32_t* depth, uint32_t* normals,
                     void* buf, cudaStream_t s);
void launch_publish(int size, int world, int rank, int dim, const int32_t* depth, const uint32_t* normals,
                    int32_t* peer_depth, uint32_t* peer_normals, cudaStream_t s);
}  // namespace mprb
