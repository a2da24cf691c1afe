// Multi-GPU exchange helpers: pack / unpack of the 64x64-px blocks a context owns (see
// include/mprb.h: mprb_exchange_*; mpr_b200/sharding.py: NativeExchange).  No reference
// counterpart - the reference is single-GPU.
#include <cstdint>
#include <cuda_runtime.h>

#include "exchange.cuh"

// pack: the 64x64-px blocks a context owns, narrowed to what carries information (a 2D frame is
// 0 / 1 -> uint8, a depth value is below the frame size -> int16, normals stay 32 bit), into one
// contiguous buffer: [depth blocks of tile 0..n-1][normal blocks of tile 0..n-1].
// unpack: the same buffers of ALL ranks, back to back, scattered into the full-size images.
// One CTA per (tile, plane); a tile's index in its owner's buffer follows the row-major order in
// which api.cu enumerates owned tiles.
namespace mprb {
namespace {

struct ExchangeGeom {
    int size;        // image side, px
    int tiles;       // 64-px tiles per side
    int world;       // ranks
    int col_step;    // 1: owner = (y + x) % world, 0: owner = y % world
    int per_rank;    // tiles per rank
    int depth_bytes; // bytes per depth pixel in transit (1 or 2)
    int planes;      // 1 (2D) or 2 (3D: + normals)
};

__device__ __forceinline__ void tile_slot(const ExchangeGeom& g, int ty, int tx, int& owner, int& k) {
    if (g.col_step) {
        owner = (ty + tx) % g.world;
        k = ty * (g.tiles / g.world) + tx / g.world;
    } else {
        owner = ty % g.world;
        k = (ty / g.world) * g.tiles + tx;
    }
}

template <bool PACK>
__global__ void __launch_bounds__(256)
k_exchange(const ExchangeGeom g, int rank, int32_t* __restrict__ depth, uint32_t* __restrict__ normals,
           unsigned char* __restrict__ buf)
{
    // PACK: blockIdx.x enumerates this rank's tiles; else: all tiles of the frame
    const int plane = blockIdx.y;
    int ty, tx, owner, k;
    if (PACK) {
        // invert tile_slot for `rank`
        k = blockIdx.x;
        owner = rank;
        if (g.col_step) {
            const int per_row = g.tiles / g.world;
            ty = k / per_row;
            tx = (((rank - ty) % g.world) + g.world) % g.world + (k % per_row) * g.world;
        } else {
            ty = (k / g.tiles) * g.world + rank;
            tx = k % g.tiles;
        }
    } else {
        ty = blockIdx.x / g.tiles;
        tx = blockIdx.x % g.tiles;
        tile_slot(g, ty, tx, owner, k);
    }
    const size_t rank_bytes = size_t(g.per_rank) * 4096 * (g.depth_bytes + (g.planes == 2 ? 4 : 0));
    unsigned char* base = buf + (PACK ? 0 : size_t(owner) * rank_bytes);
    const size_t px0 = size_t(ty) * 64 * g.size + size_t(tx) * 64;
    if (plane == 0) {
        unsigned char* dst = base + size_t(k) * 4096 * g.depth_bytes;
        for (int i = threadIdx.x; i < 4096; i += 256) {
            int32_t* p = depth + px0 + size_t(i >> 6) * g.size + (i & 63);
            if (g.depth_bytes == 1) {
                if (PACK) dst[i] = (unsigned char)*p; else *p = dst[i];
            } else {
                int16_t* d16 = reinterpret_cast<int16_t*>(dst);
                if (PACK) d16[i] = (int16_t)*p; else *p = d16[i];
            }
        }
    } else {
        uint32_t* dst = reinterpret_cast<uint32_t*>(base + size_t(g.per_rank) * 4096 * g.depth_bytes) + size_t(k) * 4096;
        for (int i = threadIdx.x; i < 4096; i += 256) {
            uint32_t* p = normals + px0 + size_t(i >> 6) * g.size + (i & 63);
            if (PACK) dst[i] = *p; else *p = dst[i];
        }
    }
}

// Multi-GPU context inside one process (api.cu: render_all): device `rank` of `world` copies the
// 64x64-px blocks it owns - tiles with (x + y) % world == rank - from its own frame into the
// primary device's frame with plain stores through the peer mapping (NVLink): 128-bit rows, one CTA
// per (tile, plane), no intermediate buffer on either side.
__global__ void __launch_bounds__(256)
k_publish(int size, int tiles, int world, int rank, const int32_t* __restrict__ depth, const uint32_t* __restrict__ normals,
          int32_t* __restrict__ peer_depth, uint32_t* __restrict__ peer_normals)
{
    const int per_row = (tiles + world - 1) / world;
    const int ty = blockIdx.x / per_row;
    const int tx = (((rank - ty) % world) + world) % world + (blockIdx.x % per_row) * world;
    if (tx >= tiles) return;
    const size_t px0 = size_t(ty) * 64 * size + size_t(tx) * 64;
    const uint4* src = reinterpret_cast<const uint4*>(blockIdx.y == 0 ? reinterpret_cast<const uint32_t*>(depth) : normals);
    uint4* dst = reinterpret_cast<uint4*>(blockIdx.y == 0 ? reinterpret_cast<uint32_t*>(peer_depth) : peer_normals);
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {          // 64 rows of 16 x 16 bytes
        const size_t o = (px0 + size_t(i >> 4) * size) / 4 + (i & 15);
        dst[o] = src[o];
    }
}

}  // namespace

void launch_publish(int size, int world, int rank, int dim, const int32_t* depth, const uint32_t* normals,
                    int32_t* peer_depth, uint32_t* peer_normals, cudaStream_t s) {
    const int tiles = size / 64;
    const int per_row = (tiles + world - 1) / world;
    k_publish<<<dim3(tiles * per_row, dim == 3 ? 2 : 1), 256, 0, s>>>(size, tiles, world, rank, depth, normals, peer_depth,
                                                                       peer_normals);
}

size_t exchange_rank_bytes(int size, int world, int dim) {
    const int tiles = size / 64;
    const size_t per_rank = size_t(tiles) * tiles / world;
    return per_rank * 4096 * (dim == 3 ? 2 + 4 : 1);
}

void launch_exchange(bool pack, int size, int world, int rank, int col_step, int dim, int32_t* depth, uint32_t* normals,
                     void* buf, cudaStream_t s) {
    ExchangeGeom g;
    g.size = size;
    g.tiles = size / 64;
    g.world = world;
    g.col_step = col_step;
    g.per_rank = g.tiles * g.tiles / world;
    g.depth_bytes = dim == 3 ? 2 : 1;
    g.planes = dim == 3 ? 2 : 1;
    if (pack) k_exchange<true><<<dim3(g.per_rank, g.planes), 256, 0, s>>>(g, rank, depth, normals, (unsigned char*)buf);
    else k_exchange<false><<<dim3(g.tiles * g.tiles, g.planes), 256, 0, s>>>(g, rank, depth, normals, (unsigned char*)buf);
}

}  // namespace mprb
