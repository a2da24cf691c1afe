// Per-warp tape staging: 512-byte arena chunks are pulled into shared memory
// with the TMA bulk-copy engine (cp.async.bulk + mbarrier) and clauses are then
// fetched with broadcast shared-memory loads.
//
// Why: a tape is warp-uniform, so fetching it clause by clause from global
// memory costs a 64-bit address computation and an L1/L2 round trip per clause
// for data every lane shares.  Tapes live in the arena as 64-cell chunks
// (reference inc/parameters.hpp:16, src/context.cu:384-413): pushed tapes are
// chunk lists linked by JUMP cells, the root tape and the tapes written by
// k_eval_root are contiguous runs.  The arena allocator here starts at a
// 64-cell boundary and hands out whole chunks, so every chunk is 512-byte
// aligned and one bulk copy moves exactly one chunk.
//
// Layout invariant relied on by the walkers: the LAST cell of every chunk a walk can reach is
// a JUMP or an end cell, so a walker only has to think about chunk boundaries when it meets
// one of those.  Pushed tapes have it by construction (cell 63 is the forward link or the end
// cell, cell 0 the back link or the header).  Contiguous tapes longer than one chunk (the root
// tape, and the tapes k_eval_root writes) are stored in the same shape: chunk 0 holds cells
// 0..62, every later chunk 62 cells at offsets 1..62, with JUMP(+1) in cell 63 and JUMP(-1)
// in cell 0 of the following chunk - exactly what a run of adjacent pushed chunks looks like.
//
// All member functions must be called by the whole warp with identical
// arguments (the stream state is warp-uniform).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "common.cuh"

namespace mprb {

constexpr int kStreamBytes = kChunk * 8 + 16;   // chunk buffer + mbarrier (+pad), per warp

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

struct TapeStream {
    uint32_t buf;            // shared-space address of the 64-cell buffer
    uint32_t bar;            // shared-space address of the mbarrier
    uint32_t phase;
    int base;                // arena index of buffer cell 0 (multiple of 64), or -1
    const uint64_t* arena;

    // storage: kStreamBytes of shared memory owned by this warp, 128-byte aligned
    __device__ __forceinline__ void init(void* storage, const uint64_t* arena_) {
        buf = smem_addr(storage);
        bar = buf + kChunk * 8;
        phase = 0;
        base = -1;
        arena = arena_;
        if ((threadIdx.x & 31) == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }

    // Makes the chunk that contains arena cell `index` resident.
    __device__ __forceinline__ void fetch(int index) {
        const int want = index & ~(kChunk - 1);
        if (want == base) return;
        __syncwarp();                                   // everyone is done reading the old chunk
        if ((threadIdx.x & 31) == 0) {
            const uint64_t* src = arena + want;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kChunk * 8)
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(buf), "l"(src), "r"(kChunk * 8), "r"(bar)
                : "memory");
        }
        // every lane waits on the barrier phase (hardware-suspended try_wait, not a spin on memory)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@!p bra WAIT_%=;\n"
            "}\n" ::"r"(bar), "r"(phase)
            : "memory");
        phase ^= 1;
        base = want;
    }

    // Clause at arena cell `index`; the chunk must be resident (fetch(index) first).
    __device__ __forceinline__ uint2 cell(int index) const {
        uint2 v;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(buf + ((index & (kChunk - 1)) << 3)));
        return v;
    }

    __device__ __forceinline__ uint2 get(int index) {
        fetch(index);
        return cell(index);
    }
};

}  // namespace mprb
