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
// Slot renaming (REMAP = true).  Slot values live in shared-memory rows, one row per slot
// id, and the rows have to be sized for the ROOT tape's slot count (up to 128) even though a
// shortened tape touches far fewer ids - scattered over the whole range, because the tape
// packer recycles slots LIFO.  For models with many slots that caps occupancy at 8 warps/SM.
// In REMAP mode every chunk is re-written on arrival into a second buffer with each slot id
// replaced by a dense ROW number, assigned on first sight per tape through a 256-byte table.
// The walkers then address rows instead of ids: the first n_rows rows are shared-memory
// rows, the (rare) rest spills to per-thread local memory.  Renaming is a bijection per
// tape, so every value, comparison and liveness test is unchanged; cells written back to the
// arena (tape pushes) are taken from the untouched raw chunk.
//
// All member functions must be called by the whole warp with identical
// arguments (the stream state is warp-uniform).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "common.cuh"

namespace mprb {

constexpr int kRemapRowsDefault = 16;           // shared-memory value rows per warp in REMAP mode
constexpr int kRemapRowsMin = 8;                //   (MPRB_REMAP_ROWS overrides; rows beyond spill to local memory)
constexpr int kStreamStridePlain = 640;         // raw chunk 512 + mbarrier, padded to 128
constexpr int kStreamStrideRemap = 1408;        // + renamed chunk 512 + table 256

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u2(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <bool REMAP>
struct TapeStream {
    uint32_t buf;            // shared-space address of the raw 64-cell chunk
    uint32_t rd;             // buffer the walkers read: renamed copy (REMAP) or the raw chunk
    uint32_t bar;            // mbarrier
    uint32_t table;          // REMAP: slot id -> row, 256 bytes, 0xFF = not seen yet
    uint32_t phase;
    uint32_t next_row;       // REMAP: rows handed out so far for the current tape
    int base;                // arena index of buffer cell 0 (multiple of 64), or -1
    const uint64_t* arena;

    static __host__ __device__ constexpr int stride() { return REMAP ? kStreamStrideRemap : kStreamStridePlain; }

    // storage: stride() bytes of shared memory owned by this warp, 128-byte aligned
    __device__ __forceinline__ void init(void* storage, const uint64_t* arena_) {
        buf = smem_addr(storage);
        bar = buf + kChunk * 8;
        rd = REMAP ? buf + 640 : buf;
        table = buf + 640 + 512;
        phase = 0;
        next_row = 0;
        base = -1;
        arena = arena_;
        if ((threadIdx.x & 31) == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }

    // Starts a new tape.  `hdr` is the root header word {0, x, y, z slot ids}; returns the same
    // word with ids replaced by rows (identity without REMAP).
    __device__ __forceinline__ uint32_t begin_tape(uint32_t hdr) {
        if (!REMAP) return hdr;
        const int lane = threadIdx.x & 31;
        __syncwarp();
        sts_u2(table + lane * 8, make_uint2(0xffffffffu, 0xffffffffu));
        __syncwarp();
        // id 0 ("no operand") -> row 0; the axes take the next rows in x, y, z order
        uint32_t out = 0, row = 1;
        uint32_t seen[3] = {0, 0, 0}, seen_row[3] = {0, 0, 0};
        #pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t id = (hdr >> (8 + 8 * k)) & 0xff;
            uint32_t r = 0;
            if (id) {
                r = row;
                bool dup = false;
                #pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (j < k && seen[j] == id) { r = seen_row[j]; dup = true; }
                if (!dup) ++row;
            }
            seen[k] = id;
            seen_row[k] = r;
            out |= r << (8 + 8 * k);
            if (lane == 0 && id) sts_u8(table + id, r);
        }
        if (lane == 0) sts_u8(table, 0);
        next_row = row;
        base = -1;                      // the renamed copy of a resident chunk belongs to the old tape
        __syncwarp();
        return out;
    }

    __device__ __forceinline__ void load_chunk(int want) {
        // Without renaming the walkers annotate the raw chunk in place (generic-proxy byte stores,
        // kernels.cu:annotate_chunk); the bulk copy below writes the same bytes through the async
        // proxy, so those stores are ordered before it explicitly.
        if (!REMAP) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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

    // REMAP: rewrites the raw chunk into the renamed buffer.  Cells in [lo, hi] (hi = the first
    // end cell at or after lo, else 63) belong to the tape being walked; with `assign` unseen
    // slot ids in them get fresh rows.  Everything else is copied verbatim (links, stale cells).
    __device__ __forceinline__ void rename(int lo, bool assign) {
        const int lane = threadIdx.x & 31;
        uint2 c[2];
        c[0] = lds_u2(buf + lane * 8);
        c[1] = lds_u2(buf + (lane + 32) * 8);
        int hi = kChunk - 1;
        if (assign) {
            const unsigned m0 = __ballot_sync(0xffffffffu, lane >= lo && (c[0].x & 0xff) == OP_END);
            const unsigned m1 = __ballot_sync(0xffffffffu, lane + 32 >= lo && (c[1].x & 0xff) == OP_END);
            hi = m0 ? __ffs(m0) - 1 : (m1 ? 31 + __ffs(m1) : kChunk - 1);
        }
        bool valid[2];
        #pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int j = lane + 32 * k;
            valid[k] = j >= lo && j <= hi && (c[k].x & 0xff) != OP_JUMP;
        }
        if (assign) {
            #pragma unroll
            for (int k = 0; k < 2; ++k) {
                #pragma unroll
                for (int b = 1; b < 4; ++b) {
                    const uint32_t id = (c[k].x >> (8 * b)) & 0xff;
                    bool need = valid[k] && lds_u8(table + id) == 0xff;
                    __syncwarp();       // orders these table reads before the leaders' writes below (racecheck: WAR)
                    unsigned m = __ballot_sync(0xffffffffu, need);
                    while (m) {
                        const int leader = __ffs(m) - 1;
                        const uint32_t lid = __shfl_sync(0xffffffffu, id, leader);
                        if (lane == leader) sts_u8(table + lid, next_row);
                        ++next_row;
                        if (id == lid) need = false;
                        m = __ballot_sync(0xffffffffu, need);
                    }
                    __syncwarp();
                }
            }
        }
        #pragma unroll
        for (int k = 0; k < 2; ++k) {
            uint32_t w = c[k].x;
            if (valid[k]) {
                const uint32_t r1 = lds_u8(table + ((w >> 8) & 0xff));
                const uint32_t r2 = lds_u8(table + ((w >> 16) & 0xff));
                const uint32_t r3 = lds_u8(table + (w >> 24));
                w = (w & 0xff) | (r1 << 8) | (r2 << 16) | (r3 << 24);
            }
            sts_u2(rd + (lane + 32 * k) * 8, make_uint2(w, c[k].y));
        }
        __syncwarp();
    }

    // Forward walking: makes the chunk of arena cell `index` resident; the walk continues at
    // index + 1, so cells up to `index` in that chunk are not part of this tape.
    // Returns true when a new chunk was brought in (false: it was already resident).
    __device__ __forceinline__ bool fetch(int index) {
        const int want = index & ~(kChunk - 1);
        if (want == base) return false;
        load_chunk(want);
        if (REMAP) rename((index & (kChunk - 1)) + 1, true);
        return true;
    }

    // Backward walking (tape push): every id was seen on the way forward, so the whole chunk
    // is renamed by lookup only.
    __device__ __forceinline__ void fetch_back(int index) {
        const int want = index & ~(kChunk - 1);
        if (want == base) return;
        load_chunk(want);
        if (REMAP) rename(0, false);
    }
};

}  // namespace mprb
