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
constexpr uint32_t kOpBounce = 31;              // no opcode of the reference (inc/gpu_opcode.hpp: 0..29); see TapeStream::row_limit
constexpr int kFloatRemapRowsDefault = 16;      // float pass with renamed slots (generated loop): MPRB_FLOAT_ROWS overrides
constexpr int kRemapRowsMin = 8;                //   (MPRB_REMAP_ROWS overrides; rows beyond spill to local memory)
// Multiples of 128 bytes: the value rows that follow the streams in shared memory are 256-byte lines
// read by a whole warp, and a misaligned line costs a third wavefront per access.
constexpr int kStreamStridePlain = 1152;        // two raw chunks 2 x 512 + two mbarriers, padded
constexpr int kStreamStrideRemap = 1920;        // + renamed chunk 512 + table 256, padded
constexpr int kStreamStrideSingle = 640;        // AHEAD = false: one raw chunk 512 + mbarrier, padded

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

// AHEAD = false drops the second raw buffer and the look-ahead: where shared memory per warp decides how
// many warps an SM holds (the float pass without renaming: ~6 KB of value rows per warp) the extra 512
// bytes cost more than the hidden copy latency returns (measured, bear 1024^3: 4.74 vs 4.98 ms).
template <bool REMAP, bool AHEAD = true>
struct TapeStream {
    uint32_t buf;            // shared-space address of the raw 64-cell chunk being walked
    uint32_t other;          // the second raw buffer: the chunk the walk will jump to next is loaded here ahead of time
    uint32_t rd;             // buffer the walkers read: renamed copy (REMAP) or the raw chunk
    uint32_t bars;           // two mbarriers, one per raw buffer: bars + 8 * (buffer index)
    uint32_t lo;             // address of raw buffer 0 (buffer index = (addr - lo) >> 9)
    uint32_t table;          // REMAP: slot id -> row, 256 bytes, 0xFF = not seen yet
    uint32_t parity;         // bit k: phase parity the next wait on buffer k has to see
    uint32_t next_row;       // REMAP: rows handed out so far for the current tape
    uint32_t row_limit;      // REMAP: clauses of the generated loops that touch a row >= row_limit get opcode
                             //   kOpBounce in the renamed copy (0xffffffff: never - the C++ walkers spill by themselves)
    uint32_t bounce_ops;     // REMAP: bit i = opcode i runs in the generated loop (only those are ever bounced)
    int base;                // arena index of cell 0 of `buf` (multiple of 64), or -1
    int pre_base;            // arena index of the chunk requested into `other`, or -1 (a request is always waited
                             // for before its buffer or barrier is used again)
    int cap;                 // arena size in cells (look-ahead targets are checked against it)
    const uint64_t* arena;

    static __host__ __device__ constexpr int stride() {
        return REMAP ? kStreamStrideRemap : (AHEAD ? kStreamStridePlain : kStreamStrideSingle);
    }

    // storage: stride() bytes of shared memory owned by this warp, 128-byte aligned
    __device__ __forceinline__ void init(void* storage, const uint64_t* arena_, int arena_cells) {
        lo = smem_addr(storage);
        buf = lo;
        other = lo + kChunk * 8;
        bars = lo + (AHEAD ? 2 : 1) * kChunk * 8;
        rd = REMAP ? bars + 16 : buf;
        table = bars + 16 + 512;
        parity = 0;
        next_row = 0;
        row_limit = 0xffffffffu;
        bounce_ops = 0;
        base = -1;
        pre_base = -1;
        cap = arena_cells;
        arena = arena_;
        if ((threadIdx.x & 31) == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars) : "memory");
            if (AHEAD) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8) : "memory");
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
        // the renamed copy of a resident chunk belongs to the old tape: forget the chunk (a raw
        // chunk that is already on its way stays valid - renaming happens when it becomes current)
        release();
        __syncwarp();
        return out;
    }

    // Drops the current chunk, handing its buffer over only once nothing is in flight into it.
    __device__ __forceinline__ void release() { base = -1; }

    // Bulk copy of the chunk at arena cell `want` into raw buffer `dst`; completion is signalled on
    // that buffer's mbarrier.  The buffers are also written with ordinary stores (the walkers annotate
    // chunks in place, kernels.cu:annotate_chunk) while the bulk copy writes through the async proxy,
    // hence the proxy fence.
    __device__ __forceinline__ void request(uint32_t dst, int want) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();                                   // everyone is done reading what was there
        if ((threadIdx.x & 31) == 0) {
            const uint32_t bar = bars + (((dst - lo) >> 9) << 3);
            const uint64_t* src = arena + want;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kChunk * 8)
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(dst), "l"(src), "r"(kChunk * 8), "r"(bar)
                : "memory");
        }
    }
    // Every lane waits for the request into `dst` (hardware-suspended try_wait, not a spin on memory).
    __device__ __forceinline__ void wait(uint32_t dst) {
        const uint32_t k = (dst - lo) >> 9;
        const uint32_t bar = bars + (k << 3);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@!p bra WAIT_%=;\n"
            "}\n" ::"r"(bar), "r"((parity >> k) & 1u)
            : "memory");
        parity ^= 1u << k;
    }

    // Makes the chunk at arena cell `want` current: it is either the one requested ahead of time
    // (swap buffers) or has to be fetched now.
    __device__ __forceinline__ void load_chunk(int want) {
        __syncwarp();                                   // everyone is done reading the old chunk (and its renamed copy)
        if (want == pre_base) {
            wait(other);
            // swap the roles of the two buffers (written as arithmetic on the buffer index, which
            // ptxas can follow on its uniform datapath; a swap through a temporary it cannot)
            buf = lo + (lo + kChunk * 8 - buf);
            other = lo + (lo + kChunk * 8 - buf);
        } else {
            if (pre_base >= 0) wait(other);             // an unused look-ahead still owns `other` and its barrier
            request(buf, want);
            wait(buf);
        }
        pre_base = -1;
        if (!REMAP) rd = buf;
        base = want;
    }

    // Look-ahead: the cell at offset `link` of the current chunk (63 walking forward, 0 walking
    // backward) is a JUMP when the tape goes on in another chunk; request that chunk into the
    // other buffer now, a whole chunk's worth of clauses before the walk gets there.  (The cell can
    // be stale data that merely looks like a JUMP - a tape that ends before cell 63 - so the target
    // is range-checked and a useless request costs one 512-byte read.)
    __device__ __forceinline__ void look_ahead(int link) {
        if (!AHEAD) return;
        const uint2 c = lds_u2(buf + link * 8);
        // (in-place hints live in bits 5-7 of the opcode byte; a JUMP never carries any.)  Written as one
        // predicate, not as early returns: with those ptxas stops treating the stream state as
        // warp-uniform, and the walkers' clause fetch / decode / dispatch fall off the uniform datapath.
        const int t = (base + link + int32_t(c.y)) & ~(kChunk - 1);
        const bool go = (c.x & 0xff) == OP_JUMP && t >= 0 && t < cap && t != base;
        if (go) request(other, t);
        pre_base = go ? t : -1;
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
                // a clause of the generated loop whose rows are not all shared-memory rows leaves the loop: the
                // walker takes its opcode from the raw chunk and runs it with the spilling row accessors
                const uint32_t op = ((bounce_ops >> (w & 31)) & 1u) && max(r1, max(r2, r3)) >= row_limit ? kOpBounce : (w & 0xff);
                w = op | (r1 << 8) | (r2 << 16) | (r3 << 24);
            }
            sts_u2(rd + (lane + 32 * k) * 8, make_uint2(w, c[k].y));
        }
        __syncwarp();
    }

    // Before the CTA retires: no bulk copy may still be on its way into this warp's buffers.
    __device__ __forceinline__ void drain() {
        if (pre_base >= 0) wait(other);
        pre_base = -1;
    }

    // Forward walking: makes the chunk of arena cell `index` resident; the walk continues at
    // index + 1, so cells up to `index` in that chunk are not part of this tape.
    // Returns true when a new chunk was brought in (false: it was already resident).
    __device__ __forceinline__ bool fetch(int index) {
        const int want = index & ~(kChunk - 1);
        if (want == base) return false;
        load_chunk(want);
        look_ahead(kChunk - 1);
        if (REMAP) rename((index & (kChunk - 1)) + 1, true);
        return true;
    }

    // Backward walking (tape push): every id was seen on the way forward, so the whole chunk
    // is renamed by lookup only.
    __device__ __forceinline__ void fetch_back(int index) {
        const int want = index & ~(kChunk - 1);
        if (want == base) return;
        load_chunk(want);
        look_ahead(0);
        if (REMAP) rename(0, false);
    }
};

}  // namespace mprb
