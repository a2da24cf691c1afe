// mprb device kernels (sm_100a).
//
// One frame = a fixed stream of persistent kernels; no host round trips.
//
//   k_eval_tiles<DIM, ROOT>   interval pass over one subdivision level:
//       a warp owns 32 tiles that share one tape (at the root level: 32
//       consecutive top-level tiles; below: half of a parent's 64 children),
//       so the clause stream and the opcode switch are warp-uniform.  Each
//       lane keeps its tile's slot values in shared memory as [slot][lane]
//       (bank-conflict free), records min/max verdicts, classifies the tile,
//       and - if ambiguous - emits its own shortened tape into the arena in
//       the reference's chunked format while the warp walks the tape backward.
//       Child coordinates, the projective transform, the occlusion pre-mask
//       and the TileNode write-back are fused in.
//   k_rank_tiles              occlusion post-mask + compaction of survivors
//   k_upsample_filled         filled image -> next level's image
//   k_eval_pixels             2D float pass: a warp owns one 8x8 tile, two pixels per lane
//   k_eval_voxels             3D float pass: a warp owns one 4x4x4 tile, two voxels per lane
//   k_eval_root<DIM>          root level, clause-parallel over an SSA / levelised root tape
//   k_normals                 per-pixel gradient pass, lanes grouped by tape
//   k_begin_frame             frame setup in one launch: control block, root tape to cell 0 of the arena, image clears
//   k_preload_tiles, k_heat_finish   brute-force frames and the work meter (analysis variants)
//
// With slot renaming (> 32 slot ids) the float pass runs the generated G = 1 loop on the renamed chunk; clauses
// that touch a row beyond the shared-memory rows come back marked kOpBounce and take the C++ accessors.
// Without slot renaming (<= 32 slot ids) the clause loops of the interval and float passes are
// generated PTX (interval_loop_ptx.inc, float_loop_ptx.inc; tools/gen_*_loop.py): one indexed
// branch per clause, operand forwarding and dead-store elision driven by hint bits that
// annotate_chunk() writes into the spare bits of the opcode byte when a chunk lands in shared
// memory.  HEAT = true instantiations are the work-metering variants; ordinary frames never run
// them.
//
// Behaviour (what is computed, bit for bit) follows the reference kernels in
// reference src/context.cu; line numbers are cited at each step.  How it is
// scheduled (work units, memory layout, fusion, queues) is specific to this
// implementation.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "deriv.cuh"
#include "ival.cuh"
#include "kernels.cuh"
#include "tape_stream.cuh"

namespace mprb {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// One lane claims the next work item for its warp.
__device__ __forceinline__ int warp_next(int32_t* head) {
    int v = 0;
    if (lane_id() == 0) v = atomicAdd(head, 1);
    return __shfl_sync(kFull, v, 0);
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned v) {
    return __reduce_add_sync(kFull, v);
}

__device__ __forceinline__ uint64_t make_jump(int32_t delta) {
    return uint64_t(OP_JUMP) | (uint64_t(uint32_t(delta)) << 32);
}

// Per-lane set of live slots for the tape push.  WIDE: 128 bits in four registers (renamed
// kernels: any slot id the reference allows).  Narrow: one register - kernels without renaming
// only ever see tapes with at most 32 slot ids (use_remap), and the push spends a third of its
// instructions on this set otherwise.
template <bool WIDE> struct SlotSet;
template <> struct SlotSet<true> {
    uint32_t w0, w1, w2, w3;
    __device__ __forceinline__ void clear() { w0 = w1 = w2 = w3 = 0; }
    __device__ __forceinline__ bool test(uint32_t s) const {
        const uint32_t k = s >> 5;
        const uint32_t v = (k == 0) ? w0 : (k == 1) ? w1 : (k == 2) ? w2 : w3;
        return (v >> (s & 31)) & 1u;
    }
    __device__ __forceinline__ void set(uint32_t s) {
        const uint32_t k = s >> 5, m = 1u << (s & 31);
        w0 |= (k == 0) ? m : 0u;
        w1 |= (k == 1) ? m : 0u;
        w2 |= (k == 2) ? m : 0u;
        w3 |= (k == 3) ? m : 0u;
    }
    __device__ __forceinline__ void reset(uint32_t s) {
        const uint32_t k = s >> 5, m = ~(1u << (s & 31));
        w0 &= (k == 0) ? m : ~0u;
        w1 &= (k == 1) ? m : ~0u;
        w2 &= (k == 2) ? m : ~0u;
        w3 &= (k == 3) ? m : ~0u;
    }
};
template <> struct SlotSet<false> {
    uint32_t w0;
    __device__ __forceinline__ void clear() { w0 = 0; }
    __device__ __forceinline__ bool test(uint32_t s) const { return (w0 >> s) & 1u; }
    __device__ __forceinline__ void set(uint32_t s) { w0 |= 1u << s; }
    __device__ __forceinline__ void reset(uint32_t s) { w0 &= ~(1u << s); }
};

// a*x + b*y + c*z + d with the reference build's rounding sequence
// (SASS of calculate_voxels / eval_pixels_d: FMUL b*y; FFMA a*x+.; FFMA c*z+.; FADD d).
__device__ __forceinline__ float dot3(float a, float x, float b, float y, float c, float z, float d) {
    return __fadd_rn(__fmaf_rn(c, z, __fmaf_rn(a, x, __fmul_rn(b, y))), d);
}
// a*x + b*y + c  (SASS of calculate_pixels: FMUL b*y; FFMA a*x+.; FADD c)
__device__ __forceinline__ float dot2(float a, float x, float b, float y, float c) {
    return __fadd_rn(__fmaf_rn(a, x, __fmul_rn(b, y)), c);
}
// ((p + 0.5) * recip - 0.5) * 2 -> FADD; FFMA; FADD t+t  (context.cu:734-736)
__device__ __forceinline__ float sample_coord(int p, float recip) {
    const float t = __fmaf_rn(__fadd_rn(float(p), 0.5f), recip, -0.5f);
    return __fadd_rn(t, t);
}

// (p / tiles_per_side - 0.5) * 2  (context.cu:93-98)
__device__ __forceinline__ float tile_edge(int p, float ftps) {
    return __fmul_rn(__fsub_rn(__fdiv_rn(float(p), ftps), 0.5f), 2.0f);
}

// Shared-memory slot rows addressed by 32-bit shared-space addresses: slot s of this lane
// is at base + s * (32 lanes * sizeof value).  The clause word already holds each slot id in
// its own byte, so `s * 256` is a mask (and a shift) away - no multiply, no 64-bit math.
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f2(uint32_t addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ uint32_t off_out2(uint32_t w) { return w & 0xff00u; }            // out * 256
__device__ __forceinline__ uint32_t off_lhs2(uint32_t w) { return (w >> 8) & 0xff00u; }     // lhs * 256
__device__ __forceinline__ uint32_t off_rhs2(uint32_t w) { return (w >> 16) & 0xff00u; }    // rhs * 256

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Where one lane keeps its slot values.  `off` is always (slot or row) * 256, i.e. a masked
// byte of the clause word.
//   REMAP = false: rows of 32 lanes in shared memory, one per slot id - conflict free, no cache
//                  misses; used while the root tape's slot count leaves enough occupancy.
//   REMAP = true:  the stream has renamed ids to dense rows (tape_stream.cuh); the first
//                  kRemapRows rows are shared-memory rows, later ones spill to local memory.
template <bool REMAP> struct Slots2 {
    uint32_t base;
    uint32_t limit;      // REMAP: rows * 256 held in shared memory (rows >= kRemapRowsMin)
    float2 loc[REMAP ? 128 - kRemapRowsMin : 1];
    __device__ __forceinline__ float2 ld(uint32_t off) const {
        if (REMAP && off >= limit) return loc[(off - limit) >> 8];
        return lds_f2(base + off);
    }
    __device__ __forceinline__ void st(uint32_t off, float2 v) {
        if (REMAP && off >= limit) loc[(off - limit) >> 8] = v;
        else sts_f2(base + off, v);
    }
};
// Normal pass: float4 values; LOCAL = true keeps them in per-thread local memory instead.
template <bool LOCAL> struct Slots4 {
    uint32_t base;
    float4 loc[LOCAL ? 128 : 1];
    __device__ __forceinline__ float4 ld(uint32_t off) const { return LOCAL ? loc[off >> 8] : lds_f4(base + (off << 1)); }
    __device__ __forceinline__ void st(uint32_t off, float4 v) { if (LOCAL) loc[off >> 8] = v; else sts_f4(base + (off << 1), v); }
};

// Bytes of dynamic shared memory a tape-walking CTA needs: slot rows + one chunk stream per warp.

template <int N> struct MatOf;
template <> struct MatOf<2> { typedef Mat3 type; };
template <> struct MatOf<3> { typedef Mat4 type; };

}  // namespace

// Opcode classes for the hints (bit i = opcode i).  FAST: handled inside the PTX loop.
constexpr uint32_t kFastOps = 0x3ffffc1cu;      // everything but END, JUMP, SIN, COS, ASIN, ACOS, ATAN
constexpr uint32_t kUsesLhs = (1u << OP_SQUARE) | (1u << OP_SQRT) | (1u << OP_NEG) | (1u << OP_ABS) | (1u << OP_ADD_LI) |
                              (1u << OP_EXP) | (1u << OP_LOG) |
                              (1u << OP_ADD_LR) | (1u << OP_MUL_LI) | (1u << OP_MUL_LR) | (1u << OP_MIN_LI) |
                              (1u << OP_MIN_LR) | (1u << OP_MAX_LI) | (1u << OP_MAX_LR) | (1u << OP_SUB_LI) |
                              (1u << OP_SUB_LR) | (1u << OP_DIV_LI) | (1u << OP_DIV_LR) | (1u << OP_COPY_LHS);
constexpr uint32_t kUsesRhs = (1u << OP_ADD_LR) | (1u << OP_MUL_LR) | (1u << OP_MIN_LR) | (1u << OP_MAX_LR) |
                              (1u << OP_SUB_IR) | (1u << OP_SUB_LR) | (1u << OP_DIV_IR) | (1u << OP_DIV_LR) |
                              (1u << OP_COPY_RHS);
static_assert(kFastOps == (((1u << 30) - 1) & ~((1u << OP_END) | (1u << OP_JUMP) | (1u << OP_SIN) | (1u << OP_COS) |
                                               (1u << OP_ASIN) | (1u << OP_ACOS) | (1u << OP_ATAN))), "kFastOps");

// Writes the forwarding hints (see tools/gen_float_loop.py, gen_interval_loop.py) into the opcode
// bytes of a freshly arrived raw chunk; FAST / LHS / RHS are the opcode classes of the loop that
// will run it.  A hint only ever relates a cell to its neighbour in memory, and the loop
// only ever runs a cell right after that neighbour (entries land after an END / JUMP / slow
// cell, which never forward), so stale cells elsewhere in the chunk do not matter.
template <uint32_t FAST, uint32_t LHS, uint32_t RHS, int SHIFT = 0>
__device__ __forceinline__ void annotate_chunk(uint32_t buf)
{
    const int lane = threadIdx.x & 31;
    uint32_t word[2];
    #pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int j = lane + 32 * k;
        const uint32_t w = lds_u32(buf + j * 8);
        const uint32_t wp = j > 0 ? lds_u32(buf + (j - 1) * 8) : 0u;
        const uint32_t wn = j < kChunk - 1 ? lds_u32(buf + (j + 1) * 8) : 0u;
        auto fast = [](uint32_t x) { return (x & 0xe0u) == 0 && ((FAST >> (x & 31u)) & 1u); };
        const uint32_t op = w & 31u;
        const uint32_t prev_out = (wp >> 8) & 0xff;
        uint32_t flags = 0;
        if (fast(w)) {
            if (fast(wp)) {
                if (((LHS >> op) & 1u) && ((w >> 16) & 0xff) == prev_out) flags |= 0x20;
                if (((RHS >> op) & 1u) && (w >> 24) == prev_out) flags |= 0x40;
            }
            if (fast(wn) && ((wn >> 8) & 0xff) == ((w >> 8) & 0xff)) flags |= 0x80;
        }
        // SHIFT > 0 (float pass, G = 2 or 4 tiles per warp): slot ids become row offsets in units of
        // 256 bytes - id * G - so that the walkers' `byte * 256` addressing needs no multiply.  Ids
        // are below 32 here (no renaming), so the bytes cannot run into each other.
        word[k] = (w & 0xff) | flags | ((w & 0xffffff00u) << SHIFT);
    }
    __syncwarp();
    if (SHIFT == 0) {
        sts_u8(buf + lane * 8, word[0]);
        sts_u8(buf + (lane + 32) * 8, word[1]);
    } else {
        sts_u32(buf + lane * 8, word[0]);
        sts_u32(buf + (lane + 32) * 8, word[1]);
    }
    __syncwarp();
}

// Opcode classes of the interval loop (tools/gen_interval_loop.py prints them).
constexpr uint32_t kIvFastOps = 0x39ffec7cu, kIvUsesLhs = 0x11bfec1cu, kIvUsesRhs = 0x20d54000u;

// The forward clause loop of the interval pass in PTX (generated: tools/gen_interval_loop.py ->
// interval_loop_ptx.inc); same scheme as run_float_clauses below.  Runs clauses from the cell after
// `cp` until one it does not handle (END, JUMP, DIV_IMM_RHS, DIV_LHS_RHS, ASIN, ACOS, ATAN, LOG) and
// returns with cp on that cell; min / max verdicts are recorded into cw / choices[] on the way.
__device__ __forceinline__ void run_interval_clauses(uint32_t& cp, uint32_t& w, uint32_t& imm, uint32_t sb, uint32_t& cw,
                                                     uint32_t& n_choice, uint32_t& any_choice, const uint32_t* choices)
{
    asm volatile(
#include "interval_loop_ptx.inc"
        : "+r"(cp), "=&r"(w), "=&r"(imm), "+r"(cw), "+r"(n_choice), "+r"(any_choice)
        : "r"(sb), "l"(choices)
        : "memory");
}

////////////////////////////////////////////////////////////////////////////////
// Interval pass

// HEAT = true is the work-metering variant behind render*_heatmap: a separate instantiation so
// that ordinary frames carry none of its registers or branches.
template <int DIM, bool ROOT, bool REMAP, bool HEAT = false>
__global__ void __launch_bounds__(kEvalThreads)
k_eval_tiles(const EvalTilesArgs a, const typename MatOf<DIM>::type mat)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    typedef TapeStream<REMAP> Stream;
    Stream ts;
    ts.init(s_dyn + warp * Stream::stride(), a.arena, a.arena_cap);
    const int n_rows = a.n_rows;       // shared-memory value rows per warp (= slot count unless REMAP)
    Slots2<REMAP> slots;
    slots.base = smem_addr(s_dyn + kEvalWarps * Stream::stride()) + (warp * n_rows * 32 + lane) * 8 - (REMAP ? 0 : 256);
    slots.limit = uint32_t(n_rows) * 256u;

    uint64_t* const arena = a.arena;
    uint32_t choices[kMaxChoices / 16];   // 2 bits per recorded min/max verdict
    // statistics accumulate per warp and are flushed once (one atomic per counter per warp)
    unsigned long long st_tiles = 0, st_cells = 0, st_ptiles = 0, st_pcells = 0, st_kept = 0, st_written = 0;

    const int n_items = ROOT ? (a.count0 + 31) / 32 : 2 * min(*a.n_parents, a.tiles_cap / 64);
    const uint32_t tps = a.tps;
    const uint64_t root_hdr = arena[0];   // axis slots always come from the root header
                                          // (context.cu:211-213)
    for (;;) {
        const int item = warp_next(a.queue);
        if (item >= n_items) break;

        // ---- which tile does this lane own? -------------------------------------
        int tile_index;     // where its TileNode lives in a.tiles
        int tape;           // arena index of the tape header (warp-uniform)
        int sx, sy, sz = 0;
        bool valid = true, inband = true;
        if (ROOT) {
            int t = item * 32 + lane;
            if (DIM == 3) t = a.count0 - 1 - t;      // highest z first: better culling
            valid = (t >= 0) && (t < a.count0);
            tile_index = t;
            tape = 0;
            const int tt = valid ? t : 0;
            sx = tt % tps;
            sy = (tt / tps) % tps;
            if (DIM == 3) sz = (tt / tps) / tps;
            inband = (sy >= a.row_begin) && (sy < a.row_end) && ((sy + a.col_step * sx) % a.row_mod == a.row_rem);
        } else {
            const int rank = item >> 1;
            // a small level: parents with a clause-parallel plan belong to k_eval_sub
            if (a.plan_of && (n_items >> 1) <= a.sub_max_parents && a.plan_of[a.pactive[rank]] >= 0) continue;
            // 3D: the upper half of the children (z = 2, 3) goes first, for the same reason
            const int sub = DIM == 3 ? ((((item & 1) ^ 1) << 5) | lane) : (((item & 1) << 5) | lane);
            const TileNode parent = a.ptiles[a.pactive[rank]];
            tape = parent.tape;
            tile_index = rank * 64 + sub;
            const int pp = parent.position;
            const int px = pp % a.ptps, py = (pp / a.ptps) % a.ptps;
            if (DIM == 3) {   // 4x4x4 children, sub = x + 4y + 16z (context.cu:579-588)
                const int pz = (pp / a.ptps) / a.ptps;
                sx = px * 4 + (sub & 3);
                sy = py * 4 + ((sub >> 2) & 3);
                sz = pz * 4 + (sub >> 4);
            } else {          // 8x8 children, sub = x + 8y (context.cu:612-617)
                sx = px * 8 + (sub & 7);
                sy = py * 8 + (sub >> 3);
            }
        }
        const int position = sx + sy * tps + (DIM == 3 ? sz * tps * tps : 0);
        const int img_index = sx + sy * tps;

        // Occlusion pre-mask (first mask_filled_tiles, context.cu:1335, :486-494)
        bool alive = valid && inband;
        if (DIM == 3 && alive) {
            if (__ldcg(&a.image[img_index]) > sz) alive = false;
        }
        if (!__any_sync(kFull, alive)) {
            if (valid) {
                a.tiles[tile_index].position = -1;
                a.tiles[tile_index].tape = tape;
                a.tiles[tile_index].next = -1;
            }
            continue;
        }

        // ---- tile box -> transformed intervals (context.cu:78-159) ---------------
        const uint32_t h = ts.begin_tape(uint32_t(root_hdr));   // axis slots (renamed to rows if REMAP)
        {
            const float ftps = float(tps);
            const ival ix = iv(tile_edge(sx, ftps), tile_edge(sx + 1, ftps));
            const ival iy = iv(tile_edge(sy, ftps), tile_edge(sy + 1, ftps));
            ival X, Y, Z;
            if (DIM == 3) {
                const ival iz = iv(tile_edge(sz, ftps), tile_edge(sz + 1, ftps));
                const float* m = mat.d;   // column major: m(r, c) = m[c * 4 + r]
                ival r[4];
                #pragma unroll
                for (int i = 0; i < 4; ++i) {
                    r[i] = iv_add(iv_add(iv_add(iv_mul(ix, m[i]), iv_mul(iy, m[4 + i])),
                                         iv_mul(iz, m[8 + i])), m[12 + i]);
                }
                X = iv_div(r[0], r[3]);
                Y = iv_div(r[1], r[3]);
                Z = iv_div(r[2], r[3]);
            } else {
                const float* m = mat.d;   // m(r, c) = m[c * 3 + r]
                ival r[3];
                #pragma unroll
                for (int i = 0; i < 3; ++i) {
                    r[i] = iv_add(iv_add(iv_mul(ix, m[i]), iv_mul(iy, m[3 + i])), m[6 + i]);
                }
                X = iv_div(r[0], r[2]);
                Y = iv_div(r[1], r[2]);
                Z = iv(a.z, a.z);
            }
            // an axis the tape never reads has slot id 0 (context.cu:210-213 writes slot 0 there,
            // harmlessly; here id 0 has no row at all)
            if (REMAP || off_out2(h)) slots.st(off_out2(h), X);
            if (REMAP || off_lhs2(h)) slots.st(off_lhs2(h), Y);
            if (REMAP || off_rhs2(h)) slots.st(off_rhs2(h), Z);
        }

        // ---- forward walk (context.cu:223-287) -------------------------------------
        // Every 64-cell chunk ends in a JUMP or the end cell (see tape_stream.cuh), so the hot
        // path is a running shared-memory pointer; chunk switches happen only at JUMP cells.
        // Without slot renaming the loop proper is generated PTX (run_interval_clauses): it comes
        // back here only for END / JUMP and the few opcodes it leaves to the C++ switch below.
        if (ts.fetch(tape) && !REMAP) annotate_chunk<kIvFastOps, kIvUsesLhs, kIvUsesRhs>(ts.buf);
        uint32_t cp = ts.rd + ((tape & (kChunk - 1)) << 3);
        uint32_t seg = cp;         // where the current chunk segment started (for statistics)
        uint32_t n_choice = 0;     // warp-uniform: how many min/max clauses seen so far
        uint32_t cw = 0;           // verdict word under construction
        uint32_t any_choice = 0;   // nonzero once a min/max was decided for this lane's tile
        unsigned cells = 0;
        uint2 d;
        for (;;) {
            if (REMAP) {
                cp += 8;
                d = lds_u2(cp);
            } else {
                run_interval_clauses(cp, d.x, d.y, slots.base, cw, n_choice, any_choice, choices);
            }
            const uint32_t w = d.x;
            const uint32_t op = w & 0xff;
            if (op <= OP_JUMP) {
                cells += (cp - seg) >> 3;
                if (op == OP_END) { --cells; break; }
                const int t = ts.base + int((cp - ts.rd) >> 3) + int32_t(d.y);
                if (ts.fetch(t) && !REMAP) annotate_chunk<kIvFastOps, kIvUsesLhs, kIvUsesRhs>(ts.buf);
                cp = ts.rd + ((t & (kChunk - 1)) << 3);
                seg = cp;
                continue;
            }
            const float imm = __uint_as_float(d.y);
            // without renaming slot id 0 ("no operand") has no row: nothing is loaded for it
            const ival L = (REMAP || off_lhs2(w)) ? slots.ld(off_lhs2(w)) : iv(0.0f, 0.0f);
            const ival R = (REMAP || off_rhs2(w)) ? slots.ld(off_rhs2(w)) : iv(0.0f, 0.0f);
            ival o;
            if (!REMAP) {
                // the opcodes the PTX loop hands back
                switch (op) {
                    case OP_ASIN:   o = iv_asin(L); break;
                    case OP_ACOS:   o = iv_acos(L); break;
                    case OP_ATAN:   o = iv_atan(L); break;
                    case OP_LOG:    o = iv_log(L); break;
                    case OP_DIV_IR: o = iv_div(imm, R); break;
                    default:        o = iv_div(L, R); break;      // OP_DIV_LR
                }
                slots.st(off_out2(w), o);
                continue;
            }
            int c = 0;
            switch (op) {
                case OP_SQUARE: o = iv_square(L); break;
                case OP_SQRT:   o = iv_sqrt(L); break;
                case OP_NEG:    o = iv_neg(L); break;
                case OP_SIN:    o = iv_sin(L); break;
                case OP_COS:    o = iv_cos(L); break;
                case OP_ASIN:   o = iv_asin(L); break;
                case OP_ACOS:   o = iv_acos(L); break;
                case OP_ATAN:   o = iv_atan(L); break;
                case OP_EXP:    o = iv_exp(L); break;
                case OP_ABS:    o = iv_abs(L); break;
                case OP_LOG:    o = iv_log(L); break;
                case OP_ADD_LI: o = iv_add(L, imm); break;
                case OP_ADD_LR: o = iv_add(L, R); break;
                case OP_MUL_LI: o = iv_mul(L, imm); break;
                case OP_MUL_LR: o = iv_mul(L, R); break;
                case OP_MIN_LI: o = iv_min(L, iv(imm, imm), c); break;
                case OP_MIN_LR: o = iv_min(L, R, c); break;
                case OP_MAX_LI: o = iv_max(L, iv(imm, imm), c); break;
                case OP_MAX_LR: o = iv_max(L, R, c); break;
                case OP_SUB_LI: o = iv_sub(L, imm); break;
                case OP_SUB_IR: o = iv_sub(imm, R); break;
                case OP_SUB_LR: o = iv_sub(L, R); break;
                case OP_DIV_LI: o = iv_div(L, imm); break;
                case OP_DIV_IR: o = iv_div(imm, R); break;
                case OP_DIV_LR: o = iv_div(L, R); break;
                case OP_COPY_IMM: o = iv(imm, imm); break;
                case OP_COPY_LHS: o = L; break;
                case OP_COPY_RHS: o = R; break;
                default: o = L; break;
            }
            if (op >= OP_MIN_LI && op <= OP_MAX_LR) {
                // Verdicts past kMaxChoices are not recorded (context.cu:257-259)
                cw |= uint32_t(c) << ((n_choice & 15) * 2);
                if ((n_choice & 15) == 15) {
                    if (n_choice < kMaxChoices) choices[n_choice >> 4] = cw;
                    cw = 0;
                }
                ++n_choice;
                any_choice |= uint32_t(c);
            }
            slots.st(off_out2(w), o);
        }
        if ((n_choice & 15) && n_choice < kMaxChoices) choices[n_choice >> 4] = cw;
        const uint2 end_raw = lds_u2(ts.buf + (cp - ts.rd));              // as stored in the arena
        const uint64_t end_cell = uint64_t(end_raw.x) | (uint64_t(end_raw.y) << 32);   // {0, result slot}
        const uint32_t i_result = (d.x >> 8) & 0xff;
        const ival result = slots.ld(off_out2(d.x));

        // ---- classify (context.cu:289-321) -----------------------------------------
        int out_position = -1;
        bool pushing = false;
        if (alive) {
            if (result.x > 0.0f) {
                // empty
            } else if (DIM == 3 && __ldcg(&a.image[img_index]) > sz) {
                // hidden below a filled tile that landed meanwhile
            } else if (result.y < 0.0f) {
                if (DIM == 3) atomicMax(&a.image[img_index], sz);
                else a.image[img_index] = 1;
            } else {
                out_position = position;
                pushing = any_choice != 0;
            }
        }
        int out_tape = tape;

        // ---- statistics ----------------------------------------------------------------
        {
            const unsigned n_alive = __popc(__ballot_sync(kFull, alive));
            st_tiles += n_alive;
            st_cells += (unsigned long long)n_alive * cells;
        }
        // Work meter: cells this tile is charged with (context.cu:1622-1633).  The root tape is
        // charged by its clause count: its chunked layout here carries JUMP cells the
        // reference's contiguous copy does not.
        unsigned heat_cells = 0;
        if (HEAT && alive) heat_cells = (tape == 0) ? unsigned(a.n_root) : cells;

        // ---- tape push: backward mark & sweep (context.cu:323-458) ---------------------
        if (__any_sync(kFull, pushing)) {
            const int cap = a.arena_cap;
            SlotSet<REMAP> live;
            live.clear();
            live.set(i_result);
            int o_idx = 0, o_off = 0;
            unsigned kept = 0;
            // Tiles of this warp that recorded the same verdicts get the same shortened tape (the
            // push is a function of the parent tape and the verdicts alone), so only the first lane
            // of each such class writes it and the others point at its copy: the logical tape of
            // every tile is what the reference's per-thread push produces (context.cu:323-458), the
            // arena holds it once, and the float pass can walk it once for several tiles.
            const bool pushed0 = pushing;
            int writer = lane;
            {
                unsigned cls = __ballot_sync(kFull, pushing);
                const int n_words = (min(n_choice, uint32_t(kMaxChoices)) + 15) >> 4;
                for (int i = 0; i < n_words; ++i) cls &= __match_any_sync(kFull, choices[i]);
                if (pushing) writer = __ffs(cls) - 1;
                pushing = pushing && writer == lane;
            }
            if (pushing) {
                if (*(volatile int32_t*)a.tape_index >= cap) {
                    pushing = false;
                } else {
                    o_idx = atomicAdd(a.tape_index, kChunk);
                    if (o_idx + kChunk >= cap) {
                        pushing = false;
                    } else {
                        o_off = kChunk - 1;
                        arena[o_idx + o_off] = end_cell;
                        kept = 1;
                    }
                }
            }
            unsigned bcells = 0;
            int ci = n_choice;
            int cw_index = -1;
            uint32_t cwb = 0;
            seg = cp;
            uint2 b;
            for (;;) {
                cp -= 8;
                b = lds_u2(cp);
                const uint32_t w = b.x;
                // Without renaming the chunk may carry the forward loop's hint bits (bits 5-7 of
                // the opcode byte, annotate_chunk): they are not part of the tape.
                const uint32_t op = REMAP ? (w & 0xff) : (w & 0x1f);
                if (op <= OP_JUMP) {
                    bcells += (seg - cp) >> 3;
                    if (op == OP_END) { --bcells; break; }
                    const int t = ts.base + int((cp - ts.rd) >> 3) + int32_t(b.y);
                    ts.fetch_back(t);
                    cp = ts.rd + ((t & (kChunk - 1)) << 3);
                    seg = cp;
                    continue;
                }
                const bool has_choice = (op >= OP_MIN_LI && op <= OP_MAX_LR);
                int choice = 0;
                if (has_choice) {
                    --ci;
                    if (ci < kMaxChoices) {
                        if ((ci >> 4) != cw_index) {
                            cw_index = ci >> 4;
                            cwb = choices[cw_index];
                        }
                        choice = (cwb >> ((ci & 15) * 2)) & 3;
                    }
                }
                const uint32_t i_out = (w >> 8) & 0xff, i_lhs = (w >> 16) & 0xff, i_rhs = w >> 24;
                if (pushing && live.test(i_out)) {
                    // Reserve the cell; open a new chunk when this one is used up.
                    --o_off;
                    bool ok = true;
                    if (o_off == 0) {
                        const int prev = o_idx;
                        if (*(volatile int32_t*)a.tape_index >= cap) {
                            ok = false;
                        } else {
                            o_idx = atomicAdd(a.tape_index, kChunk);
                            if (o_idx + kChunk >= cap) {
                                ok = false;
                            } else {
                                o_off = kChunk - 1;
                                const int32_t delta = prev - (o_idx + o_off);
                                arena[o_idx + o_off] = make_jump(delta);   // forward link
                                arena[prev] = make_jump(-delta);           // backward link
                                --o_off;
                                kept += 2;
                            }
                        }
                    }
                    if (!ok) {
                        pushing = false;   // arena exhausted: keep the parent tape
                    } else {
                        live.reset(i_out);
                        const uint2 raw = lds_u2(ts.buf + (cp - ts.rd));     // cell as stored (ids, not rows)
                        const uint64_t d64 = uint64_t(REMAP ? raw.x : (raw.x & ~0xe0u)) | (uint64_t(raw.y) << 32);
                        uint64_t e = d64;
                        bool emit = true;
                        if (choice == 0) {
                            if (i_lhs) live.set(i_lhs);
                            if (i_rhs) live.set(i_rhs);
                        } else if (choice == 1) {
                            live.set(i_lhs);
                            if (i_lhs == i_out) emit = false;
                            else e = (d64 & ~0xffull) | OP_COPY_LHS;
                        } else if (choice == 2) {
                            if (i_rhs) {
                                live.set(i_rhs);
                                if (i_rhs == i_out) emit = false;
                                else e = (d64 & ~0xffull) | OP_COPY_RHS;
                            } else {
                                e = (d64 & ~0xffull) | OP_COPY_IMM;
                            }
                        }
                        if (emit) {
                            arena[o_idx + o_off] = e;
                            ++kept;
                        } else {
                            ++o_off;   // give the reserved cell back
                        }
                    }
                }
            }
            // The walk stopped on the header cell; it goes in front (raw copy).
            if (pushing) {
                --o_off;
                const uint2 hraw = lds_u2(ts.buf + (cp - ts.rd));
                arena[o_idx + o_off] = uint64_t(hraw.x) | (uint64_t(hraw.y) << 32);
                ++kept;
                out_tape = o_idx + o_off;
            }
            {
                st_written += warp_sum(kept);                              // cells that went to the arena
                // the other members of each class take their writer's tape (or, if the arena ran
                // out under it, keep the parent tape like it does)
                const bool w_ok = __shfl_sync(kFull, pushing, writer);
                const int w_tape = __shfl_sync(kFull, out_tape, writer);
                const unsigned w_kept = __shfl_sync(kFull, kept, writer);
                if (pushed0 && writer != lane && w_ok) { out_tape = w_tape; kept = w_kept; }
                const unsigned n_push = __popc(__ballot_sync(kFull, pushed0));
                st_ptiles += n_push;
                st_pcells += (unsigned long long)n_push * bcells;
                st_kept += warp_sum(pushed0 ? kept : 0u);                  // per tile, as the reference writes them
            }
            if (HEAT && pushed0) heat_cells += (tape == 0) ? unsigned(a.n_root) : bcells;   // context.cu:1815-1826
        }
        if (HEAT) {
            // cells / px^2 onto every pixel of the footprint, in units of 1/4096 cell so that the
            // sum is exact and order-independent; the warp spreads one tile at a time (coalesced rows)
            const int px = a.heat_px;
            const int size = int(tps) * px;
            const unsigned long long units = (unsigned long long)heat_cells * unsigned(4096 / (px * px));
            unsigned m = __ballot_sync(kFull, heat_cells != 0);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int bx = __shfl_sync(kFull, sx, src) * px, by = __shfl_sync(kFull, sy, src) * px;
                const unsigned long long u = __shfl_sync(kFull, units, src);
                for (int i = lane; i < px * px; i += 32)
                    atomicAdd(&a.heat[size_t(by + i / px) * size + bx + i % px], u);
            }
        }

        if (valid) {
            a.tiles[tile_index].position = out_position;
            a.tiles[tile_index].tape = out_tape;
            a.tiles[tile_index].next = -1;
        }
    }
    ts.drain();
    if (lane == 0 && st_tiles) {
        atomicAdd(&a.ctl->stats[ST_I_TILES + a.level], st_tiles);
        atomicAdd(&a.ctl->stats[ST_I_CELLS + a.level], st_cells);
        if (st_ptiles) {
            atomicAdd(&a.ctl->stats[ST_P_TILES + a.level], st_ptiles);
            atomicAdd(&a.ctl->stats[ST_P_CELLS + a.level], st_pcells);
            atomicAdd(&a.ctl->stats[ST_P_KEPT + a.level], st_kept);
            atomicAdd(&a.ctl->stats[ST_P_WRITTEN], st_written);
        }
    }
}

////////////////////////////////////////////////////////////////////////////////
// Root level, evaluated clause-parallel.
//
// All top-level tiles run the SAME tape, and there are few of them, so walking
// it one clause at a time (k_eval_tiles<DIM, true>) leaves the machine idle
// behind a dependent chain of thousands of clauses.  Here a group of G threads
// owns ONE tile and evaluates the tape in dependency-level order instead: the
// host turns the root tape into SSA form once per Tape (every operand names the
// clause that produced it, not a reused slot) and buckets clauses by depth, so
// a level is |level| independent interval operations.  Interval results do not
// depend on evaluation order, so every value, verdict and classification is the
// one the serial walk produces.
//
// The shortened tape is produced by the same mark-and-sweep as the serial push
// (context.cu:323-458), expressed on SSA ids: marks propagate from the result
// back through the levels, then kept clauses are compacted IN TAPE ORDER with a
// group-wide prefix sum and written as one contiguous run (header, clauses, end
// cell) - a valid tape in the reference's format that simply needs no JUMP.

// Position of logical tape cell q in the chunk-terminated layout (see tape_stream.cuh); tapes
// of at most 64 cells are stored as they are.
__host__ __device__ __forceinline__ int chunked_index(int q, int n_logical) {
    if (n_logical <= kChunk || q < kChunk - 1) return q;
    const int r = q - (kChunk - 1);
    return kChunk + (r / (kChunk - 2)) * kChunk + 1 + r % (kChunk - 2);
}

__device__ __forceinline__ void group_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One clause of a clause-parallel plan (k_eval_root, k_eval_sub): the interval operations of the serial
// walk (context.cu:223-287), operands already fetched; c = the min / max verdict.
__device__ __forceinline__ ival eval_plan_clause(uint32_t op, const ival Lv, const ival Rv, float imm, int& c) {
    ival o;
    switch (op) {
        case OP_SQUARE: o = iv_square(Lv); break;
        case OP_SQRT:   o = iv_sqrt(Lv); break;
        case OP_NEG:    o = iv_neg(Lv); break;
        case OP_SIN:    o = iv_sin(Lv); break;
        case OP_COS:    o = iv_cos(Lv); break;
        case OP_ASIN:   o = iv_asin(Lv); break;
        case OP_ACOS:   o = iv_acos(Lv); break;
        case OP_ATAN:   o = iv_atan(Lv); break;
        case OP_EXP:    o = iv_exp(Lv); break;
        case OP_ABS:    o = iv_abs(Lv); break;
        case OP_LOG:    o = iv_log(Lv); break;
        case OP_ADD_LI: o = iv_add(Lv, imm); break;
        case OP_ADD_LR: o = iv_add(Lv, Rv); break;
        case OP_MUL_LI: o = iv_mul(Lv, imm); break;
        case OP_MUL_LR: o = iv_mul(Lv, Rv); break;
        case OP_MIN_LI: o = iv_min(Lv, iv(imm, imm), c); break;
        case OP_MIN_LR: o = iv_min(Lv, Rv, c); break;
        case OP_MAX_LI: o = iv_max(Lv, iv(imm, imm), c); break;
        case OP_MAX_LR: o = iv_max(Lv, Rv, c); break;
        case OP_SUB_LI: o = iv_sub(Lv, imm); break;
        case OP_SUB_IR: o = iv_sub(imm, Rv); break;
        case OP_SUB_LR: o = iv_sub(Lv, Rv); break;
        case OP_DIV_LI: o = iv_div(Lv, imm); break;
        case OP_DIV_IR: o = iv_div(imm, Rv); break;
        case OP_DIV_LR: o = iv_div(Lv, Rv); break;
        case OP_COPY_IMM: o = iv(imm, imm); break;
        case OP_COPY_RHS: o = Rv; break;
        default: o = Lv; break;                      // OP_COPY_LHS
    }
    return o;
}

template <int DIM>
__global__ void __launch_bounds__(kRootThreads)
k_eval_root(const EvalRootArgs a, const typename MatOf<DIM>::type mat)
{
    extern __shared__ __align__(16) unsigned char s_root[];
    const int G = a.group;
    const int gid = threadIdx.x / G;
    const int t = threadIdx.x % G;
    const int bar = gid + 1;
    const int n = a.n_clauses;
    const int nv = n + 4;                                   // value ids: 0 none, 1..3 axes, 3+i clause i
    unsigned char* const mine = s_root + size_t(gid) * a.smem_per_tile;
    float2* const V = reinterpret_cast<float2*>(mine);
    uint8_t* const C = mine + size_t(nv) * 8;               // min/max verdict per clause
    uint8_t* const A = C + nv;                              // liveness per value id
    int* const scratch = reinterpret_cast<int*>(mine + a.smem_per_tile - 64);
    uint64_t* const arena = a.arena;
    const uint64_t* __restrict__ const cells = a.cells;     // the Tape's own contiguous cells
    const uint32_t tps = a.tps;

    int tile = blockIdx.x * (kRootThreads / G) + gid;
    if (tile >= a.count0) return;
    if (DIM == 3) tile = a.count0 - 1 - tile;               // highest z first
    const int sx = tile % tps, sy = (tile / tps) % tps, sz = DIM == 3 ? (tile / tps) / tps : 0;
    const int img_index = sx + sy * tps;

    // Occlusion pre-mask.  The image changes under us (other tiles' atomicMax), so one
    // thread looks and the whole group follows its answer.
    if (t == 0) {
        if (a.plans) a.plan_of[tile] = -1;
        bool al = (sy >= a.row_begin) && (sy < a.row_end) && ((sy + a.col_step * sx) % a.row_mod == a.row_rem);
        if (DIM == 3 && al && __ldcg(&a.image[img_index]) > sz) al = false;
        scratch[12] = al;
        if (!al) {
            a.tiles[tile].position = -1;
            a.tiles[tile].tape = 0;
            a.tiles[tile].next = -1;
        }
    }
    group_sync(bar, G);
    if (!scratch[12]) return;

    if (t == 0) {                                            // tile box -> axis intervals
        const float ftps = float(tps);
        const ival ix = iv(tile_edge(sx, ftps), tile_edge(sx + 1, ftps));
        const ival iy = iv(tile_edge(sy, ftps), tile_edge(sy + 1, ftps));
        const float* m = mat.d;
        if (DIM == 3) {
            const ival iz = iv(tile_edge(sz, ftps), tile_edge(sz + 1, ftps));
            ival r[4];
            #pragma unroll
            for (int i = 0; i < 4; ++i)
                r[i] = iv_add(iv_add(iv_add(iv_mul(ix, m[i]), iv_mul(iy, m[4 + i])), iv_mul(iz, m[8 + i])), m[12 + i]);
            V[1] = iv_div(r[0], r[3]);
            V[2] = iv_div(r[1], r[3]);
            V[3] = iv_div(r[2], r[3]);
        } else {
            ival r[3];
            #pragma unroll
            for (int i = 0; i < 3; ++i)
                r[i] = iv_add(iv_add(iv_mul(ix, m[i]), iv_mul(iy, m[3 + i])), m[6 + i]);
            V[1] = iv_div(r[0], r[2]);
            V[2] = iv_div(r[1], r[2]);
            V[3] = iv(a.z, a.z);
        }
        V[0] = iv(0.0f, 0.0f);
        scratch[0] = 0;                                      // any verdict decided?
    }
    for (int i = t; i < nv; i += G) A[i] = 0;
    group_sync(bar, G);

    // ---- forward: one dependency level at a time ---------------------------------------
    // (every tile of the launch reads the same schedule, so it comes out of L1: fetching it levels ahead was
    // measured and only cost instructions)
    bool any_choice = false;
    const int32_t* __restrict__ const ls = a.level_start;
    const int nl = a.n_levels;
    for (int L = 0; L < nl; ++L) {
        const int k_end = ls[L + 1];
        for (int k = ls[L] + t; k < k_end; k += G) {
            const RootClause rc = a.sched[k];
            const uint32_t op = rc.op_idx & 0xff;
            const uint32_t idx = rc.op_idx >> 12;
            const float imm = rc.imm;
            const ival Lv = V[rc.lsrc];
            const ival Rv = V[rc.rsrc];
            int c = 0;
            const ival o = eval_plan_clause(op, Lv, Rv, imm, c);
            V[3 + idx] = o;
            if (op >= OP_MIN_LI && op <= OP_MAX_LR) {
                any_choice |= (c != 0);
                // verdicts past the reference's 4096-entry record count as "undecided" when
                // the tape is shortened, but still make the tile eligible (context.cu:254-263)
                C[idx] = (rc.op_idx & 0x100u) ? 0 : uint8_t(c);
            }
        }
        group_sync(bar, G);
    }
    if (any_choice) scratch[0] = 1;                          // benign same-value race
    const ival result = V[a.result_v];
    group_sync(bar, G);
    any_choice = scratch[0] != 0;

    // ---- classify (context.cu:289-321) ---------------------------------------------------
    int out_position = -1;
    bool pushing = false;
    if (result.x > 0.0f) {
        // empty
    } else if (DIM == 3 && __ldcg(&a.image[img_index]) > sz) {
        // hidden (the value may change concurrently, so let one thread decide for the group)
    } else if (result.y < 0.0f) {
        if (t == 0) {
            if (DIM == 3) atomicMax(&a.image[img_index], sz);
            else a.image[img_index] = 1;
        }
    } else {
        out_position = tile;
        pushing = any_choice;
    }
    if (DIM == 3) {   // make the racy "hidden" test group-uniform: thread 0's view wins
        if (t == 0) { scratch[1] = out_position; scratch[2] = pushing; }
        group_sync(bar, G);
        out_position = scratch[1];
        pushing = scratch[2] != 0;
    }
    int out_tape = 0;

    if (pushing) {
        // ---- mark: result -> operands, top level first ------------------------------------
        // KS[k] = clause at schedule position k stays in the shortened tape (the values are done with,
        // the flags take the last eighth of their room; a clause's liveness is final when its level is swept)
        uint8_t* const KS = mine + size_t(nv) * 6;
        const bool want_plan = a.plans != nullptr;             // frames that write no plans skip the bookkeeping
        if (t == 0) A[a.result_v] = 1;
        group_sync(bar, G);
        for (int L = nl - 1; L >= 0; --L) {
            const int k_end = ls[L + 1];
            for (int k = ls[L] + t; k < k_end; k += G) {
                const RootClause rc = a.sched[k];
                const uint32_t op = rc.op_idx & 0xff;
                const uint32_t idx = rc.op_idx >> 12;
                if (!A[3 + idx]) {
                    if (want_plan) KS[k] = 0;
                    continue;
                }
                const int c = (op >= OP_MIN_LI && op <= OP_MAX_LR) ? C[idx] : 0;
                // in the shortened tape unless the verdict's operand already sits in the output slot
                if (want_plan) KS[k] = ((c == 1 && (rc.op_idx & 0x200u)) || (c == 2 && (rc.op_idx & 0x400u))) ? 0 : 1;
                if (c == 0) { A[rc.lsrc] = 1; A[rc.rsrc] = 1; }
                else if (c == 1) { A[rc.lsrc] = 1; }
                else { A[rc.rsrc] = 1; }                     // rsrc == 0 for immediate forms
            }
            group_sync(bar, G);
        }
        // ---- sweep: compact kept clauses in tape order ---------------------------------------
        // (a live min / max whose verdict names an operand already sitting in the output slot is dropped,
        // context.cu:404-447)
        const int per = (n + G - 1) / G;
        const int i_begin = 1 + t * per, i_end = min(n + 1, i_begin + per);
        int kept = 0;
        for (int i = i_begin; i < i_end; ++i) {
            if (!A[3 + i]) continue;
            const uint32_t w = uint32_t(cells[i]);
            const uint32_t op = w & 0xff, i_out = (w >> 8) & 0xff, i_lhs = (w >> 16) & 0xff, i_rhs = w >> 24;
            const int c = (op >= OP_MIN_LI && op <= OP_MAX_LR) ? C[i] : 0;
            const bool dropped = (c == 1 && i_lhs == i_out) || (c == 2 && i_rhs != 0 && i_rhs == i_out);
            kept += dropped ? 0 : 1;
        }
        // group-wide exclusive scan of `kept`
        int incl = kept;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, o);
            if ((t & 31) >= o) incl += v;
        }
        int* const wsum = scratch + 4;                        // up to 8 warps per group
        if ((t & 31) == 31) wsum[t >> 5] = incl;
        group_sync(bar, G);
        int offset = incl - kept, total = 0;
        for (int wv = 0; wv < G / 32; ++wv) {
            const int s = wsum[wv];
            if (wv < (t >> 5)) offset += s;
            total += s;
        }
        // Logical cell q (0 = header, 1..total = clauses, total + 1 = end cell) goes to chunk
        // layout position chunked_index(q): chunk 0 holds q = 0..62, every later chunk holds 62
        // cells at offsets 1..62 between a back link (cell 0) and a forward link (cell 63).
        const int n_logical = total + 2;
        const int n_chunks = n_logical <= kChunk ? 1 : 1 + (n_logical - 63 + 61) / 62;
        // A plan for k_eval_sub goes with the tape when it is small enough for that kernel's per-tile
        // shared memory and the level can still turn out small (max_plans parents at most).
        const int plan_nv = 4 + total;
        const int plan_words = kPlanHeader + ((a.n_levels + 1 + 3) & ~3) + 5 * total;      // clauses + (at most) as many extra edges
        const bool plan_fits = a.plans != nullptr && total <= kPlanMaxClauses && total >= a.plan_min && nv >= 4 * G + 8 &&
                               sub_need_bytes(plan_nv, a.n_levels) <= a.sub_slice;
        if (t == 0) {
            const int need = n_chunks * kChunk;
            int base = -1;
            if (*(volatile int32_t*)a.tape_index < a.arena_cap) {
                base = atomicAdd(a.tape_index, need);
                if (base + need >= a.arena_cap) base = -1;     // arena exhausted: keep the root tape
            }
            scratch[3] = base;
            int pb = -1;
            if (plan_fits && base >= 0 && *(volatile int32_t*)a.plan_count < a.max_plans &&
                atomicAdd(a.plan_count, 1) < a.max_plans) {
                pb = atomicAdd(a.plan_cursor, plan_words);
                if (pb + plan_words > a.plan_cap) pb = -1;
            }
            scratch[13] = pb;
        }
        group_sync(bar, G);
        const int base = scratch[3];
        const int pb = scratch[13];
        // N[v] = value id in the shortened tape's plan (3 + its cell number), kNotInTape for a live clause
        // that was dropped; the values themselves are done with, N takes their place.
        uint16_t* const N = reinterpret_cast<uint16_t*>(mine);      // 3 + total <= kPlanMaxClauses + 3 where it is used
        constexpr uint32_t kNotInTape = 0xffffu;
        if (base >= 0) {
            int q = 1 + offset;
            for (int i = i_begin; i < i_end; ++i) {
                if (!A[3 + i]) continue;
                uint64_t d = cells[i];
                const uint32_t w = uint32_t(d);
                const uint32_t op = w & 0xff, i_out = (w >> 8) & 0xff, i_lhs = (w >> 16) & 0xff, i_rhs = w >> 24;
                const int c = (op >= OP_MIN_LI && op <= OP_MAX_LR) ? C[i] : 0;
                bool dropped = false;
                if (c == 1) {
                    if (i_lhs == i_out) dropped = true;
                    d = (d & ~0xffull) | OP_COPY_LHS;
                } else if (c == 2) {
                    if (i_rhs != 0 && i_rhs == i_out) dropped = true;
                    d = (d & ~0xffull) | (i_rhs ? OP_COPY_RHS : OP_COPY_IMM);
                }
                if (dropped) {
                    if (pb >= 0) N[3 + i] = uint16_t(kNotInTape);
                    continue;
                }
                arena[base + chunked_index(q, n_logical)] = d;
                if (pb >= 0) N[3 + i] = uint16_t(3 + q);
                ++q;
            }
            if (t == 0) {
                arena[base] = cells[0];                                              // header
                arena[base + chunked_index(total + 1, n_logical)] = cells[n + 1];    // end cell
            }
            for (int c = t; c < n_chunks; c += G) {                                  // chunk links
                if (c > 0) arena[base + c * kChunk] = make_jump(-1);
                if (c + 1 < n_chunks) arena[base + c * kChunk + kChunk - 1] = make_jump(1);
            }
            out_tape = base;
        }
        if (t == 0 && base >= 0) {                                       // a push that found no room is no push
            const unsigned long long written = (unsigned long long)(total + 2 + 2 * (n_chunks - 1));
            atomicAdd(&a.ctl->stats[ST_P_TILES], 1ull);
            atomicAdd(&a.ctl->stats[ST_P_CELLS], (unsigned long long)n);
            atomicAdd(&a.ctl->stats[ST_P_KEPT], written);
            atomicAdd(&a.ctl->stats[ST_P_WRITTEN], written);
        }
        if (pb >= 0) {
            // ---- the shortened tape's plan: its clauses in (level, opcode) order, renumbered ------------
            // An operand of a kept clause names a root clause that may have left the tape: a dropped min / max
            // (the value is then the one its output slot already held), or - for the operand a verdict
            // made unused - anything at all.  A later backward walk of the shortened tape still marks that
            // unused operand's SLOT (context.cu:386-402 looks at the slot bytes, not the opcode), which keeps
            // alive whatever kept clause wrote the slot last.  Both cases are one question: which clause of
            // the shortened tape wrote this slot last?  P[i] = the value that sat in clause i's output slot
            // before it (host-built, api.cu); following it from the operand's root producer answers it.
            uint16_t* P = reinterpret_cast<uint16_t*>(mine + size_t(nv) * 2);          // behind N
            uint16_t* P2 = reinterpret_cast<uint16_t*>(mine + size_t(nv) * 4);         // the other copy (see below)
            for (int i = t; i <= n; i += G) P[i] = a.prevw[i];
            if (t < 4) N[t] = uint16_t(t);                                // none, x, y, z
            group_sync(bar, G);                                           // N, P complete
            auto in_tape = [&](uint32_t p) { return A[p] && N[p] != kNotInTape; };
            // Slots are reused hundreds of times and most of their writers are dead in any one tile:
            // pointer jumping first (P[i] skips writers that left the tape; log2(chain) rounds), so that a
            // lookup is a hop or two.  Each round reads one copy and writes the other.
            for (;;) {
                if (t == 0) scratch[14] = 0;
                group_sync(bar, G);
                bool moved = false;
                for (int i = 1 + t; i <= n; i += G) {
                    uint32_t p = P[i];
                    if (p > 3u && !in_tape(p)) {
                        p = P[p - 3u];
                        moved = true;
                    }
                    P2[i] = uint16_t(p);
                }
                if (moved) scratch[14] = 1;
                group_sync(bar, G);
                const bool again = scratch[14] != 0;
                group_sync(bar, G);                                       // everyone has read the flag before it is reset
                uint16_t* const tmp = P;
                P = P2;
                P2 = tmp;
                if (!again) break;
            }
            auto in_plan = [&](uint32_t p) -> uint32_t {                  // root value id -> plan value id
                while (p > 3u && !in_tape(p)) p = P[p - 3u];
                return N[p];
            };
            const int k_begin = min(n, t * per), k_end = min(n, k_begin + per);
            int cnt = 0;
            for (int k = k_begin; k < k_end; ++k) cnt += KS[k];
            int pin = cnt;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, pin, o);
                if ((t & 31) >= o) pin += v;
            }
            if ((t & 31) == 31) wsum[t >> 5] = pin;
            group_sync(bar, G);
            int pp = pin - cnt;
            for (int wv = 0; wv < (t >> 5); ++wv) pp += wsum[wv];
            int* const PP = reinterpret_cast<int*>(mine + ((size_t(nv) * 7 + 3) & ~size_t(3)));   // entries before each thread's range
            PP[t] = pp;
            int32_t* const hdr = a.plans + pb;
            int32_t* const pls = hdr + kPlanHeader;
            const int sched_off = kPlanHeader + ((a.n_levels + 1 + 3) & ~3);
            RootClause* const out = reinterpret_cast<RootClause*>(hdr + sched_off);
            uint32_t* const extras = reinterpret_cast<uint32_t*>(hdr + sched_off + 4 * total);
            if (t == 0) scratch[15] = 0;                                  // how many (read again behind the next barrier)
            group_sync(bar, G);
            for (int k = k_begin; k < k_end; ++k) {
                if (!KS[k]) continue;
                const RootClause rc = a.sched[k];
                const uint32_t op = rc.op_idx & 0xff;
                const uint32_t idx = rc.op_idx >> 12;
                const int c = (op >= OP_MIN_LI && op <= OP_MAX_LR) ? C[idx] : 0;
                RootClause e;
                e.imm = rc.imm;
                if (c == 0) {
                    e.op_idx = op;
                    e.lsrc = in_plan(rc.lsrc);
                    e.rsrc = in_plan(rc.rsrc);
                } else if (c == 1 || rc.rsrc != 0) {      // the chosen operand passes through, the other is only marked
                    e.op_idx = OP_COPY_LHS;
                    e.lsrc = in_plan(c == 1 ? rc.lsrc : rc.rsrc);
                    e.rsrc = in_plan(c == 1 ? rc.rsrc : rc.lsrc);
                } else {
                    e.op_idx = OP_COPY_IMM;
                    e.lsrc = in_plan(rc.lsrc);
                    e.rsrc = 0;
                }
                e.op_idx |= (N[3 + idx] - 3u) << 12;
                out[pp++] = e;
                // the source a copy only keeps alive (k_eval_sub marks those in a pass of their own)
                const uint32_t extra = c == 0 ? 0u : ((c == 1 || rc.rsrc != 0) ? e.rsrc : e.lsrc);
                if (extra != 0) extras[atomicAdd(&scratch[15], 1)] = (N[3 + idx] - 3u) | (extra << 16);
            }
            group_sync(bar, G);                                           // PP, extras complete
            for (int L = t; L < a.n_levels; L += G) {                     // where each level starts in the plan
                const int ks = a.level_start[L];
                const int owner = ks / per;
                int v = PP[owner];
                for (int k = owner * per; k < ks; ++k) v += KS[k];
                pls[L] = v;
            }
            if (t == 0) {
                pls[a.n_levels] = total;
                hdr[PL_N] = total;
                hdr[PL_LEVELS] = a.n_levels;
                hdr[PL_RESULT] = int32_t(in_plan(uint32_t(a.result_v)));
                hdr[PL_TAPE] = base;
                hdr[PL_LOGICAL] = n_logical;
                hdr[PL_VALUES] = plan_nv;
                hdr[PL_SCHED] = sched_off;
                hdr[PL_EXTRAS] = scratch[15];
                a.plan_of[tile] = pb;
            }
        }
    }
    if (t == 0) {
        a.tiles[tile].position = out_position;
        a.tiles[tile].tape = out_tape;
        a.tiles[tile].next = -1;
        atomicAdd(&a.ctl->stats[ST_I_TILES], 1ull);
        atomicAdd(&a.ctl->stats[ST_I_CELLS], (unsigned long long)n);
    }
}

////////////////////////////////////////////////////////////////////////////////
// Small levels below the root, evaluated clause-parallel.
//
// A level with few tiles (a small frame, or one GPU's share of a frame split eight ways) leaves
// k_eval_tiles with a handful of warps, each walking a long tape one clause at a time: the frame
// waits on a dependent chain, not on the machine.  The tapes of such a level were all shortened by
// k_eval_root, which knows their dependency levels (a subset of the root tape's), so it writes a plan
// next to each tape (EvalRootArgs::plans) and here ONE WARP owns one tile and runs the plan a level
// at a time - the same scheme, values and verdicts as k_eval_root, with __syncwarp for a barrier.
// Interval results do not depend on evaluation order, so the tile's classification and its
// shortened tape (mark and sweep in tape order, written as a contiguous run like k_eval_root's)
// are those of the serial walk (context.cu:188-459).
//
// Which kernel takes a tile is a function of device-side facts both kernels read the same way:
// the level's parent count against max_parents, and whether the parent carries a plan.

constexpr int kSubThreads = kSubMaxWarps * 32;

template <int DIM>
__global__ void __launch_bounds__(kSubThreads)
k_eval_sub(const EvalSubArgs a, const typename MatOf<DIM>::type mat)
{
    extern __shared__ __align__(16) unsigned char s_sub[];
    const int n_parents = min(*a.n_parents, a.tiles_cap / 64);
    if (n_parents > a.max_parents) return;
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    unsigned char* const mine = s_sub + size_t(warp) * a.slice;
    uint64_t* const arena = a.arena;
    const uint32_t tps = a.tps;
    const int n_items = n_parents * 64;
    unsigned long long st_tiles = 0, st_cells = 0, st_ptiles = 0, st_pcells = 0, st_kept = 0;

    for (;;) {
        const int item = warp_next(a.queue);
        if (item >= n_items) break;
        const int rank = item >> 6;
        const int sub = DIM == 3 ? 63 - (item & 63) : (item & 63);      // 3D: highest z first
        const int pidx = a.pactive[rank];
        const int plan = a.plan_of[pidx];
        if (plan < 0) continue;                                          // k_eval_tiles has this parent
        const TileNode parent = a.ptiles[pidx];
        const int tile_index = rank * 64 + sub;
        int sx, sy, sz = 0;
        {
            const int pp = parent.position;
            const int px = pp % a.ptps, py = (pp / a.ptps) % a.ptps;
            if (DIM == 3) {
                const int pz = (pp / a.ptps) / a.ptps;
                sx = px * 4 + (sub & 3);
                sy = py * 4 + ((sub >> 2) & 3);
                sz = pz * 4 + (sub >> 4);
            } else {
                sx = px * 8 + (sub & 7);
                sy = py * 8 + (sub >> 3);
            }
        }
        const int position = sx + sy * tps + (DIM == 3 ? sz * tps * tps : 0);
        const int img_index = sx + sy * tps;

        // Occlusion pre-mask: lane 0 looks, the warp follows
        if (DIM == 3) {
            int seen = 0;
            if (lane == 0) seen = __ldcg(&a.image[img_index]);
            seen = __shfl_sync(kFull, seen, 0);
            if (seen > sz) {
                if (lane == 0) {
                    a.tiles[tile_index].position = -1;
                    a.tiles[tile_index].tape = parent.tape;
                    a.tiles[tile_index].next = -1;
                }
                continue;
            }
        }

        const int32_t* const hdr = a.plans + plan;
        const int n = hdr[PL_N], n_levels = hdr[PL_LEVELS], result_v = hdr[PL_RESULT];
        const int ptape = hdr[PL_TAPE], p_logical = hdr[PL_LOGICAL], nv = hdr[PL_VALUES];
        const RootClause* __restrict__ const sched = reinterpret_cast<const RootClause*>(hdr + hdr[PL_SCHED]);
        const uint32_t* __restrict__ const extras = reinterpret_cast<const uint32_t*>(hdr + hdr[PL_SCHED] + 4 * n);
        const int n_extras = hdr[PL_EXTRAS];
        float2* const V = reinterpret_cast<float2*>(mine);
        uint8_t* const C = mine + size_t(nv) * 8;
        uint8_t* const A = C + nv;
        uint8_t* const LV = A + nv;                                      // dependency level of each value
        int* const LS = reinterpret_cast<int*>(mine + ((nv * 11 + 3) & ~3));

        __syncwarp();                                                    // the previous tile is done with `mine`
        if (lane == 0) {                                                 // tile box -> axis intervals
            const float ftps = float(tps);
            const ival ix = iv(tile_edge(sx, ftps), tile_edge(sx + 1, ftps));
            const ival iy = iv(tile_edge(sy, ftps), tile_edge(sy + 1, ftps));
            const float* m = mat.d;
            if (DIM == 3) {
                const ival iz = iv(tile_edge(sz, ftps), tile_edge(sz + 1, ftps));
                ival r[4];
                #pragma unroll
                for (int i = 0; i < 4; ++i)
                    r[i] = iv_add(iv_add(iv_add(iv_mul(ix, m[i]), iv_mul(iy, m[4 + i])), iv_mul(iz, m[8 + i])), m[12 + i]);
                V[1] = iv_div(r[0], r[3]);
                V[2] = iv_div(r[1], r[3]);
                V[3] = iv_div(r[2], r[3]);
            } else {
                ival r[3];
                #pragma unroll
                for (int i = 0; i < 3; ++i)
                    r[i] = iv_add(iv_add(iv_mul(ix, m[i]), iv_mul(iy, m[3 + i])), m[6 + i]);
                V[1] = iv_div(r[0], r[2]);
                V[2] = iv_div(r[1], r[2]);
                V[3] = iv(a.z, a.z);
            }
            V[0] = iv(0.0f, 0.0f);
        }
        for (int i = lane; i < nv; i += 32) A[i] = 0;
        for (int i = lane; i <= n_levels; i += 32) LS[i] = hdr[kPlanHeader + i];
        __syncwarp();

        // ---- forward: one dependency level at a time; the next level's clause is fetched (L2) while
        // this one is evaluated --------------------------------------------------------------------
        bool any_choice = false;
        RootClause nxt = {};
        {
            const int k = LS[0] + lane;
            if (k < LS[1]) nxt = sched[k];
        }
        for (int L = 0; L < n_levels; ++L) {
            const int k0 = LS[L], k_end = LS[L + 1];
            RootClause rc = nxt;
            if (L + 1 < n_levels) {
                const int k = k_end + lane;
                if (k < LS[L + 2]) nxt = sched[k];
            }
            for (int k = k0 + lane; k < k_end; k += 32) {
                if (k >= k0 + 32) rc = sched[k];
                const uint32_t op = rc.op_idx & 0xff;
                const uint32_t idx = rc.op_idx >> 12;
                int c = 0;
                // a copy's other source is only there for the mark phase (it may sit on this very level)
                const ival Lv = op == OP_COPY_IMM ? iv(0.0f, 0.0f) : V[rc.lsrc];
                const ival Rv = op == OP_COPY_LHS ? iv(0.0f, 0.0f) : V[rc.rsrc];
                const ival o = eval_plan_clause(op, Lv, Rv, rc.imm, c);
                V[3 + idx] = o;
                LV[3 + idx] = uint8_t(L);
                if (op >= OP_MIN_LI && op <= OP_MAX_LR) {
                    any_choice |= (c != 0);
                    C[idx] = uint8_t(c);
                }
            }
            __syncwarp();
        }
        any_choice = __any_sync(kFull, any_choice);
        const ival result = V[result_v];

        // ---- classify (context.cu:289-321) ---------------------------------------------------
        int out_position = -1;
        bool pushing = false;
        {
            int verdict = 0;                                             // lane 0's view of the racy image wins
            if (lane == 0) {
                if (result.x > 0.0f) {
                    // empty
                } else if (DIM == 3 && __ldcg(&a.image[img_index]) > sz) {
                    // hidden
                } else if (result.y < 0.0f) {
                    if (DIM == 3) atomicMax(&a.image[img_index], sz);
                    else a.image[img_index] = 1;
                } else {
                    verdict = any_choice ? 2 : 1;
                }
            }
            verdict = __shfl_sync(kFull, verdict, 0);
            if (verdict) out_position = position;
            pushing = verdict == 2;
        }
        int out_tape = parent.tape;
        st_tiles += 1;
        st_cells += unsigned(p_logical - 2 + 2 * ((p_logical <= kChunk ? 1 : 1 + (p_logical - 63 + 61) / 62) - 1));

        if (pushing) {
            // ---- mark: result -> operands, top level first ------------------------------------
            if (lane == 0) A[result_v] = 1;
            __syncwarp();
            // A copy the parent's verdict left behind also keeps its unused operand's slot alive, i.e. the
            // clause that wrote that slot last in the shortened tape (the plan's extra source, see
            // k_eval_root).  Such an edge can point to a HIGHER dependency level than the copy sits on, which
            // one top-down sweep has already passed: the sweep restarts from the highest level so marked
            // until there is none (marks only grow, so the result is the backward walk's).
            // So: a top-down sweep over the operands proper (they sit on lower levels: what a level reads
            // and what it writes never meet), then one pass over all live copies that first LOOKS at their extra
            // sources and then, behind a barrier, marks them; a mark that landed on a level the sweep had
            // passed restarts the sweep from there.
            for (int top = n_levels - 1; top >= 0;) {
                RootClause ahead = {};
                {
                    const int k = LS[top] + lane;
                    if (k < LS[top + 1]) ahead = sched[k];
                }
                for (int L = top; L >= 0; --L) {
                    const int k0 = LS[L], k_end = LS[L + 1];
                    RootClause rc = ahead;
                    if (L > 0) {
                        const int k = LS[L - 1] + lane;
                        if (k < k0) ahead = sched[k];
                    }
                    for (int k = k0 + lane; k < k_end; k += 32) {
                        if (k >= k0 + 32) rc = sched[k];
                        const uint32_t op = rc.op_idx & 0xff;
                        const uint32_t idx = rc.op_idx >> 12;
                        if (!A[3 + idx]) continue;
                        const int c = (op >= OP_MIN_LI && op <= OP_MAX_LR) ? C[idx] : 0;
                        if (op == OP_COPY_IMM) continue;                          // its one source is an extra edge
                        if (c != 2) A[rc.lsrc] = 1;
                        if (c != 1 && op != OP_COPY_LHS) A[rc.rsrc] = 1;          // a copy's rsrc is an extra edge
                    }
                    __syncwarp();
                }
                int late = -1;                                                    // highest level an extra edge newly marks
                for (int k0 = 0; k0 < n_extras; k0 += 32) {
                    uint32_t extra = 0;
                    if (k0 + lane < n_extras) {
                        const uint32_t e = extras[k0 + lane];
                        if (A[3 + (e & 0xffffu)]) extra = e >> 16;
                    }
                    const bool fresh = extra != 0 && !A[extra];
                    __syncwarp();                                                 // all looked before anyone marks
                    if (fresh) {
                        A[extra] = 1;
                        if (extra > 3u) late = max(late, int(LV[extra]));
                    }
                    __syncwarp();
                }
                #pragma unroll
                for (int o = 16; o; o >>= 1) late = max(late, __shfl_xor_sync(kFull, late, o));
                top = late;
            }
            // ---- sweep: compact kept clauses in tape order (the parent's cells, as stored) ------
            const uint64_t* const cells = arena + ptape;
            const int per = (n + 31) / 32;
            const int i_begin = 1 + lane * per, i_end = min(n + 1, i_begin + per);
            int kept = 0;
            for (int i = i_begin; i < i_end; ++i) {
                if (!A[3 + i]) continue;
                const uint32_t w = uint32_t(cells[chunked_index(i, p_logical)]);
                const uint32_t op = w & 0xff, i_out = (w >> 8) & 0xff, i_lhs = (w >> 16) & 0xff, i_rhs = w >> 24;
                const int c = (op >= OP_MIN_LI && op <= OP_MAX_LR) ? C[i] : 0;
                const bool dropped = (c == 1 && i_lhs == i_out) || (c == 2 && i_rhs != 0 && i_rhs == i_out);
                kept += dropped ? 0 : 1;
            }
            int incl = kept;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += v;
            }
            const int offset = incl - kept;
            const int total = __shfl_sync(kFull, incl, 31);
            const int n_logical = total + 2;
            const int n_chunks = n_logical <= kChunk ? 1 : 1 + (n_logical - 63 + 61) / 62;
            int base = -1;
            if (lane == 0) {
                const int need = n_chunks * kChunk;
                if (*(volatile int32_t*)a.tape_index < a.arena_cap) {
                    base = atomicAdd(a.tape_index, need);
                    if (base + need >= a.arena_cap) base = -1;         // arena exhausted: keep the parent tape
                }
            }
            base = __shfl_sync(kFull, base, 0);
            if (base >= 0) {
                int q = 1 + offset;
                for (int i = i_begin; i < i_end; ++i) {
                    if (!A[3 + i]) continue;
                    uint64_t d = cells[chunked_index(i, p_logical)];
                    const uint32_t w = uint32_t(d);
                    const uint32_t op = w & 0xff, i_out = (w >> 8) & 0xff, i_lhs = (w >> 16) & 0xff, i_rhs = w >> 24;
                    const int c = (op >= OP_MIN_LI && op <= OP_MAX_LR) ? C[i] : 0;
                    if (c == 1) {
                        if (i_lhs == i_out) continue;
                        d = (d & ~0xffull) | OP_COPY_LHS;
                    } else if (c == 2) {
                        if (i_rhs != 0 && i_rhs == i_out) continue;
                        d = (d & ~0xffull) | (i_rhs ? OP_COPY_RHS : OP_COPY_IMM);
                    }
                    arena[base + chunked_index(q, n_logical)] = d;
                    ++q;
                }
                if (lane == 0) {
                    arena[base] = cells[0];                                                       // header
                    arena[base + chunked_index(total + 1, n_logical)] = cells[chunked_index(n + 1, p_logical)];   // end cell
                }
                for (int c = lane; c < n_chunks; c += 32) {
                    if (c > 0) arena[base + c * kChunk] = make_jump(-1);
                    if (c + 1 < n_chunks) arena[base + c * kChunk + kChunk - 1] = make_jump(1);
                }
                out_tape = base;
                st_ptiles += 1;
                st_pcells += unsigned(n);
                st_kept += unsigned(total + 2 + 2 * (n_chunks - 1));
            }
        }
        if (lane == 0) {
            a.tiles[tile_index].position = out_position;
            a.tiles[tile_index].tape = out_tape;
            a.tiles[tile_index].next = -1;
        }
    }
    if (lane == 0 && st_tiles) {
        atomicAdd(&a.ctl->stats[ST_I_TILES + a.level], st_tiles);
        atomicAdd(&a.ctl->stats[ST_I_CELLS + a.level], st_cells);
        atomicAdd(&a.ctl->stats[ST_I_SUB], st_tiles);
        if (st_ptiles) {
            atomicAdd(&a.ctl->stats[ST_P_TILES + a.level], st_ptiles);
            atomicAdd(&a.ctl->stats[ST_P_CELLS + a.level], st_pcells);
            atomicAdd(&a.ctl->stats[ST_P_KEPT + a.level], st_kept);
            atomicAdd(&a.ctl->stats[ST_P_WRITTEN], st_kept);
        }
    }
}

////////////////////////////////////////////////////////////////////////////////
// Post-mask + compaction (mask_filled_tiles #2, assign_next_nodes,
// subdivide / copy_active_tiles bookkeeping; context.cu:471-551, :637-651)

template <int DIM>
__global__ void __launch_bounds__(256)
k_rank_tiles(const RankArgs a)
{
    const int n_tiles = a.n_parents ? 64 * min(*a.n_parents, a.tiles_cap / 64) : a.count0;
    const int lane = lane_id();
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n_tiles; base += stride) {
        // 3D: visit each parent's 64 children highest-z first (sub = x + 4y + 16z), so that ranks -
        // and with them the order in which the next pass takes tiles - run front to back and the
        // occlusion early-outs see the surface before what lies behind it.
        int t = base + lane;
        if (DIM == 3 && a.n_parents) t = (t & ~63) | (63 - (t & 63));
        bool active = false;
        TileNode node = {-1, 0, -1};
        if (t < n_tiles) {
            node = a.tiles[t];
            if (node.position != -1) {
                active = true;
                if (DIM == 3) {
                    const int xy = node.position % (a.tps * a.tps);
                    const int z = node.position / (a.tps * a.tps);
                    if (a.image[xy] > z) {
                        active = false;
                        a.tiles[t].position = -1;
                    }
                }
            }
        }
        const unsigned m = __ballot_sync(kFull, active);
        int rank_base = 0;
        if (lane == 0 && m) rank_base = atomicAdd(a.n_active, __popc(m));
        rank_base = __shfl_sync(kFull, rank_base, 0);
        int next = -1;
        if (!a.last_level) {
            if (active) {
                const int rank = rank_base + __popc(m & ((1u << lane) - 1));
                if (((long long)rank + 1) * 64 > a.next_cap) {
                    // Next stage's tile array is too small: the frame reports MPRB_E_OVERFLOW
                    atomicOr(&a.ctl->overflow, 1 << a.level);
                } else {
                    a.active_list[rank] = t;
                    next = rank;
                }
            }
        } else if (m) {
            // Last level: the compact survivor list for the float pass (copy_active_tiles; `next`
            // stays -1), ordered so that tiles sharing a tape sit next to each other - these 32 tiles
            // are siblings, and k_eval_tiles gives siblings with the same verdicts the same tape -
            // plus one work item per run of up to gmax such tiles.
            const unsigned cls = active ? __match_any_sync(m, node.tape) : 0u;
            const int leader = __ffs(cls) - 1;
            const int within = __popc(cls & ((1u << lane) - 1));
            const int csize = __popc(cls);
            const int gmax = a.gmax;
            int tile_off = 0, item_off = 0, n_items = 0;
            unsigned leaders = __ballot_sync(kFull, active && lane == leader);
            while (leaders) {                                  // classes in order of their first lane
                const int l = __ffs(leaders) - 1;
                leaders &= leaders - 1;
                const int sz = __shfl_sync(kFull, csize, l);
                const int it = (sz + gmax - 1) / gmax;
                if (l < leader) { tile_off += sz; item_off += it; }
                n_items += it;
            }
            int item_base = 0;
            if (lane == 0) item_base = atomicAdd(a.n_items, n_items);
            item_base = __shfl_sync(kFull, item_base, 0);
            if (active) {
                const int rank = rank_base + tile_off + within;
                if ((long long)rank + 1 > a.next_cap) {
                    atomicOr(&a.ctl->overflow, 1 << a.level);
                } else {
                    a.out_tiles[rank].position = node.position;
                    a.out_tiles[rank].tape = node.tape;
                    a.out_tiles[rank].next = -1;
                    if (within % gmax == 0)
                        a.items[item_base + item_off + within / gmax] = rank * 8 + min(gmax, csize - within);
                }
            }
        }
        if (t < n_tiles) a.tiles[t].next = next;
    }
}

// Filled image of one level -> the next (copy_filled_{2d,3d}, context.cu:664-692).
// Writes every pixel, so the destination needs no separate clear.
template <int DIM>
__global__ void __launch_bounds__(256)
k_upsample_filled(const int32_t* __restrict__ prev, int32_t* __restrict__ image, int size)
{
    constexpr int F = (DIM == 3) ? 4 : 8;
    // four pixels of a row per thread (they share their parent tile: 4 divides F): one 16-byte store
    const int quads = size / 4, n = quads * size;
    int4* const out = reinterpret_cast<int4*>(image);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int x = (i % quads) * 4, y = i / quads;
        const int32_t t = prev[x / F + (y / F) * (size / F)];
        const int32_t v = t ? (DIM == 3 ? t * 4 + 3 : 1) : 0;
        out[i] = make_int4(v, v, v, v);
    }
}

////////////////////////////////////////////////////////////////////////////////
// Float pass (calculate_voxels / calculate_pixels + eval_voxels_f,
// context.cu:707-964)

// Walks one tape for the two samples this lane holds; returns the result pair.
// Clause semantics: context.cu:887-920.  There is no a*b+c shape in any clause,
// so nothing here can be contracted; the _rn intrinsics just make that explicit.
// `slots` is this lane's shared-space base address (slot s at slots + s * 256).
// The clause loop of the float pass, written in PTX (generated: tools/gen_float_loop.py ->
// float_loop_ptx.inc).
//
// The C++ form of this loop compiles to ~40 SASS instructions per clause - a compare/branch
// tree for the 30-way switch, divergence bookkeeping (BSSY/BSYNC) around it, shift+mask+add
// per operand address - and moves 7 shared-memory wavefronts per clause (clause word, two
// operand rows, one result row), which made the float pass first issue-bound (84 % issue-active,
// profiles/r01_ncu_k_eval_voxels.csv) and then, with a plain PTX loop, shared-memory bound (86 %
// of the LSU data pipe).  This version attacks both:
//   * dispatch is one indexed branch through a table (`brx.idx.uni`; every lane of the warp runs
//     the same tape, hence `.uni` and no reconvergence stack), and every handler jumps straight
//     back to the head of the loop;
//   * the tape packer reuses slots LIFO, so most clauses consume the previous clause's result
//     and overwrite its slot (bear: 60 % / 66 %).  annotate_chunk() marks those cases in the three
//     spare bits of the opcode byte when a chunk arrives, and the table has a handler per
//     (opcode, hints) that takes the operand from the result registers and skips the dead store;
//     operands a clause does not use are never loaded.
// Runs clauses starting at the cell AFTER `cp` until it meets one it does not handle - END, JUMP,
// or a trigonometric libdevice function (EXP and LOG are handlers: libdevice's own PTX) - and returns with cp on that cell and its two words
// in w / imm.  `sb` is this lane's slot base in shared space (slot s at sb + 256 s).
// G = tiles per warp (1, 2 or 4): the same loop over G f32 pairs per slot and lane (tools/gen_float_loop.py).
template <int G>
__device__ __forceinline__ void run_float_clauses(uint32_t& cp, uint32_t& w, uint32_t& imm, uint32_t sb);
template <>
__device__ __forceinline__ void run_float_clauses<1>(uint32_t& cp, uint32_t& w, uint32_t& imm, uint32_t sb)
{
    asm volatile(
#include "float_loop_ptx.inc"
        : "+r"(cp), "=&r"(w), "=&r"(imm)
        : "r"(sb)
        : "memory");
}
template <>
__device__ __forceinline__ void run_float_clauses<2>(uint32_t& cp, uint32_t& w, uint32_t& imm, uint32_t sb)
{
    asm volatile(
#include "float_loop_ptx_g2.inc"
        : "+r"(cp), "=&r"(w), "=&r"(imm)
        : "r"(sb)
        : "memory");
}
template <>
__device__ __forceinline__ void run_float_clauses<4>(uint32_t& cp, uint32_t& w, uint32_t& imm, uint32_t sb)
{
    asm volatile(
#include "float_loop_ptx_g4.inc"
        : "+r"(cp), "=&r"(w), "=&r"(imm)
        : "r"(sb)
        : "memory");
}

// G = 2 with tile 1's value rows in TENSOR MEMORY (tools/gen_float_loop.py, `tmem`): tb = address of this
// warp's column group minus 2 (slot id s -> columns 2 (s - 1), 2 (s - 1) + 1).
template <int G>
__device__ __forceinline__ void run_float_clauses_tm(uint32_t& cp, uint32_t& w, uint32_t& imm, uint32_t sb, uint32_t tb)
{
    if (G == 4) {          // tiles 0, 1 in shared memory, tiles 2, 3 in tensor memory (four columns per slot)
        asm volatile(
#include "float_loop_ptx_g4t.inc"
            : "+r"(cp), "=&r"(w), "=&r"(imm)
            : "r"(sb), "r"(tb)
            : "memory");
    } else {
        asm volatile(
#include "float_loop_ptx_g2t.inc"
            : "+r"(cp), "=&r"(w), "=&r"(imm)
            : "r"(sb), "r"(tb)
            : "memory");
    }
}
__device__ __forceinline__ void tm_st2(uint32_t taddr, float2 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float2 tm_ld2(uint32_t taddr) {      // waits for earlier stores and for the load itself
    float2 v;
    asm volatile("tcgen05.wait::st.sync.aligned;\n"
                 "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=f"(v.x), "=f"(v.y) : "r"(taddr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t tm_col(uint32_t off256) { return off256 >> 7; }   // slot * 256 -> slot * 2

// One clause in C++ (context.cu:887-920): the slot-renaming variant runs every clause through
// this, the PTX loop above only the transcendental ones.
__device__ __forceinline__ float2 float_clause(uint32_t op, float2 L, float2 R, float imm)
{
    switch (op) {
        case OP_SQUARE: return make_float2(__fmul_rn(L.x, L.x), __fmul_rn(L.y, L.y));
        case OP_SQRT:   return make_float2(sqrtf(L.x), sqrtf(L.y));
        case OP_NEG:    return make_float2(-L.x, -L.y);
        case OP_SIN:    return make_float2(sinf(L.x), sinf(L.y));
        case OP_COS:    return make_float2(cosf(L.x), cosf(L.y));
        case OP_ASIN:   return make_float2(asinf(L.x), asinf(L.y));
        case OP_ACOS:   return make_float2(acosf(L.x), acosf(L.y));
        case OP_ATAN:   return make_float2(atanf(L.x), atanf(L.y));
        case OP_EXP:    return make_float2(expf(L.x), expf(L.y));
        case OP_ABS:    return make_float2(fabsf(L.x), fabsf(L.y));
        case OP_LOG:    return make_float2(logf(L.x), logf(L.y));
        case OP_ADD_LI: return make_float2(__fadd_rn(L.x, imm), __fadd_rn(L.y, imm));
        case OP_ADD_LR: return make_float2(__fadd_rn(L.x, R.x), __fadd_rn(L.y, R.y));
        case OP_MUL_LI: return make_float2(__fmul_rn(L.x, imm), __fmul_rn(L.y, imm));
        case OP_MUL_LR: return make_float2(__fmul_rn(L.x, R.x), __fmul_rn(L.y, R.y));
        case OP_MIN_LI: return make_float2(fminf(L.x, imm), fminf(L.y, imm));
        case OP_MIN_LR: return make_float2(fminf(L.x, R.x), fminf(L.y, R.y));
        case OP_MAX_LI: return make_float2(fmaxf(L.x, imm), fmaxf(L.y, imm));
        case OP_MAX_LR: return make_float2(fmaxf(L.x, R.x), fmaxf(L.y, R.y));
        case OP_SUB_LI: return make_float2(__fsub_rn(L.x, imm), __fsub_rn(L.y, imm));
        case OP_SUB_IR: return make_float2(__fsub_rn(imm, R.x), __fsub_rn(imm, R.y));
        case OP_SUB_LR: return make_float2(__fsub_rn(L.x, R.x), __fsub_rn(L.y, R.y));
        case OP_DIV_LI: return make_float2(__fdiv_rn(L.x, imm), __fdiv_rn(L.y, imm));
        case OP_DIV_IR: return make_float2(__fdiv_rn(imm, R.x), __fdiv_rn(imm, R.y));
        case OP_DIV_LR: return make_float2(__fdiv_rn(L.x, R.x), __fdiv_rn(L.y, R.y));
        case OP_COPY_IMM: return make_float2(imm, imm);
        case OP_COPY_LHS: return L;
        case OP_COPY_RHS: return R;
        default: return L;
    }
}
__device__ __forceinline__ float2 float_clause_libdevice(uint32_t op, float2 L)
{
    switch (op) {
        case OP_SIN:  return make_float2(sinf(L.x), sinf(L.y));
        case OP_COS:  return make_float2(cosf(L.x), cosf(L.y));
        case OP_ASIN: return make_float2(asinf(L.x), asinf(L.y));
        case OP_ACOS: return make_float2(acosf(L.x), acosf(L.y));
        case OP_ATAN: return make_float2(atanf(L.x), atanf(L.y));
        case OP_EXP:  return make_float2(expf(L.x), expf(L.y));
        default:      return make_float2(logf(L.x), logf(L.y));
    }
}

// Walks one tape for the G tiles of a work item (two samples per tile and lane); r[g] receives
// tile g's result pair.  Slot rows are 32 lanes x 8 G bytes: `sb` is this lane's address in row 0
// and tile g sits 8 g bytes further.
template <bool REMAP, int G, bool TM>
__device__ __forceinline__ void walk_float(TapeStream<REMAP, REMAP>& ts, int tape, Slots2<REMAP>& slots, unsigned& cells,
                                           float2 (&r)[G], uint32_t tb)
{
    constexpr int GS = TM ? G / 2 : G;                                   // tiles per shared-memory row (TM: the other half sits in tensor memory)
    constexpr int SHIFT = GS == 4 ? 2 : (GS == 2 ? 1 : 0);
    static_assert(!REMAP || G == 1, "renamed slots: one tile per warp");
    static_assert(!TM || G == 2 || G == 4, "tensor-memory rows: half of two or four tiles per warp");
    if (ts.fetch(tape) && !REMAP) annotate_chunk<kFastOps, kUsesLhs, kUsesRhs, SHIFT>(ts.buf);
    uint32_t cp = ts.rd + ((tape & (kChunk - 1)) << 3);
    uint32_t seg = cp;
    uint32_t w, immb;
    for (;;) {
        if (REMAP) {
            // the generated loop on the renamed chunk (no forwarding hints there: every handler loads its
            // operands and stores its result); clauses that touch a spilled row come back as kOpBounce
            run_float_clauses<1>(cp, w, immb, slots.base);
        } else if (TM) {
            run_float_clauses_tm<G>(cp, w, immb, slots.base, tb);
        } else {
            run_float_clauses<G>(cp, w, immb, slots.base);
        }
        uint32_t op = w & 0xff;
        if (REMAP && op == kOpBounce) op = lds_u32(cp - ts.rd + ts.buf) & 0xff;      // its opcode, from the raw chunk
        if (op <= OP_JUMP) {
            cells += (cp - seg) >> 3;
            if (op == OP_END) { --cells; break; }
            const int t = ts.base + int((cp - ts.rd) >> 3) + int32_t(immb);
            if (ts.fetch(t) && !REMAP) annotate_chunk<kFastOps, kUsesLhs, kUsesRhs, SHIFT>(ts.buf);
            cp = ts.rd + ((t & (kChunk - 1)) << 3);
            seg = cp;
            continue;
        }
        if (REMAP) {
            const float2 L = slots.ld(off_lhs2(w));
            slots.st(off_out2(w), float_clause(op, L, slots.ld(off_rhs2(w)), __uint_as_float(immb)));
        } else if (TM) {
            #pragma unroll
            for (int g = 0; g < GS; ++g) {
                sts_f2(slots.base + off_out2(w) + 8 * g, float_clause_libdevice(op, lds_f2(slots.base + off_lhs2(w) + 8 * g)));
                tm_st2(tb + tm_col(off_out2(w)) + 2 * g, float_clause_libdevice(op, tm_ld2(tb + tm_col(off_lhs2(w)) + 2 * g)));
            }
        } else {
            #pragma unroll
            for (int g = 0; g < G; ++g) {
                const float2 L = lds_f2(slots.base + off_lhs2(w) + 8 * g);
                sts_f2(slots.base + off_out2(w) + 8 * g, float_clause_libdevice(op, L));
            }
        }
    }
    if (REMAP) {
        r[0] = slots.ld(off_out2(w));
    } else if (TM) {
        #pragma unroll
        for (int g = 0; g < GS; ++g) {
            r[g] = lds_f2(slots.base + off_out2(w) + 8 * g);
            r[(GS + g) % G] = tm_ld2(tb + tm_col(off_out2(w)) + 2 * g);
        }
    } else {
        #pragma unroll
        for (int g = 0; g < G; ++g) r[g] = lds_f2(slots.base + off_out2(w) + 8 * g);
    }
}

// A work item of the float pass: `count` (1 .. G) tiles of the compact survivor list that share one
// tape, packed as start * 8 + count (k_rank_tiles writes them).
__device__ __forceinline__ void unpack_item(int32_t code, int& start, int& count) {
    code = __shfl_sync(kFull, code, 0);      // every lane loaded the same word; the shuffle tells ptxas so
    start = code >> 3;
    count = code & 7;
}

// 2D: one warp per work item of up to G surviving 8x8 tiles, two pixels per tile and lane (y and y + 4).
template <bool REMAP, bool HEAT = false, int G = 1, bool TM = false>
__global__ void __launch_bounds__(kFloatMaxThreads)
k_eval_pixels(const EvalVoxelsArgs a, const Mat3 mat)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const int lane = lane_id();
    const int warp = __shfl_sync(kFull, int(threadIdx.x >> 5), 0);   // via shuffle: provably warp-uniform for ptxas
    typedef TapeStream<REMAP, REMAP> Stream;   // look-ahead only where slots are renamed (see tape_stream.cuh)
    Stream ts;
    ts.init(s_dyn + warp * Stream::stride(), a.arena, a.arena_cap);
    const int n_rows = a.n_rows;       // shared-memory value rows per warp (= slot count unless REMAP)
    constexpr int GS = TM ? G / 2 : G; // tiles whose rows live in shared memory (TM: the other half in tensor memory)
    Slots2<REMAP> slots;
    slots.base = smem_addr(s_dyn + (blockDim.x >> 5) * Stream::stride()) + ((warp * n_rows * 32 + lane) * GS) * 8 -
                 (REMAP ? 0 : 256 * GS);      // slot id s lives in row s - 1 (id 0 is "no operand")
    slots.limit = uint32_t(n_rows) * 256u;
    if (REMAP) { ts.row_limit = uint32_t(n_rows); ts.bounce_ops = kFastOps; }   // clauses on spilled rows leave the generated loop
    // TM: tile 1's value rows live in tensor memory.  Warp w owns TMEM lanes 32 (w % 4) .. + 31 (the
    // hardware's rule) and the column group w / 4 of the CTA's allocation, 2 n_rows columns wide.
    uint32_t tb = 0;
    // the allocation's address lands in the padding behind warp 0's chunk buffer and mbarrier (no static
    // shared memory: the kernels opt in to the whole dynamic carve-out)
    uint32_t& s_tmem = *reinterpret_cast<uint32_t*>(s_dyn + kChunk * 8 + 16);
    if (TM) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"l"((unsigned long long)__cvta_generic_to_shared(&s_tmem)), "r"(a.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tb = s_tmem + ((uint32_t(warp & 3) * 32u) << 16) + uint32_t(warp >> 2) * uint32_t(G * n_rows) - uint32_t(G);
    }
    const uint64_t* const arena = a.arena;
    unsigned long long st_tiles = 0, st_cells = 0, st_items = 0;
    const uint32_t root_hdr = uint32_t(arena[0]);
    const int n_items = min(*a.n_items, a.tiles_cap);
    const uint32_t tps = a.tps;
    const int size = tps * 8;
    const float recip = 1.0f / float(tps * 8u);
    const float* m = mat.d;
    constexpr int SHIFT = GS == 4 ? 2 : (GS == 2 ? 1 : 0);

    for (;;) {
        const int item = warp_next(a.queue);
        if (item >= n_items) break;
        int start, count;
        unpack_item(__ldg(&a.items[item]), start, count);
        const int tape = __shfl_sync(kFull, a.tiles[start].tape, 0);   // warp-uniform, and provably so
        uint32_t h = ts.begin_tape(root_hdr);
        if (!REMAP) h = (h & 0xff) | ((h & 0xffffff00u) << SHIFT);
        int px[G], py[G];
        #pragma unroll
        for (int g = 0; g < G; ++g) {
            const int position = a.tiles[start + min(g, count - 1)].position;   // short items repeat their last tile
            const int tx = position % tps, ty = (position / tps) % tps;
            px[g] = tx * 8 + (lane & 7);
            py[g] = ty * 8 + (lane >> 3);
            const float fx = sample_coord(px[g], recip);
            const float fya = sample_coord(py[g], recip), fyb = sample_coord(py[g] + 4, recip);
            const float wa = dot2(m[2], fx, m[5], fya, m[8]);
            const float wb = dot2(m[2], fx, m[5], fyb, m[8]);
            const float2 X = make_float2(dot2(m[0], fx, m[3], fya, m[6]) / wa, dot2(m[0], fx, m[3], fyb, m[6]) / wb);
            const float2 Y = make_float2(dot2(m[1], fx, m[4], fya, m[7]) / wa, dot2(m[1], fx, m[4], fyb, m[7]) / wb);
            if (REMAP) {
                slots.st(off_out2(h), X);
                slots.st(off_lhs2(h), Y);
                slots.st(off_rhs2(h), make_float2(a.z, a.z));
            } else {
                if (TM && g >= GS) {                                                    // the second half of the tiles: tensor memory
                    if (off_out2(h)) tm_st2(tb + tm_col(off_out2(h)) + 2 * (g - GS), X);
                    if (off_lhs2(h)) tm_st2(tb + tm_col(off_lhs2(h)) + 2 * (g - GS), Y);
                    if (off_rhs2(h)) tm_st2(tb + tm_col(off_rhs2(h)) + 2 * (g - GS), make_float2(a.z, a.z));
                } else {
                    if (off_out2(h)) sts_f2(slots.base + off_out2(h) + 8 * g, X);      // id 0: axis unused, no row
                    if (off_lhs2(h)) sts_f2(slots.base + off_lhs2(h) + 8 * g, Y);
                    if (off_rhs2(h)) sts_f2(slots.base + off_rhs2(h) + 8 * g, make_float2(a.z, a.z));
                }
            }
        }
        unsigned cells = 0;
        float2 r[G];
        walk_float<REMAP, G, TM>(ts, tape, slots, cells, r, tb);
        #pragma unroll
        for (int g = 0; g < G; ++g) {
            if (g >= count) break;
            if (r[g].y < 0.0f) a.image[px[g] + (py[g] + 4) * size] = 1;      // context.cu:951-962
            if (r[g].x < 0.0f) a.image[px[g] + py[g] * size] = 1;
            if (HEAT) {                                             // work / 2 on each sample (context.cu:1979-1980)
                const unsigned long long u = (unsigned long long)(tape == 0 ? unsigned(a.n_root) : cells) * 2048u;
                atomicAdd(&a.heat[px[g] + size_t(py[g]) * size], u);
                atomicAdd(&a.heat[px[g] + size_t(py[g] + 4) * size], u);
            }
        }
        st_items += 1;
        st_tiles += count;
        st_cells += (unsigned long long)cells * count;
    }
    ts.drain();
    if (lane == 0 && st_tiles) {
        atomicAdd(&a.ctl->stats[ST_F_TILES], st_tiles);
        atomicAdd(&a.ctl->stats[ST_F_CELLS], st_cells);
        atomicAdd(&a.ctl->stats[ST_F_ITEMS], st_items);
    }
    if (TM) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(a.tmem_cols) : "memory");
    }
}

// 3D: one warp per work item of up to G surviving 4x4x4 tiles, two voxels per tile and lane (z and z + 2).
// Root tiles are issued highest-z first and children inherit that order, so the
// list is roughly front-to-back and the per-lane early-out below (the
// reference's, context.cu:852-864) culls most of what lies behind the surface.
template <bool REMAP, bool HEAT = false, int G = 1, bool TM = false>
__global__ void __launch_bounds__(kFloatMaxThreads)
k_eval_voxels(const EvalVoxelsArgs a, const Mat4 mat)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const int lane = lane_id();
    const int warp = __shfl_sync(kFull, int(threadIdx.x >> 5), 0);   // via shuffle: provably warp-uniform for ptxas
    typedef TapeStream<REMAP, REMAP> Stream;   // look-ahead only where slots are renamed (see tape_stream.cuh)
    Stream ts;
    ts.init(s_dyn + warp * Stream::stride(), a.arena, a.arena_cap);
    const int n_rows = a.n_rows;       // shared-memory value rows per warp (= slot count unless REMAP)
    constexpr int GS = TM ? G / 2 : G; // tiles whose rows live in shared memory (TM: the other half in tensor memory)
    Slots2<REMAP> slots;
    slots.base = smem_addr(s_dyn + (blockDim.x >> 5) * Stream::stride()) + ((warp * n_rows * 32 + lane) * GS) * 8 -
                 (REMAP ? 0 : 256 * GS);      // slot id s lives in row s - 1 (id 0 is "no operand")
    slots.limit = uint32_t(n_rows) * 256u;
    if (REMAP) { ts.row_limit = uint32_t(n_rows); ts.bounce_ops = kFastOps; }   // clauses on spilled rows leave the generated loop
    // TM: tile 1's value rows live in tensor memory.  Warp w owns TMEM lanes 32 (w % 4) .. + 31 (the
    // hardware's rule) and the column group w / 4 of the CTA's allocation, 2 n_rows columns wide.
    uint32_t tb = 0;
    // the allocation's address lands in the padding behind warp 0's chunk buffer and mbarrier (no static
    // shared memory: the kernels opt in to the whole dynamic carve-out)
    uint32_t& s_tmem = *reinterpret_cast<uint32_t*>(s_dyn + kChunk * 8 + 16);
    if (TM) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"l"((unsigned long long)__cvta_generic_to_shared(&s_tmem)), "r"(a.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tb = s_tmem + ((uint32_t(warp & 3) * 32u) << 16) + uint32_t(warp >> 2) * uint32_t(G * n_rows) - uint32_t(G);
    }
    const uint64_t* const arena = a.arena;
    unsigned long long st_tiles = 0, st_cells = 0, st_items = 0;
    const uint32_t root_hdr = uint32_t(arena[0]);
    const int n_items = min(*a.n_items, a.tiles_cap);
    const uint32_t tps = a.tps;
    const int size = tps * 4;
    const float recip = 1.0f / float(tps * 4u);
    const float* m = mat.d;
    constexpr int SHIFT = GS == 4 ? 2 : (GS == 2 ? 1 : 0);

    for (;;) {
        const int item = warp_next(a.queue);
        if (item >= n_items) break;
        int start, count;
        unpack_item(__ldg(&a.items[item]), start, count);
        const int tape = __shfl_sync(kFull, a.tiles[start].tape, 0);   // warp-uniform, and provably so
        int px[G], py[G], pz[G];
        bool alive[G];
        bool any = false;
        #pragma unroll
        for (int g = 0; g < G; ++g) {
            const int position = a.tiles[start + min(g, count - 1)].position;   // short items repeat their last tile
            const int tx = position % tps, ty = (position / tps) % tps, tz = (position / tps) / tps;
            px[g] = tx * 4 + (lane & 3);
            py[g] = ty * 4 + ((lane >> 2) & 3);
            pz[g] = tz * 4 + (lane >> 4);            // second sample sits at pz + 2
            // This column already shows something at least as high (context.cu:852-864)
            alive[g] = g < count && __ldcg(&a.image[px[g] + py[g] * size]) < pz[g] + 2;
            any |= alive[g];
        }
        if (!__any_sync(kFull, any)) continue;

        uint32_t hdr = ts.begin_tape(root_hdr);
        if (!REMAP) hdr = (hdr & 0xff) | ((hdr & 0xffffff00u) << SHIFT);
        #pragma unroll
        for (int g = 0; g < G; ++g) {
            const float fx = sample_coord(px[g], recip), fy = sample_coord(py[g], recip);
            const float fza = sample_coord(pz[g], recip), fzb = sample_coord(pz[g] + 2, recip);
            const float wa = dot3(m[3], fx, m[7], fy, m[11], fza, m[15]);
            const float wb = dot3(m[3], fx, m[7], fy, m[11], fzb, m[15]);
            const float2 X = make_float2(dot3(m[0], fx, m[4], fy, m[8], fza, m[12]) / wa,
                                         dot3(m[0], fx, m[4], fy, m[8], fzb, m[12]) / wb);
            const float2 Y = make_float2(dot3(m[1], fx, m[5], fy, m[9], fza, m[13]) / wa,
                                         dot3(m[1], fx, m[5], fy, m[9], fzb, m[13]) / wb);
            const float2 Z = make_float2(dot3(m[2], fx, m[6], fy, m[10], fza, m[14]) / wa,
                                         dot3(m[2], fx, m[6], fy, m[10], fzb, m[14]) / wb);
            if (REMAP) {
                slots.st(off_out2(hdr), X);
                slots.st(off_lhs2(hdr), Y);
                slots.st(off_rhs2(hdr), Z);
            } else {
                if (TM && g >= GS) {                                                    // the second half of the tiles: tensor memory
                    if (off_out2(hdr)) tm_st2(tb + tm_col(off_out2(hdr)) + 2 * (g - GS), X);
                    if (off_lhs2(hdr)) tm_st2(tb + tm_col(off_lhs2(hdr)) + 2 * (g - GS), Y);
                    if (off_rhs2(hdr)) tm_st2(tb + tm_col(off_rhs2(hdr)) + 2 * (g - GS), Z);
                } else {
                    if (off_out2(hdr)) sts_f2(slots.base + off_out2(hdr) + 8 * g, X);  // id 0: axis unused, no row
                    if (off_lhs2(hdr)) sts_f2(slots.base + off_lhs2(hdr) + 8 * g, Y);
                    if (off_rhs2(hdr)) sts_f2(slots.base + off_rhs2(hdr) + 8 * g, Z);
                }
            }
        }
        unsigned cells = 0;
        float2 r[G];
        walk_float<REMAP, G, TM>(ts, tape, slots, cells, r, tb);
        #pragma unroll
        for (int g = 0; g < G; ++g) {
            if (alive[g]) {
                int* const pix = &a.image[px[g] + py[g] * size];
                // The higher sample wins when both are inside (context.cu:936-948)
                if (r[g].y < 0.0f) atomicMax(pix, pz[g] + 2);
                else if (r[g].x < 0.0f) atomicMax(pix, pz[g]);
                if (HEAT)                                           // context.cu:1962
                    atomicAdd(&a.heat[px[g] + size_t(py[g]) * size],
                              (unsigned long long)(tape == 0 ? unsigned(a.n_root) : cells) * 4096u);
            }
        }
        st_items += 1;
        st_tiles += count;
        st_cells += (unsigned long long)cells * count;
    }
    ts.drain();
    if (lane == 0 && st_tiles) {
        atomicAdd(&a.ctl->stats[ST_F_TILES], st_tiles);
        atomicAdd(&a.ctl->stats[ST_F_CELLS], st_cells);
        atomicAdd(&a.ctl->stats[ST_F_ITEMS], st_items);
    }
    if (TM) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(a.tmem_cols) : "memory");
    }
}

////////////////////////////////////////////////////////////////////////////////
// Normal pass (eval_pixels_d, context.cu:978-1132)

template <bool LOCAL>
__global__ void __launch_bounds__(kEvalThreads)
k_normals(const NormalsArgs a, const Mat4 mat)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    Slots4<LOCAL> slots;
    slots.base = smem_addr(s_dyn) + (warp * a.n_slots * 32 + lane) * 16;
    unsigned long long st_px = 0, st_cells = 0;
    const uint64_t* const arena = a.arena;
    const uint32_t h = uint32_t(arena[0]);
    const int size = a.size;
    const int tps0 = size / 64;
    const int n_items = a.n_owned * 128;        // a warp owns an 8 x 4 pixel block; 8 x 16 of them per screen tile

    for (;;) {
        const int item = warp_next(a.queue);
        if (item >= n_items) break;
        const int tile = __ldg(&a.owned[item >> 7]), blk = item & 127;
        const int px = (tile % tps0) * 64 + (blk & 7) * 8 + (lane & 7);
        const int py = (tile / tps0) * 64 + (blk >> 3) * 4 + (lane >> 3);
        const int pxy = px + py * size;
        int pz = a.image[pxy];
        int tape = -1;
        if (pz != 0) {
            // Step just in front of the surface unless that leaves the volume (context.cu:997-1005)
            if (pz < size - 1) pz += 1;
            // Deepest tile that contains this voxel (context.cu:1034-1066)
            const int t0n = size / 64;
            const int t0 = px / 64 + (py / 64) * t0n + (pz / 64) * t0n * t0n;
            const TileNode n0 = a.tiles0[t0];
            if (n0.next == -1) {
                tape = n0.tape;
            } else {
                const int t1 = n0.next * 64 + (px % 64) / 16 + ((py % 64) / 16) * 4 + ((pz % 64) / 16) * 16;
                const TileNode n1 = a.tiles1[t1];
                if (n1.next == -1) {
                    tape = n1.tape;
                } else {
                    const int t2 = n1.next * 64 + (px % 16) / 4 + ((py % 16) / 4) * 4 + ((pz % 16) / 4) * 16;
                    tape = a.tiles2[t2].tape;
                }
            }
        }
        unsigned todo = __ballot_sync(kFull, tape >= 0);
        if (!todo) continue;

        // Sample position and seed gradients (context.cu:1009-1029)
        dval sx_, sy_, sz_;
        {
            const float recip = 1.0f / float(size);
            const float fx = sample_coord(px, recip), fy = sample_coord(py, recip);
            const float fz = sample_coord(pz, recip);
            const float* m = mat.d;
            const float w = dot3(m[3], fx, m[7], fy, m[11], fz, m[15]);
            sx_ = dv(dot3(m[0], fx, m[4], fy, m[8], fz, m[12]) / w, 1.0f, 0.0f, 0.0f);
            sy_ = dv(dot3(m[1], fx, m[5], fy, m[9], fz, m[13]) / w, 0.0f, 1.0f, 0.0f);
            sz_ = dv(dot3(m[2], fx, m[6], fy, m[10], fz, m[14]) / w, 0.0f, 0.0f, 1.0f);
        }

        dval result = dv_const(0.0f);
        unsigned my_cells = 0;
        // Each lane walks ITS tile's tape.  Lanes whose pixels fall in the same 4^3 tile
        // share a tape and stay in lockstep; lanes on different tapes sit at different
        // cells, and only the opcode switch diverges (fetch, decode, operand loads and the
        // store are the same instructions for everyone).  Slots are [slot][lane] float4
        // rows, so per-lane slot indices are still bank-conflict free.
        {
            slots.st(off_out2(h), sx_);
            slots.st(off_lhs2(h), sy_);
            slots.st(off_rhs2(h), sz_);
            bool active = tape >= 0;
            int pos = active ? tape : 0;
            // The walk is a chain of dependent loads (clause -> operands -> next clause) with few
            // warps to hide it, so the NEXT clause is fetched while this one runs; only a JUMP
            // (one per 62 cells) invalidates the prefetch.  The arena has a chunk of slack behind
            // it, so reading one cell past an END cell is in bounds.  (Four cells in flight instead of
            // one changed nothing: the walk is bound by the lanes' opcode paths taking turns.)
            uint64_t ahead = __ldg(&arena[pos + 1]);
            while (__any_sync(kFull, active)) {
                if (active) {
                    const uint64_t d = ahead;
                    ++pos;
                    ahead = __ldg(&arena[pos + 1]);
                    const uint32_t w = uint32_t(d);
                    const uint32_t op = w & 0xff;
                    if (op == OP_END) {
                        result = slots.ld(off_out2(w));
                        active = false;
                    } else if (op == OP_JUMP) {
                        pos += int32_t(d >> 32);
                        ahead = __ldg(&arena[pos + 1]);
                        ++my_cells;
                    } else {
                        ++my_cells;
                        const float imm = __uint_as_float(uint32_t(d >> 32));
                        const dval L = slots.ld(off_lhs2(w));
                        const dval R = slots.ld(off_rhs2(w));
                        dval o;
                        switch (op) {   // context.cu:1081-1114
                            case OP_SQUARE: o = dv_mul(L, L); break;
                            case OP_SQRT:   o = dv_sqrt(L); break;
                            case OP_NEG:    o = dv_neg(L); break;
                            case OP_SIN:    o = dv_sin(L); break;
                            case OP_COS:    o = dv_cos(L); break;
                            case OP_ASIN:   o = dv_asin(L); break;
                            case OP_ACOS:   o = dv_acos(L); break;
                            case OP_ATAN:   o = dv_atan(L); break;
                            case OP_EXP:    o = dv_exp(L); break;
                            case OP_ABS:    o = dv_abs(L); break;
                            case OP_LOG:    o = dv_log(L); break;
                            case OP_ADD_LI: o = dv_add(L, imm); break;
                            case OP_ADD_LR: o = dv_add(L, R); break;
                            case OP_MUL_LI: o = dv_mul(L, imm); break;
                            case OP_MUL_LR: o = dv_mul(L, R); break;
                            case OP_MIN_LI: o = dv_min(L, imm); break;
                            case OP_MIN_LR: o = dv_min(L, R); break;
                            case OP_MAX_LI: o = dv_max(L, imm); break;
                            case OP_MAX_LR: o = dv_max(L, R); break;
                            case OP_SUB_LI: o = dv_sub(L, imm); break;
                            case OP_SUB_IR: o = dv_sub(imm, R); break;
                            case OP_SUB_LR: o = dv_sub(L, R); break;
                            case OP_DIV_LI: o = dv_div(L, imm); break;
                            case OP_DIV_IR: o = dv_div(imm, R); break;
                            case OP_DIV_LR: o = dv_div(L, R); break;
                            case OP_COPY_IMM: o = dv_const(imm); break;
                            case OP_COPY_LHS: o = L; break;
                            case OP_COPY_RHS: o = R; break;
                            default: o = L; break;
                        }
                        slots.st(off_out2(w), o);
                    }
                }
            }
        }

        if (tape >= 0) {
            // context.cu:1123-1131 (SASS: powf x3, FADD, FADD, sqrt; div; FFMA 127,128; F2I.U32.TRUNC)
            const float norm = sqrtf(__fadd_rn(__fadd_rn(powf(result.x, 2), powf(result.y, 2)),
                                               powf(result.z, 2)));
            const uint8_t bx = __fmaf_rn(__fdiv_rn(result.x, norm), 127.0f, 128.0f);
            const uint8_t by = __fmaf_rn(__fdiv_rn(result.y, norm), 127.0f, 128.0f);
            const uint8_t bz = __fmaf_rn(__fdiv_rn(result.z, norm), 127.0f, 128.0f);
            a.normals[pxy] = (0xFFu << 24) | (uint32_t(bz) << 16) | (uint32_t(by) << 8) | bx;
        }
        {
            st_px += __popc(__ballot_sync(kFull, tape >= 0));
            st_cells += warp_sum(tape >= 0 ? my_cells : 0u);
        }
    }
    if (lane == 0 && st_px) {
        atomicAdd(&a.ctl->stats[ST_N_PIXELS], st_px);
        atomicAdd(&a.ctl->stats[ST_N_CELLS], st_cells);
    }
}

////////////////////////////////////////////////////////////////////////////////
// Frame setup in ONE launch: clear the control block and set the arena allocation cursor (block 0), copy the
// root tape to cell 0 of the arena, clear the level-0 image and (3D) the normal image.  The arena and the
// images are managed memory (the reference exposes them as host-readable pointers); cudaMemcpyAsync /
// cudaMemsetAsync on managed ranges go through the unified-memory driver and were measured to stall the
// host for 0.3-0.6 ms every few frames (frames of 0.43 ms came out at 0.86), a kernel that touches pages
// already resident on the device does not.  `root` may be device memory or page-locked host memory (the
// host-buffer entry points stage the cells there and the kernel reads them over PCIe).
__global__ void __launch_bounds__(256)
k_begin_frame(FrameCtl* ctl, int32_t first_free, uint64_t* __restrict__ arena, const uint64_t* __restrict__ root,
              int32_t n_root, uint4* __restrict__ image0, int32_t n_image16, int32_t* __restrict__ image0_tail,
              int32_t n_tail, uint4* __restrict__ normals, long long n_normals16)
{
    if (blockIdx.x == 0) {
        int32_t* raw = reinterpret_cast<int32_t*>(ctl);
        for (int k = threadIdx.x; k < int(sizeof(FrameCtl) / 4); k += blockDim.x) raw[k] = 0;
        __syncthreads();
        if (threadIdx.x == 0) ctl->tape_cursor = first_free;
    }
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = t; i < n_root; i += stride) arena[i] = root[i];
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (long long i = t; i < n_image16; i += stride) image0[i] = zero;
    if (t < n_tail) image0_tail[t] = 0;
    for (long long i = t; i < n_normals16; i += stride) normals[i] = zero;
}

// Brute-force frames (reference preload_tiles, context.cu:45-57): every 8x8 tile goes straight
// to the float pass with the root tape.
__global__ void k_preload_tiles(TileNode* __restrict__ tiles, int32_t* __restrict__ items, int32_t count,
                                int32_t* __restrict__ n_tiles, int32_t* __restrict__ n_items)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        tiles[i].position = i;
        tiles[i].tape = 0;
        tiles[i].next = -1;
        items[i] = i * 8 + 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { *n_tiles = count; *n_items = count; }
}

// Work units -> the reference's heatmap value: cells per pixel over the clause count
// (context.cu:2140-2144).
__global__ void k_heat_finish(const unsigned long long* __restrict__ units, float* __restrict__ heat, long long n,
                              int32_t n_clauses)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        heat[i] = __fdiv_rn(float(double(units[i]) / 4096.0), float(n_clauses));
}

////////////////////////////////////////////////////////////////////////////////
// Launch wrappers

// Opt every kernel in to the device's full dynamic shared memory once.
// Dynamic shared memory of an interval / float tape-walking CTA: per warp one chunk stream plus
// 256-byte value rows - one per slot id, or kRemapRows when the stream renames slots.
bool use_remap(int n_slots);
static size_t walk_smem(int n_rows, bool remap, int warps = kEvalWarps, int group = 1, bool float_pass = false) {
    const int stream = remap ? kStreamStrideRemap : (float_pass ? kStreamStrideSingle : kStreamStridePlain);
    return size_t(warps) * (size_t(n_rows) * 256 * group + stream);
}
// Shared-memory value rows per warp: one per slot id, or a fixed budget when slots are renamed.
int walk_rows(int n_slots) {
    // Slot id 0 means "no operand" (src/tape.cpp:70) and never holds a value: ids 1 .. n_slots - 1
    // get the rows, and the walkers address row (id - 1) by starting one row early.
    if (!use_remap(n_slots)) return n_slots > 1 ? n_slots - 1 : 1;
    static const char* env = getenv("MPRB_REMAP_ROWS");
    const int rows = env ? atoi(env) : kRemapRowsDefault;
    return rows < kRemapRowsMin ? kRemapRowsMin : (rows > 128 ? 128 : rows);
}
// Shared-memory value rows per warp of the FLOAT pass.  With renamed slots its clause loop is the generated
// one, which only runs clauses whose rows are all in shared memory (the others bounce through C++ with the
// spilling accessors), so it wants more rows than the interval pass, whose C++ walker spills by itself.
// MPRB_FLOAT_ROWS overrides.
int float_rows(int n_slots) {
    if (!use_remap(n_slots)) return walk_rows(n_slots);
    static const char* env = getenv("MPRB_FLOAT_ROWS");
    const int rows = env ? atoi(env) : kFloatRemapRowsDefault;
    return std::min(std::max(rows, kRemapRowsMin), std::min(n_slots, 128));
}
// Tiles per work item of the float pass (1, 2 or 4).  Tiles of an item share their tape, and the
// clause loop is bound by fetch + dispatch, so G tiles cost little more than one - but slot rows
// grow G-fold and with them the shared memory per warp.  MPRB_FLOAT_GROUP overrides.
// Default: two tiles per item with the second tile's rows in TENSOR MEMORY (float_tmem): the shared
// memory footprint - and with it the number of resident warps - stays that of one tile.
bool float_tmem(int n_slots, bool heat) {
    if (use_remap(n_slots) || heat) return false;
    static const char* env = getenv("MPRB_FLOAT_TMEM");
    static const char* grp = getenv("MPRB_FLOAT_GROUP");
    if (env) return env[0] != '0';
    return grp == nullptr;              // an explicit group size means the shared-memory-only variants
}
int float_group(int n_slots, bool heat) {
    if (use_remap(n_slots) || heat) return 1;
    if (float_tmem(n_slots, heat)) {
        // MPRB_FLOAT_TMEM_GROUP=4: four tiles per item, two in shared memory and two in tensor memory - the same
        // fetch / decode / branch / loads / store serve four tiles (items fill to 3.7 tiles on bear), but only 19
        // warps fit an SM and the walk becomes latency-bound: bear 1024^3 float pass 5.05 ms against 3.59 ms
        static const char* tg = getenv("MPRB_FLOAT_TMEM_GROUP");
        return (tg && tg[0] == '4') ? 4 : 2;
    }
    static const char* env = getenv("MPRB_FLOAT_GROUP");
    if (env) { const int v = atoi(env); return v >= 4 ? 4 : (v >= 2 ? 2 : 1); }
    return 1;
}
// CTA shape of the float pass: the warps per CTA (and CTAs per SM) that keep the most warps resident
// given the per-warp shared memory (value rows + chunk stream), 1 KB the driver reserves per CTA, the
// limits of 32 CTAs / 64 warps per SM and - with tensor-memory rows - the 512 TMEM columns of an SM:
// a CTA allocates a power of two >= 32 columns, 2 n_rows for each group of 4 warps (the four warps of a
// group sit in the four 32-lane quarters).  MPRB_FLOAT_WARPS overrides the warps per CTA.
struct FloatShape { int warps, ctas, tmem_cols; };
static int pow2_at_least(int v) { int p = 32; while (p < v) p *= 2; return p; }
static FloatShape float_shape(int n_slots, int group, bool tmem) {
    static const char* env = getenv("MPRB_FLOAT_WARPS");
    static int smem_per_sm = 0;
    if (!smem_per_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    }
    const int rows = float_rows(n_slots);
    const size_t per_warp = walk_smem(rows, use_remap(n_slots), 1, tmem ? group / 2 : group, true);
    FloatShape best = {1, 1, 32};
    int best_resident = 0;
    for (int w = 1; w <= kFloatMaxThreads / 32; ++w) {
        if (env && w != std::min(std::max(atoi(env), 1), 32)) continue;
        int ctas = int(size_t(smem_per_sm) / (w * per_warp + 1024));
        if (ctas > 32) ctas = 32;
        if (ctas > 64 / w) ctas = 64 / w;
        int cols = 32;
        if (tmem) {
            cols = pow2_at_least(((w + 3) / 4) * group * rows);       // half of `group` tiles, two columns each, per slot
            if (cols > 512) continue;
            ctas = std::min(ctas, 512 / cols);
        }
        if (ctas * w > best_resident) { best_resident = ctas * w; best = {w, ctas, cols}; }
    }
    return best;
}
int float_warps(int n_slots, int group) { return float_shape(n_slots, group, false).warps; }
// Renaming pays once per-id rows would leave fewer than ~24 warps per SM.
bool use_remap(int n_slots) {
    static const char* force = getenv("MPRB_REMAP");
    // Kernels without renaming index a 32-bit live set by slot id (SlotSet<false>): renaming is
    // mandatory above 32 slots, whatever MPRB_REMAP says.
    if (n_slots > 32) return true;
    if (force) return force[0] == '1';
    return false;
}
// Normal pass (float4 values, no stream): shared rows, or local memory for many slots.
static size_t normals_smem(int n_slots, bool local) {
    return local ? 0 : size_t(kEvalWarps) * n_slots * 512;
}
bool use_local_normals(int n_slots) {
    static const char* force = getenv("MPRB_LOCAL_SLOTS");
    if (force) return force[0] == '1';
    return n_slots > 18;
}

// Float-pass variants without renaming: G tiles per work item, tile 1 in tensor memory or not.
template <typename F> static auto pick_float(int G, bool tmem, F f) {
    typedef std::integral_constant<bool, true> T;
    typedef std::integral_constant<bool, false> N;
    if (tmem && G == 4) return f(std::integral_constant<int, 4>(), T());
    if (tmem) return f(std::integral_constant<int, 2>(), T());
    if (G == 4) return f(std::integral_constant<int, 4>(), N());
    if (G == 2) return f(std::integral_constant<int, 2>(), N());
    return f(std::integral_constant<int, 1>(), N());
}

// Also pin the L1/shared split to "all shared" for the shared-memory variants: occupancy there
// is bounded by shared memory, and a device-wide cudaDeviceSetCacheConfig(PreferL1) made by
// other code in the process (the reference's Context constructor does that, context.cpp:47-48)
// would otherwise shrink the carve-out to one CTA per SM.  Local-slot variants prefer L1.
template <typename K>
static void opt_in(K kernel, int max_smem_optin, bool local = false) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, local ? 8 * 1024 : max_smem_optin);
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         local ? cudaSharedmemCarveoutMaxL1 : cudaSharedmemCarveoutMaxShared);
}
void init_kernels(int max_smem_optin) {
    opt_in(k_eval_tiles<2, true, false>, max_smem_optin);
    opt_in(k_eval_tiles<2, false, false>, max_smem_optin);
    opt_in(k_eval_tiles<3, true, false>, max_smem_optin);
    opt_in(k_eval_tiles<3, false, false>, max_smem_optin);
    opt_in(k_eval_tiles<2, true, true>, max_smem_optin);
    opt_in(k_eval_tiles<2, false, true>, max_smem_optin);
    opt_in(k_eval_tiles<3, true, true>, max_smem_optin);
    opt_in(k_eval_tiles<3, false, true>, max_smem_optin);
    opt_in(k_eval_root<2>, max_smem_optin);
    opt_in(k_eval_root<3>, max_smem_optin);
    opt_in(k_eval_sub<2>, max_smem_optin);
    opt_in(k_eval_sub<3>, max_smem_optin);
    opt_in(k_eval_tiles<2, true, false, true>, max_smem_optin);
    opt_in(k_eval_tiles<2, false, false, true>, max_smem_optin);
    opt_in(k_eval_tiles<3, true, false, true>, max_smem_optin);
    opt_in(k_eval_tiles<3, false, false, true>, max_smem_optin);
    opt_in(k_eval_tiles<2, true, true, true>, max_smem_optin);
    opt_in(k_eval_tiles<2, false, true, true>, max_smem_optin);
    opt_in(k_eval_tiles<3, true, true, true>, max_smem_optin);
    opt_in(k_eval_tiles<3, false, true, true>, max_smem_optin);
    opt_in(k_eval_pixels<false, true>, max_smem_optin);
    opt_in(k_eval_voxels<false, true>, max_smem_optin);
    opt_in(k_eval_pixels<true, true>, max_smem_optin);
    opt_in(k_eval_voxels<true, true>, max_smem_optin);
    for (int G = 1; G <= 4; G *= 2)
        for (int T = 0; T <= (G >= 2 ? 1 : 0); ++T)
            pick_float(G, T != 0, [&](auto g, auto t) {
                opt_in(k_eval_pixels<false, false, decltype(g)::value, decltype(t)::value>, max_smem_optin);
                opt_in(k_eval_voxels<false, false, decltype(g)::value, decltype(t)::value>, max_smem_optin);
                return 0;
            });
    opt_in(k_normals<false>, max_smem_optin);
    opt_in(k_eval_pixels<true>, max_smem_optin);
    opt_in(k_eval_voxels<true>, max_smem_optin);
    opt_in(k_normals<true>, max_smem_optin, true);
}

template <int DIM, bool ROOT>
static void launch_eval_tiles_t(const EvalTilesArgs& a, const void* mat, int grid, cudaStream_t s) {
    const bool local = use_remap(a.n_slots);
    const size_t smem = walk_smem(a.n_rows, local);
    const auto& m = *static_cast<const typename MatOf<DIM>::type*>(mat);
    if (a.heat) {
        if (local) k_eval_tiles<DIM, ROOT, true, true><<<grid, kEvalThreads, smem, s>>>(a, m);
        else k_eval_tiles<DIM, ROOT, false, true><<<grid, kEvalThreads, smem, s>>>(a, m);
    } else if (local) k_eval_tiles<DIM, ROOT, true><<<grid, kEvalThreads, smem, s>>>(a, m);
    else k_eval_tiles<DIM, ROOT, false><<<grid, kEvalThreads, smem, s>>>(a, m);
}

void launch_eval_tiles(int dim, bool root, const EvalTilesArgs& a, const void* mat, int grid, cudaStream_t s) {
    if (dim == 3) {
        if (root) launch_eval_tiles_t<3, true>(a, mat, grid, s);
        else launch_eval_tiles_t<3, false>(a, mat, grid, s);
    } else {
        if (root) launch_eval_tiles_t<2, true>(a, mat, grid, s);
        else launch_eval_tiles_t<2, false>(a, mat, grid, s);
    }
}

void launch_eval_root(int dim, const EvalRootArgs& a, const void* mat, cudaStream_t s) {
    const int tiles_per_cta = kRootThreads / a.group;
    const int grid = (a.count0 + tiles_per_cta - 1) / tiles_per_cta;
    const size_t smem = size_t(tiles_per_cta) * a.smem_per_tile;
    if (dim == 3) k_eval_root<3><<<grid, kRootThreads, smem, s>>>(a, *static_cast<const Mat4*>(mat));
    else k_eval_root<2><<<grid, kRootThreads, smem, s>>>(a, *static_cast<const Mat3*>(mat));
}

int sub_warps(int slice) {
    return std::max(1, std::min(kSubMaxWarps, (110 * 1024) / std::max(slice, 1)));
}

void launch_eval_sub(int dim, const EvalSubArgs& a, const void* mat, int grid, cudaStream_t s) {
    const int warps = sub_warps(a.slice);
    const size_t smem = size_t(warps) * a.slice;
    if (dim == 3) k_eval_sub<3><<<grid, warps * 32, smem, s>>>(a, *static_cast<const Mat4*>(mat));
    else k_eval_sub<2><<<grid, warps * 32, smem, s>>>(a, *static_cast<const Mat3*>(mat));
}

void launch_rank_tiles(int dim, const RankArgs& a, int grid, cudaStream_t s) {
    if (dim == 3) k_rank_tiles<3><<<grid, 256, 0, s>>>(a);
    else k_rank_tiles<2><<<grid, 256, 0, s>>>(a);
}

void launch_upsample_filled(int dim, const int32_t* prev, int32_t* image, int size, int grid, cudaStream_t s) {
    if (dim == 3) k_upsample_filled<3><<<grid, 256, 0, s>>>(prev, image, size);
    else k_upsample_filled<2><<<grid, 256, 0, s>>>(prev, image, size);
}

// Launches the float pass with the CTA shape of float_shape(); with tensor-memory rows the dynamic shared
// memory is padded so that no more CTAs fit an SM than its TMEM columns can serve (an allocation that
// has to wait for columns would stall a whole CTA).
static size_t float_pad_smem(const FloatShape& sh) {
    int smem_per_sm = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    // one more CTA than `ctas` must not fit: ask for just over 1 / (ctas + 1) of the SM
    return size_t(smem_per_sm) / size_t(sh.ctas + 1) - 1024 + 128;
}

void launch_eval_pixels(const EvalVoxelsArgs& a, const Mat3& mat, int grid, cudaStream_t s) {
    const bool local = use_remap(a.n_slots);
    const bool tm = a.tmem != 0;
    const int G = a.group;
    const FloatShape sh = float_shape(a.n_slots, G, tm);
    EvalVoxelsArgs b = a;
    b.tmem_cols = sh.tmem_cols;
    size_t smem = walk_smem(a.n_rows, local, sh.warps, tm ? G / 2 : G, true);
    if (tm) smem = std::max(smem, float_pad_smem(sh));
    const int fw = sh.warps;
    if (a.heat) {
        if (local) k_eval_pixels<true, true><<<grid, fw * 32, smem, s>>>(b, mat);
        else k_eval_pixels<false, true><<<grid, fw * 32, smem, s>>>(b, mat);
    } else if (local) k_eval_pixels<true><<<grid, fw * 32, smem, s>>>(b, mat);
    else pick_float(G, tm, [&](auto g, auto t) {
        k_eval_pixels<false, false, decltype(g)::value, decltype(t)::value><<<grid, fw * 32, smem, s>>>(b, mat);
        return 0;
    });
}

void launch_eval_voxels(const EvalVoxelsArgs& a, const Mat4& mat, int grid, cudaStream_t s) {
    const bool local = use_remap(a.n_slots);
    const bool tm = a.tmem != 0;
    const int G = a.group;
    const FloatShape sh = float_shape(a.n_slots, G, tm);
    EvalVoxelsArgs b = a;
    b.tmem_cols = sh.tmem_cols;
    size_t smem = walk_smem(a.n_rows, local, sh.warps, tm ? G / 2 : G, true);
    if (tm) smem = std::max(smem, float_pad_smem(sh));
    const int fw = sh.warps;
    if (a.heat) {
        if (local) k_eval_voxels<true, true><<<grid, fw * 32, smem, s>>>(b, mat);
        else k_eval_voxels<false, true><<<grid, fw * 32, smem, s>>>(b, mat);
    } else if (local) k_eval_voxels<true><<<grid, fw * 32, smem, s>>>(b, mat);
    else pick_float(G, tm, [&](auto g, auto t) {
        k_eval_voxels<false, false, decltype(g)::value, decltype(t)::value><<<grid, fw * 32, smem, s>>>(b, mat);
        return 0;
    });
}

void launch_normals(const NormalsArgs& a, const Mat4& mat, int grid, cudaStream_t s) {
    const bool local = use_local_normals(a.n_slots);
    const size_t smem = normals_smem(a.n_slots, local);
    if (local) k_normals<true><<<grid, kEvalThreads, smem, s>>>(a, mat);
    else k_normals<false><<<grid, kEvalThreads, smem, s>>>(a, mat);
}

void launch_preload_tiles(TileNode* tiles, int32_t* items, int32_t count, int32_t* n_tiles, int32_t* n_items, int grid,
                          cudaStream_t s) {
    k_preload_tiles<<<grid, 256, 0, s>>>(tiles, items, count, n_tiles, n_items);
}

void launch_heat_finish(const unsigned long long* units, float* heat, long long n, int32_t n_clauses, int grid,
                        cudaStream_t s) {
    k_heat_finish<<<grid, 256, 0, s>>>(units, heat, n, n_clauses);
}

void launch_begin_frame(FrameCtl* ctl, int32_t first_free, uint64_t* arena, const uint64_t* root, int32_t n_root_cells,
                        int32_t* image0, long long n_image0, uint32_t* normals, long long n_normals, int grid,
                        cudaStream_t s) {
    // the level-0 image has (size / 64)^d entries, a multiple of 4 only from 128 px on: its tail (< 4 entries)
    // is cleared one entry at a time
    const long long n16 = n_image0 / 4;
    k_begin_frame<<<grid, 256, 0, s>>>(ctl, first_free, arena, root, n_root_cells, reinterpret_cast<uint4*>(image0), int32_t(n16), image0 + n16 * 4,
                                       int32_t(n_image0 - n16 * 4), reinterpret_cast<uint4*>(normals), n_normals / 4);
}

template <typename K>
static int occ(K kernel, size_t smem, int threads = kEvalThreads) {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem);
    return n;
}

int occupancy_eval_tiles(int dim, bool root, int n_slots) {
    const bool local = use_remap(n_slots);
    const size_t smem = walk_smem(walk_rows(n_slots), local);
    if (dim == 3) {
        if (root) return local ? occ(k_eval_tiles<3, true, true>, smem) : occ(k_eval_tiles<3, true, false>, smem);
        return local ? occ(k_eval_tiles<3, false, true>, smem) : occ(k_eval_tiles<3, false, false>, smem);
    }
    if (root) return local ? occ(k_eval_tiles<2, true, true>, smem) : occ(k_eval_tiles<2, true, false>, smem);
    return local ? occ(k_eval_tiles<2, false, true>, smem) : occ(k_eval_tiles<2, false, false>, smem);
}

// Resident CTAs per SM of the float pass (sizes its persistent grid).
int float_ctas(int dim, int n_slots, int group, bool tmem) {
    (void)dim;
    return std::max(float_shape(n_slots, group, tmem).ctas, 1);
}

int occupancy_normals(int n_slots) {
    const bool local = use_local_normals(n_slots);
    const size_t smem = normals_smem(n_slots, local);
    return local ? occ(k_normals<true>, smem) : occ(k_normals<false>, smem);
}

}  // namespace mprb
