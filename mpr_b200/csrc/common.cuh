// Shared definitions for the mprb kernels and their host-side launcher.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mprb {

// Same 12-byte record the reference exposes as mpr::TileNode
// (reference inc/context.hpp:23-27): linear position at its level (or -1 once
// resolved), arena index of its tape header, rank among active tiles (or -1).
struct TileNode {
    int32_t position;
    int32_t tape;
    int32_t next;
};

// Clause opcodes; values are the reference's mpr::Opcode enum
// (reference inc/gpu_opcode.hpp:18-56) because they are stored in tapes.
enum : uint32_t {
    OP_END = 0, OP_JUMP = 1,
    OP_SQUARE = 2, OP_SQRT = 3, OP_NEG = 4, OP_SIN = 5, OP_COS = 6, OP_ASIN = 7,
    OP_ACOS = 8, OP_ATAN = 9, OP_EXP = 10, OP_ABS = 11, OP_LOG = 12,
    OP_ADD_LI = 13, OP_ADD_LR = 14, OP_MUL_LI = 15, OP_MUL_LR = 16,
    OP_MIN_LI = 17, OP_MIN_LR = 18, OP_MAX_LI = 19, OP_MAX_LR = 20,
    OP_SUB_LI = 21, OP_SUB_IR = 22, OP_SUB_LR = 23,
    OP_DIV_LI = 24, OP_DIV_IR = 25, OP_DIV_LR = 26,
    OP_COPY_IMM = 27, OP_COPY_LHS = 28, OP_COPY_RHS = 29,
};

constexpr int kChunk = 64;           // cells per arena chunk (reference parameters.hpp:16)
constexpr int kMaxChoices = 4096;    // recorded min/max verdicts per tile (context.cu:218,257)

// Column-major 4x4 / 3x3 transforms, passed by value as kernel arguments
// exactly like the reference passes Eigen matrices (context.cu:81,125,982).
struct Mat4 { float d[16]; };
struct Mat3 { float d[9]; };

// Indices into FrameCtl::stats (all counts are per frame).
enum : int {
    ST_I_TILES = 0,    // [+level] interval tiles evaluated (tape walked)
    ST_I_CELLS = 3,    // [+level] tape cells visited by forward walks, summed over tiles
    ST_P_TILES = 6,    // [+level] tiles that pushed a shortened tape
    ST_P_CELLS = 9,    // [+level] tape cells visited by backward walks, summed over pushing tiles
    ST_P_KEPT = 12,    // [+level] cells written by pushes (clauses + header + end + links)
    ST_F_TILES = 15,   // float-stage tiles evaluated
    ST_F_CELLS = 16,   // float-stage cells visited, summed over tiles
    ST_N_PIXELS = 17,  // normal-pass pixels evaluated
    ST_N_CELLS = 18,   // normal-pass cells visited, summed over pixels
    ST_F_ITEMS = 19,   // float-stage work items (runs of tiles sharing a tape, walked together)
    ST_P_WRITTEN = 20, // arena cells actually written by pushes (tiles with equal verdicts share one tape)
    ST_I_SUB = 21,     // interval tiles evaluated clause-parallel by k_eval_sub (part of ST_I_TILES)
    ST_COUNT = 22,
};

// Device-resident per-frame control block.  Everything the host would
// otherwise have to read back between levels lives here, so a frame is one
// uninterrupted stream of launches.
struct FrameCtl {
    int32_t n_active[4];   // [i] = tiles still ambiguous after interval level i; [3] = float-pass work items
    int32_t queue[10];     // work-queue heads, one per persistent launch
    int32_t overflow;      // bit i set: tile array of stage i+1 too small
    int32_t tape_cursor;   // arena allocation cursor in cells (the reference's *tape_index)
    int32_t plan_cursor;   // plan arena allocation cursor in words (k_eval_root -> k_eval_sub)
    int32_t plan_count;    // plans written so far (capped: only small levels use them)
    unsigned long long stats[ST_COUNT];
};

}  // namespace mprb
