// Kernel argument blocks and host-callable launchers (see kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "common.cuh"

namespace mprb {

constexpr int kEvalWarps = 4;                 // warps per CTA: interval and normal passes
// The float pass picks its CTA size at launch (float_warps): the CTA shape that packs the most
// warps into an SM's shared memory, e.g. 2 x 17 warps for bear's 23 slot rows.  The warp index is
// taken through a shuffle, which ptxas knows to be warp-uniform, so clause decode still runs on
// the uniform datapath whatever the CTA size.
constexpr int kFloatMaxThreads = 1024;
constexpr int kEvalThreads = kEvalWarps * 32;

struct EvalTilesArgs {
    uint64_t* arena;          // tape arena; root tape at cell 0
    int32_t* tape_index;      // arena allocation cursor (cells)
    int32_t arena_cap;        // arena size in cells
    int32_t* image;           // this level's filled image, tps x tps
    TileNode* tiles;          // this level's tile records (written)
    int32_t tiles_cap;        // capacity of `tiles`
    uint32_t tps;             // tiles per side at this level
    const TileNode* ptiles;   // parent level (null at the root)
    const int32_t* pactive;   // parent rank -> index into ptiles
    const int32_t* n_parents; // device count of active parents
    uint32_t ptps;            // parent tiles per side
    int32_t count0;           // root level: number of tiles
    int32_t row_begin;        // root level: this context renders tile rows
    int32_t row_end;          //   [row_begin, row_end) in y (multi-GPU sharding) ...
    int32_t row_mod;          //   ... of which only tiles with (y + col_step * x) % row_mod == row_rem
    int32_t row_rem;
    int32_t col_step;
    FrameCtl* ctl;
    int32_t* queue;           // work-queue head for this launch
    int32_t level;            // 0, 1, 2 (statistics slot / overflow bit)
    int32_t n_slots;          // slot ids used by the root tape, +1
    int32_t n_rows;           // shared-memory value rows per warp (walk_rows(n_slots))
    float z;                  // 2D only: the constant z
    // Work meter (render*_heatmap, context.cu:1513-2340); null on ordinary frames.
    unsigned long long* heat; // S x S accumulators, units of 1/4096 cell
    int32_t heat_px;          // this level's tile edge in pixels
    int32_t n_root;           // clauses of the root tape (its walk is charged without the layout's JUMPs)
    // Small levels (k_eval_sub): parents that carry a clause-parallel plan are left to that kernel
    // when the level has at most sub_max_parents parents.  Null / 0 otherwise.
    const int32_t* plan_of;
    int32_t sub_max_parents;
};

// One clause of a root tape in SSA / dependency-level order (built on the host per Tape).
struct RootClause {
    uint32_t op_idx;   // bits 0-7 opcode, bit 8 "verdict past the 4096-entry record", bits 9 / 10 "left / right operand
                       // sits in the output slot" (root plans), bits 12+ clause index
    uint32_t lsrc;     // value id of the left operand  (0 none, 1..3 x/y/z, 3+i = clause i)
    uint32_t rsrc;     // value id of the right operand
    float imm;
};

constexpr int kRootThreads = 256;

struct EvalRootArgs {
    uint64_t* arena;
    int32_t* tape_index;
    int32_t arena_cap;
    int32_t* image;           // level-0 filled image
    TileNode* tiles;          // level-0 tile records (dense, index = position)
    uint32_t tps;
    int32_t count0;
    int32_t row_begin;
    int32_t row_end;
    int32_t row_mod;
    int32_t row_rem;
    int32_t col_step;
    FrameCtl* ctl;
    const uint64_t* cells;    // the Tape's contiguous cells (header, clauses, end cell)
    const RootClause* sched;  // clauses sorted by (dependency level, opcode)
    const int32_t* level_start;   // n_levels + 1 offsets into sched
    int32_t n_levels;
    int32_t n_clauses;
    int32_t result_v;         // value id of the tape's result
    int32_t group;            // threads per tile: 32, 64, 128 or 256
    int32_t smem_per_tile;    // bytes: values + verdicts + liveness + scratch
    float z;
    // Clause-parallel plans of the shortened tapes (see k_eval_sub); plans == null: none are written.
    int32_t* plans;           // plan arena (32-bit words)
    int32_t* plan_cursor;     // allocation cursor in words (FrameCtl::plan_cursor)
    int32_t plan_cap;         // arena size in words
    int32_t* plan_of;         // per level-0 tile: word offset of its plan, -1 = none
    int32_t sub_slice;        // shared memory k_eval_sub has per tile: larger plans are not written
    int32_t* plan_count;      // FrameCtl::plan_count
    int32_t max_plans;        // no more plans than a small level has parents
    int32_t plan_min;         // shortened tapes with fewer clauses get no plan (kPlanMinClauses; tests lower it)
    const uint16_t* prevw;    // per clause i: value id that lived in its output slot before it (0 = none)
};

// A shortened tape's plan in the plan arena: kPlanHeader words, then n_levels + 1 level offsets, then
// the clauses (RootClause, 4 words each) in dependency-level order.
enum { PL_N = 0,          // clauses of the shortened tape (logical cells 1 .. n)
       PL_LEVELS = 1,     // dependency levels (those of the root tape; some may be empty)
       PL_RESULT = 2,     // value id of the result
       PL_TAPE = 3,       // arena index of the tape itself (contiguous chunked run)
       PL_LOGICAL = 4,    // its logical cell count, n + 2
       PL_VALUES = 5,     // value ids in use: 4 + n + forwarding nodes
       PL_SCHED = 6,      // word offset of the clause array from the plan's first word
       PL_EXTRAS = 7 };   // copies that keep a second source alive: (cell | source << 16) words behind the clause array
constexpr int kPlanHeader = 8;
constexpr int kPlanMaxClauses = 4000;   // < kMaxChoices: every verdict of a planned tape is recorded
constexpr int kPlanMinClauses = 160;    // shorter tapes are walked serially faster than their levels can be swept
constexpr int kSubMaxWarps = 16;

struct EvalSubArgs {
    uint64_t* arena;
    int32_t* tape_index;
    int32_t arena_cap;
    int32_t* image;           // this level's filled image
    TileNode* tiles;          // this level's tile records (written)
    int32_t tiles_cap;
    uint32_t tps;
    const TileNode* ptiles;   // parent level
    const int32_t* pactive;
    const int32_t* n_parents;
    uint32_t ptps;
    FrameCtl* ctl;
    int32_t* queue;
    int32_t level;
    const int32_t* plans;
    const int32_t* plan_of;   // indexed like ptiles
    int32_t slice;            // shared-memory bytes per warp (= per tile)
    int32_t max_parents;      // the kernel stands down when the level has more parents than this
    float z;
};

struct RankArgs {
    TileNode* tiles;          // this level's tile records
    int32_t tiles_cap;
    const int32_t* n_parents; // null at the root level
    int32_t count0;
    int32_t tps;
    const int32_t* image;     // this level's filled image
    int32_t* n_active;        // out: survivors
    int32_t* active_list;     // out: rank -> tile index (not at the last level)
    TileNode* out_tiles;      // out: compact survivor list (last level only), tiles sharing a tape adjacent
    int32_t* items;           // out: float-pass work items, start * 8 + count (last level only)
    int32_t* n_items;         // out: how many
    int32_t gmax;             // tiles per work item (float_group)
    long long next_cap;       // capacity (tiles) of the stage that receives survivors
    int32_t last_level;
    int32_t level;
    FrameCtl* ctl;
};

struct EvalVoxelsArgs {
    const uint64_t* arena;
    int32_t arena_cap;        // arena size in cells
    int32_t* image;           // full-resolution image / heightmap
    const TileNode* tiles;    // compact survivor list of the last interval level
    int32_t tiles_cap;
    const int32_t* items;     // work items: runs of up to `group` tiles sharing a tape (start * 8 + count)
    const int32_t* n_items;
    int32_t group;            // tiles per work item: 1, 2 or 4 (float_group)
    int32_t tmem;             // 1: two tiles per item, tile 1's value rows in tensor memory (float_tmem)
    int32_t tmem_cols;        // tensor-memory columns each CTA allocates (power of two, >= 32)
    uint32_t tps;             // survivor-level tiles per side (size/4 or size/8)
    FrameCtl* ctl;
    int32_t* queue;
    int32_t n_slots;
    int32_t n_rows;           // shared-memory value rows per warp (walk_rows(n_slots))
    float z;
    unsigned long long* heat; // work meter, see EvalTilesArgs
    int32_t n_root;
};

struct NormalsArgs {
    const uint64_t* arena;
    const int32_t* image;     // heightmap
    uint32_t* normals;
    int32_t size;
    const int32_t* owned;     // the 64x64-px screen tiles this context renders (y * tiles_per_side + x) ...
    int32_t n_owned;          // ... and how many
    const TileNode* tiles0;
    const TileNode* tiles1;
    const TileNode* tiles2;
    FrameCtl* ctl;
    int32_t* queue;
    int32_t n_slots;
};

void init_kernels(int max_smem_optin);
void launch_preload_tiles(TileNode* tiles, int32_t* items, int32_t count, int32_t* n_tiles, int32_t* n_items, int grid,
                          cudaStream_t s);
void launch_heat_finish(const unsigned long long* units, float* heat, long long n, int32_t n_clauses, int grid,
                        cudaStream_t s);
// clears the control block, copies the root tape to cell 0 of the arena and clears
// the level-0 image and the normal image (n_normals: 0 in 2D, else a multiple of 4) - one launch, no driver memset
void launch_begin_frame(FrameCtl* ctl, int32_t first_free, uint64_t* arena, const uint64_t* root, int32_t n_root_cells,
                        int32_t* image0, long long n_image0, uint32_t* normals, long long n_normals, int grid,
                        cudaStream_t s);
void launch_eval_tiles(int dim, bool root, const EvalTilesArgs& a, const void* mat, int grid, cudaStream_t s);
void launch_eval_root(int dim, const EvalRootArgs& a, const void* mat, cudaStream_t s);
// Shared memory a tile of k_eval_sub needs for a plan with nv value ids and n_levels levels.
__host__ __device__ inline int sub_need_bytes(int nv, int n_levels) { return ((nv * 11 + 3) & ~3) + 4 * (n_levels + 1); }
// Warps per CTA of k_eval_sub for a per-tile slice (two CTAs share an SM).
int sub_warps(int slice);
bool use_remap(int n_slots);      // tapes with too many slots for shared-memory rows go through slot renaming
void launch_eval_sub(int dim, const EvalSubArgs& a, const void* mat, int grid, cudaStream_t s);
void launch_rank_tiles(int dim, const RankArgs& a, int grid, cudaStream_t s);
void launch_upsample_filled(int dim, const int32_t* prev, int32_t* image, int size, int grid, cudaStream_t s);
void launch_eval_pixels(const EvalVoxelsArgs& a, const Mat3& mat, int grid, cudaStream_t s);
void launch_eval_voxels(const EvalVoxelsArgs& a, const Mat4& mat, int grid, cudaStream_t s);
void launch_normals(const NormalsArgs& a, const Mat4& mat, int grid, cudaStream_t s);

// Resident CTAs per SM for a given slot count (sizes the persistent grids).
int walk_rows(int n_slots);
int float_rows(int n_slots);     // value rows per warp of the float pass
int float_group(int n_slots, bool heat);
bool float_tmem(int n_slots, bool heat);
int float_ctas(int dim, int n_slots, int group, bool tmem);
int float_warps(int n_slots, int group);
int occupancy_eval_tiles(int dim, bool root, int n_slots);
int occupancy_normals(int n_slots);

}  // namespace mprb
