/* mprb -- C ABI of the B200 tile-recursive implicit-surface renderer.
 *
 * This is the drop-in boundary for the one hot path of mkeeter/mpr:
 * mpr::Context::render2D / render3D and the kernels beneath them.  The
 * reference has no FFI for this path -- its boundary is the C++ struct surface
 * of inc/tape.hpp, inc/context.hpp and inc/util.hpp -- so each entry point
 * below names the reference member it stands behind.  The C++ headers in
 * mpr_b200/inc/ re-create those structs on top of this ABI (see
 * INTEGRATION.md); plain pointers and sizes only, no C++ or torch types.
 *
 * Conventions: every function returning int yields 0 on success and a
 * non-zero MPRB_E_* code otherwise; mprb_last_error() describes the most
 * recent failure on the calling thread.  Matrices are column-major floats,
 * the layout Eigen::Matrix3f / Matrix4f use in the reference.  All buffers
 * handed out by mprb_ctx_buffers() are CUDA managed memory: dereferenceable on
 * the host after a render call returns (calls return only after the device
 * has finished, like the reference's cudaDeviceSynchronize at
 * src/context.cu:1279 / :1457) and on the device that owns the context.
 */
#ifndef MPRB_H
#define MPRB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPRB_OK 0
#define MPRB_E_CUDA 1        /* a CUDA runtime call failed */
#define MPRB_E_ARG 2         /* invalid argument */
#define MPRB_E_OVERFLOW 3    /* a tile list outgrew its array: every level's list is allocated once for the
                                worst case (64 children per surviving parent) but capped at 64 Mi tiles
                                (805 MB); beyond that the frame is incomplete (mprb_frame_stats.overflow
                                names the level) and the context stays usable */
#define MPRB_E_PARSE 4       /* malformed .frep input */

typedef struct mprb_ctx mprb_ctx;
typedef struct mprb_tape mprb_tape;

/* Same layout as mpr::TileNode (reference inc/context.hpp:23-27). */
typedef struct mprb_tile_node {
    int32_t position; /* linear tile index at its level, or -1 once resolved */
    int32_t tape;     /* index of this tile's tape header in tape_data */
    int32_t next;     /* rank among active tiles (children at next*64+i), or -1 */
} mprb_tile_node;

/* Mirrors the data members of mpr::Context (reference inc/context.hpp:60-72). */
typedef struct mprb_buffers {
    int32_t image_size_px;
    int32_t* filled[4];           /* stages[i].filled; [3] is the final image / heightmap */
    mprb_tile_node* tiles[4];     /* stages[i].tiles */
    uint64_t tile_array_size[4];  /* stages[i].tile_array_size: entries valid after the last frame */
    uint64_t* tape_data;          /* subtape arena; the root tape is at cell 0 */
    int32_t* tape_index;          /* cells of the arena in use after the last frame (page-locked host memory) */
    int32_t* num_active_tiles;    /* survivors of the last interval level (page-locked host memory) */
    uint32_t* normals;            /* 0xFFzzyyxx per pixel (3D only) */
} mprb_buffers;

typedef struct mprb_ctx_opts {
    int32_t device;        /* CUDA device ordinal; -1 = current device */
    int64_t num_subtapes;  /* arena size in 64-cell chunks; 0 = 640000
                              (reference inc/parameters.hpp:18-22; 6400000 with BIG_SERVER) */
    /* Multi-GPU sharding: this context renders only the level-0 tile rows
     * [row_begin, row_end) (64-pixel rows in y); row_end = 0 means "all". */
    int32_t row_begin;
    int32_t row_end;
    /* Cyclic variant: of those rows, only rows y with y % row_mod == row_rem (row_mod <= 1: all).
     * Interleaving tile rows across GPUs balances load far better than contiguous bands. */
    int32_t row_mod;
    int32_t row_rem;
    /* Tile-cyclic variant: with col_step = 1 a level-0 tile (x, y) belongs to this context iff
     * (y + x) % row_mod == row_rem - diagonal stripes of 64x64-px screen columns, the finest
     * independent unit (a 3D column keeps all its z tiles).  0 = whole rows, as above. */
    int32_t col_step;
    /* One context over n_gpus devices of this process (device, device + 1, ...; peer access to
     * `device` required): each renders the screen columns (x + y) % n_gpus == its index and writes
     * its blocks of the final image / normals straight into `device`'s buffers over NVLink, so after a
     * render call stages[3].filled and normals hold the whole frame there (the other stages describe
     * device 0's share only).  0 = take the count from the environment variable MPRB_GPUS (default 1),
     * which is how the reference's unchanged drivers, whose mpr::Context(int) knows nothing of
     * devices, are put on several GPUs.  Not to be combined with the row_* options. */
    int32_t n_gpus;
} mprb_ctx_opts;

/* Per-frame counters, filled by the render calls (device-side; no extra syncs). */
typedef struct mprb_frame_stats {
    int32_t n_active[3];      /* ambiguous tiles left after interval level 0, 1, 2 */
    int32_t tape_index;       /* arena cells in use */
    int32_t overflow;         /* bit i: stage i+1 tile array was too small */
    uint64_t i_tiles[3];      /* interval tiles whose tape was walked, per level */
    uint64_t i_cells[3];      /* tape cells visited by those walks */
    uint64_t p_tiles[3];      /* tiles that wrote a shortened tape */
    uint64_t p_cells[3];      /* tape cells visited by the backward walks */
    uint64_t p_kept[3];       /* arena cells written by pushes */
    uint64_t f_tiles;         /* float-stage tiles (64 samples each) */
    uint64_t f_cells;         /* tape cells visited by the float stage */
    uint64_t n_pixels;        /* normal-pass pixels */
    uint64_t n_cells;         /* tape cells visited by the normal pass */
    float gpu_ms;             /* device time of the whole frame (CUDA events) */
    float kernel_ms[12];      /* device time per launch (only when timing is enabled) */
    int32_t n_launches;       /* kernels launched for the frame */
    uint64_t f_items;         /* float-stage work items: runs of up to 2 (4) tiles that share one tape */
    uint64_t p_written;       /* arena cells actually written by pushes: sibling tiles whose verdicts
                                 agree share ONE copy of their shortened tape (p_kept counts it per tile) */
    uint64_t i_sub_tiles;     /* of i_tiles: tiles of a small level evaluated clause-parallel, one warp per tile
                                 on the plan written with the parent's shortened tape (k_eval_sub) */
} mprb_frame_stats;

/* ---- context: mpr::Context::Context(int32_t) (src/context.cpp:16-49) ------------ */
int mprb_ctx_create(int32_t image_size_px, const mprb_ctx_opts* opts, mprb_ctx** out);
void mprb_ctx_destroy(mprb_ctx* ctx);
int mprb_ctx_buffers(mprb_ctx* ctx, mprb_buffers* out);
/* Enables per-kernel CUDA-event timing (fills kernel_ms); off by default. */
int mprb_ctx_set_timing(mprb_ctx* ctx, int enabled);

/* ---- tape: mpr::Tape (inc/tape.hpp:24-30, upload at src/tape.cpp:223-227) -------- */
int mprb_tape_create(const uint64_t* host_cells, int32_t n_cells, mprb_tape** out);
void mprb_tape_destroy(mprb_tape* tape);
const uint64_t* mprb_tape_data(const mprb_tape* tape);  /* managed memory (Tape::data) */
int32_t mprb_tape_length(const mprb_tape* tape);        /* Tape::length */
int32_t mprb_tape_num_slots(const mprb_tape* tape);

/* ---- frames: mpr::Context::render2D / render3D (src/context.cu:1136, :1282) ------ */
int mprb_render2d(mprb_ctx* ctx, const mprb_tape* tape, const float mat3_colmajor[9], float z);
int mprb_render3d(mprb_ctx* ctx, const mprb_tape* tape, const float mat4_colmajor[16]);

/* Host-buffer variants: the tape cells come from (pinned or pageable) host
 * memory and the results are copied into host buffers before returning.
 * image_out / depth_out: size*size int32; normals_out: size*size uint32.  Any output pointer
 * may be NULL to skip that download (multi-GPU callers gather bands on the device first). */
int mprb_render2d_host(mprb_ctx* ctx, const uint64_t* host_cells, int32_t n_cells,
                       const float mat3_colmajor[9], float z, int32_t* image_out);
int mprb_render3d_host(mprb_ctx* ctx, const uint64_t* host_cells, int32_t n_cells,
                       const float mat4_colmajor[16], int32_t* depth_out, uint32_t* normals_out);

int mprb_frame_stats_get(mprb_ctx* ctx, mprb_frame_stats* out);

/* ---- analysis variants ---------------------------------------------------------------- */
/* Context::render2D_brute (src/context.cu:1461-1508): no subdivision; every 8x8 tile of the
 * frame is evaluated point by point with the full tape. */
int mprb_render2d_brute(mprb_ctx* ctx, const mprb_tape* tape, const float mat3_colmajor[9], float z);
/* Context::render2D_heatmap / render3D_heatmap (src/context.cu:1984-2340): a normal frame that
 * also accumulates the amortised work per pixel (tape cells walked by every tile covering the
 * pixel, divided by the tile's pixel count), normalised by the clause count of the tape.
 * *heatmap_out: size*size floats of managed memory owned by the caller (mprb_free_device). */
int mprb_render2d_heatmap(mprb_ctx* ctx, const mprb_tape* tape, const float mat3_colmajor[9], float z,
                          float** heatmap_out);
int mprb_render3d_heatmap(mprb_ctx* ctx, const mprb_tape* tape, const float mat4_colmajor[16],
                          float** heatmap_out);

/* ---- multi-GPU exchange helpers (no reference counterpart; the reference is single-GPU) ---- */
/* For contexts created with row_mod = world > 1 over the whole frame (row_begin = 0, row_end = all,
 * tiles per side divisible by world).  After a frame, mprb_exchange_pack writes the 64x64-px blocks
 * this context owns - depth narrowed to uint8 (2D) / int16 (3D), then the normals in 3D - into
 * `dst_device` (mprb_exchange_bytes bytes); all-gather those buffers in rank order (NCCL, MPI, ...)
 * and hand the result to mprb_exchange_unpack, which scatters every rank's blocks into this
 * context's full-size image and normals.  `stream` is a cudaStream_t (NULL = the context's own);
 * the calls are asynchronous on it.  dim = 2 or 3: the kind of frame last rendered. */
size_t mprb_exchange_bytes(const mprb_ctx* ctx, int dim);
int mprb_exchange_pack(mprb_ctx* ctx, int dim, void* dst_device, void* stream);
int mprb_exchange_unpack(mprb_ctx* ctx, int dim, const void* src_device, void* stream);
/* Stores the 64x64-px blocks this context owns (all of them without sharding) of the frame last
 * rendered straight into full-size images somewhere else: S*S int32 depth / filled values and, for
 * dim = 3, S*S packed normals (null = skip).  The destinations may be page-locked host memory the
 * device can address (cudaHostAlloc, or cudaHostRegister'ed pages - e.g. a shared-memory segment that
 * the other ranks' processes registered too, so that N GPUs fill ONE host frame over N PCIe links at
 * once and nobody gathers it on a device first), or memory of a peer device with access enabled.  One
 * launch of stores on the context's stream; returns when they have landed. */
int mprb_ctx_publish(mprb_ctx* ctx, int dim, int32_t* dst_image, uint32_t* dst_normals);

/* ---- post-effects: mpr::Effects (inc/effects.hpp:21-37, src/effects.cu:229-297) ---- */
typedef struct mprb_effects mprb_effects;
/* ssao_kernel: 64x3, ssao_rvecs: 256x3, both column-major floats (the Eigen members the
 * reference's Effects constructor fills with rand()). */
int mprb_effects_create(const float* ssao_kernel_64x3, const float* ssao_rvecs_256x3, mprb_effects** out);
void mprb_effects_destroy(mprb_effects* fx);
/* Effects::drawSSAO / drawShaded on the context's last 3D frame; the result is `image`. */
int mprb_effects_draw_ssao(mprb_effects* fx, mprb_ctx* ctx);
int mprb_effects_draw_shaded(mprb_effects* fx, mprb_ctx* ctx);
/* Effects::image / Effects::tmp (managed memory, size*size int32 each). */
int mprb_effects_buffers(mprb_effects* fx, int32_t** image, int32_t** tmp);

/* ---- host helpers ----------------------------------------------------------------- */
/* .frep bytes -> packed tape (libfive archive reader + the packer that restates
 * src/tape.cpp).  *cells_out is malloc'd; release with mprb_free(). */
int mprb_tape_from_frep(const uint8_t* bytes, size_t n_bytes, int simplify,
                        uint64_t** cells_out, int32_t* n_cells_out, int32_t* n_slots_out);
void mprb_free(void* p);
/* Releases device / managed memory (cudaFree); counterpart of CUDA_FREE in inc/util.hpp. */
void mprb_free_device(void* p);
/* Managed allocation; counterpart of CUDA_MALLOC (cudaMallocManaged) in inc/util.hpp:26-33. */
int mprb_malloc_managed(size_t n_bytes, void** out);

const char* mprb_last_error(void);
const char* mprb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MPRB_H */
