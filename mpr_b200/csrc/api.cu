// C-ABI entry points of libmprb: context and tape management, and the
// per-frame launch sequence (see include/mprb.h for the contract).
//
// Reference counterparts: Context::Context (src/context.cpp:16-49),
// Context::render2D (src/context.cu:1136-1280), Context::render3D
// (src/context.cu:1282-1458), Tape upload (src/tape.cpp:223-227).
//
// Unlike the reference, a frame never returns to the host between levels: tile
// counts live in a device control block (FrameCtl), every tape-walking kernel
// is a persistent grid pulling work items from a device queue, and tile arrays
// are sized for the worst case up front.  The only host<->device traffic of a
// frame is the kernel launches, one 304-byte control-block read-back at the
// end, and (host-buffer variants only) the tape upload and the image download.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/mprb.h"
#include "common.cuh"
#include "host/mprb_host.hpp"
#include "exchange.cuh"
#include "postfx.cuh"
#include "kernels.cuh"
#include "libfive/tree/archive.hpp"

using namespace mprb;

static_assert(sizeof(mprb_tile_node) == sizeof(TileNode), "TileNode layout");

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define MPRB_CUDA(expr)                                                              \
    do {                                                                             \
        cudaError_t e_ = (expr);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            return fail(MPRB_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                         \
        }                                                                            \
    } while (0)

constexpr long long kMaxStageTiles = 64ll << 20;   // cap on any one tile array
constexpr int kMaxLaunches = 12;

template <typename T>
cudaError_t managed_alloc(T** out, size_t count, int device) {
    void* p = nullptr;
    cudaError_t e = cudaMallocManaged(&p, sizeof(T) * std::max<size_t>(count, 1));
    if (e != cudaSuccess) return e;
    // Keep pages on the GPU; host reads after a frame migrate on demand.
    cudaMemAdvise(p, sizeof(T) * std::max<size_t>(count, 1), cudaMemAdviseSetPreferredLocation, device);
    *out = static_cast<T*>(p);
    return cudaSuccess;
}

}  // namespace

// Device-side copies of a tape: made once per device that renders it (a multi-GPU context walks the
// same Tape on every device; a plain cudaMalloc on "whatever device is current" would leave the others
// dereferencing memory they cannot reach).
struct TapeDev {
    uint64_t* cells = nullptr;       // the contiguous cells (k_eval_root's sweep reads them)
    uint64_t* chunked = nullptr;     // the same tape in chunk-terminated layout (tape_stream.cuh)
    RootClause* sched = nullptr;     // clause-parallel plan for the root level (see k_eval_root)
    int32_t* level_start = nullptr;
    uint16_t* prevw = nullptr;       // per clause: the value id its output slot held before it
    int32_t group = 0;               // threads per tile; 0 = no plan (serial root walk)
    int32_t smem_per_tile = 0;
};

struct mprb_tape {
    uint64_t* cells = nullptr;   // managed; the cells as given (Tape::data)
    int32_t length = 0;
    int32_t n_chunked = 0;
    std::vector<uint64_t> host_cells;
    std::vector<uint64_t> host_chunked;
    int32_t n_slots = 0;
    std::vector<RootClause> sched;       // host copy of the root plan
    std::vector<int32_t> level_start;
    std::vector<uint16_t> prevw;
    int32_t n_levels = 0;
    int32_t result_v = 0;
    mutable std::mutex mu;
    mutable std::map<int, TapeDev> dev;  // by CUDA device ordinal
};

struct mprb_ctx {
    int device = 0;
    int size = 0;
    int sm_count = 0;
    int row_begin = 0, row_end = 0, row_mod = 1, row_rem = 0, col_step = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    cudaEvent_t ev_k[kMaxLaunches + 1] = {};
    bool timing = false;

    // API-visible (managed) buffers
    int32_t* filled[4] = {};
    TileNode* tiles[4] = {};
    long long tiles_cap[4] = {};
    uint64_t tile_array_size[4] = {};
    uint64_t* arena = nullptr;
    long long arena_cells = 0;
    int32_t* tape_index = nullptr;
    int32_t* num_active_tiles = nullptr;
    uint32_t* normals = nullptr;

    // internal
    int32_t* active_list[3] = {};
    int32_t* float_items = nullptr;  // device: work items of the float pass (runs of tiles sharing a tape)
    FrameCtl* ctl = nullptr;
    FrameCtl* ctl_host = nullptr;    // pinned
    uint64_t* stage_cells = nullptr; // pinned staging for host tapes
    int32_t stage_cells_cap = 0;
    bool serial_root = false;        // debugging / A-B switch: MPRB_SERIAL_ROOT=1
    // Clause-parallel plans of the tapes k_eval_root shortens, for small levels (k_eval_sub)
    int32_t* plans = nullptr;        // device: plan arena (32-bit words)
    long long plans_words = 0;
    int32_t* plan_of = nullptr;      // device: per level-0 tile, word offset of its plan or -1
    long long plan_of_cap = 0;
    int sub_waves = -1;              // k_eval_sub serves levels of up to this many waves of tiles; -1 = from the tape's
                                     // shape (see render), MPRB_SUB_WAVES overrides (0 = off)
    // What the last frame of (hint_tape, hint_dim) looked like: plans cost k_eval_root time and k_eval_sub a
    // launch, so both are skipped while the level below the root is large or its tapes are too short to
    // plan (looked at again every 32nd frame; the frame itself is the same either way).
    const void* hint_tape = nullptr;
    int hint_dim = 0;
    int hint_parents = 0, hint_plans = 0;
    unsigned frame_no = 0;
    int32_t* owned_tiles = nullptr;  // device: level-0 screen tiles (y * tiles_per_side + x) this context renders
    int n_owned = 0;
    unsigned long long* heat_units = nullptr;   // work meter of render*_heatmap (device, S*S), lazily allocated
    // Host-buffer entry points: the per-tape root plan is cached across frames as long as
    // the caller keeps passing the same cells (the cells themselves are re-uploaded every frame).
    mprb_tape* host_plan = nullptr;
    std::vector<uint64_t> host_plan_cells;

    mprb_frame_stats stats = {};
    std::map<long long, int> occ_cache;

    // multi-GPU: sub-contexts on the other devices (owned by this, the primary, context)
    std::vector<mprb_ctx*> peers;
    cudaEvent_t ev_body = nullptr;   // the frame's kernels are enqueued up to here (no timing)
    // exchange helpers may run on a caller's stream: the next frame waits for them
    cudaEvent_t ev_ext = nullptr;
    bool ext_pending = false;
};

namespace {

int tape_num_slots(const uint64_t* cells, int32_t n) {
    int mx = 0;
    for (int32_t i = 0; i < n; ++i) {
        const uint32_t w = uint32_t(cells[i]);
        const uint32_t op = w & 0xff;
        if (op == OP_JUMP) continue;
        mx = std::max<int>(mx, (w >> 8) & 0xff);
        mx = std::max<int>(mx, (w >> 16) & 0xff);
        mx = std::max<int>(mx, w >> 24);
    }
    return mx + 1;
}

int validate_tape(const uint64_t* cells, int32_t n) {
    if (!cells || n < 2) return fail(MPRB_E_ARG, "tape needs at least a header and an end cell");
    if ((cells[0] & 0xff) != 0 || (cells[n - 1] & 0xff) != 0)
        return fail(MPRB_E_ARG, "tape must start with a header cell and finish with an end cell");
    for (int32_t i = 1; i + 1 < n; ++i) {
        const uint32_t op = uint32_t(cells[i]) & 0xff;
        if (op < OP_SQUARE || op > OP_COPY_RHS)
            return fail(MPRB_E_ARG, "tape cell %d has opcode %u, which is not a clause", i, op);
    }
    if (tape_num_slots(cells, n) > 128)
        return fail(MPRB_E_ARG, "tape uses more than 128 slots (the reference kernels hold 128)");
    return MPRB_OK;
}

// Contiguous tape -> chunk-terminated layout (see tape_stream.cuh).
std::vector<uint64_t> chunk_layout(const uint64_t* cells, int32_t n) {
    if (n <= kChunk) return std::vector<uint64_t>(cells, cells + n);
    const int n_chunks = 1 + (n - 63 + 61) / 62;
    std::vector<uint64_t> out(size_t(n_chunks) * kChunk, 0);
    int last = 0;
    for (int q = 0; q < n; ++q) {
        int idx = q;
        if (q >= kChunk - 1) {
            const int r = q - (kChunk - 1);
            idx = kChunk + (r / (kChunk - 2)) * kChunk + 1 + r % (kChunk - 2);
        }
        out[idx] = cells[q];
        last = idx;
    }
    const uint64_t fwd = uint64_t(OP_JUMP) | (uint64_t(uint32_t(1)) << 32);
    const uint64_t back = uint64_t(OP_JUMP) | (uint64_t(uint32_t(-1)) << 32);
    for (int c = 0; c < n_chunks; ++c) {
        if (c > 0) out[size_t(c) * kChunk] = back;
        if (c + 1 < n_chunks) out[size_t(c) * kChunk + kChunk - 1] = fwd;
    }
    out.resize(size_t(last) + 1);
    return out;
}

// Root tape -> SSA + dependency levels.  Value ids: 0 = none, 1..3 = x, y, z, 3+i = clause i.
// Within a level clauses are ordered by opcode so that neighbouring lanes mostly run the
// same interval operator.
struct RootPlan {
    std::vector<RootClause> sched;
    std::vector<int32_t> level_start;
    std::vector<uint16_t> prevw;      // [i] = value id that sat in clause i's output slot before it wrote there
    int result_v = 0;
};

RootPlan build_root_plan(const uint64_t* cells, int32_t n_cells) {
    const int n = n_cells - 2;
    RootPlan p;
    std::vector<int> writer(256, 0);          // slot -> value id that currently lives there
    const uint32_t hdr = uint32_t(cells[0]);
    // same binding order as the kernels: x, then y, then z (a later axis wins a shared slot)
    const int ax[3] = {int((hdr >> 8) & 0xff), int((hdr >> 16) & 0xff), int(hdr >> 24)};
    for (int k = 0; k < 3; ++k) if (ax[k]) writer[ax[k]] = 1 + k;
    std::vector<int> depth(n + 4, 0);
    std::vector<RootClause> byidx(n + 1);
    p.prevw.assign(n + 1, 0);
    int n_choice = 0, max_depth = 0;
    for (int i = 1; i <= n; ++i) {
        const uint64_t d = cells[i];
        const uint32_t w = uint32_t(d);
        const uint32_t op = w & 0xff, out = (w >> 8) & 0xff, lhs = (w >> 16) & 0xff, rhs = w >> 24;
        RootClause rc;
        rc.lsrc = lhs ? uint32_t(writer[lhs]) : 0u;
        rc.rsrc = rhs ? uint32_t(writer[rhs]) : 0u;
        uint32_t hi = uint32_t(d >> 32);
        memcpy(&rc.imm, &hi, 4);
        const bool is_choice = op >= OP_MIN_LI && op <= OP_MAX_LR;
        const bool past_cap = is_choice && n_choice >= kMaxChoices;
        n_choice += is_choice;
        // bits 9 / 10: the left / right operand's slot is the output slot (a verdict for it drops the clause)
        rc.op_idx = op | (past_cap ? 0x100u : 0u) | (lhs == out ? 0x200u : 0u) | (rhs != 0 && rhs == out ? 0x400u : 0u) |
                    (uint32_t(i) << 12);
        byidx[i] = rc;
        const int dep = 1 + std::max(depth[rc.lsrc], depth[rc.rsrc]);
        depth[3 + i] = dep;
        max_depth = std::max(max_depth, dep);
        p.prevw[i] = uint16_t(std::min(writer[out], 0xffff));
        writer[out] = 3 + i;
    }
    p.result_v = writer[(uint32_t(cells[n + 1]) >> 8) & 0xff];
    std::vector<std::vector<int>> levels(max_depth + 1);
    for (int i = 1; i <= n; ++i) levels[depth[3 + i]].push_back(i);
    p.level_start.push_back(0);
    for (int L = 1; L <= max_depth; ++L) {
        auto& v = levels[L];
        std::stable_sort(v.begin(), v.end(), [&](int x, int y) {
            return (byidx[x].op_idx & 0xff) < (byidx[y].op_idx & 0xff);
        });
        for (int i : v) p.sched.push_back(byidx[i]);
        p.level_start.push_back(int32_t(p.sched.size()));
    }
    return p;
}

// The device-side copies of `t` on `device` (which must be current); null on a CUDA error.
const TapeDev* tape_on(const mprb_tape* t, int device) {
    std::lock_guard<std::mutex> lock(t->mu);
    auto it = t->dev.find(device);
    if (it != t->dev.end()) return &it->second;
    TapeDev d;
    auto upload = [](auto** dst, const auto& v) {
        typedef typename std::remove_reference<decltype(v[0])>::type T;
        cudaError_t e = cudaMalloc(dst, sizeof(T) * std::max<size_t>(v.size(), 1));
        if (e == cudaSuccess && !v.empty()) e = cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
        return e;
    };
    cudaError_t e = upload(&d.cells, t->host_cells);
    if (e == cudaSuccess) e = upload(&d.chunked, t->host_chunked);
    const int n = t->length - 2;
    if (e == cudaSuccess && !t->sched.empty()) {
        // Threads per tile ~ clauses per level, so that a level is one pass; bounded by what one
        // CTA's shared memory holds on THIS device.
        int group = 32;
        while (group < kRootThreads && group < n / std::max(t->n_levels, 1)) group *= 2;
        const int per_tile = ((n + 4) * 10 + 64 + 15) / 16 * 16;
        int max_smem = 0;
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        while (group < kRootThreads && (kRootThreads / group) * per_tile > max_smem) group *= 2;
        if (per_tile <= max_smem) {
            e = upload(&d.sched, t->sched);
            if (e == cudaSuccess) e = upload(&d.level_start, t->level_start);
            if (e == cudaSuccess && n + 4 < 0xffff) e = upload(&d.prevw, t->prevw);
            d.group = group;
            d.smem_per_tile = per_tile;
        }
    }
    if (e != cudaSuccess) {
        if (d.cells) cudaFree(d.cells);
        if (d.chunked) cudaFree(d.chunked);
        if (d.sched) cudaFree(d.sched);
        if (d.level_start) cudaFree(d.level_start);
        if (d.prevw) cudaFree(d.prevw);
        fail(MPRB_E_CUDA, "tape upload to device %d: %s", device, cudaGetErrorString(e));
        return nullptr;
    }
    return &(t->dev[device] = d);
}

int ensure_stage(mprb_ctx* c, int stage, long long cap) {
    if (c->tiles_cap[stage] >= cap) return MPRB_OK;
    if (c->tiles[stage]) cudaFree(c->tiles[stage]);
    c->tiles[stage] = nullptr;
    MPRB_CUDA(managed_alloc(&c->tiles[stage], size_t(cap), c->device));
    c->tiles_cap[stage] = cap;
    if (stage < 3) {
        if (c->active_list[stage]) cudaFree(c->active_list[stage]);
        c->active_list[stage] = nullptr;
        MPRB_CUDA(cudaMalloc(&c->active_list[stage], sizeof(int32_t) * size_t(cap)));
    } else {
        if (c->float_items) cudaFree(c->float_items);
        c->float_items = nullptr;
        MPRB_CUDA(cudaMalloc(&c->float_items, sizeof(int32_t) * size_t(cap)));
    }
    return MPRB_OK;
}

int cached_occupancy(mprb_ctx* c, int kind, int dim, bool root, int n_slots, int group = 1) {
    const long long key = (long long)kind << 40 | (long long)dim << 32 | (long long)root << 24 | (long long)group << 16 | n_slots;
    auto itr = c->occ_cache.find(key);
    if (itr != c->occ_cache.end()) return itr->second;
    int n = 0;
    if (kind == 0) n = occupancy_eval_tiles(dim, root, n_slots);
    else if (kind == 1) n = float_ctas(dim, n_slots, group, float_tmem(n_slots, false) && group >= 2);
    else n = occupancy_normals(n_slots);
    n = std::max(n, 1);
    c->occ_cache[key] = n;
    return n;
}

struct Timer {
    mprb_ctx* c;
    int n = 0;
    void mark() {
        if (c->timing && n <= kMaxLaunches) cudaEventRecord(c->ev_k[n], c->stream);
        ++n;
    }
};

// The frame proper.  `cells` may be a device/managed pointer (async D2D copy)
// or a host pointer (async H2D copy through pinned staging).
int render(mprb_ctx* c, int dim, const mprb_tape* plan, bool cells_on_host, const float* matrix, float z,
           bool heat = false, bool brute = false)
{
    if (!c || !plan) return fail(MPRB_E_ARG, "null context or tape");
    const int S = c->size;
    cudaStream_t s = c->stream;
    MPRB_CUDA(cudaSetDevice(c->device));
    const TapeDev* td = tape_on(plan, c->device);
    if (!td) return MPRB_E_CUDA;
    const int32_t n_cells = plan->length;
    const int32_t n_chunked = plan->n_chunked;
    const int n_slots = plan->n_slots;
    if (n_chunked + kChunk >= c->arena_cells) return fail(MPRB_E_ARG, "tape longer than the arena");

    const int tps0 = S / 64;
    const long long count0 = dim == 3 ? (long long)tps0 * tps0 * tps0 : (long long)tps0 * tps0;
    // Stage layout.  3D: 64^3 -> 16^3 -> 4^3 -> voxels (stages 0,1,2,3).
    // 2D: 64^2 -> 8^2 -> pixels (stages 0,2,3), as in the reference.
    const int n_levels = dim == 3 ? 3 : 2;
    int stage_of[3] = {0, dim == 3 ? 1 : 2, 2};
    int px_of[3] = {64, dim == 3 ? 16 : 8, 4};
    {
        long long cap = count0;
        for (int l = 1; l < n_levels; ++l) {
            cap = std::min(cap * 64, kMaxStageTiles);
            if (int e = ensure_stage(c, stage_of[l], cap)) return e;
        }
        if (int e = ensure_stage(c, 3, cap)) return e;   // compact survivor list
    }

    Mat4 m4;
    Mat3 m3;
    if (dim == 3) memcpy(m4.d, matrix, sizeof(m4.d));
    else memcpy(m3.d, matrix, sizeof(m3.d));
    const void* mat = dim == 3 ? static_cast<const void*>(&m4) : static_cast<const void*>(&m3);

    if (c->ext_pending) {            // an exchange helper touched the frame buffers on another stream
        MPRB_CUDA(cudaStreamWaitEvent(s, c->ev_ext, 0));
        c->ext_pending = false;
    }
    cudaEventRecord(c->ev_begin, s);
    Timer tm{c};
    tm.mark();

    // Root tape to cell 0 of the arena (context.cu:1139-1142), in chunk-terminated layout;
    // pushed tapes start at the next 64-cell boundary so every chunk is 512-byte aligned.
    const int32_t first_free = (n_chunked + kChunk - 1) / kChunk * kChunk;
    // One launch clears the control block, copies the root tape and clears the level-0 and normal images: the
    // arena and the images are managed memory, and driver copies / memsets on managed ranges stall the host
    // (see k_begin_frame).
    const uint64_t* root_src = td->chunked;
    if (cells_on_host) {
        // Host-buffer entry points pay the upload every frame: host cells -> page-locked staging, which the
        // kernel reads over PCIe.
        if (c->stage_cells_cap < n_chunked) {
            if (c->stage_cells) cudaFreeHost(c->stage_cells);
            c->stage_cells = nullptr;
            MPRB_CUDA(cudaMallocHost(&c->stage_cells, sizeof(uint64_t) * size_t(n_chunked)));
            c->stage_cells_cap = n_chunked;
        }
        memcpy(c->stage_cells, plan->host_chunked.data(), sizeof(uint64_t) * size_t(n_chunked));
        root_src = c->stage_cells;
    }
    {
        const long long n_normals = dim == 3 ? (long long)S * S : 0;
        const long long work16 = n_chunked + n_normals / 4 + (long long)tps0 * tps0 / 4;
        const int grid = int(std::max<long long>(1, std::min<long long>(c->sm_count * 8, (work16 + 255) / 256)));
        launch_begin_frame(c->ctl, first_free, c->arena, root_src, n_chunked, c->filled[0], (long long)tps0 * tps0,
                           c->normals, n_normals, grid, s);
    }
    unsigned long long* heat_units = nullptr;
    if (heat) {
        if (!c->heat_units) MPRB_CUDA(cudaMalloc(&c->heat_units, sizeof(unsigned long long) * size_t(S) * S));
        heat_units = c->heat_units;
        MPRB_CUDA(cudaMemsetAsync(heat_units, 0, sizeof(unsigned long long) * size_t(S) * S, s));
    }

    // Small levels below the root run clause-parallel (k_eval_sub) on plans k_eval_root writes with the
    // tapes it shortens.  A tile of that kernel gets a fixed slice of shared memory, sized for the root
    // tape (no shortened tape needs more) but at most 24 KB; the level is "small" while its tiles make at
    // most sub_waves waves of that kernel.
    const bool clause_parallel_root = td->group > 0 && !c->serial_root && !heat && !brute;
    int sub_slice = 0, sub_max_parents = 0, sub_grid = 0;
    const bool hinted = c->hint_tape == static_cast<const void*>(plan) && c->hint_dim == dim && c->sub_waves < 100000 &&
                        (c->frame_no++ & 31u) != 31u;
    // How many waves of k_eval_sub still beat one lane per tile walking the tape: a tile there costs about its
    // levels (~1000 cycles each: fetch, divergent interval operators, two sweeps) plus ~30 cycles per clause,
    // against two walks of the tape at ~400 (renamed slots: ~800) cycles a clause here; half the root tape's
    // length stands in for the shortened tapes'.  bear (72 levels x 7.6 clauses): 1 wave; prospero (22 x 275): 8.
    int sub_waves = c->sub_waves;
    if (sub_waves < 0) {
        const double half = 0.5 * (n_cells - 2);
        const double serial = half * (use_remap(n_slots) ? 800.0 : 400.0);
        const double parallel = 1000.0 * plan->n_levels + 30.0 * half;
        sub_waves = std::max(1, std::min(8, int(serial / parallel)));
    }
    if (clause_parallel_root && sub_waves > 0 && n_levels >= 2 && td->prevw && plan->n_levels < 256) {
        const int need = sub_need_bytes(n_cells - 2 + 4, plan->n_levels);
        sub_slice = std::min(std::max((need + 15) / 16 * 16, 2048), 24576);
        const int warps = sub_warps(sub_slice);
        const int ctas = (2 * (warps * sub_slice + 1024) <= 227 * 1024) ? 2 : 1;
        sub_grid = c->sm_count * ctas;
        sub_max_parents = std::max(1, sub_waves * sub_grid * warps / 64);
        if (hinted && (c->hint_parents > sub_max_parents + sub_max_parents / 8 || c->hint_plans == 0)) sub_slice = 0;
    }
    if (sub_slice) {
        const long long per_plan = kPlanHeader + plan->n_levels + 4 + 5LL * (sub_slice / 10 + 1);
        // never more plans than root tiles; the arena running out only sends tiles to the serial kernel
        const long long want = std::min<long long>(per_plan * (std::min<long long>(sub_max_parents, count0) + 8), 64LL << 20);
        if (c->plans_words < want) {
            if (c->plans) cudaFree(c->plans);
            c->plans = nullptr;
            c->plans_words = 0;
            MPRB_CUDA(cudaMalloc(&c->plans, sizeof(int32_t) * size_t(want)));
            c->plans_words = want;
        }
        if (c->plan_of_cap < count0) {
            if (c->plan_of) cudaFree(c->plan_of);
            c->plan_of = nullptr;
            c->plan_of_cap = 0;
            MPRB_CUDA(cudaMalloc(&c->plan_of, sizeof(int32_t) * size_t(count0)));
            c->plan_of_cap = count0;
        }
    }

    int q = 0;
    const int small_grid = c->sm_count * 4;
    const int group = float_group(n_slots, heat);      // tiles per work item of the float pass
    if (brute) {
        // Context::render2D_brute (context.cu:1461-1508): no interval levels; every 8x8 tile goes
        // to the float pass with the root tape.
        const long long count = (long long)(S / 8) * (S / 8);
        if (int e = ensure_stage(c, 3, count)) return e;
        MPRB_CUDA(cudaMemsetAsync(c->filled[3], 0, sizeof(int32_t) * size_t(S) * S, s));
        launch_preload_tiles(c->tiles[3], c->float_items, int32_t(count), &c->ctl->n_active[n_levels - 1],
                             &c->ctl->n_active[3], int(std::min<long long>(small_grid, (count + 255) / 256)), s);
        tm.mark();
    }
    for (int l = 0; l < n_levels && !brute; ++l) {
        const int st = stage_of[l];
        const bool root = (l == 0);
        const bool last = (l == n_levels - 1);
        const int tps = S / px_of[l];

        EvalTilesArgs ea = {};
        ea.arena = c->arena;
        ea.tape_index = &c->ctl->tape_cursor;
        ea.arena_cap = int32_t(c->arena_cells);
        ea.image = c->filled[st];
        ea.tiles = c->tiles[st];
        ea.tiles_cap = int32_t(std::min<long long>(c->tiles_cap[st], INT32_MAX));
        ea.tps = uint32_t(tps);
        if (!root) {
            ea.ptiles = c->tiles[stage_of[l - 1]];
            ea.pactive = c->active_list[stage_of[l - 1]];
            ea.n_parents = &c->ctl->n_active[l - 1];
            ea.ptps = uint32_t(S / px_of[l - 1]);
        }
        ea.count0 = int32_t(count0);
        ea.row_begin = c->row_begin;
        ea.row_end = c->row_end;
        ea.row_mod = c->row_mod;
        ea.row_rem = c->row_rem;
        ea.col_step = c->col_step;
        ea.ctl = c->ctl;
        ea.queue = &c->ctl->queue[q++];
        ea.level = l;
        ea.n_slots = n_slots;
        ea.n_rows = walk_rows(n_slots);
        ea.z = z;
        ea.heat = heat_units;
        ea.heat_px = px_of[l];
        ea.n_root = n_cells - 2;
        if (root && clause_parallel_root) {
            EvalRootArgs ra = {};
            ra.arena = c->arena;
            ra.tape_index = &c->ctl->tape_cursor;
            ra.arena_cap = int32_t(c->arena_cells);
            ra.image = c->filled[st];
            ra.tiles = c->tiles[st];
            ra.tps = uint32_t(tps);
            ra.count0 = int32_t(count0);
            ra.row_begin = c->row_begin;
            ra.row_end = c->row_end;
            ra.row_mod = c->row_mod;
            ra.row_rem = c->row_rem;
            ra.col_step = c->col_step;
            ra.ctl = c->ctl;
            ra.cells = td->cells;
            ra.sched = td->sched;
            ra.level_start = td->level_start;
            ra.n_levels = plan->n_levels;
            ra.n_clauses = n_cells - 2;
            ra.result_v = plan->result_v;
            // with plans to write, more threads per tile: the plan's clauses are fetched from L2 one by one
            // per thread, and a frame that wants plans has few root tiles anyway (bear 256^3 level 0 with
            // plans: 0.16 ms at 32 threads per tile, 0.11 at 128; 4096 root tiles are slower at 64 than at 32)
            ra.group = sub_slice ? std::max(td->group, std::min(kRootThreads, 128)) : td->group;
            ra.smem_per_tile = td->smem_per_tile;
            ra.z = z;
            if (sub_slice) {
                ra.plans = c->plans;
                ra.plan_cursor = &c->ctl->plan_cursor;
                ra.plan_cap = int32_t(std::min<long long>(c->plans_words, INT32_MAX));
                ra.plan_of = c->plan_of;
                ra.sub_slice = sub_slice;
                ra.plan_count = &c->ctl->plan_count;
                ra.max_plans = sub_max_parents;
                ra.plan_min = c->sub_waves >= 100000 ? 0 : kPlanMinClauses;     // the test setting plans everything
                ra.prevw = td->prevw;
            }
            launch_eval_root(dim, ra, mat, s);
        } else {
            if (l == 1 && sub_slice) {
                // tiles whose parent carries a plan, when the level is small; the serial kernel skips those
                EvalSubArgs sa = {};
                sa.arena = c->arena;
                sa.tape_index = &c->ctl->tape_cursor;
                sa.arena_cap = int32_t(c->arena_cells);
                sa.image = c->filled[st];
                sa.tiles = c->tiles[st];
                sa.tiles_cap = ea.tiles_cap;
                sa.tps = uint32_t(tps);
                sa.ptiles = ea.ptiles;
                sa.pactive = ea.pactive;
                sa.n_parents = ea.n_parents;
                sa.ptps = ea.ptps;
                sa.ctl = c->ctl;
                sa.queue = &c->ctl->queue[q++];
                sa.level = l;
                sa.plans = c->plans;
                sa.plan_of = c->plan_of;
                sa.slice = sub_slice;
                sa.max_parents = sub_max_parents;
                sa.z = z;
                launch_eval_sub(dim, sa, mat, sub_grid, s);
                ea.plan_of = c->plan_of;
                ea.sub_max_parents = sub_max_parents;
            }
            int grid = c->sm_count * cached_occupancy(c, 0, dim, root, n_slots);
            if (root) {
                const long long items = (count0 + 31) / 32;
                grid = int(std::min<long long>(grid, (items + kEvalWarps - 1) / kEvalWarps));
            }
            launch_eval_tiles(dim, root, ea, mat, std::max(grid, 1), s);
        }
        tm.mark();

        RankArgs ra = {};
        ra.tiles = c->tiles[st];
        ra.tiles_cap = ea.tiles_cap;
        ra.n_parents = root ? nullptr : &c->ctl->n_active[l - 1];
        ra.count0 = int32_t(count0);
        ra.tps = tps;
        ra.image = c->filled[st];
        ra.n_active = &c->ctl->n_active[l];
        ra.active_list = c->active_list[st];
        ra.out_tiles = c->tiles[3];
        ra.items = c->float_items;
        ra.n_items = &c->ctl->n_active[3];
        ra.gmax = group;
        ra.next_cap = last ? c->tiles_cap[3] : c->tiles_cap[stage_of[l + 1]];
        ra.last_level = last ? 1 : 0;
        ra.level = l;
        ra.ctl = c->ctl;
        int rgrid = small_grid;
        if (root) rgrid = int(std::min<long long>(small_grid, (count0 + 255) / 256));
        launch_rank_tiles(dim, ra, std::max(rgrid, 1), s);
        tm.mark();

        const int next_stage = last ? 3 : stage_of[l + 1];
        const int next_size = last ? S : S / px_of[l + 1];
        const long long px = (long long)next_size * next_size;
        launch_upsample_filled(dim, c->filled[st], c->filled[next_stage], next_size,
                               int(std::min<long long>(c->sm_count * 8, (px + 255) / 256)), s);
        tm.mark();
    }

    {
        EvalVoxelsArgs va = {};
        va.arena = c->arena;
        va.arena_cap = int32_t(c->arena_cells);
        va.image = c->filled[3];
        va.tiles = c->tiles[3];
        va.tiles_cap = int32_t(std::min<long long>(c->tiles_cap[3], INT32_MAX));
        va.items = c->float_items;
        va.n_items = &c->ctl->n_active[3];
        va.group = brute ? 1 : group;
        va.tmem = (!brute && float_tmem(n_slots, heat)) ? 1 : 0;
        va.tps = uint32_t(S / px_of[n_levels - 1]);
        va.ctl = c->ctl;
        va.queue = &c->ctl->queue[q++];
        va.n_slots = n_slots;
        va.n_rows = float_rows(n_slots);
        va.z = z;
        va.heat = heat_units;
        va.n_root = n_cells - 2;
        const int grid = c->sm_count * cached_occupancy(c, 1, dim, false, n_slots, va.group);
        if (dim == 2) launch_eval_pixels(va, m3, grid, s);
        else launch_eval_voxels(va, m4, grid, s);
        tm.mark();
    }
    if (dim == 3) {
        NormalsArgs na = {};
        na.arena = c->arena;
        na.image = c->filled[3];
        na.normals = c->normals;
        na.size = S;
        na.owned = c->owned_tiles;
        na.n_owned = c->n_owned;
        na.tiles0 = c->tiles[0];
        na.tiles1 = c->tiles[1];
        na.tiles2 = c->tiles[2];
        na.ctl = c->ctl;
        na.queue = &c->ctl->queue[q++];
        na.n_slots = n_slots;
        launch_normals(na, m4, c->sm_count * cached_occupancy(c, 2, 3, false, n_slots), s);
        tm.mark();
    }
    c->stats.n_launches = tm.n - 1;
    c->hint_tape = plan;
    c->hint_dim = dim;
    MPRB_CUDA(cudaGetLastError());
    return MPRB_OK;
}

// Closes the frame on the context's stream: counters to the host, end-of-frame event.
int end_frame(mprb_ctx* c) {
    MPRB_CUDA(cudaSetDevice(c->device));
    MPRB_CUDA(cudaMemcpyAsync(c->ctl_host, c->ctl, sizeof(FrameCtl), cudaMemcpyDeviceToHost, c->stream));
    MPRB_CUDA(cudaEventRecord(c->ev_end, c->stream));
    return MPRB_OK;
}

// One frame on every device of the context.  A multi-GPU context (mprb_ctx_opts::n_gpus / MPRB_GPUS)
// is a primary plus one sub-context per further device, each rendering the 64x64-px screen columns
// (x + y) % N == its index on its own stream, arena and tile lists - nothing is exchanged during the
// frame.  When a peer is done it writes the depth and normal blocks it owns STRAIGHT INTO the
// primary's full-size images over NVLink (k_publish: peer-to-peer stores, no staging buffer, no
// pack / unpack), and the primary's stream waits for those stores; so stages[3].filled and normals
// hold the whole frame on the primary device, which is all a caller of the reference's surface
// ever looks at.  The tile lists and per-level images of a multi-GPU context cover the primary's
// columns only.
int render_all(mprb_ctx* c, int dim, const mprb_tape* plan, bool cells_on_host, const float* matrix, float z,
               bool heat = false, bool brute = false)
{
    if (!c || !plan) return fail(MPRB_E_ARG, "null context or tape");
    if (c->peers.empty()) {
        if (int e = render(c, dim, plan, cells_on_host, matrix, z, heat, brute)) return e;
        return end_frame(c);
    }
    if (heat || brute) return fail(MPRB_E_ARG, "the analysis variants (brute, heatmap) run on single-GPU contexts only");
    if (int e = render(c, dim, plan, cells_on_host, matrix, z)) return e;
    MPRB_CUDA(cudaEventRecord(c->ev_body, c->stream));           // the primary no longer touches its images
    for (mprb_ctx* p : c->peers) {
        if (int e = render(p, dim, plan, cells_on_host, matrix, z)) return e;
        MPRB_CUDA(cudaStreamWaitEvent(p->stream, c->ev_body, 0));
        launch_publish(p->size, p->row_mod, p->row_rem, dim, p->filled[3], p->normals, c->filled[3], c->normals,
                       p->stream);
        MPRB_CUDA(cudaEventRecord(p->ev_body, p->stream));
        MPRB_CUDA(cudaGetLastError());
    }
    MPRB_CUDA(cudaSetDevice(c->device));
    for (mprb_ctx* p : c->peers) MPRB_CUDA(cudaStreamWaitEvent(c->stream, p->ev_body, 0));
    for (mprb_ctx* p : c->peers)
        if (int e = end_frame(p)) return e;
    return end_frame(c);
}

const mprb_tape* host_plan_for(mprb_ctx* c, const uint64_t* cells, int32_t n) {
    if (c->host_plan && int32_t(c->host_plan_cells.size()) == n &&
        memcmp(c->host_plan_cells.data(), cells, sizeof(uint64_t) * size_t(n)) == 0)
        return c->host_plan;
    if (c->host_plan) mprb_tape_destroy(c->host_plan);
    c->host_plan = nullptr;
    if (mprb_tape_create(cells, n, &c->host_plan) != MPRB_OK) return nullptr;
    c->host_plan_cells.assign(cells, cells + n);
    return c->host_plan;
}

// Waits for the frame and publishes the counters.
int finish_one(mprb_ctx* c, int dim);
int finish(mprb_ctx* c, int dim) {
    int rc = finish_one(c, dim);
    for (mprb_ctx* p : c->peers) {
        const int e = finish_one(p, dim);
        if (e && !rc) rc = e;
        // a multi-GPU frame reports the work of all its devices
        mprb_frame_stats& a = c->stats;
        const mprb_frame_stats& b = p->stats;
        for (int i = 0; i < 3; ++i) {
            a.n_active[i] += b.n_active[i];
            a.i_tiles[i] += b.i_tiles[i];
            a.i_cells[i] += b.i_cells[i];
            a.p_tiles[i] += b.p_tiles[i];
            a.p_cells[i] += b.p_cells[i];
            a.p_kept[i] += b.p_kept[i];
        }
        a.f_tiles += b.f_tiles;
        a.f_cells += b.f_cells;
        a.f_items += b.f_items;
        a.p_written += b.p_written;
        a.i_sub_tiles += b.i_sub_tiles;
        a.n_pixels += b.n_pixels;
        a.n_cells += b.n_cells;
        a.overflow |= b.overflow;
        a.n_launches += b.n_launches + 1;      // + its publish kernel
    }
    if (!c->peers.empty()) MPRB_CUDA(cudaSetDevice(c->device));
    return rc;
}
int finish_one(mprb_ctx* c, int dim) {
    MPRB_CUDA(cudaSetDevice(c->device));
    MPRB_CUDA(cudaStreamSynchronize(c->stream));
    const FrameCtl& f = *c->ctl_host;
    mprb_frame_stats& st = c->stats;
    const int n_launches = st.n_launches;
    memset(&st, 0, sizeof(st));
    st.n_launches = n_launches;
    const int n_levels = dim == 3 ? 3 : 2;
    for (int i = 0; i < 3; ++i) {
        st.n_active[i] = f.n_active[i];
        st.i_tiles[i] = f.stats[ST_I_TILES + i];
        st.i_cells[i] = f.stats[ST_I_CELLS + i];
        st.p_tiles[i] = f.stats[ST_P_TILES + i];
        st.p_cells[i] = f.stats[ST_P_CELLS + i];
        st.p_kept[i] = f.stats[ST_P_KEPT + i];
    }
    st.f_tiles = f.stats[ST_F_TILES];
    st.f_cells = f.stats[ST_F_CELLS];
    st.n_pixels = f.stats[ST_N_PIXELS];
    st.n_cells = f.stats[ST_N_CELLS];
    st.f_items = f.stats[ST_F_ITEMS];
    st.p_written = f.stats[ST_P_WRITTEN];
    st.i_sub_tiles = f.stats[ST_I_SUB];
    st.overflow = f.overflow;
    c->hint_parents = f.n_active[0];
    c->hint_plans = f.plan_count;
    cudaEventElapsedTime(&st.gpu_ms, c->ev_begin, c->ev_end);
    if (c->timing) {
        for (int i = 0; i < n_launches && i < kMaxLaunches; ++i)
            cudaEventElapsedTime(&st.kernel_ms[i], c->ev_k[i], c->ev_k[i + 1]);
    }
    // Host-visible mirrors of the reference's members.
    const int tps0 = c->size / 64;
    c->tile_array_size[0] = dim == 3 ? uint64_t(tps0) * tps0 * tps0 : uint64_t(tps0) * tps0;
    if (dim == 3) {
        c->tile_array_size[1] = uint64_t(f.n_active[0]) * 64;
        c->tile_array_size[2] = uint64_t(f.n_active[1]) * 64;
        c->tile_array_size[3] = uint64_t(f.n_active[2]);
    } else {
        c->tile_array_size[2] = uint64_t(f.n_active[0]) * 64;
        c->tile_array_size[3] = uint64_t(f.n_active[1]);
    }
    st.tape_index = f.tape_cursor;
    *c->tape_index = f.tape_cursor;
    *c->num_active_tiles = f.n_active[n_levels - 1];
    if (f.overflow)
        return fail(MPRB_E_OVERFLOW, "tile list overflow (mask 0x%x): more than %lld tiles at one level",
                    f.overflow, kMaxStageTiles);
    return MPRB_OK;
}

}  // namespace

////////////////////////////////////////////////////////////////////////////////

extern "C" {

const char* mprb_last_error(void) { return g_error.c_str(); }
const char* mprb_version(void) { return "mprb 0.1 (sm_100a)"; }

static int create_one(int32_t image_size_px, const mprb_ctx_opts* opts, mprb_ctx** out);

int mprb_ctx_create(int32_t image_size_px, const mprb_ctx_opts* opts, mprb_ctx** out) {
    if (!out) return fail(MPRB_E_ARG, "null out pointer");
    *out = nullptr;
    int n_gpus = opts ? opts->n_gpus : 0;
    if (n_gpus <= 0) {                       // callers of the reference's surface (mpr::Context) pass none
        const char* env = getenv("MPRB_GPUS");
        n_gpus = env ? atoi(env) : 1;
    }
    if (n_gpus <= 1) return create_one(image_size_px, opts, out);

    // ---- one context spanning n_gpus devices of this process ------------------------------------
    if (opts && (opts->row_mod > 1 || opts->row_begin != 0 || opts->row_end != 0))
        return fail(MPRB_E_ARG, "n_gpus > 1 shards the frame itself; do not combine it with row_* options");
    int first = opts ? opts->device : -1, n_dev = 0;
    if (first < 0) MPRB_CUDA(cudaGetDevice(&first));
    MPRB_CUDA(cudaGetDeviceCount(&n_dev));
    if (first + n_gpus > n_dev)
        return fail(MPRB_E_ARG, "n_gpus = %d from device %d, but only %d devices are visible", n_gpus, first, n_dev);
    for (int i = 1; i < n_gpus; ++i) {
        int can = 0;
        MPRB_CUDA(cudaDeviceCanAccessPeer(&can, first + i, first));
        if (!can) return fail(MPRB_E_ARG, "device %d cannot write device %d's memory (no peer access)", first + i, first);
    }
    mprb_ctx_opts o = {};
    if (opts) o = *opts;
    o.n_gpus = 1;
    o.row_begin = 0;
    o.row_end = 0;
    o.row_mod = n_gpus;
    o.col_step = 1;
    mprb_ctx* primary = nullptr;
    for (int i = 0; i < n_gpus; ++i) {
        o.device = first + i;
        o.row_rem = i;
        mprb_ctx* sub = nullptr;
        if (int e = create_one(image_size_px, &o, &sub)) {
            if (primary) mprb_ctx_destroy(primary);
            return e;
        }
        if (i == 0) {
            primary = sub;
            continue;
        }
        primary->peers.push_back(sub);
        cudaError_t pe = cudaDeviceEnablePeerAccess(first, 0);          // current device = first + i
        if (pe == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); pe = cudaSuccess; }
        if (pe != cudaSuccess) {
            mprb_ctx_destroy(primary);
            return fail(MPRB_E_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", first + i, first, cudaGetErrorString(pe));
        }
        // the peers store into the primary's (managed) frame buffers: map them there, do not migrate
        const size_t n = size_t(image_size_px) * image_size_px;
        cudaMemAdvise(primary->filled[3], sizeof(int32_t) * n, cudaMemAdviseSetAccessedBy, first + i);
        cudaMemAdvise(primary->normals, sizeof(uint32_t) * n, cudaMemAdviseSetAccessedBy, first + i);
    }
    cudaGetLastError();
    MPRB_CUDA(cudaSetDevice(first));
    *out = primary;
    return MPRB_OK;
}

static int create_one(int32_t image_size_px, const mprb_ctx_opts* opts, mprb_ctx** out) {
    *out = nullptr;
    if (image_size_px < 64 || image_size_px % 64 != 0)
        return fail(MPRB_E_ARG, "image_size_px must be a positive multiple of 64 (got %d)", image_size_px);
    int device = opts ? opts->device : -1;
    if (device < 0) MPRB_CUDA(cudaGetDevice(&device));
    MPRB_CUDA(cudaSetDevice(device));

    mprb_ctx* c = new mprb_ctx;
    c->device = device;
    c->size = image_size_px;
    const int tps0 = image_size_px / 64;
    c->serial_root = getenv("MPRB_SERIAL_ROOT") != nullptr;
    if (const char* e = getenv("MPRB_SUB_WAVES")) c->sub_waves = std::max(0, atoi(e));
    c->row_begin = opts ? opts->row_begin : 0;
    c->row_end = (opts && opts->row_end > 0) ? opts->row_end : tps0;
    c->row_mod = (opts && opts->row_mod > 1) ? opts->row_mod : 1;
    c->row_rem = (opts && opts->row_mod > 1) ? opts->row_rem : 0;
    c->col_step = (opts && opts->row_mod > 1 && opts->col_step) ? 1 : 0;
    if (c->row_rem < 0 || c->row_rem >= c->row_mod) {
        delete c;
        return fail(MPRB_E_ARG, "row_rem must lie in [0, row_mod)");
    }
    if (c->row_begin < 0 || c->row_end > tps0 || c->row_begin >= c->row_end) {
        delete c;
        return fail(MPRB_E_ARG, "tile-row band [%d, %d) is outside [0, %d)", opts ? opts->row_begin : 0,
                    opts ? opts->row_end : 0, tps0);
    }
    int max_smem = 0;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    init_kernels(max_smem);

    auto bail = [&](cudaError_t e, const char* what) {
        fail(MPRB_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
        mprb_ctx_destroy(c);
        return MPRB_E_CUDA;
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    cudaEventCreate(&c->ev_begin);
    cudaEventCreate(&c->ev_end);
    cudaEventCreateWithFlags(&c->ev_body, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_ext, cudaEventDisableTiming);
    for (auto& ev : c->ev_k) cudaEventCreate(&ev);

    // Filled images: (S/64)^2, (S/16)^2, (S/4)^2, S^2 (context.cpp:21-26)
    for (int i = 0; i < 4; ++i) {
        const size_t side = size_t(image_size_px) / (64 >> (2 * i));
        if ((e = managed_alloc(&c->filled[i], side * side, device)) != cudaSuccess)
            return bail(e, "filled image");
    }
    if ((e = managed_alloc(&c->normals, size_t(image_size_px) * image_size_px, device)) != cudaSuccess)
        return bail(e, "normals");
    {   // The screen tiles this context owns, for the per-pixel normal pass
        std::vector<int32_t> owned;
        for (int ty = c->row_begin; ty < c->row_end; ++ty)
            for (int tx = 0; tx < tps0; ++tx)
                if ((ty + c->col_step * tx) % c->row_mod == c->row_rem) owned.push_back(ty * tps0 + tx);
        c->n_owned = int(owned.size());
        if ((e = cudaMalloc(&c->owned_tiles, sizeof(int32_t) * std::max<size_t>(owned.size(), 1))) != cudaSuccess)
            return bail(e, "owned tile list");
        if ((e = cudaMemcpy(c->owned_tiles, owned.data(), sizeof(int32_t) * owned.size(), cudaMemcpyHostToDevice)) !=
            cudaSuccess)
            return bail(e, "owned tile list");
    }
    const long long chunks = (opts && opts->num_subtapes > 0) ? opts->num_subtapes : 640000;
    c->arena_cells = chunks * kChunk;
    if (c->arena_cells > INT32_MAX) {
        mprb_ctx_destroy(c);
        return fail(MPRB_E_ARG, "num_subtapes too large for 32-bit tape indices");
    }
    if ((e = managed_alloc(&c->arena, size_t(c->arena_cells) + kChunk, device)) != cudaSuccess)
        return bail(e, "tape arena");
    // Host-side mirrors of two device counters (the device copies live in FrameCtl)
    // The two counters the reference keeps in managed memory are written by the HOST after every frame and never
    // read by a kernel here.  As managed allocations they shared unified-memory pages with small device-side
    // arrays (stage-0 tiles, level-0 image, tape cells): each host write pulled such a page to the host, the next
    // frame's first kernel faulted it back (+0.4 ms on a 0.43 ms frame, most frames once the driver's thrashing
    // heuristics settled).  Page-locked host memory is just as readable through the same pointers.
    if ((e = cudaMallocHost(&c->tape_index, sizeof(int32_t))) != cudaSuccess) return bail(e, "tape_index");
    if ((e = cudaMallocHost(&c->num_active_tiles, sizeof(int32_t))) != cudaSuccess)
        return bail(e, "num_active_tiles");
    *c->tape_index = 0;
    *c->num_active_tiles = 0;
    // Stage 0 holds every 64^3 tile of the volume (context.cpp:39-43)
    const long long count0 = (long long)tps0 * tps0 * tps0;
    c->tiles_cap[0] = 0;
    {
        TileNode* p = nullptr;
        if ((e = managed_alloc(&p, size_t(count0), device)) != cudaSuccess) return bail(e, "stage 0 tiles");
        c->tiles[0] = p;
        c->tiles_cap[0] = count0;
        if ((e = cudaMalloc(&c->active_list[0], sizeof(int32_t) * size_t(count0))) != cudaSuccess)
            return bail(e, "active list");
    }
    if ((e = cudaMalloc(&c->ctl, sizeof(FrameCtl))) != cudaSuccess) return bail(e, "control block");
    if ((e = cudaMallocHost(&c->ctl_host, sizeof(FrameCtl))) != cudaSuccess) return bail(e, "pinned control block");
    memset(c->ctl_host, 0, sizeof(FrameCtl));
    cudaDeviceSynchronize();
    *out = c;
    return MPRB_OK;
}

void mprb_ctx_destroy(mprb_ctx* c) {
    if (!c) return;
    for (mprb_ctx* p : c->peers) mprb_ctx_destroy(p);
    c->peers.clear();
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 4; ++i) {
        if (c->filled[i]) cudaFree(c->filled[i]);
        if (c->tiles[i]) cudaFree(c->tiles[i]);
    }
    for (int i = 0; i < 3; ++i) if (c->active_list[i]) cudaFree(c->active_list[i]);
    if (c->float_items) cudaFree(c->float_items);
    if (c->arena) cudaFree(c->arena);
    if (c->tape_index) cudaFreeHost(c->tape_index);
    if (c->num_active_tiles) cudaFreeHost(c->num_active_tiles);
    if (c->normals) cudaFree(c->normals);
    if (c->ctl) cudaFree(c->ctl);
    if (c->ctl_host) cudaFreeHost(c->ctl_host);
    if (c->stage_cells) cudaFreeHost(c->stage_cells);
    if (c->heat_units) cudaFree(c->heat_units);
    if (c->owned_tiles) cudaFree(c->owned_tiles);
    if (c->plans) cudaFree(c->plans);
    if (c->plan_of) cudaFree(c->plan_of);
    if (c->host_plan) mprb_tape_destroy(c->host_plan);
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_end) cudaEventDestroy(c->ev_end);
    if (c->ev_body) cudaEventDestroy(c->ev_body);
    if (c->ev_ext) cudaEventDestroy(c->ev_ext);
    for (auto& ev : c->ev_k) if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int mprb_ctx_buffers(mprb_ctx* c, mprb_buffers* out) {
    if (!c || !out) return fail(MPRB_E_ARG, "null argument");
    out->image_size_px = c->size;
    for (int i = 0; i < 4; ++i) {
        out->filled[i] = c->filled[i];
        out->tiles[i] = reinterpret_cast<mprb_tile_node*>(c->tiles[i]);
        out->tile_array_size[i] = c->tile_array_size[i];
    }
    out->tape_data = c->arena;
    out->tape_index = c->tape_index;
    out->num_active_tiles = c->num_active_tiles;
    out->normals = c->normals;
    return MPRB_OK;
}

int mprb_ctx_set_timing(mprb_ctx* c, int enabled) {
    if (!c) return fail(MPRB_E_ARG, "null context");
    c->timing = enabled != 0;
    return MPRB_OK;
}

int mprb_tape_create(const uint64_t* host_cells, int32_t n_cells, mprb_tape** out) {
    if (!out) return fail(MPRB_E_ARG, "null out pointer");
    *out = nullptr;
    if (int e = validate_tape(host_cells, n_cells)) return e;
    mprb_tape* t = new mprb_tape;
    cudaError_t e = cudaMallocManaged(&t->cells, sizeof(uint64_t) * size_t(n_cells));
    if (e == cudaSuccess)
        e = cudaMemcpy(t->cells, host_cells, sizeof(uint64_t) * size_t(n_cells), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (t->cells) cudaFree(t->cells);
        delete t;
        return fail(MPRB_E_CUDA, "tape upload: %s", cudaGetErrorString(e));
    }
    t->length = n_cells;
    t->n_slots = tape_num_slots(host_cells, n_cells);
    t->host_cells.assign(host_cells, host_cells + n_cells);
    t->host_chunked = chunk_layout(host_cells, n_cells);
    t->n_chunked = int32_t(t->host_chunked.size());
    if (n_cells > 2 && n_cells - 2 < (1 << 20)) {
        RootPlan plan = build_root_plan(host_cells, n_cells);
        if (plan.result_v != 0) {
            t->n_levels = int(plan.level_start.size()) - 1;
            t->result_v = plan.result_v;
            t->sched.swap(plan.sched);
            t->level_start.swap(plan.level_start);
            t->prevw.swap(plan.prevw);
        }
    }
    // Device-side copies are made per device on first use (tape_on); make the current device's now
    // so that an out-of-memory condition surfaces here.
    int dev = 0;
    cudaGetDevice(&dev);
    if (!tape_on(t, dev)) {
        mprb_tape_destroy(t);
        return MPRB_E_CUDA;
    }
    *out = t;
    return MPRB_OK;
}

void mprb_tape_destroy(mprb_tape* t) {
    if (!t) return;
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : t->dev) {
        cudaSetDevice(kv.first);
        TapeDev& d = kv.second;
        if (d.cells) cudaFree(d.cells);
        if (d.chunked) cudaFree(d.chunked);
        if (d.sched) cudaFree(d.sched);
        if (d.level_start) cudaFree(d.level_start);
        if (d.prevw) cudaFree(d.prevw);
    }
    cudaSetDevice(cur);
    if (t->cells) cudaFree(t->cells);
    delete t;
}

const uint64_t* mprb_tape_data(const mprb_tape* t) { return t ? t->cells : nullptr; }
int32_t mprb_tape_length(const mprb_tape* t) { return t ? t->length : 0; }
int32_t mprb_tape_num_slots(const mprb_tape* t) { return t ? t->n_slots : 0; }

int mprb_render2d(mprb_ctx* c, const mprb_tape* t, const float mat3[9], float z) {
    if (!c || !t || !mat3) return fail(MPRB_E_ARG, "null argument");
    if (int e = render_all(c, 2, t, false, mat3, z)) return e;
    return finish(c, 2);
}

int mprb_render3d(mprb_ctx* c, const mprb_tape* t, const float mat4[16]) {
    if (!c || !t || !mat4) return fail(MPRB_E_ARG, "null argument");
    if (int e = render_all(c, 3, t, false, mat4, 0.0f)) return e;
    return finish(c, 3);
}

int mprb_render2d_brute(mprb_ctx* c, const mprb_tape* t, const float mat3[9], float z) {
    if (!c || !t || !mat3) return fail(MPRB_E_ARG, "null argument");
    if (int e = render_all(c, 2, t, false, mat3, z, false, true)) return e;
    return finish(c, 2);
}

// Managed S*S float result, owned by the caller (the reference returns a Ptr<float[]>).
static int heatmap_out(mprb_ctx* c, const mprb_tape* t, float** out) {
    const size_t n = size_t(c->size) * c->size;
    float* h = nullptr;
    MPRB_CUDA(managed_alloc(&h, n, c->device));
    launch_heat_finish(c->heat_units, h, (long long)n, t->length - 2,
                       int(std::min<size_t>(size_t(c->sm_count) * 8, (n + 255) / 256)), c->stream);
    *out = h;
    return MPRB_OK;
}

int mprb_render2d_heatmap(mprb_ctx* c, const mprb_tape* t, const float mat3[9], float z, float** heatmap) {
    if (!c || !t || !mat3 || !heatmap) return fail(MPRB_E_ARG, "null argument");
    if (int e = render_all(c, 2, t, false, mat3, z, true)) return e;
    if (int e = heatmap_out(c, t, heatmap)) return e;
    const int e = finish(c, 2);
    if (e) {              // e.g. a tile list overflowed: the caller gets no half-valid buffer to free
        cudaFree(*heatmap);
        *heatmap = nullptr;
    }
    return e;
}

int mprb_render3d_heatmap(mprb_ctx* c, const mprb_tape* t, const float mat4[16], float** heatmap) {
    if (!c || !t || !mat4 || !heatmap) return fail(MPRB_E_ARG, "null argument");
    if (int e = render_all(c, 3, t, false, mat4, 0.0f, true)) return e;
    if (int e = heatmap_out(c, t, heatmap)) return e;
    const int e = finish(c, 3);
    if (e) {              // e.g. a tile list overflowed: the caller gets no half-valid buffer to free
        cudaFree(*heatmap);
        *heatmap = nullptr;
    }
    return e;
}

int mprb_render2d_host(mprb_ctx* c, const uint64_t* host_cells, int32_t n_cells,
                       const float mat3[9], float z, int32_t* image_out) {
    if (!c || !mat3) return fail(MPRB_E_ARG, "null argument");
    if (int e = validate_tape(host_cells, n_cells)) return e;
    const mprb_tape* t = host_plan_for(c, host_cells, n_cells);
    if (!t) return MPRB_E_CUDA;
    if (int e = render_all(c, 2, t, true, mat3, z)) return e;
    const size_t n = size_t(c->size) * c->size;
    if (image_out)
        MPRB_CUDA(cudaMemcpyAsync(image_out, c->filled[3], sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    return finish(c, 2);
}

int mprb_render3d_host(mprb_ctx* c, const uint64_t* host_cells, int32_t n_cells,
                       const float mat4[16], int32_t* depth_out, uint32_t* normals_out) {
    if (!c || !mat4) return fail(MPRB_E_ARG, "null argument");
    if (int e = validate_tape(host_cells, n_cells)) return e;
    const mprb_tape* t = host_plan_for(c, host_cells, n_cells);
    if (!t) return MPRB_E_CUDA;
    if (int e = render_all(c, 3, t, true, mat4, 0.0f)) return e;
    const size_t n = size_t(c->size) * c->size;
    if (depth_out)
        MPRB_CUDA(cudaMemcpyAsync(depth_out, c->filled[3], sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    if (normals_out)
        MPRB_CUDA(cudaMemcpyAsync(normals_out, c->normals, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    return finish(c, 3);
}

// Address a kernel on `device` may store through for `p`, or null: device / managed memory as it is,
// page-locked host memory by its mapping.
static void* device_address(void* p) {
    if (!p) return nullptr;
    cudaPointerAttributes at = {};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return p;
    if (at.type == cudaMemoryTypeHost) return at.devicePointer;
    return nullptr;
}

int mprb_ctx_publish(mprb_ctx* c, int dim, int32_t* dst_image, uint32_t* dst_normals) {
    if (!c) return fail(MPRB_E_ARG, "null context");
    if (dim != 2 && dim != 3) return fail(MPRB_E_ARG, "dim must be 2 or 3");
    if (!c->peers.empty()) return fail(MPRB_E_ARG, "a multi-GPU context holds its whole frame on the primary device");
    const int tiles = c->size / 64;
    const int world = std::max(c->row_mod, 1);
    if (c->row_begin != 0 || c->row_end != tiles || (world > 1 && c->col_step != 1))
        return fail(MPRB_E_ARG, "publish needs the tile-cyclic sharding (col_step = 1) over the whole frame, or none");
    MPRB_CUDA(cudaSetDevice(c->device));
    int32_t* const d_img = static_cast<int32_t*>(device_address(dst_image));
    uint32_t* const d_nrm = dim == 3 && dst_normals ? static_cast<uint32_t*>(device_address(dst_normals)) : nullptr;
    if (!d_img || (dim == 3 && dst_normals && !d_nrm))
        return fail(MPRB_E_ARG, "destination is neither device memory nor page-locked host memory (cudaHostRegister it)");
    launch_publish(c->size, world, c->row_rem, d_nrm ? 3 : 2, c->filled[3], c->normals, d_img, d_nrm, c->stream);
    MPRB_CUDA(cudaGetLastError());
    MPRB_CUDA(cudaStreamSynchronize(c->stream));
    return MPRB_OK;
}

static int exchange_check(const mprb_ctx* c, int dim) {
    if (!c) return fail(MPRB_E_ARG, "null context");
    if (dim != 2 && dim != 3) return fail(MPRB_E_ARG, "dim must be 2 or 3");
    const int tiles = c->size / 64;
    if (c->row_mod <= 1 || c->row_begin != 0 || c->row_end != tiles || tiles % c->row_mod != 0)
        return fail(MPRB_E_ARG, "exchange needs row_mod = world > 1 over the whole frame, tiles per side divisible by world");
    if (dim == 3 && c->size > 32768)
        return fail(MPRB_E_ARG, "exchange carries depth as int16: image side must not exceed 32768");
    return MPRB_OK;
}

size_t mprb_exchange_bytes(const mprb_ctx* c, int dim) {
    if (exchange_check(c, dim)) return 0;
    return exchange_rank_bytes(c->size, c->row_mod, dim);
}

int mprb_exchange_pack(mprb_ctx* c, int dim, void* dst, void* stream) {
    if (int e = exchange_check(c, dim)) return e;
    if (!dst) return fail(MPRB_E_ARG, "null buffer");
    MPRB_CUDA(cudaSetDevice(c->device));
    launch_exchange(true, c->size, c->row_mod, c->row_rem, c->col_step, dim, c->filled[3], c->normals, dst,
                    stream ? cudaStream_t(stream) : c->stream);
    MPRB_CUDA(cudaGetLastError());
    if (stream && cudaStream_t(stream) != c->stream) {      // order the next frame after this launch
        MPRB_CUDA(cudaEventRecord(c->ev_ext, cudaStream_t(stream)));
        c->ext_pending = true;
    }
    return MPRB_OK;
}

int mprb_exchange_unpack(mprb_ctx* c, int dim, const void* src, void* stream) {
    if (int e = exchange_check(c, dim)) return e;
    if (!src) return fail(MPRB_E_ARG, "null buffer");
    MPRB_CUDA(cudaSetDevice(c->device));
    launch_exchange(false, c->size, c->row_mod, c->row_rem, c->col_step, dim, c->filled[3], c->normals,
                    const_cast<void*>(src), stream ? cudaStream_t(stream) : c->stream);
    MPRB_CUDA(cudaGetLastError());
    if (stream && cudaStream_t(stream) != c->stream) {      // order the next frame after this launch
        MPRB_CUDA(cudaEventRecord(c->ev_ext, cudaStream_t(stream)));
        c->ext_pending = true;
    }
    return MPRB_OK;
}

struct mprb_effects {
    float* kernel = nullptr;     // device, 64x3
    float* rvecs = nullptr;      // device, 256x3
    int32_t* image = nullptr;    // managed
    int32_t* tmp = nullptr;      // managed
    int size = 0;
};

static int effects_resize(mprb_effects* fx, mprb_ctx* c) {
    if (fx->size == c->size) return MPRB_OK;
    if (fx->image) cudaFree(fx->image);
    if (fx->tmp) cudaFree(fx->tmp);
    fx->image = fx->tmp = nullptr;
    const size_t n = size_t(c->size) * c->size;
    MPRB_CUDA(managed_alloc(&fx->image, n, c->device));
    MPRB_CUDA(managed_alloc(&fx->tmp, n, c->device));
    fx->size = c->size;
    return MPRB_OK;
}

int mprb_effects_create(const float* kernel, const float* rvecs, mprb_effects** out) {
    if (!kernel || !rvecs || !out) return fail(MPRB_E_ARG, "null argument");
    mprb_effects* fx = new mprb_effects;
    cudaError_t e = cudaMalloc(&fx->kernel, sizeof(float) * 64 * 3);
    if (e == cudaSuccess) e = cudaMalloc(&fx->rvecs, sizeof(float) * 256 * 3);
    if (e == cudaSuccess) e = cudaMemcpy(fx->kernel, kernel, sizeof(float) * 64 * 3, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(fx->rvecs, rvecs, sizeof(float) * 256 * 3, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        mprb_effects_destroy(fx);
        return fail(MPRB_E_CUDA, "effects: %s", cudaGetErrorString(e));
    }
    *out = fx;
    return MPRB_OK;
}

void mprb_effects_destroy(mprb_effects* fx) {
    if (!fx) return;
    if (fx->kernel) cudaFree(fx->kernel);
    if (fx->rvecs) cudaFree(fx->rvecs);
    if (fx->image) cudaFree(fx->image);
    if (fx->tmp) cudaFree(fx->tmp);
    delete fx;
}

int mprb_effects_draw_ssao(mprb_effects* fx, mprb_ctx* c) {
    if (!fx || !c) return fail(MPRB_E_ARG, "null argument");
    MPRB_CUDA(cudaSetDevice(c->device));
    if (int e = effects_resize(fx, c)) return e;
    launch_clear_pair(fx->tmp, fx->image, (long long)c->size * c->size, c->stream);
    launch_draw_ssao(c->filled[3], c->normals, fx->kernel, fx->rvecs, c->size, fx->tmp, c->stream);
    launch_blur_ssao(c->filled[3], fx->tmp, c->size, fx->image, c->stream);
    MPRB_CUDA(cudaStreamSynchronize(c->stream));
    return MPRB_OK;
}

int mprb_effects_draw_shaded(mprb_effects* fx, mprb_ctx* c) {
    if (!fx || !c) return fail(MPRB_E_ARG, "null argument");
    MPRB_CUDA(cudaSetDevice(c->device));
    if (int e = effects_resize(fx, c)) return e;
    launch_clear_pair(fx->tmp, fx->image, (long long)c->size * c->size, c->stream);
    launch_draw_ssao(c->filled[3], c->normals, fx->kernel, fx->rvecs, c->size, fx->image, c->stream);
    launch_blur_ssao(c->filled[3], fx->image, c->size, fx->tmp, c->stream);
    launch_draw_shaded(c->filled[3], c->normals, fx->tmp, c->size, fx->image, c->stream);
    MPRB_CUDA(cudaStreamSynchronize(c->stream));
    return MPRB_OK;
}

int mprb_effects_buffers(mprb_effects* fx, int32_t** image, int32_t** tmp) {
    if (!fx) return fail(MPRB_E_ARG, "null argument");
    if (image) *image = fx->image;
    if (tmp) *tmp = fx->tmp;
    return MPRB_OK;
}

int mprb_frame_stats_get(mprb_ctx* c, mprb_frame_stats* out) {
    if (!c || !out) return fail(MPRB_E_ARG, "null argument");
    *out = c->stats;
    return MPRB_OK;
}

int mprb_tape_from_frep(const uint8_t* bytes, size_t n_bytes, int simplify,
                        uint64_t** cells_out, int32_t* n_cells_out, int32_t* n_slots_out) {
    if (!bytes || !cells_out || !n_cells_out) return fail(MPRB_E_ARG, "null argument");
    std::istringstream in(std::string(reinterpret_cast<const char*>(bytes), n_bytes), std::ios::binary);
    libfive::Cache::setSimplify(simplify != 0);
    auto archive = libfive::Archive::deserialize(in);
    libfive::Cache::setSimplify(true);
    if (archive.shapes.empty() || !archive.shapes.front().tree.id())
        return fail(MPRB_E_PARSE, "no shape in archive");
    int n_slots = 0;
    const auto tape = pack_tape(archive.shapes.front().tree, &n_slots);
    uint64_t* p = static_cast<uint64_t*>(malloc(sizeof(uint64_t) * tape.size()));
    if (!p) return fail(MPRB_E_ARG, "out of memory");
    memcpy(p, tape.data(), sizeof(uint64_t) * tape.size());
    *cells_out = p;
    *n_cells_out = int32_t(tape.size());
    if (n_slots_out) *n_slots_out = n_slots;
    return MPRB_OK;
}

void mprb_free(void* p) { free(p); }
void mprb_free_device(void* p) { if (p) cudaFree(p); }

int mprb_malloc_managed(size_t n_bytes, void** out) {
    if (!out) return fail(MPRB_E_ARG, "null argument");
    *out = nullptr;
    MPRB_CUDA(cudaMallocManaged(out, n_bytes ? n_bytes : 1));
    return MPRB_OK;
}

}  // extern "C"
