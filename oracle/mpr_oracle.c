/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the
 * product path (mpr_b200/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / reference legs may use it, and only as the checker
 * or as the timed CPU baseline.
 *
 * A plain-C, CPU restatement of the reference renderer's algorithm for the hot
 * path (reference src/context.cu, inc/gpu_interval.hpp, inc/gpu_deriv.hpp).
 * It keeps the reference's own decomposition -- one tile per "thread", a
 * 128-entry slot array, per-tile choice bits, back-to-front chunked tape
 * pushes claimed with a fetch-add -- so that it can be compared with both the
 * reference build (oracle/_ref, on a GPU) and the B200 kernels (which are
 * organised differently).  Each function cites the reference lines it follows.
 *
 * PARITY STATUS: the reference ships no golden vectors or known-answer tests
 * for this path.  This restatement is pinned against
 * outputs of the unmodified reference CUDA code run on a B200
 * (tests/golden/, minted by tools/gpu_check.py through oracle/_ref).
 * Known, documented divergence: the float transcendentals (sinf, cosf, asinf,
 * acosf, atanf, expf, logf, powf) come from glibc here and from CUDA libdevice
 * in the reference; they may differ in the last ulp, which can flip isolated
 * pixels for models that use them (bear, gears).  +, -, *, /, sqrt with
 * directed rounding are exact on both sides.
 */
#define _GNU_SOURCE
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#pragma STDC FENV_ACCESS ON

#define CHUNK 64
#define MAX_CHOICES 4096

enum {
    OP_END = 0, OP_JUMP = 1, OP_SQUARE, OP_SQRT, OP_NEG, OP_SIN, OP_COS, OP_ASIN, OP_ACOS,
    OP_ATAN, OP_EXP, OP_ABS, OP_LOG, OP_ADD_LI, OP_ADD_LR, OP_MUL_LI, OP_MUL_LR, OP_MIN_LI,
    OP_MIN_LR, OP_MAX_LI, OP_MAX_LR, OP_SUB_LI, OP_SUB_IR, OP_SUB_LR, OP_DIV_LI, OP_DIV_IR,
    OP_DIV_LR, OP_COPY_IMM, OP_COPY_LHS, OP_COPY_RHS
};

typedef struct { float lo, hi; } ival;
typedef struct { float dx, dy, dz, v; } dval;
typedef struct { int32_t position, tape, next; } tile_t;

typedef struct {
    int size;
    int32_t* filled[4];
    tile_t* tiles[4];
    size_t tile_count[4];     /* entries valid after the last frame */
    size_t tile_cap[4];
    uint64_t* arena;
    int64_t arena_cells;
    int32_t tape_index;
    uint32_t* normals;
    /* work counters of the last frame */
    uint64_t work_interval;   /* interval clause evaluations */
    uint64_t work_float;      /* float voxel*clause evaluations */
    uint64_t* heat;           /* optional S*S work meter, units of 1/4096 cell (render*_heatmap) */
} oracle_ctx;

/* ---- clause fields (reference inc/clause.hpp:18-23) ------------------------------- */
static inline unsigned c_op(uint64_t d) { return (unsigned)(d & 0xff); }
static inline unsigned c_out(uint64_t d) { return (unsigned)((d >> 8) & 0xff); }
static inline unsigned c_lhs(uint64_t d) { return (unsigned)((d >> 16) & 0xff); }
static inline unsigned c_rhs(uint64_t d) { return (unsigned)((d >> 24) & 0xff); }
static inline float c_imm(uint64_t d) { uint32_t u = (uint32_t)(d >> 32); float f; memcpy(&f, &u, 4); return f; }
static inline int32_t c_jump(uint64_t d) { return (int32_t)(uint32_t)(d >> 32); }

/* ---- directed rounding ------------------------------------------------------------ */
/* The optimiser must not fold or reorder these; the empty asm pins each value. */
static inline float pin(float x) { __asm__ volatile("" : "+x"(x)); return x; }
static inline float add_rd(float a, float b) { fesetround(FE_DOWNWARD); float r = pin(pin(a) + pin(b)); return r; }
static inline float add_ru(float a, float b) { fesetround(FE_UPWARD); float r = pin(pin(a) + pin(b)); return r; }
static inline float sub_rd(float a, float b) { fesetround(FE_DOWNWARD); float r = pin(pin(a) - pin(b)); return r; }
static inline float sub_ru(float a, float b) { fesetround(FE_UPWARD); float r = pin(pin(a) - pin(b)); return r; }
static inline float mul_rd(float a, float b) { fesetround(FE_DOWNWARD); float r = pin(pin(a) * pin(b)); return r; }
static inline float mul_ru(float a, float b) { fesetround(FE_UPWARD); float r = pin(pin(a) * pin(b)); return r; }
static inline float div_rd(float a, float b) { fesetround(FE_DOWNWARD); float r = pin(pin(a) / pin(b)); return r; }
static inline float div_ru(float a, float b) { fesetround(FE_UPWARD); float r = pin(pin(a) / pin(b)); return r; }
static inline float sqrt_rd(float a) { fesetround(FE_DOWNWARD); float r = pin(sqrtf(pin(a))); return r; }
static inline float sqrt_ru(float a) { fesetround(FE_UPWARD); float r = pin(sqrtf(pin(a))); return r; }
/* CUDA fminf/fmaxf: a NaN operand yields the other operand. */
static inline float cmin(float a, float b) { return isnan(a) ? b : isnan(b) ? a : (a < b ? a : b); }
static inline float cmax(float a, float b) { return isnan(a) ? b : isnan(b) ? a : (a > b ? a : b); }
/* Transcendentals are evaluated in round-to-nearest (the reference applies no
 * effective directed rounding to them, see ival.cuh header). */
static inline float near1(float (*f)(float), float x) { fesetround(FE_TONEAREST); return f(x); }

static inline ival iv(float lo, float hi) { ival r = {lo, hi}; return r; }

/* ---- interval operators (reference inc/gpu_interval.hpp) -------------------------- */
static ival iv_neg(ival a) { return iv(-a.hi, -a.lo); }                                  /* :65-67 */
static ival iv_add(ival a, ival b) { return iv(add_rd(a.lo, b.lo), add_ru(a.hi, b.hi)); }   /* :71-73 */
static ival iv_addf(ival a, float c) { return iv(add_rd(a.lo, c), add_ru(a.hi, c)); }       /* :75-77 */
static ival iv_sub(ival a, ival b) { return iv(sub_rd(a.lo, b.hi), sub_ru(a.hi, b.lo)); }   /* :284-286 */
static ival iv_subf(ival a, float c) { return iv(sub_rd(a.lo, c), sub_ru(a.hi, c)); }       /* :288-290 */
static ival iv_fsub(float c, ival b) { return iv(sub_rd(c, b.hi), sub_ru(c, b.lo)); }       /* :292-294 */

static ival iv_mul(ival x, ival y) {                                                      /* :85-146 */
    if (x.lo < 0.0f) {
        if (x.hi > 0.0f) {
            if (y.lo < 0.0f) {
                if (y.hi > 0.0f) return iv(cmin(mul_rd(x.lo, y.hi), mul_rd(x.hi, y.lo)),
                                           cmax(mul_ru(x.lo, y.lo), mul_ru(x.hi, y.hi)));
                return iv(mul_rd(x.hi, y.lo), mul_ru(x.lo, y.lo));
            }
            if (y.hi > 0.0f) return iv(mul_rd(x.lo, y.hi), mul_ru(x.hi, y.hi));
            return iv(0.0f, 0.0f);
        }
        if (y.lo < 0.0f) {
            if (y.hi > 0.0f) return iv(mul_rd(x.lo, y.hi), mul_ru(x.lo, y.lo));
            return iv(mul_rd(x.hi, y.hi), mul_ru(x.lo, y.lo));
        }
        if (y.hi > 0.0f) return iv(mul_rd(x.lo, y.hi), mul_ru(x.hi, y.lo));
        return iv(0.0f, 0.0f);
    }
    if (x.hi > 0.0f) {
        if (y.lo < 0.0f) {
            if (y.hi > 0.0f) return iv(mul_rd(x.hi, y.lo), mul_ru(x.hi, y.hi));
            return iv(mul_rd(x.hi, y.lo), mul_ru(x.lo, y.hi));
        }
        if (y.hi > 0.0f) return iv(mul_rd(x.lo, y.lo), mul_ru(x.hi, y.hi));
        return iv(0.0f, 0.0f);
    }
    return iv(0.0f, 0.0f);
}
static ival iv_mulf(ival x, float c) {                                                    /* :148-154 */
    if (c < 0.0f) return iv(mul_rd(x.hi, c), mul_ru(x.lo, c));
    return iv(mul_rd(x.lo, c), mul_ru(x.hi, c));
}
static ival iv_div(ival x, ival y) {                                                      /* :162-190 */
    if (y.lo <= 0.0f && y.hi >= 0.0f) return iv(-INFINITY, INFINITY);
    if (x.hi < 0.0f) {
        if (y.hi < 0.0f) return iv(div_rd(x.hi, y.lo), div_ru(x.lo, y.hi));
        return iv(div_rd(x.lo, y.lo), div_ru(x.hi, y.hi));
    } else if (x.lo < 0.0f) {
        if (y.hi < 0.0f) return iv(div_rd(x.hi, y.hi), div_ru(x.lo, y.hi));
        return iv(div_rd(x.lo, y.lo), div_ru(x.hi, y.lo));
    }
    if (y.hi < 0.0f) return iv(div_rd(x.hi, y.hi), div_ru(x.lo, y.lo));
    return iv(div_rd(x.lo, y.hi), div_ru(x.hi, y.lo));
}
static ival iv_divf(ival x, float c) {                                                    /* :192-200 */
    if (c < 0.0f) return iv(div_rd(x.hi, c), div_ru(x.lo, c));
    if (c > 0.0f) return iv(div_rd(x.lo, c), div_ru(x.hi, c));
    return iv(-INFINITY, INFINITY);
}
static ival iv_min(ival x, ival y, int* choice) {                                         /* :208-228 */
    if (x.hi < y.lo) { *choice = 1; return x; }
    if (y.hi < x.lo) { *choice = 2; return y; }
    return iv(cmin(x.lo, y.lo), cmin(x.hi, y.hi));
}
static ival iv_max(ival x, ival y, int* choice) {                                         /* :232-252 */
    if (x.lo > y.hi) { *choice = 1; return x; }
    if (y.lo > x.hi) { *choice = 2; return y; }
    return iv(cmax(x.lo, y.lo), cmax(x.hi, y.hi));
}
static ival iv_square(ival x) {                                                           /* :256-266 */
    if (x.hi < 0.0f) return iv(mul_rd(x.hi, x.hi), mul_ru(x.lo, x.lo));
    if (x.lo > 0.0f) return iv(mul_rd(x.lo, x.lo), mul_ru(x.hi, x.hi));
    if (-x.lo > x.hi) return iv(0.0f, mul_ru(x.lo, x.lo));
    return iv(0.0f, mul_ru(x.hi, x.hi));
}
static ival iv_abs(ival x) {                                                              /* :268-276 */
    if (x.lo >= 0.0f) return x;
    if (x.hi < 0.0f) return iv_neg(x);
    return iv(0.0f, cmax(-x.lo, x.hi));
}
static ival iv_sqrt(ival x) {                                                             /* :296-304 */
    if (x.hi < 0.0f) return iv(NAN, NAN);
    if (x.lo <= 0.0f) return iv(0.0f, sqrt_ru(x.hi));
    return iv(sqrt_rd(x.lo), sqrt_ru(x.hi));
}
static ival iv_acos(ival x) {                                                             /* :306-314 */
    if (x.hi < -1.0f || x.lo > 1.0f) return iv(NAN, NAN);
    return iv(near1(acosf, x.hi), near1(acosf, x.lo));
}
static ival iv_asin(ival x) {                                                             /* :316-324 */
    if (x.hi < -1.0f || x.lo > 1.0f) return iv(NAN, NAN);
    return iv(near1(asinf, x.lo), near1(asinf, x.hi));
}
static ival iv_atan(ival x) { return iv(near1(atanf, x.lo), near1(atanf, x.hi)); }        /* :326-330 */
static ival iv_exp(ival x) { return iv(near1(expf, x.lo), near1(expf, x.hi)); }           /* :332-336 */
static ival iv_log(ival x) {                                                              /* :382-391 */
    if (x.hi < 0.0f) return iv(NAN, NAN);
    if (x.lo <= 0.0f) return iv(0.0f, near1(logf, x.hi));
    return iv(near1(logf, x.lo), near1(logf, x.hi));
}

/* ---- tile coordinate helpers ------------------------------------------------------- */
static inline void unpack(int32_t pos, int32_t tps, int* x, int* y, int* z, int* w) {     /* context.cu:24-30 */
    *x = pos % tps; *y = (pos / tps) % tps; *z = (pos / tps) / tps; *w = pos % (tps * tps);
}

/* calculate_intervals_{2d,3d} (context.cu:78-159); all in round-to-nearest
 * except the interval operators themselves. */
static float tile_edge(int p, int tps) {
    fesetround(FE_TONEAREST);
    return (pin((float)p / (float)tps) - 0.5f) * 2.0f;
}
static void tile_intervals(int dim, int32_t position, int tps, const float* mat, float zc, ival out[3]) {
    int x, y, z, w;
    unpack(position, tps, &x, &y, &z, &w);
    const ival ix = iv(tile_edge(x, tps), tile_edge(x + 1, tps));
    const ival iy = iv(tile_edge(y, tps), tile_edge(y + 1, tps));
    if (dim == 3) {
        const ival iz = iv(tile_edge(z, tps), tile_edge(z + 1, tps));
        ival r[4];
        for (int i = 0; i < 4; ++i)   /* mat(i, j) = mat[j * 4 + i] */
            r[i] = iv_addf(iv_add(iv_add(iv_mulf(ix, mat[i]), iv_mulf(iy, mat[4 + i])), iv_mulf(iz, mat[8 + i])), mat[12 + i]);
        out[0] = iv_div(r[0], r[3]); out[1] = iv_div(r[1], r[3]); out[2] = iv_div(r[2], r[3]);
    } else {
        ival r[3];
        for (int i = 0; i < 3; ++i)   /* mat(i, j) = mat[j * 3 + i] */
            r[i] = iv_addf(iv_add(iv_mulf(ix, mat[i]), iv_mulf(iy, mat[3 + i])), mat[6 + i]);
        out[0] = iv_div(r[0], r[2]); out[1] = iv_div(r[1], r[2]); out[2] = iv(zc, zc);
    }
}

static inline int32_t claim_chunk(oracle_ctx* c) {
    return __atomic_fetch_add(&c->tape_index, CHUNK, __ATOMIC_RELAXED);
}
static inline void image_max(int32_t* p, int32_t v) {
    int32_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
}

/* eval_tiles_i for one tile (context.cu:188-459).  Returns clauses evaluated. */
/* Work meter of the *_heatmap variants (context.cu:1622-1633, :1815-1826): a tile spreads the
 * cells it walked (JUMP cells included, as `work++` sits before the switch there) evenly over
 * its footprint: cells / px^2 per pixel.  Kept in integer units of 1/4096 cell so that the sum
 * does not depend on the order of the additions. */
static void heat_tile(oracle_ctx* c, int tps, int32_t position, unsigned cells) {
    if (!c->heat || !cells) return;
    const int px = c->size / tps;
    const uint64_t units = (uint64_t)cells * (uint64_t)(4096 / (px * px));
    const int x0 = (position % tps) * px, y0 = ((position / tps) % tps) * px;
    for (int y = 0; y < px; ++y)
        for (int x = 0; x < px; ++x)
            __atomic_fetch_add(&c->heat[(size_t)(y0 + y) * c->size + x0 + x], units, __ATOMIC_RELAXED);
}
/* cells a forward walk of this tape visits before the end cell, JUMPs included */
static unsigned tape_cells(const uint64_t* tape_data, int32_t tape) {
    const uint64_t* data = &tape_data[tape];
    unsigned n = 0;
    for (;;) {
        const uint64_t d = *++data;
        if (!c_op(d)) return n;
        ++n;
        if (c_op(d) == OP_JUMP) data += c_jump(d);
    }
}

static unsigned eval_tile_interval(oracle_ctx* c, int dim, int32_t* image, int tps, tile_t* tile,
                                   const ival values[3])
{
    uint64_t* const tape_data = c->arena;
    const int64_t cap = c->arena_cells;
    if (tile->position == -1) return 0;                                       /* :206-208 */

    ival slots[256];
    const uint64_t hdr = tape_data[0];
    slots[c_out(hdr)] = values[0];                                            /* :210-213 */
    slots[c_lhs(hdr)] = values[1];
    slots[c_rhs(hdr)] = values[2];

    const uint64_t* data = &tape_data[tile->tape];
    uint32_t choices[256];
    memset(choices, 0, sizeof(choices));
    int choice_index = 0;
    int has_any_choice = 0;
    unsigned work = 0;
    unsigned hcells = 0;                   /* work as the heatmap variant counts it */
    const int32_t hpos = tile->position;

    for (;;) {                                                                /* :223-287 */
        const uint64_t d = *++data;
        const unsigned op = c_op(d);
        if (!op) break;
        ++hcells;
        if (op == OP_JUMP) { data += c_jump(d); continue; }
        ++work;
        const ival lhs = slots[c_lhs(d)], rhs = slots[c_rhs(d)];
        const float imm = c_imm(d);
        ival out;
        int ch = 0, is_choice = 0;
        switch (op) {
            case OP_SQUARE: out = iv_square(lhs); break;
            case OP_SQRT: out = iv_sqrt(lhs); break;
            case OP_NEG: out = iv_neg(lhs); break;
            case OP_SIN: case OP_COS: out = iv(-1.0f, 1.0f); break;          /* gpu_interval.hpp:353 */
            case OP_ASIN: out = iv_asin(lhs); break;
            case OP_ACOS: out = iv_acos(lhs); break;
            case OP_ATAN: out = iv_atan(lhs); break;
            case OP_EXP: out = iv_exp(lhs); break;
            case OP_ABS: out = iv_abs(lhs); break;
            case OP_LOG: out = iv_log(lhs); break;
            case OP_ADD_LI: out = iv_addf(lhs, imm); break;
            case OP_ADD_LR: out = iv_add(lhs, rhs); break;
            case OP_MUL_LI: out = iv_mulf(lhs, imm); break;
            case OP_MUL_LR: out = iv_mul(lhs, rhs); break;
            case OP_MIN_LI: out = iv_min(lhs, iv(imm, imm), &ch); is_choice = 1; break;
            case OP_MIN_LR: out = iv_min(lhs, rhs, &ch); is_choice = 1; break;
            case OP_MAX_LI: out = iv_max(lhs, iv(imm, imm), &ch); is_choice = 1; break;
            case OP_MAX_LR: out = iv_max(lhs, rhs, &ch); is_choice = 1; break;
            case OP_SUB_LI: out = iv_subf(lhs, imm); break;
            case OP_SUB_IR: out = iv_fsub(imm, rhs); break;
            case OP_SUB_LR: out = iv_sub(lhs, rhs); break;
            case OP_DIV_LI: out = iv_divf(lhs, imm); break;
            case OP_DIV_IR: out = iv_div(iv(imm, imm), rhs); break;
            case OP_DIV_LR: out = iv_div(lhs, rhs); break;
            case OP_COPY_IMM: out = iv(imm, imm); break;
            case OP_COPY_LHS: out = lhs; break;
            case OP_COPY_RHS: out = rhs; break;
            default: out = lhs; break;
        }
        if (is_choice) {                                                      /* :254-263 */
            if (choice_index < MAX_CHOICES) choices[choice_index / 16] |= (uint32_t)ch << ((choice_index % 16) * 2);
            choice_index++;
            has_any_choice |= (ch != 0);
        }
        slots[c_out(d)] = out;
    }

    const unsigned i_out = c_out(*data);                                      /* :290 */
    heat_tile(c, tps, hpos, hcells);                                          /* :1622-1633 */
    hcells = 0;
    int x, y, z, w;
    unpack(tile->position, tps, &x, &y, &z, &w);
    if (slots[i_out].lo > 0.0f) { tile->position = -1; return work; }         /* empty :293-296 */
    if (dim == 3 && __atomic_load_n(&image[w], __ATOMIC_RELAXED) > z) {       /* masked :299-305 */
        tile->position = -1; return work;
    }
    if (slots[i_out].hi < 0.0f) {                                             /* filled :308-317 */
        tile->position = -1;
        if (dim == 3) image_max(&image[w], z); else image[w] = 1;
        return work;
    }
    if (!has_any_choice) return work;                                         /* :319-321 */

    /* ---- push (:323-458) ---- */
    uint8_t active[256];
    memset(active, 0, sizeof(active));
    active[i_out] = 1;
    if (__atomic_load_n(&c->tape_index, __ATOMIC_RELAXED) >= cap) return work;
    int32_t out_index = claim_chunk(c);
    int32_t out_offset = CHUNK;
    if ((int64_t)out_index + out_offset >= cap) return work;
    out_offset--;
    tape_data[out_index + out_offset] = *data;

    for (;;) {
        uint64_t d = *--data;
        const unsigned op = c_op(d);
        if (!op) break;
        ++hcells;
        if (op == OP_JUMP) { data += c_jump(d); continue; }
        const int has_choice = op >= OP_MIN_LI && op <= OP_MAX_LR;
        choice_index -= has_choice;
        const unsigned o = c_out(d);
        if (!active[o]) continue;
        const int choice = (has_choice && choice_index < MAX_CHOICES)
            ? (int)((choices[choice_index / 16] >> ((choice_index % 16) * 2)) & 3) : 0;
        --out_offset;
        if (out_offset == 0) {
            const int32_t prev_index = out_index;
            if (__atomic_load_n(&c->tape_index, __ATOMIC_RELAXED) >= cap) { heat_tile(c, tps, hpos, hcells); return work; }
            out_index = claim_chunk(c);
            out_offset = CHUNK;
            if ((int64_t)out_index + out_offset >= cap) { heat_tile(c, tps, hpos, hcells); return work; }
            --out_offset;
            const int32_t delta = prev_index - (out_index + out_offset);
            tape_data[out_index + out_offset] = (uint64_t)OP_JUMP | ((uint64_t)(uint32_t)delta << 32);
            tape_data[prev_index] = (uint64_t)OP_JUMP | ((uint64_t)(uint32_t)(-delta) << 32);
            --out_offset;
        }
        active[o] = 0;
        if (choice == 0) {
            if (c_lhs(d)) active[c_lhs(d)] = 1;
            if (c_rhs(d)) active[c_rhs(d)] = 1;
        } else if (choice == 1) {
            active[c_lhs(d)] = 1;
            if (c_lhs(d) == o) { ++out_offset; continue; }
            d = (d & ~0xffull) | OP_COPY_LHS;
        } else if (choice == 2) {
            if (c_rhs(d)) {
                active[c_rhs(d)] = 1;
                if (c_rhs(d) == o) { ++out_offset; continue; }
                d = (d & ~0xffull) | OP_COPY_RHS;
            } else {
                d = (d & ~0xffull) | OP_COPY_IMM;
            }
        }
        tape_data[out_index + out_offset] = d;
    }
    heat_tile(c, tps, hpos, hcells);                                          /* :1815-1826 */
    out_offset--;
    tape_data[out_index + out_offset] = *data;
    tile->tape = out_index + out_offset;
    return work;
}

/* ---- float stage --------------------------------------------------------------------- */
/* Sample positions with the rounding sequence of the reference build
 * (SASS of calculate_voxels / calculate_pixels / eval_pixels_d for sm_100a):
 *   t = fma(p + 0.5, 1/size, -0.5); f = t + t
 *   a*x + b*y + c*z + d  ->  fma(c, z, fma(a, x, b*y)) + d                          */
static inline float sample_coord(int p, float recip) { float t = fmaf((float)p + 0.5f, recip, -0.5f); return t + t; }
static inline float dot3(float a, float x, float b, float y, float cc, float z, float d) {
    return fmaf(cc, z, fmaf(a, x, b * y)) + d;
}
static inline float dot2(float a, float x, float b, float y, float cc) { return fmaf(a, x, b * y) + cc; }

/* eval_voxels_f clause semantics for one sample (context.cu:874-927) */
static float eval_tape_float(const uint64_t* tape_data, int32_t tape, float X, float Y, float Z, unsigned* work) {
    float slots[256];
    const uint64_t hdr = tape_data[0];
    slots[c_out(hdr)] = X; slots[c_lhs(hdr)] = Y; slots[c_rhs(hdr)] = Z;
    const uint64_t* data = &tape_data[tape];
    for (;;) {
        const uint64_t d = *++data;
        const unsigned op = c_op(d);
        if (!op) break;
        if (op == OP_JUMP) { data += c_jump(d); continue; }
        ++*work;
        const float l = slots[c_lhs(d)], r = slots[c_rhs(d)], imm = c_imm(d);
        float o;
        switch (op) {
            case OP_SQUARE: o = l * l; break;
            case OP_SQRT: o = sqrtf(l); break;
            case OP_NEG: o = -l; break;
            case OP_SIN: o = sinf(l); break;
            case OP_COS: o = cosf(l); break;
            case OP_ASIN: o = asinf(l); break;
            case OP_ACOS: o = acosf(l); break;
            case OP_ATAN: o = atanf(l); break;
            case OP_EXP: o = expf(l); break;
            case OP_ABS: o = fabsf(l); break;
            case OP_LOG: o = logf(l); break;
            case OP_ADD_LI: o = l + imm; break;
            case OP_ADD_LR: o = l + r; break;
            case OP_MUL_LI: o = l * imm; break;
            case OP_MUL_LR: o = l * r; break;
            case OP_MIN_LI: o = cmin(l, imm); break;
            case OP_MIN_LR: o = cmin(l, r); break;
            case OP_MAX_LI: o = cmax(l, imm); break;
            case OP_MAX_LR: o = cmax(l, r); break;
            case OP_SUB_LI: o = l - imm; break;
            case OP_SUB_IR: o = imm - r; break;
            case OP_SUB_LR: o = l - r; break;
            case OP_DIV_LI: o = l / imm; break;
            case OP_DIV_IR: o = imm / r; break;
            case OP_DIV_LR: o = l / r; break;
            case OP_COPY_IMM: o = imm; break;
            case OP_COPY_LHS: o = l; break;
            case OP_COPY_RHS: o = r; break;
            default: o = l; break;
        }
        slots[c_out(d)] = o;
    }
    return slots[c_out(*data)];
}

/* ---- derivative stage (reference inc/gpu_deriv.hpp; contraction per SASS) ------------- */
static inline dval dvm(float v, float dx, float dy, float dz) { dval r = {dx, dy, dz, v}; return r; }
static inline dval dvc(float v) { return dvm(v, 0, 0, 0); }
static dval dv_mul(dval a, dval b) {
    return dvm(a.v * b.v, fmaf(a.dx, b.v, b.dx * a.v), fmaf(a.dy, b.v, b.dy * a.v), fmaf(a.dz, b.v, b.dz * a.v));
}
static dval dv_div(dval a, dval b) {
    const float d = powf(b.v, 2);
    return dvm(a.v / b.v, fmaf(b.v, a.dx, -(a.v * b.dx)) / d, fmaf(b.v, a.dy, -(a.v * b.dy)) / d,
               fmaf(b.v, a.dz, -(a.v * b.dz)) / d);
}
static dval dv_fdiv(float c, dval b) {
    const float d = powf(b.v, 2);
    return dvm(c / b.v, (-c * b.dx) / d, (-c * b.dy) / d, (-c * b.dz) / d);
}

static dval eval_tape_deriv(const uint64_t* tape_data, int32_t tape, dval X, dval Y, dval Z) {
    dval slots[256];
    memset(slots, 0, sizeof(slots));
    const uint64_t hdr = tape_data[0];
    slots[c_out(hdr)] = X; slots[c_lhs(hdr)] = Y; slots[c_rhs(hdr)] = Z;
    const uint64_t* data = &tape_data[tape];
    for (;;) {
        const uint64_t d = *++data;
        const unsigned op = c_op(d);
        if (!op) break;
        if (op == OP_JUMP) { data += c_jump(d); continue; }
        const dval l = slots[c_lhs(d)], r = slots[c_rhs(d)];
        const float imm = c_imm(d);
        dval o;
        switch (op) {   /* context.cu:1081-1114 */
            case OP_SQUARE: o = dv_mul(l, l); break;
            case OP_SQRT: { const float s = sqrtf(l.v), dd = s + s; o = dvm(s, l.dx / dd, l.dy / dd, l.dz / dd); break; }
            case OP_NEG: o = dvm(-l.v, -l.dx, -l.dy, -l.dz); break;
            case OP_SIN: { const float cc = cosf(l.v); o = dvm(sinf(l.v), cc * l.dx, cc * l.dy, cc * l.dz); break; }
            case OP_COS: { const float s = -sinf(l.v); o = dvm(cosf(l.v), s * l.dx, s * l.dy, s * l.dz); break; }
            case OP_ASIN: { const float dd = sqrtf(fmaf(-l.v, l.v, 1.0f)); o = dvm(asinf(l.v), l.dx / dd, l.dy / dd, l.dz / dd); break; }
            case OP_ACOS: { const float dd = -sqrtf(fmaf(-l.v, l.v, 1.0f)); o = dvm(acosf(l.v), l.dx / dd, l.dy / dd, l.dz / dd); break; }
            case OP_ATAN: { const float dd = fmaf(l.v, l.v, 1.0f); o = dvm(atanf(l.v), l.dx / dd, l.dy / dd, l.dz / dd); break; }
            case OP_EXP: { const float e = expf(l.v); o = dvm(e, e * l.dx, e * l.dy, e * l.dz); break; }
            case OP_ABS: o = (l.v < 0.0f) ? dvm(-l.v, -l.dx, -l.dy, -l.dz) : l; break;
            case OP_LOG: o = dvm(logf(l.v), l.dx / l.v, l.dy / l.v, l.dz / l.v); break;
            case OP_ADD_LI: o = dvm(l.v + imm, l.dx, l.dy, l.dz); break;
            case OP_ADD_LR: o = dvm(l.v + r.v, l.dx + r.dx, l.dy + r.dy, l.dz + r.dz); break;
            case OP_MUL_LI: o = dvm(l.v * imm, l.dx * imm, l.dy * imm, l.dz * imm); break;
            case OP_MUL_LR: o = dv_mul(l, r); break;
            case OP_MIN_LI: o = (l.v < imm) ? l : dvc(imm); break;
            case OP_MIN_LR: o = (l.v < r.v) ? l : r; break;
            case OP_MAX_LI: o = (l.v >= imm) ? l : dvc(imm); break;
            case OP_MAX_LR: o = (l.v >= r.v) ? l : r; break;
            case OP_SUB_LI: o = dvm(l.v - imm, l.dx, l.dy, l.dz); break;
            case OP_SUB_IR: o = dvm(imm - r.v, -r.dx, -r.dy, -r.dz); break;
            case OP_SUB_LR: o = dvm(l.v - r.v, l.dx - r.dx, l.dy - r.dy, l.dz - r.dz); break;
            case OP_DIV_LI: o = dvm(l.v / imm, l.dx / imm, l.dy / imm, l.dz / imm); break;
            case OP_DIV_IR: o = dv_fdiv(imm, r); break;
            case OP_DIV_LR: o = dv_div(l, r); break;
            case OP_COPY_IMM: o = dvc(imm); break;
            case OP_COPY_LHS: o = l; break;
            case OP_COPY_RHS: o = r; break;
            default: o = l; break;
        }
        slots[c_out(d)] = o;
    }
    return slots[c_out(*data)];
}

/* ---- frame driver (Context::render2D / render3D, context.cu:1136-1458) ---------------- */
static void ensure_tiles(oracle_ctx* c, int stage, size_t n) {
    if (c->tile_cap[stage] < n) {
        free(c->tiles[stage]);
        c->tiles[stage] = (tile_t*)malloc(sizeof(tile_t) * (n ? n : 1));
        c->tile_cap[stage] = n;
    }
}

static void render(oracle_ctx* c, int dim, const uint64_t* tape, int32_t n_cells, const float* mat, float zc, int threads,
                   int brute)
{
    const int S = c->size;
    (void)threads;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    memcpy(c->arena, tape, sizeof(uint64_t) * (size_t)n_cells);
    c->tape_index = n_cells;                                                  /* :1139, :1285 */
    for (int i = 0; i < 4; ++i) {
        const size_t side = (size_t)S / (64 >> (2 * i));
        memset(c->filled[i], 0, sizeof(int32_t) * side * side);
    }
    memset(c->normals, 0, sizeof(uint32_t) * (size_t)S * S);
    if (c->heat) memset(c->heat, 0, sizeof(uint64_t) * (size_t)S * S);
    uint64_t work_i = 0, work_f = 0;

    const int n_levels = dim == 3 ? 3 : 2;
    const int stage_of[3] = {0, dim == 3 ? 1 : 2, 2};
    const int px_of[3] = {64, dim == 3 ? 16 : 8, 4};
    const int split = dim == 3 ? 4 : 8;

    size_t count = 1;
    if (brute) {                                                              /* render2D_brute :1461-1508 */
        count = (size_t)(S / 8) * (S / 8);
        ensure_tiles(c, 3, count);
        for (size_t i = 0; i < count; ++i) {
            c->tiles[3][i].position = (int32_t)i; c->tiles[3][i].tape = 0; c->tiles[3][i].next = -1;
        }
        c->tile_count[3] = count;
    }
    for (int i = 0; i < dim && !brute; ++i) count *= (size_t)(S / 64);
    if (!brute) ensure_tiles(c, 0, count);
    for (size_t i = 0; i < count && !brute; ++i) {                            /* preload_tiles :45 */
        c->tiles[0][i].position = (int32_t)i; c->tiles[0][i].tape = 0; c->tiles[0][i].next = -1;
    }
    if (!brute) c->tile_count[0] = count;

    for (int l = 0; l < n_levels && !brute; ++l) {
        const int st = stage_of[l];
        const int tps = S / px_of[l];
        tile_t* tiles = c->tiles[st];
        int32_t* image = c->filled[st];
        const long n = (long)count;

        if (dim == 3) {                                                       /* mask_filled_tiles :1335 */
            for (long i = 0; i < n; ++i) {
                if (tiles[i].position == -1) continue;
                int x, y, z, w; unpack(tiles[i].position, tps, &x, &y, &z, &w);
                if (image[w] > z) tiles[i].position = -1;
            }
        }
        #pragma omp parallel for schedule(dynamic, 16) reduction(+:work_i)
        for (long i = 0; i < n; ++i) {
            if (tiles[i].position == -1) continue;
            ival v[3];
            tile_intervals(dim, tiles[i].position, tps, mat, zc, v);
            work_i += eval_tile_interval(c, dim, image, tps, &tiles[i], v);
        }
        /* post-mask + ranks (:1359-1371); serial, so ranks follow list order */
        int32_t active = 0;
        for (long i = 0; i < n; ++i) {
            if (tiles[i].position != -1 && dim == 3) {
                int x, y, z, w; unpack(tiles[i].position, tps, &x, &y, &z, &w);
                if (image[w] > z) tiles[i].position = -1;
            }
            tiles[i].next = (tiles[i].position != -1) ? active++ : -1;
        }
        const int last = (l == n_levels - 1);
        const int nst = last ? 3 : stage_of[l + 1];
        const size_t next_count = last ? (size_t)active : (size_t)active * 64;
        ensure_tiles(c, nst, next_count);
        tile_t* out = c->tiles[nst];
        if (!last) {                                                          /* subdivide :564-631 */
            const int ntps = tps * split;
            for (long i = 0; i < n; ++i) {
                if (tiles[i].next == -1) continue;
                int x, y, z, w; unpack(tiles[i].position, tps, &x, &y, &z, &w);
                for (int sub = 0; sub < 64; ++sub) {
                    int sx, sy, sz = 0;
                    if (dim == 3) { sx = x * 4 + sub % 4; sy = y * 4 + (sub / 4) % 4; sz = z * 4 + sub / 16; }
                    else { sx = x * 8 + sub % 8; sy = y * 8 + sub / 8; }
                    tile_t* t = &out[(size_t)tiles[i].next * 64 + sub];
                    t->position = sx + sy * ntps + sz * ntps * ntps;
                    t->tape = tiles[i].tape;
                    t->next = -1;
                }
            }
        } else {                                                              /* copy_active_tiles :637-651 */
            for (long i = 0; i < n; ++i) {
                if (tiles[i].next == -1) continue;
                tile_t* t = &out[tiles[i].next];
                t->position = tiles[i].position; t->tape = tiles[i].tape; t->next = -1;
                tiles[i].next = -1;
            }
        }
        c->tile_count[nst] = next_count;
        {                                                                     /* copy_filled :664-692 */
            const int nsize = last ? S : S / px_of[l + 1];
            int32_t* nimg = c->filled[nst];
            for (int y = 0; y < nsize; ++y)
                for (int x = 0; x < nsize; ++x) {
                    const int32_t t = image[x / split + (y / split) * (nsize / split)];
                    if (t) nimg[x + y * nsize] = dim == 3 ? t * 4 + 3 : 1;
                }
        }
        count = next_count;
    }

    /* float stage (calculate_voxels/pixels + eval_voxels_f, :707-964) */
    {
        fesetround(FE_TONEAREST);
        const tile_t* tiles = c->tiles[3];
        int32_t* image = c->filled[3];
        const int tps = S / split;
        const float recip = 1.0f / (float)(unsigned)(tps * split);
        const long n = (long)count;
        #pragma omp parallel for schedule(dynamic, 64) reduction(+:work_f)
        for (long i = 0; i < n; ++i) {
            fesetround(FE_TONEAREST);
            int tx, ty, tz, tw; unpack(tiles[i].position, tps, &tx, &ty, &tz, &tw);
            unsigned work = 0;
            const uint64_t hcells = c->heat ? tape_cells(c->arena, tiles[i].tape) : 0;
            for (int lane = 0; lane < 32; ++lane) {
                if (dim == 3) {
                    const int px = tx * 4 + lane % 4, py = ty * 4 + (lane / 4) % 4, pz = tz * 4 + lane / 16;
                    int32_t* pix = &image[px + py * S];
                    if (__atomic_load_n(pix, __ATOMIC_RELAXED) >= pz + 2) continue;      /* :861 */
                    if (c->heat)                                                           /* :1962 */
                        __atomic_fetch_add(&c->heat[px + (size_t)py * S], hcells * 4096u, __ATOMIC_RELAXED);
                    const float fx = sample_coord(px, recip), fy = sample_coord(py, recip);
                    float val[2];
                    for (int k = 0; k < 2; ++k) {
                        const float fz = sample_coord(pz + 2 * k, recip);
                        const float w = dot3(mat[3], fx, mat[7], fy, mat[11], fz, mat[15]);
                        val[k] = eval_tape_float(c->arena, tiles[i].tape,
                            dot3(mat[0], fx, mat[4], fy, mat[8], fz, mat[12]) / w,
                            dot3(mat[1], fx, mat[5], fy, mat[9], fz, mat[13]) / w,
                            dot3(mat[2], fx, mat[6], fy, mat[10], fz, mat[14]) / w, &work);
                    }
                    if (val[1] < 0.0f) image_max(pix, pz + 2);                              /* :936-948 */
                    else if (val[0] < 0.0f) image_max(pix, pz);
                } else {
                    const int px = tx * 8 + lane % 8, py = ty * 8 + lane / 8;
                    const float fx = sample_coord(px, recip);
                    if (c->heat) {                                                         /* :1979-1980 */
                        c->heat[px + (size_t)py * S] += hcells * 2048u;
                        c->heat[px + (size_t)(py + 4) * S] += hcells * 2048u;
                    }
                    for (int k = 0; k < 2; ++k) {
                        const float fy = sample_coord(py + 4 * k, recip);
                        const float w = dot2(mat[2], fx, mat[5], fy, mat[8]);
                        const float v = eval_tape_float(c->arena, tiles[i].tape,
                            dot2(mat[0], fx, mat[3], fy, mat[6]) / w,
                            dot2(mat[1], fx, mat[4], fy, mat[7]) / w, zc, &work);
                        if (v < 0.0f) image[px + (py + 4 * k) * S] = 1;                      /* :951-962 */
                    }
                }
            }
            work_f += work;
        }
    }

    /* normals (eval_pixels_d, :978-1132) */
    if (dim == 3) {
        const int32_t* image = c->filled[3];
        const tile_t *t0 = c->tiles[0], *t1 = c->tiles[1], *t2 = c->tiles[2];
        const float recip = 1.0f / (float)(unsigned)S;
        const int n0 = S / 64;
        #pragma omp parallel for schedule(dynamic, 8)
        for (int py = 0; py < S; ++py) {
            fesetround(FE_TONEAREST);
            for (int px = 0; px < S; ++px) {
                int pz = image[px + py * S];
                if (pz == 0) continue;
                if (pz < S - 1) pz += 1;
                const float fx = sample_coord(px, recip), fy = sample_coord(py, recip), fz = sample_coord(pz, recip);
                const float w = dot3(mat[3], fx, mat[7], fy, mat[11], fz, mat[15]);
                const dval X = dvm(dot3(mat[0], fx, mat[4], fy, mat[8], fz, mat[12]) / w, 1, 0, 0);
                const dval Y = dvm(dot3(mat[1], fx, mat[5], fy, mat[9], fz, mat[13]) / w, 0, 1, 0);
                const dval Z = dvm(dot3(mat[2], fx, mat[6], fy, mat[10], fz, mat[14]) / w, 0, 0, 1);
                int32_t tape;
                const tile_t a = t0[px / 64 + (py / 64) * n0 + (pz / 64) * n0 * n0];
                if (a.next == -1) tape = a.tape;
                else {
                    const tile_t b = t1[a.next * 64 + (px % 64) / 16 + ((py % 64) / 16) * 4 + ((pz % 64) / 16) * 16];
                    if (b.next == -1) tape = b.tape;
                    else tape = t2[b.next * 64 + (px % 16) / 4 + ((py % 16) / 4) * 4 + ((pz % 16) / 4) * 16].tape;
                }
                const dval r = eval_tape_deriv(c->arena, tape, X, Y, Z);
                const float norm = sqrtf((powf(r.dx, 2) + powf(r.dy, 2)) + powf(r.dz, 2));
                const uint8_t bx = (uint8_t)(uint32_t)fmaf(r.dx / norm, 127.0f, 128.0f);
                const uint8_t by = (uint8_t)(uint32_t)fmaf(r.dy / norm, 127.0f, 128.0f);
                const uint8_t bz = (uint8_t)(uint32_t)fmaf(r.dz / norm, 127.0f, 128.0f);
                c->normals[px + py * S] = (0xFFu << 24) | ((uint32_t)bz << 16) | ((uint32_t)by << 8) | bx;
            }
        }
    }
    fesetround(FE_TONEAREST);
    c->work_interval = work_i;
    c->work_float = work_f;
}

/* ---- exported API ---------------------------------------------------------------------- */
oracle_ctx* mpro_create(int size, int64_t num_subtapes) {
    oracle_ctx* c = (oracle_ctx*)calloc(1, sizeof(oracle_ctx));
    c->size = size;
    for (int i = 0; i < 4; ++i) {
        const size_t side = (size_t)size / (64 >> (2 * i));
        c->filled[i] = (int32_t*)calloc(side * side, sizeof(int32_t));
    }
    c->normals = (uint32_t*)calloc((size_t)size * size, sizeof(uint32_t));
    c->arena_cells = (num_subtapes > 0 ? num_subtapes : 640000) * CHUNK;
    c->arena = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)c->arena_cells + CHUNK));
    return c;
}
void mpro_destroy(oracle_ctx* c) {
    if (!c) return;
    for (int i = 0; i < 4; ++i) { free(c->filled[i]); free(c->tiles[i]); }
    free(c->normals); free(c->arena); free(c->heat); free(c);
}
void mpro_render2d(oracle_ctx* c, const uint64_t* tape, int32_t n, const float* mat3, float z, int threads) {
    render(c, 2, tape, n, mat3, z, threads, 0);
}
void mpro_render3d(oracle_ctx* c, const uint64_t* tape, int32_t n, const float* mat4, int threads) {
    render(c, 3, tape, n, mat4, 0.0f, threads, 0);
}
void mpro_render2d_brute(oracle_ctx* c, const uint64_t* tape, int32_t n, const float* mat3, float z, int threads) {
    render(c, 2, tape, n, mat3, z, threads, 1);
}
int32_t* mpro_filled(oracle_ctx* c, int stage) { return c->filled[stage]; }
tile_t* mpro_tiles(oracle_ctx* c, int stage) { return c->tiles[stage]; }
uint64_t mpro_tile_count(oracle_ctx* c, int stage) { return c->tile_count[stage]; }
uint64_t* mpro_arena(oracle_ctx* c) { return c->arena; }
int32_t mpro_tape_index(oracle_ctx* c) { return c->tape_index; }
uint32_t* mpro_normals(oracle_ctx* c) { return c->normals; }
/* Switches the work meter on (context.cu:1984-2340, render2D_heatmap / render3D_heatmap): after
 * the next render mpro_heat() holds, per pixel, 4096 x the amortised cells walked; dividing by
 * 4096 * (tape length - 2) gives the reference's heatmap value. */
void mpro_set_heat(oracle_ctx* c, int on) {
    if (on && !c->heat) c->heat = (uint64_t*)calloc((size_t)c->size * c->size, sizeof(uint64_t));
    if (!on && c->heat) { free(c->heat); c->heat = NULL; }
}
uint64_t* mpro_heat(oracle_ctx* c) { return c->heat; }
uint64_t mpro_work_interval(oracle_ctx* c) { return c->work_interval; }
uint64_t mpro_work_float(oracle_ctx* c) { return c->work_float; }
int mpro_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Logical tape = the non-JUMP cells a reader sees from header to end cell
 * (reader protocol: context.cu:223-229, benchmark/tape_shortening.cpp:65-72).
 * Arena addresses differ run to run, so tapes are compared through this hash. */
static uint64_t mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    return h ^ (h >> 33);
}
uint64_t mpro_tape_hash(const uint64_t* arena, int32_t start, int32_t* len_out) {
    uint64_t h = mix(0x1234567ull, arena[start]);
    int32_t len = 0;
    const uint64_t* data = &arena[start];
    for (;;) {
        const uint64_t d = *++data;
        if (c_op(d) == OP_JUMP) { data += c_jump(d); continue; }
        h = mix(h, d);
        if (!c_op(d)) break;
        ++len;
    }
    if (len_out) *len_out = len;
    return h;
}
void mpro_tape_hashes(const uint64_t* arena, const int32_t* starts, int64_t n, uint64_t* hashes, int32_t* lens) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) hashes[i] = mpro_tape_hash(arena, starts[i], lens ? &lens[i] : 0);
}
/* Copies one logical tape (header, clauses, end cell) into out; returns cells written. */
int32_t mpro_tape_flatten(const uint64_t* arena, int32_t start, uint64_t* out, int32_t out_cap) {
    int32_t n = 0;
    if (n < out_cap) out[n] = arena[start];
    ++n;
    const uint64_t* data = &arena[start];
    for (;;) {
        const uint64_t d = *++data;
        if (c_op(d) == OP_JUMP) { data += c_jump(d); continue; }
        if (n < out_cap) out[n] = d;
        ++n;
        if (!c_op(d)) break;
    }
    return n;
}

/* ---- unit-test hooks: single interval operators ------------------------------------------ */
/* op codes follow the clause opcodes; unary ops ignore b; *_LI / *_IR forms take the immediate
 * in b[0].  out = {lo, hi}; returns the min/max verdict (0 for other ops). */
int mpro_interval_op(int op, const float* a, const float* b, float* out) {
    const ival x = iv(a[0], a[1]);
    const ival y = iv(b[0], b[1]);
    const float imm = b[0];
    ival r = x;
    int ch = 0;
    switch (op) {
        case OP_SQUARE: r = iv_square(x); break;
        case OP_SQRT: r = iv_sqrt(x); break;
        case OP_NEG: r = iv_neg(x); break;
        case OP_SIN: case OP_COS: r = iv(-1.0f, 1.0f); break;
        case OP_ASIN: r = iv_asin(x); break;
        case OP_ACOS: r = iv_acos(x); break;
        case OP_ATAN: r = iv_atan(x); break;
        case OP_EXP: r = iv_exp(x); break;
        case OP_ABS: r = iv_abs(x); break;
        case OP_LOG: r = iv_log(x); break;
        case OP_ADD_LI: r = iv_addf(x, imm); break;
        case OP_ADD_LR: r = iv_add(x, y); break;
        case OP_MUL_LI: r = iv_mulf(x, imm); break;
        case OP_MUL_LR: r = iv_mul(x, y); break;
        case OP_MIN_LI: r = iv_min(x, iv(imm, imm), &ch); break;
        case OP_MIN_LR: r = iv_min(x, y, &ch); break;
        case OP_MAX_LI: r = iv_max(x, iv(imm, imm), &ch); break;
        case OP_MAX_LR: r = iv_max(x, y, &ch); break;
        case OP_SUB_LI: r = iv_subf(x, imm); break;
        case OP_SUB_IR: r = iv_fsub(imm, x); break;
        case OP_SUB_LR: r = iv_sub(x, y); break;
        case OP_DIV_LI: r = iv_divf(x, imm); break;
        case OP_DIV_IR: r = iv_div(iv(imm, imm), x); break;
        case OP_DIV_LR: r = iv_div(x, y); break;
        default: break;
    }
    fesetround(FE_TONEAREST);
    out[0] = r.lo; out[1] = r.hi;
    return ch;
}

/* Evaluates a whole tape at n points in float (the float-stage clause semantics). */
void mpro_eval_points(const uint64_t* tape, const float* xyz, int64_t n, float* out) {
    fesetround(FE_TONEAREST);
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        unsigned work = 0;
        out[i] = eval_tape_float(tape, 0, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &work);
    }
}

/* ------------------------------------------------------------------------------------------
 * Post-effects (reference src/effects.cu).  PARITY UNPINNED: effects.cu needs Eigen in device
 * code and cannot be compiled here, so these follow the source text only; floating point is
 * evaluated operation by operation in round-to-nearest without contraction.
 * ------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } fx_v3;
static fx_v3 fx_mk(float x, float y, float z) { fx_v3 v = {x, y, z}; return v; }
static float fx_dot(fx_v3 a, fx_v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static fx_v3 fx_norm(fx_v3 a) {                    /* Eigen normalized(): untouched when the norm is 0 */
    float z = fx_dot(a, a);
    if (z > 0.0f) { float n = sqrtf(z); return fx_mk(a.x / n, a.y / n, a.z / n); }
    return a;
}
static float fx_ndc(float p, int size) { return 2.0f * ((p + 0.5f) / (float)size - 0.5f); }
/* CUDA float -> unsigned conversion: saturating, NaN -> 0 */
static unsigned fx_f2u(float f) {
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (unsigned)f;
}
static int32_t fx_f2i(float f) {                   /* CUDA float -> int: saturating, NaN -> 0 */
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 0x7fffffff;
    if (f <= -2147483648.0f) return (int32_t)0x80000000;
    return (int32_t)f;
}
static fx_v3 fx_unpack_normal(uint32_t n) {
    return fx_norm(fx_mk((float)(n & 0xFF) - 128.0f, (float)((n >> 8) & 0xFF) - 128.0f,
                         (float)((n >> 16) & 0xFF) - 128.0f));
}

/* draw_ssao, effects.cu:17-89.  kernel 64x3 / rvecs 256x3 column-major. */
void mpro_fx_draw_ssao(const int32_t* depth, const uint32_t* norm, const float* kernel, const float* rvecs,
                       int size, int32_t* output) {
    const float RADIUS = 0.1f;
    #pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < size; ++y)
        for (int x = 0; x < size; ++x) {
            const int h = depth[x + y * size];
            if (!h) continue;
            const fx_v3 pos = fx_mk(fx_ndc((float)x, size), fx_ndc((float)y, size), fx_ndc((float)h, size));
            const fx_v3 normal = fx_unpack_normal(norm[x + y * size]);
            const int ri = (x % 16) * 16 + (y % 16);          /* threadIdx % 16 with 16x16 blocks */
            const fx_v3 rvec = fx_mk(rvecs[ri], rvecs[256 + ri], rvecs[512 + ri]);
            const float rn = fx_dot(rvec, normal);
            const fx_v3 tangent = fx_norm(fx_mk(rvec.x - normal.x * rn, rvec.y - normal.y * rn, rvec.z - normal.z * rn));
            const fx_v3 bitangent = fx_mk(normal.y * tangent.z - normal.z * tangent.y,
                                          normal.z * tangent.x - normal.x * tangent.z,
                                          normal.x * tangent.y - normal.y * tangent.x);
            float occlusion = 0.0f;
            for (int i = 0; i < 64; ++i) {
                const fx_v3 k = fx_mk(kernel[i], kernel[64 + i], kernel[128 + i]);
                const fx_v3 r = fx_mk(tangent.x * k.x + bitangent.x * k.y + normal.x * k.z,
                                      tangent.y * k.x + bitangent.y * k.y + normal.y * k.z,
                                      tangent.z * k.x + bitangent.z * k.y + normal.z * k.z);
                const fx_v3 sp = fx_mk(r.x * RADIUS + pos.x, r.y * RADIUS + pos.y, r.z * RADIUS + pos.z);
                const unsigned px = fx_f2u((sp.x / 2.0f + 0.5f) * (float)size);
                const unsigned py = fx_f2u((sp.y / 2.0f + 0.5f) * (float)size);
                const unsigned actual_h = (px < (unsigned)size && py < (unsigned)size) ? (unsigned)depth[px + py * size] : 0u;
                const float actual_z = 2.0f * (((float)actual_h + 0.5f) / (float)size - 0.5f);
                const float dz = fabsf(sp.z - actual_z);
                if (dz < RADIUS) {
                    occlusion += (sp.z <= actual_z) ? 1.0f : 0.0f;
                } else if (dz < RADIUS * 2.0f) {
                    if (sp.z <= actual_z) {
                        const float t = (RADIUS - (dz - RADIUS)) / RADIUS;
                        occlusion += t * t;                    /* powf(t, 2.0f) */
                    }
                }
            }
            occlusion = (float)(1.0 - (double)(occlusion / 64.0f));
            output[x + y * size] = (uint8_t)fx_f2u(occlusion * 255.0f);
        }
}

/* blur_ssao, effects.cu:93-155 (including the origin-relative second loop). */
void mpro_fx_blur_ssao(const int32_t* image, const int32_t* ssao, int size, int32_t* output) {
    #pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < size; ++y)
        for (int x = 0; x < size; ++x) {
            float best = 1000000.0f, value = 0.0f;
            for (unsigned q = 0; q < 4; ++q) {
                const int xmin = (q & 1) ? 0 : -2, ymin = (q & 2) ? 0 : -2;
                float sum = 0.0f, count = 0.0f;
                for (int i = 0; i <= 2; ++i)
                    for (int j = 0; j <= 2; ++j) {
                        const int tx = x + xmin + i, ty = y + ymin + j;
                        if (tx >= 0 && tx < size && ty >= 0 && ty < size && image[tx + ty * size]) {
                            sum += (float)ssao[tx + ty * size];
                            count += 1.0f;
                        }
                    }
                const float mean = sum / count;
                float stdev = 0.0f;
                for (int i = 0; i <= 2; ++i)
                    for (int j = 0; j <= 2; ++j) {
                        const int tx = xmin + i, ty = ymin + j;
                        if (tx >= 0 && tx < size && ty >= 0 && ty < size && image[tx + ty * size]) {
                            const float d = mean - (float)ssao[tx + ty * size];
                            stdev += d * d;
                        }
                    }
                stdev /= count - 1.0f;
                stdev = sqrtf(stdev);
                if (stdev < best) { best = stdev; value = mean; }
            }
            output[x + y * size] = fx_f2i(value);
        }
}

/* draw_shaded, effects.cu:159-225. */
void mpro_fx_draw_shaded(const int32_t* depth, const uint32_t* norm, const int32_t* ssao, int size, int32_t* output) {
    #pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < size; ++y)
        for (int x = 0; x < size; ++x) {
            const int h = depth[x + y * size];
            if (!h) continue;
            const uint8_t s = (uint8_t)ssao[x + y * size];
            const fx_v3 normal = fx_unpack_normal(norm[x + y * size]);
            const fx_v3 pos = fx_mk(fx_ndc((float)x, size), fx_ndc((float)y, size), fx_ndc((float)h, size));
            const fx_v3 ld = fx_norm(fx_mk(5.0f - pos.x, 5.0f - pos.y, 10.0f - pos.z));
            float light = fmaxf(0.0f, fx_dot(ld, normal)) * 0.8f;
            light *= (float)s / 255.0f;
            light += 0.2f;
            if (light < 0.0f) light = 0.0f; else if (light > 1.0f) light = 1.0f;
            const uint32_t color = (uint8_t)fx_f2u(light * 255.0f);
            output[x + y * size] = (int32_t)((0xFFu << 24) | (color << 16) | (color << 8) | color);
        }
}

/* Effects::drawSSAO (effects.cu:253-275) and Effects::drawShaded (effects.cu:277-297). */
void mpro_fx_ssao(const int32_t* depth, const uint32_t* norm, const float* kernel, const float* rvecs, int size,
                  int32_t* tmp, int32_t* image) {
    memset(tmp, 0, sizeof(int32_t) * (size_t)size * size);
    memset(image, 0, sizeof(int32_t) * (size_t)size * size);
    mpro_fx_draw_ssao(depth, norm, kernel, rvecs, size, tmp);
    mpro_fx_blur_ssao(depth, tmp, size, image);
}
void mpro_fx_shaded(const int32_t* depth, const uint32_t* norm, const float* kernel, const float* rvecs, int size,
                    int32_t* tmp, int32_t* image) {
    memset(tmp, 0, sizeof(int32_t) * (size_t)size * size);
    memset(image, 0, sizeof(int32_t) * (size_t)size * size);
    mpro_fx_draw_ssao(depth, norm, kernel, rvecs, size, image);
    mpro_fx_blur_ssao(depth, image, size, tmp);
    mpro_fx_draw_shaded(depth, norm, tmp, size, image);
}
