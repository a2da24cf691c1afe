// Tree -> packed clause tape (host only).
//
// Produces the exact cell sequence the reference's Tape constructor produces
// for the same Tree (reference src/tape.cpp:21-228):
//   cell 0      header    {op 0, bytes 1..3 = slots bound to X, Y, Z (0 = unused)}
//   cells 1..n  clauses   in orderedDfs order, 8 bytes each (inc/clause.hpp:18-23)
//   cell n+1    end       {op 0, byte 1 = slot holding the result}
// Slots come from a LIFO free list; an operand whose last use is the current
// clause is released *before* the output slot is chosen, so a clause may write
// the slot it reads (tape.cpp:199-212).  Slot 0 is never handed out.
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

#include "libfive/tree/tree.hpp"
#include "mprb_host.hpp"

namespace mprb {

namespace {

enum GpuOp : uint8_t {  // mirrors mpr::Opcode (inc/gpu_opcode.hpp:18-56)
    G_INVALID = 0, G_JUMP, G_SQUARE, G_SQRT, G_NEG, G_SIN, G_COS, G_ASIN, G_ACOS, G_ATAN,
    G_EXP, G_ABS, G_LOG, G_ADD_LI, G_ADD_LR, G_MUL_LI, G_MUL_LR, G_MIN_LI, G_MIN_LR,
    G_MAX_LI, G_MAX_LR, G_SUB_LI, G_SUB_IR, G_SUB_LR, G_DIV_LI, G_DIV_IR, G_DIV_LR,
    G_COPY_IMM, G_COPY_LHS, G_COPY_RHS,
};

inline uint64_t cell(uint8_t op, uint8_t out, uint8_t lhs, uint8_t rhs, float imm) {
    uint32_t lo = uint32_t(op) | (uint32_t(out) << 8) | (uint32_t(lhs) << 16) |
                  (uint32_t(rhs) << 24);
    uint32_t hi;
    memcpy(&hi, &imm, 4);
    return uint64_t(lo) | (uint64_t(hi) << 32);
}

// kind: 1 = unary, 2 = commutative binary, 3 = non-commutative binary, 0 = no clause
struct OpMap { int kind; uint8_t base; };
OpMap classify(libfive::Opcode::Opcode op) {
    using namespace libfive::Opcode;
    switch (op) {
        case OP_SQUARE: return {1, G_SQUARE};
        case OP_SQRT: return {1, G_SQRT};
        case OP_NEG: return {1, G_NEG};
        case OP_SIN: return {1, G_SIN};
        case OP_COS: return {1, G_COS};
        case OP_ASIN: return {1, G_ASIN};
        case OP_ACOS: return {1, G_ACOS};
        case OP_ATAN: return {1, G_ATAN};
        case OP_EXP: return {1, G_EXP};
        case OP_ABS: return {1, G_ABS};
        case OP_LOG: return {1, G_LOG};
        case OP_ADD: return {2, G_ADD_LI};
        case OP_MUL: return {2, G_MUL_LI};
        case OP_MIN: return {2, G_MIN_LI};
        case OP_MAX: return {2, G_MAX_LI};
        case OP_SUB: return {3, G_SUB_LI};
        case OP_DIV: return {3, G_DIV_LI};
        default: return {0, 0};
    }
}

}  // namespace

std::vector<uint64_t> pack_tape(const libfive::Tree& tree, int* num_slots_out) {
    typedef libfive::Tree::Id Id;
    const auto ordered = tree.orderedDfs();

    // Pass 1: last use of every operand, axes present, nodes that get a clause.
    std::map<Id, Id> last_used;
    Id axes[3] = {nullptr, nullptr, nullptr};
    std::vector<Id> clauses;
    clauses.reserve(ordered.size());
    for (const auto& c : ordered) {
        using namespace libfive::Opcode;
        if (c->op == VAR_X) axes[0] = c.id();
        else if (c->op == VAR_Y) axes[1] = c.id();
        else if (c->op == VAR_Z) axes[2] = c.id();
        const OpMap m = classify(c->op);
        if (!m.kind) continue;
        if (m.kind >= 2) last_used[c->rhs.get()] = c.id();
        last_used[c->lhs.get()] = c.id();
        clauses.push_back(c.id());
    }

    // Pass 2: slot assignment and clause emission.
    std::vector<uint8_t> free_slots;
    std::map<Id, uint8_t> bound;
    unsigned num_slots = 1;  // slot 0 means "no operand"
    auto take = [&](Id id) -> uint8_t {
        uint8_t s = 0;
        if (!free_slots.empty()) {
            s = free_slots.back();
            free_slots.pop_back();
        } else if (num_slots == 255) {
            fprintf(stderr, "Ran out of slots!\n");
        } else {
            s = uint8_t(num_slots++);
        }
        bound[id] = s;
        return s;
    };
    auto reg = [&](Id id) -> uint8_t {
        auto itr = bound.find(id);
        if (itr == bound.end()) {
            fprintf(stderr, "Could not find bound slots %i\n", int(id->op));
            return 0;
        }
        return itr->second;
    };

    std::vector<uint64_t> flat;
    flat.reserve(clauses.size() + 2);
    {
        uint8_t ax[3] = {0, 0, 0};
        for (int i = 0; i < 3; ++i) {
            if (axes[i]) ax[i] = take(axes[i]);
        }
        flat.push_back(cell(0, ax[0], ax[1], ax[2], 0.0f));
    }

    for (Id c : clauses) {
        const OpMap m = classify(c->op);
        const Id l = c->lhs.get();
        const Id r = c->rhs.get();
        uint8_t op = 0, i_lhs = 0, i_rhs = 0;
        float imm = 0.0f;
        if (m.kind == 1) {
            op = m.base;
            i_lhs = reg(l);
        } else {
            const bool lconst = l->op == libfive::Opcode::CONSTANT;
            const bool rconst = r->op == libfive::Opcode::CONSTANT;
            if (m.kind == 2) {  // base = *_LHS_IMM, base+1 = *_LHS_RHS
                if (lconst) { op = m.base; i_lhs = reg(r); imm = l->value; }
                else if (rconst) { op = m.base; i_lhs = reg(l); imm = r->value; }
                else { op = uint8_t(m.base + 1); i_lhs = reg(l); i_rhs = reg(r); }
            } else {            // base = *_LHS_IMM, +1 = *_IMM_RHS, +2 = *_LHS_RHS
                if (lconst) { op = uint8_t(m.base + 1); i_rhs = reg(r); imm = l->value; }
                else if (rconst) { op = m.base; i_lhs = reg(l); imm = r->value; }
                else { op = uint8_t(m.base + 2); i_lhs = reg(l); i_rhs = reg(r); }
            }
        }
        // Release operands whose last reader is this clause (lhs first, then
        // rhs, matching the free-list order of the reference).
        const Id ops[2] = {l, r};
        for (Id h : ops) {
            if (h && h->op != libfive::Opcode::CONSTANT && last_used[h] == c) {
                auto itr = bound.find(h);
                if (itr != bound.end()) {  // guards lhs == rhs
                    free_slots.push_back(itr->second);
                    bound.erase(itr);
                }
            }
        }
        const uint8_t out = take(c);
        flat.push_back(cell(op, out, i_lhs, i_rhs, imm));
    }

    flat.push_back(cell(0, reg(ordered.back().id()), 0, 0, 0.0f));
    if (num_slots_out) *num_slots_out = int(num_slots);
    return flat;
}

}  // namespace mprb
