// Minimal libfive-compatible expression DAG (host only).
//
// libfive itself cannot be built in this environment (it needs Eigen, Boost
// and libpng), yet mpr::Tape's constructor takes a `const libfive::Tree&`
// (reference inc/tape.hpp:25) and every benchmark driver builds or loads one.
// This header provides the slice of libfive::Tree that src/tape.cpp and
// benchmark/*.cpp rely on (include/libfive/tree/tree.hpp:32-248): hash-consed
// immutable nodes, X/Y/Z, float constants, arithmetic operators, remap(),
// orderedDfs(), and ->op / ->lhs / ->rhs / ->value / ->rank access.
//
// Hash-consing follows libfive's Cache (src/tree/cache.cpp:74-149): constants
// are unique per value, operations per (op, lhs, rhs); operations whose
// operands are all constant are folded at construction; identity and
// commutative re-balancing rules (cache.cpp:323-470) are applied.  The affine
// collection pass (cache.cpp:472-504) is NOT implemented.
#pragma once
#include <cstdint>
#include <iosfwd>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "libfive/tree/opcode.hpp"

namespace libfive {

class Tree {
public:
    struct Tree_ {
        Opcode::Opcode op;
        uint8_t flags;
        unsigned rank;
        float value;
        std::shared_ptr<Tree_> lhs;
        std::shared_ptr<Tree_> rhs;
        uint64_t serial = 0;        // creation order; stands in for libfive's pointer-ordered maps
        ~Tree_();
    };
    typedef const Tree_* Id;
    enum Flags { FLAG_LOCATION_AGNOSTIC = (1 << 1) };

    Tree(float v);
    Tree(double v) : Tree(static_cast<float>(v)) {}
    Tree(int v) : Tree(static_cast<float>(v)) {}
    explicit Tree(Opcode::Opcode op, Tree a = Tree(), Tree b = Tree());

    static Tree X() { return Tree(Opcode::VAR_X); }
    static Tree Y() { return Tree(Opcode::VAR_Y); }
    static Tree Z() { return Tree(Opcode::VAR_Z); }
    static Tree Invalid() { return Tree(); }

    const std::shared_ptr<Tree_>& operator->() const { return ptr; }
    bool operator==(const Tree& o) const { return ptr.get() == o.ptr.get(); }
    Id id() const { return ptr.get(); }
    Tree lhs() const { return Tree(ptr->lhs); }
    Tree rhs() const { return Tree(ptr->rhs); }
    Tree operator-() const;

    // Substitutes the three axes (tree.cpp "remap"); everything else is
    // rebuilt bottom-up through the cache, so simplification re-applies.
    Tree remap(Tree X, Tree Y, Tree Z) const;

    // Children-before-parents order used by the tape packer
    // (libfive/libfive/src/tree/tree.cpp:146-187): a depth-first walk that
    // emits a node once all of its users have been emitted, then reversed.
    std::vector<Tree> orderedDfs() const;

    // .frep loader (archive with exactly one shape).
    static Tree deserialize(std::istream& in);
    static Tree load(const std::string& filename);

    explicit Tree(std::shared_ptr<Tree_> t) : ptr(std::move(t)) {}

protected:
    Tree() {}
    std::shared_ptr<Tree_> ptr;
    friend class Cache;
};

// Stand-in for libfive::Cache::instance(), which tape.cpp holds as a lock.
class Cache {
public:
    struct Handle {};
    static Handle instance() { return Handle(); }
    // Controls whether identity/commutative simplification runs (default on).
    static void setSimplify(bool on);
};

}  // namespace libfive

#define MPRB_TREE_UNARY(F) libfive::Tree F(const libfive::Tree& a)
MPRB_TREE_UNARY(square); MPRB_TREE_UNARY(sqrt); MPRB_TREE_UNARY(abs);
MPRB_TREE_UNARY(sin); MPRB_TREE_UNARY(cos); MPRB_TREE_UNARY(tan);
MPRB_TREE_UNARY(asin); MPRB_TREE_UNARY(acos); MPRB_TREE_UNARY(atan);
MPRB_TREE_UNARY(log); MPRB_TREE_UNARY(exp);
#undef MPRB_TREE_UNARY
#define MPRB_TREE_BINARY(F) libfive::Tree F(const libfive::Tree& a, const libfive::Tree& b)
MPRB_TREE_BINARY(operator+); MPRB_TREE_BINARY(operator*); MPRB_TREE_BINARY(min);
MPRB_TREE_BINARY(max); MPRB_TREE_BINARY(operator-); MPRB_TREE_BINARY(operator/);
MPRB_TREE_BINARY(atan2); MPRB_TREE_BINARY(pow); MPRB_TREE_BINARY(nth_root);
MPRB_TREE_BINARY(mod); MPRB_TREE_BINARY(nanfill); MPRB_TREE_BINARY(compare);
#undef MPRB_TREE_BINARY
