// Host-side expression DAG + .frep reader (see shim/libfive/tree/tree.hpp).
//
// Behavioural references (libfive submodule of the reference repo):
//   hash-consing / folding ........ libfive/libfive/src/tree/cache.cpp:38-149
//   identity rules ................ cache.cpp:323-431
//   commutative re-balancing ...... cache.cpp:433-470
//   orderedDfs .................... libfive/libfive/src/tree/tree.cpp:146-187
//   remap ......................... tree.cpp:189-235
//   archive format ................ libfive/libfive/src/tree/deserializer.cpp:38-199
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <istream>
#include <list>
#include <map>
#include <iostream>
#include <tuple>

#include "libfive/tree/archive.hpp"
#include "libfive/tree/tree.hpp"

namespace libfive {

////////////////////////////////////////////////////////////////////////////////
// Opcode table

namespace Opcode {

struct Info { Opcode op; int nargs; const char* name; const char* scm; const char* sym; };
static const Info kInfo[] = {
    {INVALID, -1, "INVALID", "", ""},
    {CONSTANT, 0, "CONSTANT", "", ""},
    {VAR_X, 0, "VAR_X", "x", "x"}, {VAR_Y, 0, "VAR_Y", "y", "y"}, {VAR_Z, 0, "VAR_Z", "z", "z"},
    {VAR_FREE, 0, "VAR_FREE", "var-free", ""},
    {CONST_VAR, 1, "CONST_VAR", "const-var", ""},
    {OP_SQUARE, 1, "OP_SQUARE", "square", "square"}, {OP_SQRT, 1, "OP_SQRT", "sqrt", "sqrt"},
    {OP_NEG, 1, "OP_NEG", "neg", "-"}, {OP_SIN, 1, "OP_SIN", "sin", "sin"},
    {OP_COS, 1, "OP_COS", "cos", "cos"}, {OP_TAN, 1, "OP_TAN", "tan", "tan"},
    {OP_ASIN, 1, "OP_ASIN", "asin", "asin"}, {OP_ACOS, 1, "OP_ACOS", "acos", "acos"},
    {OP_ATAN, 1, "OP_ATAN", "atan", "atan"}, {OP_EXP, 1, "OP_EXP", "exp", "exp"},
    {OP_ABS, 1, "OP_ABS", "abs", "abs"}, {OP_LOG, 1, "OP_LOG", "log", "log"},
    {OP_RECIP, 1, "OP_RECIP", "recip", "recip"},
    {OP_ADD, 2, "OP_ADD", "add", "+"}, {OP_MUL, 2, "OP_MUL", "mul", "*"},
    {OP_MIN, 2, "OP_MIN", "min", "min"}, {OP_MAX, 2, "OP_MAX", "max", "max"},
    {OP_SUB, 2, "OP_SUB", "sub", "-"}, {OP_DIV, 2, "OP_DIV", "div", "/"},
    {OP_ATAN2, 2, "OP_ATAN2", "atan2", "atan2"}, {OP_POW, 2, "OP_POW", "pow", "pow"},
    {OP_NTH_ROOT, 2, "OP_NTH_ROOT", "nth-root", "nth-root"},
    {OP_MOD, 2, "OP_MOD", "mod", "mod"}, {OP_NANFILL, 2, "OP_NANFILL", "nanfill", "nanfill"},
    {OP_COMPARE, 2, "OP_COMPARE", "compare", "compare"},
    {ORACLE, 0, "ORACLE", "oracle", ""},
};

static const Info* find(Opcode op) {
    for (const auto& i : kInfo) {
        if (i.op == op) return &i;
    }
    return nullptr;
}

size_t args(Opcode op) {
    const Info* i = find(op);
    return (i && i->nargs >= 0) ? size_t(i->nargs) : size_t(-1);
}
bool isCommutative(Opcode op) {
    return op == OP_ADD || op == OP_MUL || op == OP_MIN || op == OP_MAX;
}
std::string toString(Opcode op) { const Info* i = find(op); return i ? i->name : ""; }
std::string toScmString(Opcode op) { const Info* i = find(op); return i ? i->scm : ""; }
std::string toOpString(Opcode op) { const Info* i = find(op); return i ? i->sym : ""; }

}  // namespace Opcode

////////////////////////////////////////////////////////////////////////////////
// Hash-consing tables

namespace {

typedef std::shared_ptr<Tree::Tree_> Node;
typedef std::tuple<int, const Tree::Tree_*, const Tree::Tree_*> Key;

struct Tables {
    std::map<float, std::weak_ptr<Tree::Tree_>> constants;
    std::weak_ptr<Tree::Tree_> nan_constant;
    std::map<Key, std::weak_ptr<Tree::Tree_>> ops;
    bool simplify = true;
    uint64_t next_serial = 1;
};

Tables& tables() {
    static Tables* t = new Tables;  // intentionally leaked: trees may outlive statics
    return *t;
}

Node constant(float v) {
    Tables& T = tables();
    if (std::isnan(v)) {
        Node n = T.nan_constant.lock();
        if (!n) {
            n.reset(new Tree::Tree_{Opcode::CONSTANT, Tree::FLAG_LOCATION_AGNOSTIC, 0, v,
                                    nullptr, nullptr});
            n->serial = T.next_serial++;
            T.nan_constant = n;
        }
        return n;
    }
    auto itr = T.constants.find(v);
    if (itr != T.constants.end()) {
        if (Node n = itr->second.lock()) return n;
    }
    Node n(new Tree::Tree_{Opcode::CONSTANT, Tree::FLAG_LOCATION_AGNOSTIC, 0, v, nullptr, nullptr});
    n->serial = T.next_serial++;
    T.constants[v] = n;
    return n;
}

float fold(Opcode::Opcode op, float a, float b) {
    using namespace Opcode;
    switch (op) {
        case OP_SQUARE: return a * a;
        case OP_SQRT: return std::sqrt(a);
        case OP_NEG: return -a;
        case OP_SIN: return std::sin(a);
        case OP_COS: return std::cos(a);
        case OP_TAN: return std::tan(a);
        case OP_ASIN: return std::asin(a);
        case OP_ACOS: return std::acos(a);
        case OP_ATAN: return std::atan(a);
        case OP_EXP: return std::exp(a);
        case OP_ABS: return std::fabs(a);
        case OP_LOG: return std::log(a);
        case OP_RECIP: return 1.0f / a;
        case CONST_VAR: return a;
        case OP_ADD: return a + b;
        case OP_MUL: return a * b;
        case OP_MIN: return (b < a) ? b : a;   // Eigen cwiseMin
        case OP_MAX: return (a < b) ? b : a;   // Eigen cwiseMax
        case OP_SUB: return a - b;
        case OP_DIV: return a / b;
        case OP_ATAN2: return std::atan2(a, b);
        case OP_POW: return std::pow(a, b);
        case OP_NTH_ROOT:
            return (a < 0 && (int(b) & 1)) ? -std::pow(-a, 1.0f / b) : std::pow(a, 1.0f / b);
        case OP_MOD: {
            float r = std::fmod(a, b);
            while (r < 0) r += std::fabs(b);
            return (b == 0) ? std::nanf("") : r;
        }
        case OP_NANFILL: return std::isnan(a) ? b : a;
        case OP_COMPARE: return (a < b) ? -1.0f : (a > b) ? 1.0f : 0.0f;
        default: return std::nanf("");
    }
}

Node operation(Opcode::Opcode op, Node lhs, Node rhs);

Node checkIdentity(Opcode::Opcode op, const Node& a, const Node& b) {
    using namespace Opcode;
    const auto op_a = a ? a->op : INVALID;
    const auto op_b = b ? b->op : INVALID;
    const bool ca = op_a == CONSTANT, cb = op_b == CONSTANT;
    switch (op) {
        case OP_NEG: if (op_a == OP_NEG) return a->lhs; break;
        case OP_ABS: if (op_a == OP_ABS) return a; break;
        case OP_ADD:
            if (ca && a->value == 0) return b;
            if (cb && b->value == 0) return a;
            if (op_b == OP_NEG) return operation(OP_SUB, a, b->lhs);
            break;
        case OP_SUB:
            if (ca && a->value == 0) return operation(OP_NEG, b, nullptr);
            if (cb && b->value == 0) return a;
            break;
        case OP_MUL:
            if (ca) {
                if (a->value == 0) return a;
                if (a->value == 1) return b;
                if (a->value == -1) return operation(OP_NEG, b, nullptr);
            }
            if (cb) {
                if (b->value == 0) return b;
                if (b->value == 1) return a;
                if (b->value == -1) return operation(OP_NEG, a, nullptr);
            } else if (a == b) {
                return operation(OP_SQUARE, a, nullptr);
            }
            break;
        case OP_POW:
        case OP_NTH_ROOT:
            if (cb && b->value == 1) return a;
            break;
        case OP_MIN:
        case OP_MAX:
            if (a == b) return a;
            break;
        default: break;
    }
    return nullptr;
}

Node checkCommutative(Opcode::Opcode op, const Node& a, const Node& b) {
    if (!Opcode::isCommutative(op)) return nullptr;
    const unsigned al = a->lhs ? a->lhs->rank : 0, ar = a->rhs ? a->rhs->rank : 0;
    const unsigned bl = b->lhs ? b->lhs->rank : 0, br = b->rhs ? b->rhs->rank : 0;
    if (a->op == op) {
        if (al > b->rank) return operation(op, a->lhs, operation(op, a->rhs, b));
        if (ar > b->rank) return operation(op, a->rhs, operation(op, a->lhs, b));
    } else if (b->op == op) {
        if (bl > a->rank) return operation(op, b->lhs, operation(op, b->rhs, a));
        if (br > a->rank) return operation(op, b->rhs, operation(op, b->lhs, a));
    }
    return nullptr;
}

// Affine collapse (libfive cache.cpp:185-301, :463-501): an ADD/SUB whose two sides share a term
// is rewritten as sum(positive groups) - sum(negative groups), terms grouped by coefficient.
// libfive keys its term maps on node ADDRESSES, so the order of terms inside a group depends on
// the allocator; creation order is used here, which is what a bump-style allocator gives.
struct BySerial {
    bool operator()(const Node& a, const Node& b) const { return a->serial < b->serial; }
};
typedef std::map<Node, float, BySerial> Affine;

Affine asAffine(const Node& n) {
    using namespace Opcode;
    Affine out;
    if (n->op == OP_ADD) {
        out = asAffine(n->lhs);
        for (const auto& i : asAffine(n->rhs)) {
            auto f = out.find(i.first);
            if (f == out.end()) out.insert(i); else f->second += i.second;
        }
    } else if (n->op == OP_SUB) {
        out = asAffine(n->lhs);
        for (const auto& i : asAffine(n->rhs)) {
            auto f = out.find(i.first);
            if (f == out.end()) out.insert({i.first, -i.second}); else f->second -= i.second;
        }
    } else if (n->op == OP_NEG) {
        for (const auto& i : asAffine(n->lhs)) out.insert({i.first, -i.second});
    } else if (n->op == OP_MUL) {
        if (n->lhs->op == CONSTANT) {
            for (const auto& i : asAffine(n->rhs)) out.insert({i.first, i.second * n->lhs->value});
        } else if (n->rhs->op == CONSTANT) {
            for (const auto& i : asAffine(n->lhs)) out.insert({i.first, i.second * n->rhs->value});
        } else {
            out.insert({n, 1.0f});
        }
    } else if (n->op == OP_DIV) {
        if (n->rhs->op == CONSTANT) {
            for (const auto& i : asAffine(n->lhs)) out.insert({i.first, i.second / n->rhs->value});
        } else {
            out.insert({n, 1.0f});
        }
    } else if (n->op == CONSTANT) {
        out.insert({constant(1.0f), n->value});
    } else {
        out.insert({n, 1.0f});
    }
    return out;
}

Node fromAffine(const Affine& ns) {
    std::map<float, std::list<Node>> cs;
    for (const auto& n : ns) cs[n.second].push_back(n.first);
    typedef std::list<std::pair<float, std::list<Node>>> Groups;
    Groups pos, neg;
    for (const auto& c : cs) {
        if (c.first < 0) neg.push_back({-c.first, c.second});
        else if (c.first > 0) pos.push_back(c);
    }
    auto accumulate = [](const Groups& vs) {
        Node out = constant(0.0f);
        for (const auto& v : vs) {
            Node cur = constant(0.0f);
            for (const auto& n : v.second) cur = operation(Opcode::OP_ADD, cur, n);
            out = operation(Opcode::OP_ADD, out, operation(Opcode::OP_MUL, cur, constant(v.first)));
        }
        return out;
    };
    return operation(Opcode::OP_SUB, accumulate(pos), accumulate(neg));
}

Node checkAffine(Opcode::Opcode op, const Node& a_, const Node& b_) {
    if (op != Opcode::OP_ADD && op != Opcode::OP_SUB) return nullptr;
    Affine a = asAffine(a_);
    const Affine b = asAffine(b_);
    bool overlap = false;
    for (const auto& k : b) {
        auto itr = a.find(k.first);
        if (itr != a.end()) {
            if (op == Opcode::OP_ADD) itr->second += k.second; else itr->second -= k.second;
            overlap = true;
        } else {
            a.insert({k.first, op == Opcode::OP_ADD ? k.second : -k.second});
        }
    }
    return overlap ? fromAffine(a) : nullptr;
}

Node operation(Opcode::Opcode op, Node lhs, Node rhs) {
    Tables& T = tables();
    if (T.simplify && Opcode::args(op) >= 1) {
        if (Node t = checkIdentity(op, lhs, rhs)) return t;
        if (Opcode::args(op) == 2) {
            if (Node t = checkCommutative(op, lhs, rhs)) return t;
            if (Node t = checkAffine(op, lhs, rhs)) return t;
        }
    }
    // All-constant operands fold to a constant.
    if ((lhs || rhs) && (!lhs || lhs->op == Opcode::CONSTANT) &&
        (!rhs || rhs->op == Opcode::CONSTANT)) {
        return constant(fold(op, lhs ? lhs->value : 0.0f, rhs ? rhs->value : 0.0f));
    }
    const Key k(int(op), lhs.get(), rhs.get());
    auto itr = T.ops.find(k);
    if (itr != T.ops.end()) {
        if (Node n = itr->second.lock()) return n;
    }
    const bool agnostic = (!lhs || (lhs->flags & Tree::FLAG_LOCATION_AGNOSTIC)) &&
                          (!rhs || (rhs->flags & Tree::FLAG_LOCATION_AGNOSTIC)) &&
                          op != Opcode::VAR_X && op != Opcode::VAR_Y && op != Opcode::VAR_Z;
    const unsigned rank = std::max(lhs ? lhs->rank + 1 : 0u, rhs ? rhs->rank + 1 : 0u);
    Node n(new Tree::Tree_{op, uint8_t(agnostic ? Tree::FLAG_LOCATION_AGNOSTIC : 0), rank,
                           std::nanf(""), lhs, rhs});
    n->serial = T.next_serial++;
    T.ops[k] = n;
    return n;
}

}  // namespace

Tree::Tree_::~Tree_() {}

void Cache::setSimplify(bool on) { tables().simplify = on; }

Tree::Tree(float v) : ptr(constant(v)) {}

Tree::Tree(Opcode::Opcode op, Tree a, Tree b) : ptr(operation(op, a.ptr, b.ptr)) {}

Tree Tree::operator-() const { return Tree(Opcode::OP_NEG, *this); }

std::vector<Tree> Tree::orderedDfs() const {
    // Pass 1: count how many times each node is reached (= number of uses
    // along all paths that a plain stack walk takes).  Pass 2: repeat the
    // same walk and emit a node the last time it is reached; reverse.
    std::map<Id, unsigned> count;
    std::vector<std::shared_ptr<Tree_>> todo = {ptr};
    while (!todo.empty()) {
        auto t = todo.back();
        todo.pop_back();
        if (!t) continue;
        count[t.get()]++;
        if (t->lhs) todo.push_back(t->lhs);
        if (t->rhs) todo.push_back(t->rhs);
    }
    std::vector<Tree> out;
    out.reserve(count.size());
    todo = {ptr};
    while (!todo.empty()) {
        auto t = todo.back();
        todo.pop_back();
        if (!t) continue;
        if (t->lhs) todo.push_back(t->lhs);
        if (t->rhs) todo.push_back(t->rhs);
        if (--count[t.get()] == 0) out.push_back(Tree(t));
    }
    std::reverse(out.begin(), out.end());
    return out;
}

Tree Tree::remap(Tree X_, Tree Y_, Tree Z_) const {
    // NOTE: the pass-1 walk above re-visits shared subtrees, which is
    // exponential on some DAGs; remap uses a memoised post-order instead.
    std::map<Id, std::shared_ptr<Tree_>> done;
    std::vector<std::pair<std::shared_ptr<Tree_>, bool>> stack = {{ptr, false}};
    while (!stack.empty()) {
        auto top = stack.back();
        stack.pop_back();
        const auto& t = top.first;
        if (!t || done.count(t.get())) continue;
        if (!top.second) {
            stack.push_back({t, true});
            stack.push_back({t->lhs, false});
            stack.push_back({t->rhs, false});
            continue;
        }
        std::shared_ptr<Tree_> r;
        switch (t->op) {
            case Opcode::VAR_X: r = X_.ptr; break;
            case Opcode::VAR_Y: r = Y_.ptr; break;
            case Opcode::VAR_Z: r = Z_.ptr; break;
            case Opcode::CONSTANT: r = t; break;
            default:
                r = operation(t->op, t->lhs ? done[t->lhs.get()] : nullptr,
                              t->rhs ? done[t->rhs.get()] : nullptr);
        }
        done[t.get()] = r;
    }
    return Tree(done[ptr.get()]);
}

////////////////////////////////////////////////////////////////////////////////
// Archive reader

namespace {

struct Reader {
    std::istream& in;
    template <typename T> T bytes() {
        T t = T();
        in.read(reinterpret_cast<char*>(&t), sizeof(t));
        return t;
    }
    std::string str() {
        std::string out;
        if (in.eof() || in.get() != '"') {
            std::cerr << "mprb frep reader: expected opening quote\n";
            return out;
        }
        while (!in.eof()) {
            char c = char(in.get());
            if (c == '"') break;
            if (c == '\\') {
                if (!in.eof()) out.push_back(char(in.get()));
            } else {
                out.push_back(c);
            }
        }
        return out;
    }
};

}  // namespace

Archive Archive::deserialize(std::istream& in) {
    Archive out;
    Reader r{in};
    std::vector<Tree> nodes;  // ids are file order, shared across shapes
    while (true) {
        char tag;
        in.get(tag);
        if (in.eof()) break;
        if (tag != 'T' && tag != 't') {
            std::cerr << "mprb frep reader: unexpected shape tag " << int(tag) << "\n";
            break;
        }
        Shape s;
        s.name = r.str();
        s.doc = r.str();
        if (tag == 't') {
            s.tree = nodes.at(r.bytes<uint32_t>());
        } else {
            while (true) {
                const uint8_t op_ = r.bytes<uint8_t>();
                if (in.eof() || op_ == 0xFF) break;
                const auto op = Opcode::Opcode(op_);
                const size_t nargs = Opcode::args(op);
                if (op == Opcode::CONSTANT) {
                    nodes.push_back(Tree(r.bytes<float>()));
                } else if (op == Opcode::ORACLE || nargs == size_t(-1)) {
                    std::cerr << "mprb frep reader: unsupported opcode " << int(op_) << "\n";
                    return out;
                } else if (nargs == 2) {
                    // Right operand is stored first (serializer.cpp:63-64)
                    const uint32_t rhs = r.bytes<uint32_t>();
                    const uint32_t lhs = r.bytes<uint32_t>();
                    nodes.push_back(Tree(op, nodes.at(lhs), nodes.at(rhs)));
                } else if (nargs == 1) {
                    const uint32_t lhs = r.bytes<uint32_t>();
                    nodes.push_back(Tree(op, nodes.at(lhs)));
                } else {
                    nodes.push_back(Tree(op));
                }
            }
            if (nodes.empty()) return out;
            s.tree = nodes.back();
        }
        // Free-variable names, terminated by 0xFF
        while (!in.eof()) {
            const uint8_t b = r.bytes<uint8_t>();
            if (in.eof() || b == 0xFF) break;
            in.unget();
            std::string name = r.str();
            const uint32_t idx = r.bytes<uint32_t>();
            if (idx < nodes.size()) s.vars[nodes[idx].id()] = name;
        }
        out.shapes.push_back(s);
    }
    return out;
}

Tree Tree::deserialize(std::istream& in) {
    auto a = Archive::deserialize(in);
    return a.shapes.empty() ? Tree::Invalid() : a.shapes.front().tree;
}

Tree Tree::load(const std::string& filename) {
    std::ifstream f(filename, std::ios::in | std::ios::binary);
    return f.is_open() ? deserialize(f) : Tree::Invalid();
}

}  // namespace libfive

////////////////////////////////////////////////////////////////////////////////
// Free-function operators (global namespace, as in libfive)

#define MPRB_U(F, OP) \
    libfive::Tree F(const libfive::Tree& a) { return libfive::Tree(libfive::Opcode::OP, a); }
MPRB_U(square, OP_SQUARE) MPRB_U(sqrt, OP_SQRT) MPRB_U(abs, OP_ABS) MPRB_U(sin, OP_SIN)
MPRB_U(cos, OP_COS) MPRB_U(tan, OP_TAN) MPRB_U(asin, OP_ASIN) MPRB_U(acos, OP_ACOS)
MPRB_U(atan, OP_ATAN) MPRB_U(log, OP_LOG) MPRB_U(exp, OP_EXP)
#undef MPRB_U
#define MPRB_B(F, OP)                                                       \
    libfive::Tree F(const libfive::Tree& a, const libfive::Tree& b) {       \
        return libfive::Tree(libfive::Opcode::OP, a, b);                    \
    }
MPRB_B(operator+, OP_ADD) MPRB_B(operator*, OP_MUL) MPRB_B(min, OP_MIN) MPRB_B(max, OP_MAX)
MPRB_B(operator-, OP_SUB) MPRB_B(operator/, OP_DIV) MPRB_B(atan2, OP_ATAN2) MPRB_B(pow, OP_POW)
MPRB_B(nth_root, OP_NTH_ROOT) MPRB_B(mod, OP_MOD) MPRB_B(nanfill, OP_NANFILL)
MPRB_B(compare, OP_COMPARE)
#undef MPRB_B
