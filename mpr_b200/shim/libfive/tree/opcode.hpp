// Minimal libfive-compatible opcode table (host only).
//
// Only what mpr::Tape construction and the reference's benchmark drivers
// touch.  Numbering is libfive's *packed* scheme, which the reference forces
// on at configure time (reference CMakeLists.txt:6-8) and which the .frep
// fixtures are written in (libfive/libfive/include/libfive/tree/opcode.hpp:63-101).
#pragma once
#include <cstddef>
#include <string>

namespace libfive {
namespace Opcode {

enum Opcode {
    INVALID = 0,
    CONSTANT = 1, VAR_X = 2, VAR_Y = 3, VAR_Z = 4, VAR_FREE = 5, CONST_VAR = 6,
    OP_SQUARE = 7, OP_SQRT = 8, OP_NEG = 9, OP_SIN = 10, OP_COS = 11, OP_TAN = 12,
    OP_ASIN = 13, OP_ACOS = 14, OP_ATAN = 15, OP_EXP = 16, OP_ABS = 17, OP_LOG = 18,
    OP_RECIP = 19,
    OP_ADD = 20, OP_MUL = 21, OP_MIN = 22, OP_MAX = 23, OP_SUB = 24, OP_DIV = 25,
    OP_ATAN2 = 26, OP_POW = 27, OP_NTH_ROOT = 28, OP_MOD = 29, OP_NANFILL = 30,
    OP_COMPARE = 31,
    ORACLE = 32,
    LAST_OP = 33,
};

// Number of operands (0, 1, 2), or size_t(-1) for INVALID / LAST_OP.
size_t args(Opcode op);
bool isCommutative(Opcode op);
std::string toString(Opcode op);      // "OP_ADD"
std::string toScmString(Opcode op);   // "add"
std::string toOpString(Opcode op);    // "+"

}  // namespace Opcode
}  // namespace libfive
