// Host-side helpers shared by the C-ABI implementation (not public API).
#pragma once
#include <cstdint>
#include <vector>

namespace libfive { class Tree; }

namespace mprb {

// Tree -> packed tape; restates reference src/tape.cpp:21-228.
std::vector<uint64_t> pack_tape(const libfive::Tree& tree, int* num_slots_out);

}  // namespace mprb
