// Reference header name kept so that `#include "gpu_opcode.hpp"` in the drivers resolves; see mpr.hpp.
#pragma once
#include "mpr.hpp"
