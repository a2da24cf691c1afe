// Reference header name kept so that `#include "tape.hpp"` in the drivers resolves; see mpr.hpp.
#pragma once
#include "mpr.hpp"
