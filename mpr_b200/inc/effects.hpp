// Reference header name kept so that `#include "effects.hpp"` in the drivers resolves; see mpr.hpp.
#pragma once
#include "mpr.hpp"
