// libfive::Cache stand-in lives in tree.hpp (src/tape.cpp:13 includes this name).
#pragma once
#include "libfive/tree/tree.hpp"
