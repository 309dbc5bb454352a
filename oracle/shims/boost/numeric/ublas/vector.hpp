// oracle/shims: see matrix.hpp (TEST INFRASTRUCTURE)
#pragma once
#include "matrix.hpp"
