#pragma once
// include shim: build_physical_plan is declared with the operator mirror (bosql_operator.hpp)
#include "exec/operator.hpp"
