#pragma once
// include shim: the reference's exec/operator.hpp = the product's operator mirror, plus the two declarations the
// reference keeps in that header although they are not operators: its text formatter (out of scope, taken from the
// reference itself) and the driver loop run_query (src/exec/execution.cpp, compiled unmodified on top of this header).
#include "bosql_operator.hpp"
#include "exec/formatter.hpp"

namespace bosql {
void run_query(std::unique_ptr<Operator> root, const std::vector<std::string>& col_names, const std::vector<TypeId>& col_types,
               Formatter& formatter, const Dictionary* dict);
}
