#pragma once
// include shim: the reference's logical layer (out of scope) is compiled from its own sources on top of the product's
// AST / LogicalOp declarations; get_output_schema (src/logical/planner.cpp:167) is the one declaration those lack.
#include <tuple>

#include "bosql_sql.hpp"

namespace bosql {
std::tuple<std::vector<std::string>, std::vector<TypeId>, const Dictionary*> get_output_schema(const LogicalOp* plan, const Catalog& catalog);
}
