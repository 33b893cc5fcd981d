// csv_loader.hpp — CSV ingest with the reference's type inference (src/storage/csv_loader.cpp:7-166), §8f N1.
#pragma once

#include <istream>
#include <string>
#include <utility>

#include "bosql_types.hpp"

namespace bosql {

// Same contract as the reference's load_csv (include/storage/csv_loader.h:18-19): header line, comma separated, no
// quoting; per column DATE32 (all values 8 characters, stoi in [19000000, 21000000]) else INT64 (all values stod-integral)
// else DOUBLE (all values stod-parsable) else STRING (dictionary ids in first-seen order); min/max/ndv recorded.
std::pair<Table, TableMeta> load_csv(const std::string& filename);
std::pair<Table, TableMeta> load_csv(std::istream& stream);

}  // namespace bosql
