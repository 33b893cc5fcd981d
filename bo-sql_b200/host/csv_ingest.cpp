// csv_ingest.cpp — CSV ingest with the reference's type inference (declared in bosql_types.hpp; src/storage/csv_loader.cpp:7-166).
//
// The reference keeps every cell as a std::string in a vector<vector<string>> and parses each column up to four times
// (src/storage/csv_loader.cpp:26-38, 48-162; 1.85 s for a 1 M-row file).  Here the file is read once into one buffer whose
// separators are overwritten with NULs, so every cell is a C string in place; a column is classified in one pass that
// runs the same libc conversions the reference's std::stoi / std::stod wrap (strtol / strtod: prefix parsing, errno range
// errors), which keeps the inferred types and values identical.  Columns then go to the device at first use.
#include "bosql_types.hpp"

#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <limits>
#include <sstream>
#include <unordered_set>

namespace bosql {

namespace {

// std::stoi: strtol, throws when nothing converts or the value leaves int's range
bool stoi_like(const char* s, int& out) {
    errno = 0;
    char* end = nullptr;
    long v = std::strtol(s, &end, 10);
    if (end == s) return false;
    if (errno == ERANGE || v < std::numeric_limits<int>::min() || v > std::numeric_limits<int>::max()) return false;
    out = static_cast<int>(v);
    return true;
}

// std::stod: strtod, throws when nothing converts or on ERANGE
bool stod_like(const char* s, double& out) {
    errno = 0;
    char* end = nullptr;
    double v = std::strtod(s, &end);
    if (end == s) return false;
    if (errno == ERANGE) return false;
    out = v;
    return true;
}

}  // namespace

std::pair<Table, TableMeta> load_csv(std::istream& stream) {
    std::string buf((std::istreambuf_iterator<char>(stream)), std::istreambuf_iterator<char>());
    Table table;
    table.dict = std::make_shared<Dictionary>();
    std::vector<ColumnMeta> metas;

    // ---- split into lines and cells in place ----------------------------------------------------------------
    // std::getline semantics: a line ends at '\n' (a final line without one still counts); a cell ends at ','; a
    // trailing comma does not open an empty last cell (the reference's inner getline stops at end of line).
    std::vector<std::string> headers;
    std::vector<std::vector<const char*>> cols;     // cols[c][r] -> NUL-terminated cell
    std::vector<std::vector<size_t>> lens;          // cell lengths (the size() == 8 test for dates)
    std::vector<const char*> row_cells;
    std::vector<size_t> row_lens;
    size_t pos = 0, n_rows = 0;
    bool first_line = true;
    while (pos < buf.size()) {
        size_t eol = buf.find('\n', pos);
        if (eol == std::string::npos) eol = buf.size();
        const size_t line_begin = pos, line_end = eol;
        pos = eol + 1;
        if (!first_line && line_begin == line_end) continue;           // empty data lines are skipped (:28)
        row_cells.clear();
        row_lens.clear();
        size_t c0 = line_begin;
        while (c0 < line_end) {
            size_t comma = buf.find(',', c0);
            if (comma == std::string::npos || comma > line_end) comma = line_end;
            row_cells.push_back(buf.data() + c0);
            row_lens.push_back(comma - c0);
            if (comma < buf.size()) buf[comma] = '\0';
            c0 = comma + 1;
        }
        if (line_end < buf.size()) buf[line_end] = '\0';
        if (first_line) {
            first_line = false;
            for (size_t i = 0; i < row_cells.size(); ++i) headers.emplace_back(row_cells[i], row_lens[i]);
            cols.resize(headers.size());
            lens.resize(headers.size());
            continue;
        }
        if (row_cells.size() != headers.size()) throw std::runtime_error("Row size mismatch");
        for (size_t c = 0; c < headers.size(); ++c) {
            cols[c].push_back(row_cells[c]);
            lens[c].push_back(row_lens[c]);
        }
        ++n_rows;
    }

    // ---- classify and convert each column -------------------------------------------------------------------
    for (size_t c = 0; c < headers.size(); ++c) {
        TableColumn column;
        column.name = headers[c];
        ColumnMeta meta(headers[c], TypeId::STRING);
        const auto& cells = cols[c];

        // DATE32 (:48-82)
        bool all_date = n_rows > 0;
        for (size_t r = 0; r < n_rows && all_date; ++r) {
            int d;
            if (lens[c][r] != 8 || !stoi_like(cells[r], d)) all_date = false;
            else if (d < 19000000 || d > 21000000) all_date = false;
        }
        if (all_date) {
            std::vector<Date32> data(n_rows);
            Date32 lo = std::numeric_limits<Date32>::max(), hi = std::numeric_limits<Date32>::min();
            for (size_t r = 0; r < n_rows; ++r) {
                int d = 0;
                stoi_like(cells[r], d);
                data[r] = d;
                lo = std::min(lo, d);
                hi = std::max(hi, d);
            }
            meta.type = TypeId::DATE32;
            meta.stats.min_date = lo;
            meta.stats.max_date = hi;
            meta.stats.ndv = std::unordered_set<Date32>(data.begin(), data.end()).size();
            column.data = std::make_unique<ColumnVector<Date32>>(std::move(data));
            table.columns.push_back(std::move(column));
            metas.push_back(std::move(meta));
            continue;
        }

        // INT64, parsed THROUGH double like the reference (:85-118): exact only up to 2^53
        std::vector<double> as_f(n_rows);
        bool all_f64 = n_rows > 0, all_i64 = n_rows > 0;
        for (size_t r = 0; r < n_rows && all_f64; ++r) {
            if (!stod_like(cells[r], as_f[r])) {
                all_f64 = all_i64 = false;
                break;
            }
            const double v = as_f[r];
            if (v != std::floor(v) || v < static_cast<double>(std::numeric_limits<i64>::min()) ||
                v > static_cast<double>(std::numeric_limits<i64>::max()))
                all_i64 = false;
        }
        if (all_i64) {
            std::vector<i64> data(n_rows);
            i64 lo = std::numeric_limits<i64>::max(), hi = std::numeric_limits<i64>::min();
            for (size_t r = 0; r < n_rows; ++r) {
                const double v = as_f[r];
                data[r] = v >= 9223372036854775808.0 ? std::numeric_limits<i64>::min() : static_cast<i64>(v);   // x86 cast of 2^63
                lo = std::min(lo, data[r]);
                hi = std::max(hi, data[r]);
            }
            meta.type = TypeId::INT64;
            meta.stats.min_i64 = lo;
            meta.stats.max_i64 = hi;
            meta.stats.ndv = std::unordered_set<i64>(data.begin(), data.end()).size();
            column.data = std::make_unique<ColumnVector<i64>>(std::move(data));
            table.columns.push_back(std::move(column));
            metas.push_back(std::move(meta));
            continue;
        }
        if (all_f64) {                                               // DOUBLE (:121-149)
            double lo = std::numeric_limits<double>::max(), hi = std::numeric_limits<double>::lowest();
            for (double v : as_f) {
                lo = std::min(lo, v);
                hi = std::max(hi, v);
            }
            meta.type = TypeId::DOUBLE;
            meta.stats.min_f64 = lo;
            meta.stats.max_f64 = hi;
            // std::set<double> semantics: -0.0 and 0.0 are one value, every NaN is "equivalent" to everything else
            std::unordered_set<double> uniq;
            bool has_nan = false;
            for (double v : as_f) {
                if (std::isnan(v)) has_nan = true;
                else uniq.insert(v == 0.0 ? 0.0 : v);
            }
            meta.stats.ndv = uniq.size() + ((has_nan && uniq.empty()) ? 1 : 0);
            column.data = std::make_unique<ColumnVector<double>>(std::move(as_f));
            table.columns.push_back(std::move(column));
            metas.push_back(std::move(meta));
            continue;
        }
        // STRING: dictionary ids in first-seen order (:152-161)
        std::vector<StrId> data(n_rows);
        for (size_t r = 0; r < n_rows; ++r) data[r] = table.dict->get_or_add(std::string(cells[r], lens[c][r]));
        meta.stats.ndv = std::unordered_set<StrId>(data.begin(), data.end()).size();
        column.data = std::make_unique<ColumnVector<StrId>>(std::move(data));
        table.columns.push_back(std::move(column));
        metas.push_back(std::move(meta));
    }
    TableMeta table_meta("", std::move(metas), n_rows);
    return {std::move(table), std::move(table_meta)};
}

std::pair<Table, TableMeta> load_csv(const std::string& filename) {
    std::ifstream file(filename, std::ios::binary);
    if (!file.is_open()) throw std::runtime_error("Cannot open file: " + filename);
    return load_csv(file);
}

}  // namespace bosql
