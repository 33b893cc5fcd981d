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

// Fast paths for the cells that make up almost every numeric file: plain decimals.  They are taken only when the WHOLE
// cell is [+-]digits[.digits] with at most 15 significant digits and at most 22 fractional digits; then the value is
// mantissa / 10^k with both operands exact in binary64, so the one IEEE division is correctly rounded (Clinger's fast
// path) - the same bits glibc's correctly rounded strtod returns.  Anything else (exponents, hex, inf/nan, blanks, a
// trailing '\r' or other suffix, long digit strings) goes to strtol / strtod below, whose prefix semantics the reference's
// std::stoi / std::stod have.
const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

bool fast_decimal(const char* s, size_t len, double& out) {
    size_t i = 0;
    bool neg = false;
    if (i < len && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
    uint64_t mant = 0;
    int digits = 0, sig = 0, frac = 0;
    for (; i < len && s[i] >= '0' && s[i] <= '9'; ++i, ++digits) {
        mant = mant * 10 + static_cast<uint64_t>(s[i] - '0');
        if (mant) ++sig;
        if (sig > 15) return false;
    }
    if (i < len && s[i] == '.') {
        ++i;
        for (; i < len && s[i] >= '0' && s[i] <= '9'; ++i, ++digits, ++frac) {
            mant = mant * 10 + static_cast<uint64_t>(s[i] - '0');
            if (mant) ++sig;
            if (sig > 15 || frac >= 22) return false;
        }
    }
    if (i != len || digits == 0) return false;
    const double v = static_cast<double>(mant) / kPow10[frac];
    out = neg ? -v : v;
    return true;
}

bool fast_int8(const char* s, size_t len, int& out) {       // the DATE32 test only ever accepts 8-character cells
    if (len != 8) return false;
    int v = 0;
    for (size_t i = 0; i < 8; ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (s[i] - '0');
    }
    out = v;
    return true;
}

// number of distinct values (std::set / unordered_set size in the reference): open addressing over the 64-bit patterns
template <typename T>
size_t count_distinct(const std::vector<T>& v) {
    if (v.empty()) return 0;
    size_t cap = 16;
    while (cap < v.size() * 2) cap <<= 1;
    std::vector<uint64_t> slots(cap, 0);
    std::vector<uint8_t> used(cap, 0);
    size_t n = 0;
    for (const T& x : v) {
        uint64_t k = 0;
        std::memcpy(&k, &x, sizeof(T));
        uint64_t h = k * 0x9E3779B97F4A7C15ull;
        h ^= h >> 32;
        size_t i = static_cast<size_t>(h) & (cap - 1);
        while (used[i] && slots[i] != k) i = (i + 1) & (cap - 1);
        if (!used[i]) {
            used[i] = 1;
            slots[i] = k;
            ++n;
        }
    }
    return n;
}

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

std::pair<Table, TableMeta> load_csv(std::istream& stream) { return load_csv(stream, nullptr); }

std::pair<Table, TableMeta> load_csv(std::istream& stream, std::shared_ptr<Dictionary> shared_dict) {
    std::string buf;
    {
        std::ostringstream all;                       // one bulk copy through the stream buffer (not a per-character iterator)
        all << stream.rdbuf();
        buf = std::move(all).str();
    }
    Table table;
    table.dict = shared_dict ? std::move(shared_dict) : std::make_shared<Dictionary>();
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
            // bounded by the line: the last cell of a line must not scan on to the next comma anywhere in the file
            const void* hit = std::memchr(buf.data() + c0, ',', line_end - c0);
            const size_t comma = hit ? static_cast<size_t>(static_cast<const char*>(hit) - buf.data()) : line_end;
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
        std::vector<Date32> dates;
        if (all_date && lens[c][0] == 8) dates.resize(n_rows);
        for (size_t r = 0; r < n_rows && all_date; ++r) {
            int d;
            if (lens[c][r] != 8 || !(fast_int8(cells[r], 8, d) || stoi_like(cells[r], d))) all_date = false;
            else if (d < 19000000 || d > 21000000) all_date = false;
            else dates[r] = d;
        }
        if (all_date) {
            std::vector<Date32> data = std::move(dates);
            Date32 lo = std::numeric_limits<Date32>::max(), hi = std::numeric_limits<Date32>::min();
            for (Date32 d : data) {
                lo = std::min(lo, d);
                hi = std::max(hi, d);
            }
            meta.type = TypeId::DATE32;
            meta.stats.min_date = lo;
            meta.stats.max_date = hi;
            meta.stats.ndv = count_distinct(data);
            column.data = std::make_unique<ColumnVector<Date32>>(std::move(data));
            table.columns.push_back(std::move(column));
            metas.push_back(std::move(meta));
            continue;
        }

        // INT64, parsed THROUGH double like the reference (:85-118): exact only up to 2^53
        std::vector<double> as_f(n_rows);
        bool all_f64 = n_rows > 0, all_i64 = n_rows > 0;
        for (size_t r = 0; r < n_rows && all_f64; ++r) {
            if (!fast_decimal(cells[r], lens[c][r], as_f[r]) && !stod_like(cells[r], as_f[r])) {
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
            meta.stats.ndv = count_distinct(data);
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
            std::vector<double> keys;
            keys.reserve(as_f.size());
            bool has_nan = false;
            for (double v : as_f) {
                if (std::isnan(v)) has_nan = true;
                else keys.push_back(v == 0.0 ? 0.0 : v);
            }
            const size_t uniq = count_distinct(keys);
            meta.stats.ndv = uniq + ((has_nan && uniq == 0) ? 1 : 0);
            column.data = std::make_unique<ColumnVector<double>>(std::move(as_f));
            table.columns.push_back(std::move(column));
            metas.push_back(std::move(meta));
            continue;
        }
        // STRING: dictionary ids in first-seen order (:152-161)
        std::vector<StrId> data(n_rows);
        for (size_t r = 0; r < n_rows; ++r) data[r] = table.dict->get_or_add(std::string(cells[r], lens[c][r]));
        meta.stats.ndv = count_distinct(data);
        column.data = std::make_unique<ColumnVector<StrId>>(std::move(data));
        table.columns.push_back(std::move(column));
        metas.push_back(std::move(meta));
    }
    TableMeta table_meta("", std::move(metas), n_rows);
    return {std::move(table), std::move(table_meta)};
}

std::pair<Table, TableMeta> load_csv(const std::string& filename) { return load_csv(filename, nullptr); }

std::pair<Table, TableMeta> load_csv(const std::string& filename, std::shared_ptr<Dictionary> shared_dict) {
    std::ifstream file(filename, std::ios::binary);
    if (!file.is_open()) throw std::runtime_error("Cannot open file: " + filename);
    return load_csv(file, std::move(shared_dict));
}

}  // namespace bosql
