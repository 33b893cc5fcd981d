// oracle/ref_shim.cpp — TEST INFRASTRUCTURE, not product code.
//
// A C-ABI wrapper around the UNMODIFIED reference executor (bolu-atx/bo-sql),
// compiled by oracle/Makefile together with the reference's own sources where
// they lie under /root/reference into oracle/_ref/libbosql_ref.so.  Nothing from
// the reference is copied into this repository; this file only calls its public
// API:
//   parse_sql                      include/parser/parser.h:57
//   LogicalPlanner::build_logical_plan   include/logical/planner.h:16
//   build_physical_plan            include/exec/physical_planner.h:11
//   Operator::{open,next,close}    include/exec/operator.hpp:17-31
//   Table / Dictionary / Catalog   include/storage/table.h:20, dictionary.h:11, catalog/catalog.h:46
//
// The shim lets pytest build tables from raw typed arrays, run a SQL string and
// read back RAW typed result columns (the reference's own text formatter prints
// DOUBLE with six decimals, src/exec/execution.cpp:31, so text cannot carry parity).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.

#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "catalog/catalog.h"
#include "exec/operator.hpp"
#include "exec/physical_planner.h"
#include "logical/planner.h"
#include "parser/parser.h"
#include "storage/csv_loader.h"

using namespace bosql;

namespace {

thread_local std::string g_err;

struct RefDict {
    std::shared_ptr<Dictionary> dict = std::make_shared<Dictionary>();
};

struct RefTable {
    Table table;
    std::vector<ColumnMeta> metas;
    size_t rows = 0;
};

struct RefCatalog {
    Catalog catalog;
};

struct RefResult {
    std::vector<std::string> names;
    std::vector<TypeId> types;
    // one byte buffer per output column, elements packed at their natural width
    std::vector<std::vector<unsigned char>> cols;
    size_t rows = 0;
    double seconds = 0.0;       // open() .. last next() .. close()
    Dictionary* dict = nullptr; // dictionary the root operator reports
};

size_t width_of(TypeId t) {
    return (t == TypeId::INT64 || t == TypeId::DOUBLE) ? 8 : 4;
}

template <typename T>
std::unique_ptr<Column> make_col(const void* data, size_t n) {
    const T* p = static_cast<const T*>(data);
    return std::make_unique<ColumnVector<T>>(std::vector<T>(p, p + n));
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void* ref_dict_create() { return new RefDict(); }
void ref_dict_destroy(void* d) { delete static_cast<RefDict*>(d); }
// Dictionary::get_or_add, src/storage/dictionary.cpp:5
unsigned ref_dict_get_or_add(void* d, const char* s) {
    return static_cast<RefDict*>(d)->dict->get_or_add(s);
}
size_t ref_dict_size(void* d) { return static_cast<RefDict*>(d)->dict->strings.size(); }
const char* ref_dict_get(void* d, unsigned id) {
    auto& v = static_cast<RefDict*>(d)->dict->strings;
    return id < v.size() ? v[id].c_str() : nullptr;
}

void* ref_catalog_create() { return new RefCatalog(); }
void ref_catalog_destroy(void* c) { delete static_cast<RefCatalog*>(c); }

// A table under construction. `dict` may be shared between tables (the reference's own
// join tests do so, tests/test_execution.cpp:116-123) or NULL for a fresh dictionary.
void* ref_table_create(const char* name, void* dict) {
    auto* t = new RefTable();
    t->table.name = name;
    t->table.dict = dict ? static_cast<RefDict*>(dict)->dict : std::make_shared<Dictionary>();
    return t;
}

// type: TypeId ordinal (include/types.h:17) 0=INT64 1=DOUBLE 2=STRING(ids) 3=DATE32
int ref_table_add_column(void* tp, const char* name, int type, const void* data, size_t n) {
    try {
        auto* t = static_cast<RefTable*>(tp);
        if (!t->table.columns.empty() && n != t->rows) {
            g_err = "column length mismatch";
            return 1;
        }
        std::unique_ptr<Column> col;
        switch (static_cast<TypeId>(type)) {
            case TypeId::INT64: col = make_col<int64_t>(data, n); break;
            case TypeId::DOUBLE: col = make_col<double>(data, n); break;
            case TypeId::STRING: col = make_col<uint32_t>(data, n); break;
            case TypeId::DATE32: col = make_col<int32_t>(data, n); break;
            default: g_err = "bad type"; return 1;
        }
        t->table.columns.push_back({name, std::move(col)});
        t->metas.emplace_back(name, static_cast<TypeId>(type));
        t->rows = n;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// Catalog::register_table, src/catalog/catalog.cpp:5.  Consumes the table handle.
int ref_catalog_register(void* cp, void* tp) {
    try {
        auto* c = static_cast<RefCatalog*>(cp);
        std::unique_ptr<RefTable> t(static_cast<RefTable*>(tp));
        TableMeta meta(t->table.name, std::move(t->metas), t->rows);
        c->catalog.register_table(std::move(t->table), std::move(meta));
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// load_csv, src/storage/csv_loader.cpp:168 — registered under `name`
// (the CLI hard-codes "table", src/cli/main.cpp:105).
int ref_catalog_load_csv(void* cp, const char* path, const char* name) {
    try {
        auto* c = static_cast<RefCatalog*>(cp);
        auto [table, meta] = load_csv(path);
        table.name = name;
        meta.name = name;
        c->catalog.register_table(std::move(table), std::move(meta));
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// Catalog introspection so tests can read what load_csv inferred.
int ref_catalog_table_info(void* cp, const char* name, size_t* rows, size_t* ncols) {
    auto* c = static_cast<RefCatalog*>(cp);
    auto meta = c->catalog.get_table_meta(name);
    if (!meta.has_value()) { g_err = "no such table"; return 1; }
    *rows = meta->row_count;
    *ncols = meta->columns.size();
    return 0;
}

int ref_catalog_column_info(void* cp, const char* name, size_t i, const char** col_name, int* type,
                            const void** data, long long* min_i, long long* max_i,
                            double* min_f, double* max_f, size_t* ndv) {
    auto* c = static_cast<RefCatalog*>(cp);
    auto meta = c->catalog.get_table_meta(name);
    auto tab = c->catalog.get_table_data(name);
    if (!meta.has_value() || !tab.has_value() || i >= meta->columns.size()) { g_err = "bad column"; return 1; }
    const auto& cm = meta->columns[i];
    *col_name = cm.name.c_str();
    *type = static_cast<int>(cm.type);
    *ndv = cm.stats.ndv;
    *min_f = cm.stats.min_f64; *max_f = cm.stats.max_f64;
    if (cm.type == TypeId::DATE32) { *min_i = cm.stats.min_date; *max_i = cm.stats.max_date; }
    else { *min_i = cm.stats.min_i64; *max_i = cm.stats.max_i64; }
    const Column* col = tab->columns[i].data.get();
    switch (cm.type) {
        case TypeId::INT64: *data = dynamic_cast<const ColumnVector<int64_t>*>(col)->data.data(); break;
        case TypeId::DOUBLE: *data = dynamic_cast<const ColumnVector<double>*>(col)->data.data(); break;
        case TypeId::STRING: *data = dynamic_cast<const ColumnVector<uint32_t>*>(col)->data.data(); break;
        case TypeId::DATE32: *data = dynamic_cast<const ColumnVector<int32_t>*>(col)->data.data(); break;
    }
    return 0;
}

size_t ref_catalog_dict_size(void* cp, const char* name) {
    auto tab = static_cast<RefCatalog*>(cp)->catalog.get_table_data(name);
    return (tab.has_value() && tab->dict) ? tab->dict->strings.size() : 0;
}
const char* ref_catalog_dict_get(void* cp, const char* name, unsigned id) {
    auto tab = static_cast<RefCatalog*>(cp)->catalog.get_table_data(name);
    if (!tab.has_value() || !tab->dict || id >= tab->dict->strings.size()) return nullptr;
    return tab->dict->strings[id].c_str();
}

// parse → logical plan → physical plan → open/next/close, exactly as execute_select_sql does
// (src/cli/main.cpp:40-57) but keeping typed columns.  Returns NULL on error (message in
// ref_last_error) — the reference signals every error as std::runtime_error.
void* ref_query(void* cp, const char* sql) {
    try {
        auto* c = static_cast<RefCatalog*>(cp);
        SelectStmt stmt = parse_sql(sql);
        LogicalPlanner planner;
        auto logical = planner.build_logical_plan(stmt);
        auto root = build_physical_plan(logical.get(), c->catalog);

        auto res = std::make_unique<RefResult>();
        res->names = root->output_names();
        res->types = root->output_types();
        res->dict = root->dictionary();
        res->cols.resize(res->types.size());

        auto t0 = std::chrono::steady_clock::now();
        root->open();
        ExecBatch batch;
        while (root->next(batch)) {
            for (size_t j = 0; j < batch.columns.size() && j < res->cols.size(); ++j) {
                const auto& s = batch.columns[j];
                size_t w = width_of(s.type);
                const auto* p = static_cast<const unsigned char*>(s.data);
                res->cols[j].insert(res->cols[j].end(), p, p + w * batch.length);
            }
            res->rows += batch.length;
        }
        root->close();
        auto t1 = std::chrono::steady_clock::now();
        res->seconds = std::chrono::duration<double>(t1 - t0).count();
        return res.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

size_t ref_result_rows(void* r) { return static_cast<RefResult*>(r)->rows; }
size_t ref_result_cols(void* r) { return static_cast<RefResult*>(r)->types.size(); }
double ref_result_seconds(void* r) { return static_cast<RefResult*>(r)->seconds; }
const char* ref_result_name(void* r, size_t i) { return static_cast<RefResult*>(r)->names[i].c_str(); }
int ref_result_type(void* r, size_t i) { return static_cast<int>(static_cast<RefResult*>(r)->types[i]); }
const void* ref_result_data(void* r, size_t i) { return static_cast<RefResult*>(r)->cols[i].data(); }
int ref_result_has_dict(void* r) { return static_cast<RefResult*>(r)->dict != nullptr; }
const char* ref_result_dict_get(void* r, unsigned id) {
    auto* d = static_cast<RefResult*>(r)->dict;
    return (d && id < d->strings.size()) ? d->strings[id].c_str() : nullptr;
}
void ref_result_free(void* r) { delete static_cast<RefResult*>(r); }

// The plan's printed shape (LogicalOp::to_string, src/logical/logical.cpp) for plan-shape checks.
int ref_explain(const char* sql, char* out, size_t cap) {
    try {
        SelectStmt stmt = parse_sql(sql);
        LogicalPlanner planner;
        auto logical = planner.build_logical_plan(stmt);
        std::string s = logical->to_string();
        std::strncpy(out, s.c_str(), cap - 1);
        out[cap - 1] = 0;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

}  // extern "C"
