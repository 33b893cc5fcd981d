// bosql_types.hpp — host-side data model of the drop-in: the same names, fields and meaning as the
// reference's include/types.h, include/exec/execution_types.hpp, include/storage/{dictionary,table}.h and
// include/catalog/catalog.h, so code written against the reference's operator interface compiles against
// this header unchanged.  What is new is the device side: every Column can carry an HBM mirror
// (`device`), and a column may exist ONLY in HBM (DeviceColumn: synthetic tables generated on the GPU).
#pragma once

#include <algorithm>
#include <cstdint>
#include <istream>
#include <memory>
#include <optional>
#include <span>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

struct bq_col;

namespace bosql {

using i64 = int64_t;
using f64 = double;
using StrId = uint32_t;   // dictionary id
using Date32 = int32_t;   // YYYYMMDD

// ordinals are part of the C ABI (BQ_INT64.. in include/bosql_b200.h)  — reference: include/types.h:17
enum class TypeId { INT64, DOUBLE, STRING, DATE32 };

inline size_t type_width(TypeId t) { return (t == TypeId::INT64 || t == TypeId::DOUBLE) ? 8 : 4; }

// reference: include/types.h:20-69 (tagged 16-byte scalar; kept for API compatibility — the GPU path never boxes)
union DatumValue {
    int64_t i64_val;
    double f64_val;
    StrId str_id;
    Date32 date32_val;
};

struct Datum {
    TypeId type;
    DatumValue value;

    int64_t as_i64() const { return check(TypeId::INT64).value.i64_val; }
    double as_f64() const { return check(TypeId::DOUBLE).value.f64_val; }
    StrId as_str() const { return check(TypeId::STRING).value.str_id; }
    Date32 as_date32() const { return check(TypeId::DATE32).value.date32_val; }

    static Datum from_i64(int64_t v) { Datum d{TypeId::INT64, {}}; d.value.i64_val = v; return d; }
    static Datum from_f64(double v) { Datum d{TypeId::DOUBLE, {}}; d.value.f64_val = v; return d; }
    static Datum from_str(StrId v) { Datum d{TypeId::STRING, {}}; d.value.str_id = v; return d; }
    static Datum from_date32(Date32 v) { Datum d{TypeId::DATE32, {}}; d.value.date32_val = v; return d; }

private:
    const Datum& check(TypeId want) const {
        if (type != want) throw std::runtime_error("Type mismatch");
        return *this;
    }
};

struct ColumnType {
    TypeId type_id;
    std::string name;
    ColumnType(TypeId id, std::string col_name = "") : type_id(id), name(std::move(col_name)) {}
    bool operator==(const ColumnType& o) const { return type_id == o.type_id; }   // name is not compared
    bool operator!=(const ColumnType& o) const { return !(*this == o); }
};

template <typename T>
class OptionalRef {
    const T* ptr = nullptr;
public:
    OptionalRef() = default;
    OptionalRef(const T& ref) : ptr(&ref) {}
    bool has_value() const { return ptr != nullptr; }
    explicit operator bool() const { return has_value(); }
    const T& value() const { if (!ptr) throw std::bad_optional_access(); return *ptr; }
    const T* operator->() const { return ptr; }
    const T& operator*() const { return value(); }
};

template <typename T> TypeId type_id_for();
template <> inline TypeId type_id_for<int64_t>() { return TypeId::INT64; }
template <> inline TypeId type_id_for<double>() { return TypeId::DOUBLE; }
template <> inline TypeId type_id_for<int32_t>() { return TypeId::DATE32; }
template <> inline TypeId type_id_for<uint32_t>() { return TypeId::STRING; }

// HBM mirror of a column: created at first use by an operator (or at load for device-only columns).
struct DeviceMirror {
    bq_col* handle = nullptr;
    size_t rows = 0;
    const void* host_data = nullptr;   // what was uploaded (re-upload if the vector moved or grew)
    ~DeviceMirror();
};

// reference: include/types.h:127-145
struct Column {
    virtual ~Column() {}
    virtual TypeId type() const = 0;
    virtual size_t size() const = 0;
    virtual const void* host_data() const = 0;          // nullptr when the column lives only in HBM
    mutable std::shared_ptr<DeviceMirror> device;        // "columns become device-resident after load"
};

template <typename T>
struct ColumnVector : public Column {
    std::vector<T> data;
    explicit ColumnVector(size_t reserve = 0) { data.reserve(reserve); }
    explicit ColumnVector(std::vector<T> d) : data(std::move(d)) {}
    TypeId type() const override { return type_id_for<T>(); }
    size_t size() const override { return data.size(); }
    const void* host_data() const override { return data.data(); }
    void append(const T& v) { data.push_back(v); }
};

// A host column over memory the caller owns (e.g. pinned buffers): no copy into a std::vector, and uploads run at
// full PCIe speed.  The caller keeps the memory alive and unchanged while the table is registered.
struct BorrowedColumn : public Column {
    TypeId type_id;
    const void* ptr;
    size_t rows;
    BorrowedColumn(TypeId t, const void* p, size_t n) : type_id(t), ptr(p), rows(n) {}
    TypeId type() const override { return type_id; }
    size_t size() const override { return rows; }
    const void* host_data() const override { return ptr; }
};

// A column that exists only on the device (synthetic 1 B-row tables never touch host memory).
struct DeviceColumn : public Column {
    TypeId type_id;
    size_t rows;
    DeviceColumn(TypeId t, bq_col* handle, size_t n, bool take_ownership);
    TypeId type() const override { return type_id; }
    size_t size() const override { return rows; }
    const void* host_data() const override { return nullptr; }
    bq_col* handle() const { return (device && device->handle) ? device->handle : borrowed_; }
private:
    bq_col* borrowed_ = nullptr;   // handle owned by the caller (bench.py keeps synthetic columns alive itself)
};

struct RecordBatch {
    std::vector<ColumnType> schema;
    std::vector<std::unique_ptr<Column>> columns;
    RecordBatch(std::vector<ColumnType> s) : schema(std::move(s)) { columns.reserve(schema.size()); }
    size_t num_rows() const { return columns.empty() ? 0 : columns[0]->size(); }
    size_t num_columns() const { return columns.size(); }
    template <typename T> void add_column(std::unique_ptr<ColumnVector<T>> col) { columns.push_back(std::move(col)); }
    Column* get_column(size_t index) const { return columns.at(index).get(); }
    const ColumnType& get_column_type(size_t index) const { return schema.at(index); }
};

// reference: include/exec/execution_types.hpp:11-34 — the batch contract of Operator::next
struct ColumnSlice {
    const void* data;
    TypeId type;
    size_t length;
    std::shared_ptr<void> owner;
};

struct ExecBatch {
    std::vector<ColumnSlice> columns;
    size_t length = 0;
    void clear() { columns.clear(); length = 0; }
};

template <typename T>
std::span<const T> get_col(const ExecBatch& batch, size_t i) {
    if (batch.columns[i].type != type_id_for<T>()) throw std::runtime_error("Type mismatch");
    return {reinterpret_cast<const T*>(batch.columns[i].data), batch.columns[i].length};
}

// reference: include/storage/dictionary.h:11-17, src/storage/dictionary.cpp:5-12 (first-seen ids from 0).
// An index makes get_or_add O(1); ids are assigned exactly as the reference assigns them.
class Dictionary {
public:
    std::vector<std::string> strings;
    StrId get_or_add(const std::string& s);
    const std::string& get(StrId id) const { return strings[id]; }
private:
    std::unordered_map<std::string, StrId> index_;
    size_t indexed_ = 0;
};

// reference: include/storage/table.h:14-30
struct TableColumn {
    std::string name;
    std::unique_ptr<Column> data;
};

struct Table {
    std::string name;
    std::vector<TableColumn> columns;
    std::shared_ptr<Dictionary> dict;
    size_t get_column_index(const std::string& col_name) const;
    const Column& get_column_data(const std::string& col_name) const;
};

// reference: include/catalog/catalog.h:16-62
struct ColumnStats {
    i64 min_i64 = 0, max_i64 = 0;
    f64 min_f64 = 0.0, max_f64 = 0.0;
    Date32 min_date = 0, max_date = 0;
    size_t ndv = 0;
};

struct ColumnMeta {
    std::string name;
    TypeId type;
    ColumnStats stats;
    ColumnMeta(std::string n, TypeId t, size_t ndv = 0) : name(std::move(n)), type(t) { stats.ndv = ndv; }
};

struct TableMeta {
    std::string name;
    std::vector<ColumnMeta> columns;
    size_t row_count = 0;
    TableMeta() = default;
    TableMeta(std::string n, std::vector<ColumnMeta> cols, size_t rows)
        : name(std::move(n)), columns(std::move(cols)), row_count(rows) {}
};

class Catalog {
    std::unordered_map<std::string, std::pair<Table, TableMeta>> tables_;
public:
    void register_table(Table table, TableMeta&& table_meta);
    OptionalRef<const Table> get_table_data(const std::string& name) const;
    OptionalRef<const TableMeta> get_table_meta(const std::string& name) const;
    std::vector<std::string> list_tables() const;
};

// ---- CSV ingest (host/csv_ingest.cpp) -----------------------------------------------------------------
// The reference's load_csv contract (include/storage/csv_loader.h:18-19, src/storage/csv_loader.cpp:7-166): header line,
// comma separated, no quoting; per column DATE32 (all values 8 characters, stoi in [19000000, 21000000]) else INT64 (all
// values stod-integral) else DOUBLE (all values stod-parsable) else STRING (dictionary ids in first-seen order); min / max /
// ndv recorded in the TableMeta.
std::pair<Table, TableMeta> load_csv(std::istream& stream);
std::pair<Table, TableMeta> load_csv(const std::string& filename);
// The same loader appending to a dictionary that other tables already use: string ids stay comparable
// across tables, so a join whose two sides both carry STRING columns decodes correctly - the reference gives every table a
// dictionary of its own and HashJoin keeps only the left one (src/exec/operator.cpp:694-704).
std::pair<Table, TableMeta> load_csv(std::istream& stream, std::shared_ptr<Dictionary> shared_dict);
std::pair<Table, TableMeta> load_csv(const std::string& filename, std::shared_ptr<Dictionary> shared_dict);

}  // namespace bosql
