// capi.cpp — include/bosql_b200_exec.h over the C++ operator layer.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <typeinfo>

#include "bosql_b200_exec.h"
#include "bosql_operator.hpp"
#include "bosql_types.hpp"
#include "exchange.hpp"
#include "gpu_device.hpp"

using namespace bosql;

struct bqx_dict {
    std::shared_ptr<Dictionary> dict = std::make_shared<Dictionary>();
};
struct bqx_table {
    Table table;
    std::vector<ColumnMeta> metas;
    size_t rows = 0;
};
struct bqx_catalog {
    Catalog catalog;
};
struct bqx_plan {
    std::unique_ptr<LogicalOp> logical;
    std::unique_ptr<Operator> root;
    ExecBatch batch;
    std::string kind;
};
struct bqx_result {
    std::vector<std::vector<unsigned char>> cols;
    std::vector<std::shared_ptr<void>> whole;       // set instead of cols when the root handed over its host copy
    size_t rows = 0;
    double seconds = 0.0;
};

namespace {

thread_local std::string g_err;

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

template <typename T>
std::unique_ptr<Column> host_column(const void* data, size_t n) {
    const T* p = static_cast<const T*>(data);
    return std::make_unique<ColumnVector<T>>(std::vector<T>(p, p + n));
}

ParseOptions parse_options(unsigned flags) {
    ParseOptions o;
    o.between = flags & 1u;
    o.decimal_literals = flags & 2u;
    o.negative_literals = flags & 4u;
    o.keywords_any_case = flags & 8u;
    return o;
}

const char* kind_of(const Operator* op) {
    if (dynamic_cast<const ColumnarScan*>(op)) return "ColumnarScan";
    if (dynamic_cast<const Selection*>(op)) return "Selection";
    if (dynamic_cast<const Project*>(op)) return "Project";
    if (dynamic_cast<const HashJoin*>(op)) return "HashJoin";
    if (dynamic_cast<const HashAggregate*>(op)) return "HashAggregate";
    if (dynamic_cast<const OrderBy*>(op)) return "OrderBy";
    if (dynamic_cast<const Limit*>(op)) return "Limit";
    return "Operator";
}

}  // namespace

extern "C" {

const char* bqx_last_error(void) { return g_err.c_str(); }

int bqx_init(int device) {
    return guarded([&] { gpu::init_context(device); });
}

int bqx_set_exchange(const bqx_exchange* x) {
    return guarded([&] {
        gpu::Exchange& e = gpu::exchange();
        e.native = false;
        if (!x || x->world <= 1) {
            e.active = false;
            e.fn = bqx_exchange{};
            return;
        }
        if (x->rank < 0 || x->rank >= x->world) throw std::runtime_error("exchange: rank out of range");
        if (!x->all_gather || !x->all_gather_v || !x->all_to_all_v || !x->all_reduce_sum_u32 || !x->host_all_gather_i64)
            throw std::runtime_error("exchange: every collective must be supplied");
        e.fn = *x;
        e.active = true;
    });
}

namespace {
// the native exchange table: plain functions over bq_comm_* on this process's context (`user` is unused)
int native_all_gather(void*, const void* send, void* recv, size_t bytes, void*) { return bq_comm_all_gather(gpu::context(), send, recv, bytes); }
int native_all_gather_v(void*, const void* send, void* recv, const int64_t* bytes_by_rank, void*) {
    return bq_comm_all_gather_v(gpu::context(), send, recv, bytes_by_rank);
}
int native_all_to_all_v(void*, const void* send, const int64_t* sb, void* recv, const int64_t* rb, void*) {
    return bq_comm_all_to_all_v(gpu::context(), send, sb, recv, rb);
}
int native_sum_u32(void*, void* buf, size_t words, void*) { return bq_comm_all_reduce_sum_u32(gpu::context(), buf, words); }
int native_host_gather(void*, const int64_t* mine, int32_t n, int64_t* all) { return bq_comm_host_all_gather_i64(gpu::context(), mine, n, all); }

void install_native(int world, int rank, const void* id128, int keep_sharded) {
    if (bq_comm_init(gpu::context(), world, rank, id128)) gpu::throw_last_error();
    gpu::Exchange& e = gpu::exchange();
    e.fn = bqx_exchange{};
    e.fn.world = world;
    e.fn.rank = rank;
    e.fn.keep_sharded = keep_sharded;
    e.fn.all_gather = native_all_gather;
    e.fn.all_gather_v = native_all_gather_v;
    e.fn.all_to_all_v = native_all_to_all_v;
    e.fn.all_reduce_sum_u32 = native_sum_u32;
    e.fn.host_all_gather_i64 = native_host_gather;
    e.native = true;
    e.active = world > 1;
}
}  // namespace

int bqx_comm_unique_id(void* id128) {
    return guarded([&] {
        if (bq_comm_unique_id(id128)) gpu::throw_last_error();
    });
}

int bqx_comm_init(int world, int rank, const void* id128, int keep_sharded) {
    return guarded([&] { install_native(world, rank, id128, keep_sharded); });
}

int bqx_comm_init_file(const char* path, int world, int rank, int keep_sharded) {
    return guarded([&] {
        if (world < 0) world = std::getenv("WORLD_SIZE") ? std::atoi(std::getenv("WORLD_SIZE")) : 1;
        if (rank < 0) rank = std::getenv("RANK") ? std::atoi(std::getenv("RANK")) : 0;
        unsigned char id[BQ_COMM_ID_BYTES];
        const std::string file(path);
        if (rank == 0) {
            if (bq_comm_unique_id(id)) gpu::throw_last_error();
            const std::string tmp = file + ".tmp";
            FILE* f = std::fopen(tmp.c_str(), "wb");
            if (!f || std::fwrite(id, 1, sizeof id, f) != sizeof id) throw std::runtime_error("cannot write the rendezvous file " + tmp);
            std::fclose(f);
            if (std::rename(tmp.c_str(), file.c_str())) throw std::runtime_error("cannot publish the rendezvous file " + file);
        } else {
            bool got = false;
            for (int tries = 0; tries < 6000 && !got; ++tries) {           // up to a minute
                if (FILE* f = std::fopen(file.c_str(), "rb")) {
                    got = std::fread(id, 1, sizeof id, f) == sizeof id;
                    std::fclose(f);
                }
                if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(10));
            }
            if (!got) throw std::runtime_error("rendezvous file " + file + " did not appear");
        }
        install_native(world, rank, id, keep_sharded);
    });
}

int bqx_comm_stats(uint64_t* calls5, uint64_t* bytes_sent) {
    return guarded([&] {
        if (bq_comm_stats(gpu::context(), calls5, bytes_sent)) gpu::throw_last_error();
    });
}

int bqx_exchange_keep_sharded(int on) {
    gpu::exchange().fn.keep_sharded = on ? 1 : 0;
    return 0;
}

bq_ctx* bqx_context(void) {
    try {
        return gpu::context();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

bqx_dict* bqx_dict_create(void) { return new bqx_dict(); }
void bqx_dict_destroy(bqx_dict* d) { delete d; }
uint32_t bqx_dict_get_or_add(bqx_dict* d, const char* s) { return d->dict->get_or_add(s); }
size_t bqx_dict_size(const bqx_dict* d) { return d->dict->strings.size(); }
const char* bqx_dict_get(const bqx_dict* d, uint32_t id) {
    return id < d->dict->strings.size() ? d->dict->strings[id].c_str() : nullptr;
}

bqx_catalog* bqx_catalog_create(void) { return new bqx_catalog(); }
void bqx_catalog_destroy(bqx_catalog* c) { delete c; }

bqx_table* bqx_table_create(const char* name, bqx_dict* dict) {
    auto* t = new bqx_table();
    t->table.name = name;
    t->table.dict = dict ? dict->dict : std::make_shared<Dictionary>();
    return t;
}

int bqx_table_add_column(bqx_table* t, const char* name, int type, const void* data, size_t n) {
    return guarded([&] {
        if (!t->table.columns.empty() && n != t->rows) throw std::runtime_error("column length mismatch");
        std::unique_ptr<Column> col;
        switch (static_cast<TypeId>(type)) {
            case TypeId::INT64: col = host_column<int64_t>(data, n); break;
            case TypeId::DOUBLE: col = host_column<double>(data, n); break;
            case TypeId::STRING: col = host_column<uint32_t>(data, n); break;
            case TypeId::DATE32: col = host_column<int32_t>(data, n); break;
            default: throw std::runtime_error("Unknown column type");
        }
        t->table.columns.push_back({name, std::move(col)});
        t->metas.emplace_back(name, static_cast<TypeId>(type));
        t->rows = n;
    });
}

int bqx_table_add_borrowed_column(bqx_table* t, const char* name, int type, const void* data, size_t n) {
    return guarded([&] {
        if (!t->table.columns.empty() && n != t->rows) throw std::runtime_error("column length mismatch");
        if (type < 0 || type > 3) throw std::runtime_error("Unknown column type");
        t->table.columns.push_back({name, std::make_unique<BorrowedColumn>(static_cast<TypeId>(type), data, n)});
        t->metas.emplace_back(name, static_cast<TypeId>(type));
        t->rows = n;
    });
}

int bqx_table_add_device_column(bqx_table* t, const char* name, bq_col* col, int take_ownership) {
    return guarded([&] {
        const size_t n = bq_col_size(col);
        if (!t->table.columns.empty() && n != t->rows) throw std::runtime_error("column length mismatch");
        const TypeId type = static_cast<TypeId>(bq_col_type(col));
        t->table.columns.push_back({name, std::make_unique<DeviceColumn>(type, col, n, take_ownership != 0)});
        t->metas.emplace_back(name, type);
        t->rows = n;
    });
}

int bqx_table_set_stats(bqx_table* t, const char* column, int64_t min_i, int64_t max_i, double min_f, double max_f, size_t ndv) {
    return guarded([&] {
        for (auto& m : t->metas)
            if (m.name == column) {
                m.stats.ndv = ndv;
                m.stats.min_f64 = min_f;
                m.stats.max_f64 = max_f;
                if (m.type == TypeId::DATE32) {
                    m.stats.min_date = static_cast<Date32>(min_i);
                    m.stats.max_date = static_cast<Date32>(max_i);
                } else {
                    m.stats.min_i64 = min_i;
                    m.stats.max_i64 = max_i;
                }
                return;
            }
        throw std::runtime_error(std::string("Column not found: ") + column);
    });
}

int bqx_catalog_register(bqx_catalog* c, bqx_table* tp) {
    return guarded([&] {
        std::unique_ptr<bqx_table> t(tp);
        TableMeta meta(t->table.name, std::move(t->metas), t->rows);
        c->catalog.register_table(std::move(t->table), std::move(meta));
    });
}

int bqx_catalog_load_csv(bqx_catalog* c, const char* path, const char* name) {
    return guarded([&] {
        auto [table, meta] = load_csv(std::string(path));
        table.name = name;
        meta.name = name;
        c->catalog.register_table(std::move(table), std::move(meta));
    });
}

int bqx_catalog_table_info(bqx_catalog* c, const char* name, size_t* rows, size_t* ncols) {
    return guarded([&] {
        OptionalRef<const TableMeta> m = c->catalog.get_table_meta(name);
        if (!m.has_value()) throw std::runtime_error(std::string("Table not found: ") + name);
        *rows = m->row_count;
        *ncols = m->columns.size();
    });
}

int bqx_catalog_column_info(bqx_catalog* c, const char* name, size_t i, const char** col_name, int* type, const void** data,
                            int64_t* min_i, int64_t* max_i, double* min_f, double* max_f, size_t* ndv) {
    return guarded([&] {
        OptionalRef<const TableMeta> m = c->catalog.get_table_meta(name);
        OptionalRef<const Table> t = c->catalog.get_table_data(name);
        if (!m.has_value() || !t.has_value() || i >= m->columns.size()) throw std::runtime_error("bad column");
        const ColumnMeta& cm = m->columns[i];
        *col_name = cm.name.c_str();
        *type = static_cast<int>(cm.type);
        *ndv = cm.stats.ndv;
        *min_f = cm.stats.min_f64;
        *max_f = cm.stats.max_f64;
        *min_i = cm.type == TypeId::DATE32 ? cm.stats.min_date : cm.stats.min_i64;
        *max_i = cm.type == TypeId::DATE32 ? cm.stats.max_date : cm.stats.max_i64;
        *data = t->columns[i].data->host_data();
    });
}

size_t bqx_catalog_dict_size(bqx_catalog* c, const char* name) {
    OptionalRef<const Table> t = c->catalog.get_table_data(name);
    return (t.has_value() && t->dict) ? t->dict->strings.size() : 0;
}

const char* bqx_catalog_dict_get(bqx_catalog* c, const char* name, uint32_t id) {
    OptionalRef<const Table> t = c->catalog.get_table_data(name);
    if (!t.has_value() || !t->dict || id >= t->dict->strings.size()) return nullptr;
    return t->dict->strings[id].c_str();
}

int bqx_catalog_evict_device(bqx_catalog* c, const char* table) {
    return guarded([&] {
        OptionalRef<const Table> t = c->catalog.get_table_data(table);
        if (!t.has_value()) throw std::runtime_error(std::string("Table not found: ") + table);
        for (const auto& col : t->columns)
            if (col.data->host_data()) col.data->device.reset();
    });
}

int bqx_plan_create(bqx_catalog* c, const char* sql, unsigned parse_flags, bqx_plan** out) {
    return guarded([&] {
        SelectStmt stmt = parse_sql(sql, parse_options(parse_flags));
        auto p = std::make_unique<bqx_plan>();
        LogicalPlanner planner;
        p->logical = planner.build_logical_plan(stmt);
        p->root = build_physical_plan(p->logical.get(), c->catalog);
        p->kind = kind_of(p->root.get());
        *out = p.release();
    });
}

void bqx_plan_destroy(bqx_plan* p) { delete p; }
size_t bqx_plan_columns(const bqx_plan* p) { return p->root->output_names().size(); }
const char* bqx_plan_column_name(const bqx_plan* p, size_t i) { return p->root->output_names().at(i).c_str(); }
int bqx_plan_column_type(const bqx_plan* p, size_t i) { return static_cast<int>(p->root->output_types().at(i)); }
int bqx_plan_has_dict(const bqx_plan* p) { return p->root->dictionary() != nullptr; }
const char* bqx_plan_dict_get(const bqx_plan* p, uint32_t id) {
    Dictionary* d = p->root->dictionary();
    return (d && id < d->strings.size()) ? d->strings[id].c_str() : nullptr;
}
const char* bqx_plan_root_kind(const bqx_plan* p) { return p->kind.c_str(); }

int bqx_plan_open(bqx_plan* p) {
    return guarded([&] { p->root->open(); });
}

int bqx_plan_next(bqx_plan* p, const void** col_data, size_t n_cols, size_t* rows, int* end_of_stream) {
    return guarded([&] {
        if (!p->root->next(p->batch)) {
            *rows = 0;
            *end_of_stream = 1;
            return;
        }
        *end_of_stream = 0;
        *rows = p->batch.length;
        for (size_t i = 0; i < n_cols && i < p->batch.columns.size(); ++i) col_data[i] = p->batch.columns[i].data;
    });
}

int bqx_plan_close(bqx_plan* p) {
    return guarded([&] { p->root->close(); });
}

int bqx_plan_run(bqx_plan* p, bqx_result** out) {
    return guarded([&] {
        auto res = std::make_unique<bqx_result>();
        const auto& types = p->root->output_types();
        res->cols.resize(types.size());
        auto t0 = std::chrono::steady_clock::now();
        p->root->open();
        ExecBatch batch;
        bool first = true;
        while (p->root->next(batch)) {
            if (first && p->root->host_result(res->whole, res->rows)) break;     // zero-copy: the batches alias these columns
            first = false;
            for (size_t j = 0; j < batch.columns.size() && j < res->cols.size(); ++j) {
                const auto& s = batch.columns[j];
                const auto* b = static_cast<const unsigned char*>(s.data);
                res->cols[j].insert(res->cols[j].end(), b, b + type_width(s.type) * batch.length);
            }
            res->rows += batch.length;
        }
        p->root->close();
        res->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        *out = res.release();
    });
}

int bqx_plan_run_device(bqx_plan* p, bq_rel** out) {
    return guarded([&] {
        p->root->open();
        gpu::DeviceRelationPtr rel;
        try {
            rel = p->root->device_result();
        } catch (...) {
            p->root->close();
            throw;
        }
        p->root->close();
        // An owning relation for the caller.  Columns that this result alone holds (the output of an aggregate, a sort, a
        // gather) are handed over as they are; anything shared - a scan's table columns, a view into one - is copied
        // device to device, so the table keeps its columns.
        const bool sole_owner = rel.use_count() == 1;
        std::vector<bq_col*> cols;
        for (auto& c : rel->cols) {
            bq_col* v = nullptr;
            if (sole_owner && c.use_count() == 1 && c->owns && !c->parent && bq_col_owns(c->h) && bq_col_size(c->h) == rel->rows) {
                v = c->h;
                c->owns = false;              // the handle leaves with the result
            } else if (bq_slice(gpu::context(), c->h, 0, rel->rows, &v)) {
                for (bq_col* done : cols) bq_col_free(gpu::context(), done);      // handed-over and copied columns alike are ours by now
                gpu::throw_last_error();
            }
            cols.push_back(v);
        }
        gpu::check(bq_rel_create(gpu::context(), cols.data(), static_cast<int>(cols.size()), out));
    });
}

size_t bqx_result_rows(const bqx_result* r) { return r->rows; }
size_t bqx_result_cols(const bqx_result* r) { return r->cols.size(); }
double bqx_result_seconds(const bqx_result* r) { return r->seconds; }
const void* bqx_result_data(const bqx_result* r, size_t i) { return r->whole.empty() ? r->cols.at(i).data() : r->whole.at(i).get(); }
void bqx_result_free(bqx_result* r) { delete r; }

int bqx_explain(const char* sql, unsigned parse_flags, char* out, size_t cap) {
    return guarded([&] {
        SelectStmt stmt = parse_sql(sql, parse_options(parse_flags));
        LogicalPlanner planner;
        std::string s = planner.build_logical_plan(stmt)->to_string();
        std::strncpy(out, s.c_str(), cap - 1);
        out[cap - 1] = 0;
    });
}

}  // extern "C"
