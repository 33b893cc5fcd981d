// gpu_device.cpp — process-wide context, device mirrors of table columns, storage-layer methods.
#include "gpu_device.hpp"

#include <cstdlib>
#include <mutex>

namespace bosql {

// ---- storage (reference: src/storage/dictionary.cpp, src/storage/table.cpp, src/catalog/catalog.cpp) ----
StrId Dictionary::get_or_add(const std::string& s) {
    // `strings` is public and may have been filled directly: index whatever is new first (first occurrence wins,
    // like the reference's std::find)
    for (; indexed_ < strings.size(); ++indexed_) index_.emplace(strings[indexed_], static_cast<StrId>(indexed_));
    auto it = index_.find(s);
    if (it != index_.end()) return it->second;
    strings.push_back(s);
    StrId id = static_cast<StrId>(strings.size() - 1);
    index_.emplace(s, id);
    indexed_ = strings.size();
    return id;
}

size_t Table::get_column_index(const std::string& col_name) const {
    for (size_t i = 0; i < columns.size(); ++i)
        if (columns[i].name == col_name) return i;
    throw std::runtime_error("Column not found: " + col_name);
}

const Column& Table::get_column_data(const std::string& col_name) const { return *columns[get_column_index(col_name)].data; }

void Catalog::register_table(Table table, TableMeta&& table_meta) {
    std::string name = table_meta.name;
    tables_[name] = {std::move(table), std::move(table_meta)};
}

OptionalRef<const Table> Catalog::get_table_data(const std::string& name) const {
    auto it = tables_.find(name);
    if (it == tables_.end()) return {};
    return OptionalRef<const Table>(it->second.first);
}

OptionalRef<const TableMeta> Catalog::get_table_meta(const std::string& name) const {
    auto it = tables_.find(name);
    if (it == tables_.end()) return {};
    return OptionalRef<const TableMeta>(it->second.second);
}

std::vector<std::string> Catalog::list_tables() const {
    std::vector<std::string> out;
    for (const auto& kv : tables_) out.push_back(kv.first);
    return out;
}

DeviceMirror::~DeviceMirror() {
    if (handle) bq_col_free(nullptr, handle);
}

DeviceColumn::DeviceColumn(TypeId t, bq_col* handle, size_t n, bool take_ownership) : type_id(t), rows(n) {
    device = std::make_shared<DeviceMirror>();
    device->handle = take_ownership ? handle : nullptr;
    device->rows = n;
    borrowed_ = take_ownership ? nullptr : handle;
}

namespace gpu {

static bq_ctx* g_ctx = nullptr;
static std::mutex g_mu;

void init_context(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx) return;
    if (bq_ctx_create(device, &g_ctx)) throw std::runtime_error(bq_last_error());
}

bq_ctx* context() {
    if (!g_ctx) {
        int dev = 0;
        if (const char* e = std::getenv("BOSQL_DEVICE")) dev = std::atoi(e);
        else if (const char* r = std::getenv("LOCAL_RANK")) dev = std::atoi(r);
        init_context(dev);
    }
    return g_ctx;
}

void shutdown_context() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx) bq_ctx_destroy(g_ctx);
    g_ctx = nullptr;
}

void throw_last_error() { throw std::runtime_error(bq_last_error()); }

DevCol::~DevCol() {
    if (h && owns) bq_col_free(nullptr, h);
}

DevColPtr adopt(bq_col* h) { return std::make_shared<DevCol>(h, true); }

DevColPtr view_of(const DevColPtr& col, size_t begin, size_t end) {
    if (begin == 0 && end == col->rows()) return col;
    const size_t w = type_width(col->type());
    bq_col* h = nullptr;
    check(bq_col_wrap(context(), static_cast<int>(col->type()), static_cast<char*>(bq_col_ptr(col->h)) + begin * w, end - begin, &h));
    DevColPtr v = std::make_shared<DevCol>(h, true);      // owns the (non-owning) handle, not the memory
    v->parent = col;
    return v;
}

namespace {
struct PinnedPool {
    struct Buf { void* p; size_t bytes; };
    std::vector<Buf> free_list;
    size_t cached = 0;
    static constexpr size_t kMaxCached = 8ull << 30;
    ~PinnedPool() {
        for (auto& b : free_list) bq_host_free(b.p);
    }
    void* take(size_t& bytes) {
        // small results (a few groups) share 64 KB buffers, large ones are rounded to 2 MB: either way the next query
        // finds its buffer in the free list and no page is pinned on the query path
        const size_t grain = bytes <= (64u << 10) ? (64u << 10) : (2u << 20);
        bytes = (bytes + grain - 1) / grain * grain;
        size_t best = free_list.size();
        for (size_t i = 0; i < free_list.size(); ++i)
            if (free_list[i].bytes >= bytes && free_list[i].bytes <= 2 * bytes &&
                (best == free_list.size() || free_list[i].bytes < free_list[best].bytes)) best = i;
        if (best != free_list.size()) {
            Buf b = free_list[best];
            free_list.erase(free_list.begin() + static_cast<long>(best));
            cached -= b.bytes;
            bytes = b.bytes;
            return b.p;
        }
        void* p = nullptr;
        check(bq_host_alloc(bytes, &p));
        return p;
    }
    void give(void* p, size_t bytes) {
        if (cached + bytes > kMaxCached) {
            bq_host_free(p);
            return;
        }
        free_list.push_back({p, bytes});
        cached += bytes;
    }
};
PinnedPool& pinned_pool() {
    static PinnedPool* pool = new PinnedPool();      // leaked on purpose: results may outlive static destruction order
    return *pool;
}
}  // namespace

std::shared_ptr<void> host_buffer(size_t bytes) {
    // always pinned: a copy into pageable memory is staged by the driver and blocks the calling thread (0.17 ms measured for
    // a 20-row result); pinned buffers let all columns of a result be enqueued and waited for once
    size_t got = bytes ? bytes : 1;
    void* p = pinned_pool().take(got);
    return std::shared_ptr<void>(p, [got](void* q) { pinned_pool().give(q, got); });
}

DevColPtr mirror_of(const Column& col) {
    if (const auto* dc = dynamic_cast<const DeviceColumn*>(&col)) {
        bq_col* h = dc->handle();
        if (!h) throw std::runtime_error("device column without a handle");
        return std::make_shared<DevCol>(h, false);
    }
    const void* host = col.host_data();
    const size_t n = col.size();
    if (!col.device || col.device->rows != n || col.device->host_data != host || !col.device->handle) {
        auto m = std::make_shared<DeviceMirror>();
        check(bq_col_upload(context(), static_cast<int>(col.type()), host, n, &m->handle));
        m->rows = n;
        m->host_data = host;
        col.device = m;
    }
    // the mirror stays owned by the Column; keep it alive through the DevCol's lifetime
    auto keep = col.device;
    return DevColPtr(new DevCol(keep->handle, false), [keep](DevCol* p) { delete p; });
}

DeviceRelationPtr relation_from(bq_rel* rel) {
    auto out = std::make_shared<DeviceRelation>();
    out->rows = bq_rel_rows(rel);
    // the shell owns its columns; take them over one by one by re-wrapping through bq_rel_create's inverse:
    // bq_rel_release() hands the handles back without freeing them.
    int n = bq_rel_cols(rel);
    std::vector<bq_col*> hs(n);
    bq_rel_release(rel, hs.data());
    for (int i = 0; i < n; ++i) out->cols.push_back(adopt(hs[i]));
    return out;
}

void split_conjuncts(const Expr* e, Dictionary* dict, std::vector<Conjunct>& out) {
    if (e->type == ExprType::BINARY_OP && e->op == BinaryOp::AND) {
        split_conjuncts(e->left.get(), dict, out);
        split_conjuncts(e->right.get(), dict, out);
        return;
    }
    out.push_back({e->clone(), dict});
}

}  // namespace gpu
}  // namespace bosql
