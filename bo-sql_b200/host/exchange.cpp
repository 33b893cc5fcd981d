// exchange.cpp — see exchange.hpp.
#include "exchange.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace bosql::gpu {

namespace {

size_t width_of(TypeId t) { return (t == TypeId::INT64 || t == TypeId::DOUBLE) ? 8 : 4; }
size_t align16(size_t b) { return (b + 15) & ~static_cast<size_t>(15); }

void xcheck(int rc, const char* what) {
    if (rc) throw std::runtime_error(std::string("exchange callback failed: ") + what);
}

DevColPtr alloc_bytes(size_t bytes) {
    bq_col* h = nullptr;
    check(bq_col_alloc(context(), BQ_INT64, (bytes + 7) / 8, &h));
    return adopt(h);
}

DevColPtr alloc_col(TypeId t, size_t n) {
    bq_col* h = nullptr;
    check(bq_col_alloc(context(), static_cast<int>(t), n, &h));
    return adopt(h);
}

char* ptr_of(const DevColPtr& c) { return static_cast<char*>(bq_col_ptr(c->h)); }

}  // namespace

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

PhaseTrace::PhaseTrace() {
    const char* e = std::getenv("BOSQL_TRACE");
    on = e && *e && *e != '0';
    if (on) {
        bq_ctx_sync(context());
        last = now_ms();
    }
}

void PhaseTrace::mark(const char* what) {
    if (!on) return;
    bq_ctx_sync(context());
    const double t = now_ms();
    size_t reserved = 0, used = 0;
    bq_ctx_pool_stats(context(), &reserved, &used);
    std::fprintf(stderr, "[bosql trace rank %d] %-28s %8.3f ms   pool %.2f / %.2f GiB   peer blocks mapped %zu\n",
                 exchange().active ? exchange().rank() : 0, what, t - last, used / 1073741824.0, reserved / 1073741824.0,
                 bq_ctx_ipc_mappings(context()));
    last = now_ms();
}

Exchange& exchange() {
    static Exchange x;
    return x;
}

void agree_on(const std::function<void()>& step) {
    Exchange& x = exchange();
    if (!x.active) {
        step();
        return;
    }
    std::string err;
    try {
        step();
    } catch (const std::exception& e) {
        err = e.what();
    }
    int64_t worst = 0;
    for (int64_t f : x.host_gather({err.empty() ? 0 : (err.find("Division by zero") != std::string::npos ? 1 : 2)})) worst = std::max(worst, f);
    if (!err.empty()) throw std::runtime_error(err);
    if (worst == 1) throw std::runtime_error("Division by zero");      // src/exec/expression.cpp:52, raised by another rank's rows
    if (worst) throw std::runtime_error("expression evaluation failed on another rank");
}

std::vector<int64_t> Exchange::host_gather(const std::vector<int64_t>& mine) {
    std::vector<int64_t> all(mine.size() * static_cast<size_t>(world()));
    check(bq_ctx_sync(context()));
    xcheck(fn.host_all_gather_i64(fn.user, mine.data(), static_cast<int32_t>(mine.size()), all.data()), "host_all_gather_i64");
    return all;
}

int64_t Exchange::host_sum(int64_t v) {
    int64_t s = 0;
    for (int64_t x : host_gather({v})) s += x;
    return s;
}

bool Exchange::host_all(bool v) {
    for (int64_t x : host_gather({v ? 1 : 0}))
        if (!x) return false;
    return true;
}

void Exchange::minmax(int64_t& lo, int64_t& hi) {
    auto all = host_gather({lo, hi});
    bool any = false;
    for (int r = 0; r < world(); ++r) {
        int64_t l = all[2 * r], h = all[2 * r + 1];
        if (l > h) continue;
        if (!any) { lo = l; hi = h; any = true; }
        else { lo = std::min(lo, l); hi = std::max(hi, h); }
    }
    if (!any) { lo = 0; hi = -1; }
}

void Exchange::sum_words(void* device_words, size_t n_words) {
    xcheck(fn.all_reduce_sum_u32(fn.user, device_words, n_words, bq_ctx_stream(context())), "all_reduce_sum_u32");
}

DevColPtr Exchange::all_gather_column(const DevColPtr& col, size_t rows, const std::vector<int64_t>& rows_by_rank) {
    const TypeId t = col->type();
    const size_t w = width_of(t);
    size_t total = 0;
    std::vector<int64_t> bytes(rows_by_rank.size());
    for (size_t r = 0; r < rows_by_rank.size(); ++r) {
        bytes[r] = rows_by_rank[r] * static_cast<int64_t>(w);
        total += static_cast<size_t>(rows_by_rank[r]);
    }
    if (static_cast<int64_t>(rows) != rows_by_rank[static_cast<size_t>(rank())]) throw std::runtime_error("internal: all_gather_column row count mismatch");
    DevColPtr out = alloc_col(t, total);
    xcheck(fn.all_gather_v(fn.user, ptr_of(col), ptr_of(out), bytes.data(), bq_ctx_stream(context())), "all_gather_v");
    return out;
}

GatheredPartials::~GatheredPartials() {
    for (bq_rel* r : parts) bq_rel_free(context(), r);     // the columns are non-owning views into `buffer`
}

void gather_partials(const DeviceRelation* local, int error_flags, bool has_key, TypeId key_type, int64_t capacity,
                     GatheredPartials& out) {
    Exchange& x = exchange();
    bq_ctx* ctx = context();
    std::vector<TypeId> types;
    if (has_key) types.push_back(key_type);
    types.insert(types.end(), {TypeId::INT64, TypeId::DOUBLE, TypeId::DOUBLE});
    const size_t cnt_col = has_key ? 1 : 0;
    size_t my_rows = error_flags ? 1 : (local ? local->rows : 0);
    if (capacity >= 0 && static_cast<int64_t>(my_rows) > capacity) throw std::runtime_error("internal: partial state exceeds its capacity");

    const int W = x.world();
    std::vector<int64_t> rows_by_rank(static_cast<size_t>(W), capacity);
    if (capacity < 0) rows_by_rank = x.host_gather({static_cast<int64_t>(my_rows)});
    auto block_of = [&](size_t n) {
        size_t b = 0;
        for (TypeId t : types) b += align16(width_of(t) * n);
        return b;
    };
    auto offsets_of = [&](size_t n) {
        std::vector<size_t> off;
        size_t b = 0;
        for (TypeId t : types) {
            off.push_back(b);
            b += align16(width_of(t) * n);
        }
        return off;
    };

    const size_t my_n = static_cast<size_t>(rows_by_rank[static_cast<size_t>(x.rank())]);
    const size_t my_block = block_of(my_n);
    DevColPtr send = alloc_bytes(std::max<size_t>(my_block, 16));
    check(bq_zero_bytes(ctx, ptr_of(send), my_block));
    auto my_off = offsets_of(my_n);
    DevColPtr poison;
    if (error_flags) {
        int64_t neg = -static_cast<int64_t>(error_flags);
        bq_col* h = nullptr;
        check(bq_col_upload(ctx, BQ_INT64, &neg, 1, &h));
        poison = adopt(h);
        check(bq_copy_bytes(ctx, ptr_of(send) + my_off[cnt_col], ptr_of(poison), 8));
    } else if (local && local->rows) {
        for (size_t c = 0; c < types.size(); ++c)
            check(bq_copy_bytes(ctx, ptr_of(send) + my_off[c], bq_col_ptr(local->cols[c]->h), width_of(types[c]) * local->rows));
    }

    size_t total = 0;
    std::vector<int64_t> bytes(static_cast<size_t>(W));
    std::vector<size_t> base(static_cast<size_t>(W));
    for (int r = 0; r < W; ++r) {
        base[static_cast<size_t>(r)] = total;
        bytes[static_cast<size_t>(r)] = static_cast<int64_t>(block_of(static_cast<size_t>(rows_by_rank[static_cast<size_t>(r)])));
        total += static_cast<size_t>(bytes[static_cast<size_t>(r)]);
    }
    out.buffer = alloc_bytes(std::max<size_t>(total, 16));
    void* stream = bq_ctx_stream(ctx);
    if (capacity >= 0) xcheck(x.fn.all_gather(x.fn.user, ptr_of(send), ptr_of(out.buffer), my_block, stream), "all_gather");
    else xcheck(x.fn.all_gather_v(x.fn.user, ptr_of(send), ptr_of(out.buffer), bytes.data(), stream), "all_gather_v");

    for (int r = 0; r < W; ++r) {
        const size_t n = static_cast<size_t>(rows_by_rank[static_cast<size_t>(r)]);
        auto off = offsets_of(n);
        std::vector<bq_col*> views;
        for (size_t c = 0; c < types.size(); ++c) {
            bq_col* v = nullptr;
            check(bq_col_wrap(ctx, static_cast<int>(types[c]), ptr_of(out.buffer) + base[static_cast<size_t>(r)] + off[c], n, &v));
            views.push_back(v);
        }
        bq_rel* rel = nullptr;
        check(bq_rel_create(ctx, views.data(), static_cast<int>(views.size()), &rel));
        out.parts.push_back(rel);
    }
    // `send` may be released here: the allocator is stream ordered, and the collective was enqueued on the same stream
}

void all_gather_fold_state(bq_agg_state* state) {
    Exchange& x = exchange();
    bq_ctx* ctx = context();
    void* mine = nullptr;
    size_t bytes = 0;
    check(bq_agg_state_dense(state, &mine, &bytes));
    if (!mine) throw std::runtime_error("internal: a hash-table state cannot be exchanged as a dense block");
    DevColPtr all = alloc_bytes(bytes * static_cast<size_t>(x.world()));
    xcheck(x.fn.all_gather(x.fn.user, mine, ptr_of(all), bytes, bq_ctx_stream(ctx)), "all_gather");
    check(bq_agg_state_fold(ctx, state, ptr_of(all), x.world()));
    // `all` may be released here: the allocator is stream ordered and the fold was enqueued on the same stream
}

// Partition count shared by both shuffle implementations: a power of two >= world (x8 when world is not a power of two, so
// the contiguous runs handed to the ranks stay balanced); hash bits [40, 40 + log2 P) are disjoint from the bits the local
// tables use (top bits: L2 partition, bottom bits: slot).
static int shuffle_log2_parts(int W) {
    int log2p = 0;
    while ((1 << log2p) < W) ++log2p;
    if ((1 << log2p) != W) log2p += 3;
    if (log2p > 10) throw std::runtime_error("too many ranks for one partition pass");
    return log2p;
}

static Shuffled shuffle_collective(const DevColPtr& key, const std::vector<DevColPtr>& payload, size_t rows) {
    Exchange& x = exchange();
    bq_ctx* ctx = context();
    const int W = x.world();
    // the two-step form: partition into a send buffer (one contiguous run per rank), then the host's all-to-all
    const int log2p = shuffle_log2_parts(W);
    const int P = 1 << log2p;
    std::vector<const bq_col*> pay;
    for (const auto& c : payload) pay.push_back(c->h);
    bq_col *pk = nullptr, *off = nullptr, *pp[2] = {nullptr, nullptr};
    PhaseTrace trace;
    check(bq_partition(ctx, key->h, pay.data(), static_cast<int>(pay.size()), 0, rows, log2p, 40, &pk, pp, &off));
    trace.mark("shuffle: partition by rank");
    DevColPtr part_key = adopt(pk), offsets = adopt(off);
    std::vector<DevColPtr> part_pay;
    for (size_t i = 0; i < payload.size(); ++i) part_pay.push_back(adopt(pp[i]));
    std::vector<int64_t> host_off(static_cast<size_t>(P) + 1);
    check(bq_col_read(ctx, offsets->h, 0, static_cast<size_t>(P) + 1, host_off.data()));

    std::vector<int64_t> send_rows(static_cast<size_t>(W));
    for (int r = 0; r < W; ++r) {
        const size_t p0 = static_cast<size_t>(r) * P / W, p1 = static_cast<size_t>(r + 1) * P / W;
        send_rows[static_cast<size_t>(r)] = host_off[p1] - host_off[p0];
    }
    auto matrix = x.host_gather(send_rows);                 // matrix[s*W + d] = rows rank s sends to rank d
    std::vector<int64_t> recv_rows(static_cast<size_t>(W));
    size_t total = 0;
    for (int s = 0; s < W; ++s) {
        recv_rows[static_cast<size_t>(s)] = matrix[static_cast<size_t>(s) * W + static_cast<size_t>(x.rank())];
        total += static_cast<size_t>(recv_rows[static_cast<size_t>(s)]);
    }
    if (total > 0xFFFFFFFFull) throw std::runtime_error("a rank would own more than 2^32 rows after the shuffle");

    Shuffled out;
    out.rows = total;
    void* stream = bq_ctx_stream(ctx);
    auto move = [&](const DevColPtr& src) {
        const size_t w = width_of(src->type());
        std::vector<int64_t> sb(static_cast<size_t>(W)), rb(static_cast<size_t>(W));
        for (int r = 0; r < W; ++r) {
            sb[static_cast<size_t>(r)] = send_rows[static_cast<size_t>(r)] * static_cast<int64_t>(w);
            rb[static_cast<size_t>(r)] = recv_rows[static_cast<size_t>(r)] * static_cast<int64_t>(w);
        }
        DevColPtr dst = alloc_col(src->type(), total);
        trace.mark("shuffle:   alloc recv");
        xcheck(x.fn.all_to_all_v(x.fn.user, ptr_of(src), sb.data(), ptr_of(dst), rb.data(), stream), "all_to_all_v");
        return dst;
    };
    trace.mark("shuffle: size exchange");
    out.key = move(part_key);
    trace.mark("shuffle:   all-to-all key");
    for (const auto& c : part_pay) out.payload.push_back(move(c));
    trace.mark("shuffle:   all-to-all payload");
    return out;
}

// The shuffle as ONE pass: count, agree on sizes, then the partitioning kernel writes every row straight into the owning
// rank's receive buffer over NVLink (peer memory mapped through CUDA IPC) - no send buffer, no collective on the data path.
static Shuffled shuffle_peer_write(const DevColPtr& key, const std::vector<DevColPtr>& payload, size_t rows,
                                   const std::vector<int64_t>& hot_keys, bool keep_hot_local) {
    Exchange& x = exchange();
    bq_ctx* ctx = context();
    const int W = x.world(), me = x.rank();
    const int log2p = shuffle_log2_parts(W);
    const int P = 1 << log2p;
    PhaseTrace trace;
    const bool hot = !hot_keys.empty();
    const size_t tables = hot ? 2 * static_cast<size_t>(P) : static_cast<size_t>(P);      // entry P = the hot partition
    std::vector<int64_t> counts(tables);
    bq_part_plan* plan = nullptr;
    if (hot) check(bq_partition_count_hot(ctx, key->h, 0, rows, log2p, 40, hot_keys.data(), static_cast<int>(hot_keys.size()), counts.data(), &plan));
    else check(bq_partition_count(ctx, key->h, 0, rows, log2p, 40, counts.data(), &plan));
    struct PlanGuard {
        bq_part_plan* p;
        ~PlanGuard() { bq_part_plan_free(p); }
    } guard{plan};
    trace.mark("shuffle: count");

    auto first_part = [&](int r) { return static_cast<size_t>(r) * P / W; };
    std::vector<int64_t> send_rows(static_cast<size_t>(W), 0);
    for (int r = 0; r < W; ++r)
        for (size_t q = first_part(r); q < first_part(r + 1); ++q) send_rows[static_cast<size_t>(r)] += counts[q];
    const int64_t my_hot = hot ? counts[static_cast<size_t>(P)] : 0;
    send_rows.push_back(my_hot);
    const size_t M = static_cast<size_t>(W) + 1;
    auto matrix = x.host_gather(send_rows);                 // matrix[s*M + d] = rows rank s sends to rank d; [s*M + W] = its hot rows
    size_t cold = 0, hot_before_me = 0, hot_all = 0;
    for (int s = 0; s < W; ++s) {
        cold += static_cast<size_t>(matrix[static_cast<size_t>(s) * M + static_cast<size_t>(me)]);
        if (s < me) hot_before_me += static_cast<size_t>(matrix[static_cast<size_t>(s) * M + W]);
        hot_all += static_cast<size_t>(matrix[static_cast<size_t>(s) * M + W]);
    }
    // the hot rows follow the hashed ones: mine only, or every rank's in rank order
    const size_t total = cold + (keep_hot_local ? static_cast<size_t>(my_hot) : hot_all);
    const size_t my_hot_at = cold + (keep_hot_local ? 0 : hot_before_me);
    if (total > 0xFFFFFFFFull) throw std::runtime_error("a rank would own more than 2^32 rows after the shuffle");

    // receive buffers (blocks of their own, exportable) and their handles
    const size_t n_cols = 1 + payload.size();
    std::vector<DevColPtr> recv;
    std::vector<int64_t> my_handles(n_cols * 8);
    for (size_t c = 0; c < n_cols; ++c) {
        const TypeId t = c == 0 ? key->type() : payload[c - 1]->type();
        bq_col* h = nullptr;
        check(bq_col_alloc_shared(ctx, static_cast<int>(t), total, &h));
        recv.push_back(adopt(h));
        check(bq_col_ipc_export(ctx, h, &my_handles[c * 8]));
    }
    // (this exchange is also the barrier that says: every receive buffer exists and its previous contents are dead)
    auto handles = x.host_gather(my_handles);
    std::vector<std::vector<char*>> peer(static_cast<size_t>(W), std::vector<char*>(n_cols, nullptr));
    for (int r = 0; r < W; ++r)
        for (size_t c = 0; c < n_cols; ++c) {
            if (r == me) {
                peer[static_cast<size_t>(r)][c] = ptr_of(recv[c]);
            } else {
                void* p = nullptr;
                check(bq_ipc_open(ctx, &handles[(static_cast<size_t>(r) * n_cols + c) * 8], &p));
                peer[static_cast<size_t>(r)][c] = static_cast<char*>(p);
            }
        }
    trace.mark("shuffle: sizes + handles");

    // where my rows of partition q start inside the owning rank's buffers
    std::vector<void*> dest(tables * 3, nullptr);
    auto type_of = [&](size_t c) { return c == 0 ? key->type() : payload[c - 1]->type(); };
    for (int d = 0; d < W; ++d) {
        size_t row = 0;
        for (int s = 0; s < me; ++s) row += static_cast<size_t>(matrix[static_cast<size_t>(s) * M + static_cast<size_t>(d)]);
        for (size_t q = first_part(d); q < first_part(d + 1); ++q) {
            for (size_t c = 0; c < n_cols; ++c) dest[c * tables + q] = peer[static_cast<size_t>(d)][c] + row * width_of(type_of(c));
            row += static_cast<size_t>(counts[q]);
        }
    }
    if (hot)
        for (size_t c = 0; c < n_cols; ++c) {
            dest[c * tables + static_cast<size_t>(P)] = ptr_of(recv[c]) + my_hot_at * width_of(type_of(c));
            for (size_t q = static_cast<size_t>(P) + 1; q < tables; ++q) dest[c * tables + q] = ptr_of(recv[c]);      // never written (0 rows)
        }
    std::vector<const bq_col*> pay;
    for (const auto& c : payload) pay.push_back(c->h);
    check(bq_partition_scatter(ctx, plan, pay.data(), static_cast<int>(pay.size()), dest.data(), dest.data() + tables, dest.data() + 2 * tables));
    // every rank's writes have landed once every rank's kernel has finished
    x.host_gather({0});
    trace.mark("shuffle: scatter over NVLink");
    if (hot && !keep_hot_local && hot_all) {
        // replicate the hot rows: every rank's slice of the tail, gathered in place (rank order)
        std::vector<int64_t> bytes(static_cast<size_t>(W));
        for (size_t c = 0; c < n_cols; ++c) {
            const size_t w = width_of(type_of(c));
            for (int s = 0; s < W; ++s) bytes[static_cast<size_t>(s)] = matrix[static_cast<size_t>(s) * M + W] * static_cast<int64_t>(w);
            xcheck(x.fn.all_gather_v(x.fn.user, ptr_of(recv[c]) + my_hot_at * w, ptr_of(recv[c]) + cold * w, bytes.data(), bq_ctx_stream(ctx)),
                   "all_gather_v");
        }
        trace.mark("shuffle: replicate hot rows");
    }

    Shuffled out;
    out.rows = total;
    out.key = recv[0];
    for (size_t c = 1; c < n_cols; ++c) out.payload.push_back(recv[c]);
    return out;
}

Shuffled shuffle_by_key(const DevColPtr& key, const std::vector<DevColPtr>& payload, size_t rows, const std::vector<int64_t>& hot_keys,
                        bool keep_hot_local) {
    // $BOSQL_SHUFFLE=collective keeps the two-step form (partition into a send buffer, then the host's all-to-all): the
    // comparison point for the fused peer-write kernel, and the way out on a box without CUDA IPC between the ranks
    const char* mode = std::getenv("BOSQL_SHUFFLE");
    if (mode && std::string(mode) == "collective") {
        if (!hot_keys.empty()) throw std::runtime_error("BOSQL_SHUFFLE=collective has no hot-key handling");
        return shuffle_collective(key, payload, rows);
    }
    return shuffle_peer_write(key, payload, rows, hot_keys, keep_hot_local);
}

DeviceRelationPtr all_gather_relation(const DeviceRelationPtr& local, const std::vector<TypeId>& types) {
    Exchange& x = exchange();
    auto rows_by_rank = x.host_gather({static_cast<int64_t>(local->rows)});
    auto out = std::make_shared<DeviceRelation>();
    for (int64_t n : rows_by_rank) out->rows += static_cast<size_t>(n);
    for (size_t c = 0; c < types.size(); ++c) {
        DevColPtr src = local->cols[c];
        if (!src) src = alloc_col(types[c], 0);
        out->cols.push_back(x.all_gather_column(src, local->rows, rows_by_rank));
    }
    return out;
}

}  // namespace bosql::gpu
