/* bosql_b200_exec.h — C ABI of the operator layer (libbosql_b200_exec.so).
 *
 * The reference is a C++ program; its "plugin interface" for the hot path is the Operator class family of
 * include/exec/operator.hpp:17-218, constructed only by build_physical_plan (src/exec/physical_planner.cpp:9)
 * and consumed only through open()/next()/close() (src/exec/execution.cpp:14-59).  The C++ mirror of that
 * interface lives in bo-sql_b200/host/bosql_operator.hpp.  This header exposes the same life cycle to non-C++
 * callers (pytest, bench.py): build tables from typed arrays, plan a SQL string exactly as
 * execute_select_sql does (src/cli/main.cpp:40-57: parse_sql -> build_logical_plan -> build_physical_plan),
 * then open / next / close, or run to completion into typed result columns.
 *
 * Status convention: 0 = ok, nonzero = error with the message in bqx_last_error() (the reference's only error
 * channel is std::runtime_error).  Handles are opaque.  Nothing here falls back to the CPU.
 */
#ifndef BOSQL_B200_EXEC_H
#define BOSQL_B200_EXEC_H

#include <stddef.h>
#include <stdint.h>

#include "bosql_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bqx_dict bqx_dict;         /* shared_ptr<Dictionary>   (include/storage/dictionary.h:11) */
typedef struct bqx_table bqx_table;       /* a Table under construction (include/storage/table.h:20)    */
typedef struct bqx_catalog bqx_catalog;   /* Catalog                  (include/catalog/catalog.h:46)    */
typedef struct bqx_plan bqx_plan;         /* root Operator of a physical plan                           */
typedef struct bqx_result bqx_result;     /* drained output of a plan: typed host columns               */

const char* bqx_last_error(void);
/* Select the GPU of this process (one process per GPU). Optional: default $BOSQL_DEVICE, $LOCAL_RANK, 0. */
int bqx_init(int device);
/* The kernel-layer context of this process (for callers that mix both layers, e.g. bench.py's generator). */
bq_ctx* bqx_context(void);

bqx_dict* bqx_dict_create(void);
void bqx_dict_destroy(bqx_dict* d);
uint32_t bqx_dict_get_or_add(bqx_dict* d, const char* s);      /* src/storage/dictionary.cpp:5 */
size_t bqx_dict_size(const bqx_dict* d);
const char* bqx_dict_get(const bqx_dict* d, uint32_t id);

bqx_catalog* bqx_catalog_create(void);
void bqx_catalog_destroy(bqx_catalog* c);

/* dict may be NULL (fresh dictionary) or shared between tables (tests/test_execution.cpp:116-123). */
bqx_table* bqx_table_create(const char* name, bqx_dict* dict);
/* host column: copied into a ColumnVector<T>; uploaded to HBM on first use by an operator */
int bqx_table_add_column(bqx_table* t, const char* name, int type, const void* data, size_t n);
/* host column over caller-owned memory (e.g. pinned, bq_host_alloc): not copied; must outlive the catalog */
int bqx_table_add_borrowed_column(bqx_table* t, const char* name, int type, const void* data, size_t n);
/* device-resident column (synthetic tables generated in HBM); take_ownership: the table frees it */
int bqx_table_add_device_column(bqx_table* t, const char* name, bq_col* col, int take_ownership);
/* catalog statistics of a column (ColumnStats, include/catalog/catalog.h:16-21); integers for INT64/DATE32,
 * doubles for DOUBLE; ndv = 0 when unknown */
int bqx_table_set_stats(bqx_table* t, const char* column, int64_t min_i, int64_t max_i, double min_f, double max_f, size_t ndv);
/* Catalog::register_table (src/catalog/catalog.cpp:5); consumes the table handle */
int bqx_catalog_register(bqx_catalog* c, bqx_table* t);

/* load_csv (src/storage/csv_loader.cpp:168) + register under `name` (the reference CLI registers "table", src/cli/main.cpp:105) */
int bqx_catalog_load_csv(bqx_catalog* c, const char* path, const char* name);
/* what the loader inferred: rows / columns, then per column name, type, host data, statistics (ColumnStats) */
int bqx_catalog_table_info(bqx_catalog* c, const char* name, size_t* rows, size_t* ncols);
int bqx_catalog_column_info(bqx_catalog* c, const char* name, size_t i, const char** col_name, int* type, const void** data,
                            int64_t* min_i, int64_t* max_i, double* min_f, double* max_f, size_t* ndv);
size_t bqx_catalog_dict_size(bqx_catalog* c, const char* name);
const char* bqx_catalog_dict_get(bqx_catalog* c, const char* name, uint32_t id);

/* Drop the HBM mirrors of a registered table's host columns: the next query uploads them again (bench.py's
 * end-to-end leg pays the host->device copy inside every timed step this way). */
int bqx_catalog_evict_device(bqx_catalog* c, const char* table);

/* extensions of the SQL front end, off by default: bit 0 BETWEEN, bit 1 decimal literals,
 * bit 2 negative literals (-5, -1.5), bit 3 keywords in any case */
int bqx_plan_create(bqx_catalog* c, const char* sql, unsigned parse_flags, bqx_plan** out);
void bqx_plan_destroy(bqx_plan* p);
size_t bqx_plan_columns(const bqx_plan* p);
const char* bqx_plan_column_name(const bqx_plan* p, size_t i);   /* Operator::output_names() */
int bqx_plan_column_type(const bqx_plan* p, size_t i);           /* Operator::output_types() */
int bqx_plan_has_dict(const bqx_plan* p);                        /* Operator::dictionary() != nullptr */
const char* bqx_plan_dict_get(const bqx_plan* p, uint32_t id);
const char* bqx_plan_root_kind(const bqx_plan* p);               /* class name of the root operator */

/* Operator::open / next / close.  next fills up to n_cols column pointers valid until the following call on this
 * plan and sets *rows (0 and return value 1 = end of stream; a batch never has 0 rows). */
int bqx_plan_open(bqx_plan* p);
int bqx_plan_next(bqx_plan* p, const void** col_data, size_t n_cols, size_t* rows, int* end_of_stream);
int bqx_plan_close(bqx_plan* p);

/* open .. next* .. close into typed columns; seconds = host wall time of that interval */
int bqx_plan_run(bqx_plan* p, bqx_result** out);
/* open .. (the root's whole output) .. close, left in HBM: *out is a kernel-layer relation (bq_rel_*; the caller frees it with
 * bq_rel_free).  For results that feed further device work or are too large to page out (a sharded 10^8-group table). */
int bqx_plan_run_device(bqx_plan* p, bq_rel** out);
size_t bqx_result_rows(const bqx_result* r);
size_t bqx_result_cols(const bqx_result* r);
double bqx_result_seconds(const bqx_result* r);
const void* bqx_result_data(const bqx_result* r, size_t i);
void bqx_result_free(bqx_result* r);

/* ---- multi-GPU exchange ------------------------------------------------------------------------------
 * One process per GPU, each holding a ROW SHARD of every table under the same name and schema.  The host supplies the
 * collectives (torch.distributed over NCCL in bosql_b200.distributed; any NCCL/MPI host can fill the same table); the
 * operators call them at the exchange points of a plan:
 *   scan -> selection -> aggregate      local fused kernel, partial states all-gathered, merged in rank order
 *   bitmap (semi) join                  per-rank bitmaps over the global key domain summed (= OR: keys are unique)
 *   other joins                         build side all-gathered (broadcast join), probe side never moves
 *   high-cardinality GROUP BY           rows hash-partitioned by key and exchanged all-to-all, then aggregated locally
 * Every rank must run the same statements in the same order.  Catalog statistics (min / max / ndv) passed with
 * bqx_table_set_stats describe the WHOLE table, not the shard.  The reference is one process (no counterpart).
 * All device pointers are ordered on `stream` (the context's cudaStream_t); byte counts are exact. */
typedef struct bqx_exchange {
    void* user;
    int32_t world, rank;
    int32_t keep_sharded;    /* nonzero: a shuffled GROUP BY leaves each rank with the groups it owns (no final gather) */
    int32_t pad;
    /* recv[r*bytes .. (r+1)*bytes) = rank r's send[0 .. bytes) */
    int (*all_gather)(void* user, const void* send, void* recv, size_t bytes, void* stream);
    /* rank r contributes bytes_by_rank[r] bytes; recv holds them back to back in rank order */
    int (*all_gather_v)(void* user, const void* send, void* recv, const int64_t* bytes_by_rank, void* stream);
    /* send holds send_bytes[r] bytes for each rank r back to back; recv receives recv_bytes[r] from each rank r */
    int (*all_to_all_v)(void* user, const void* send, const int64_t* send_bytes, void* recv, const int64_t* recv_bytes, void* stream);
    /* in-place element-wise sum of 32-bit words over all ranks */
    int (*all_reduce_sum_u32)(void* user, void* buf, size_t words, void* stream);
    /* HOST exchange of n int64 per rank: all[r*n + i] = rank r's mine[i] (row counts, min/max, status words) */
    int (*host_all_gather_i64)(void* user, const int64_t* mine, int32_t n, int64_t* all);
} bqx_exchange;
/* Installs (copies) the table for this process; NULL or world <= 1 returns to single-GPU execution. */
int bqx_set_exchange(const bqx_exchange* x);

/* The native exchange: NCCL inside the library (bq_comm_*, include/bosql_b200.h), so that nothing but C++ runs between
 * a plan's open() and its collectives.  Rank 0 calls bqx_comm_unique_id and passes the 128 bytes to the other ranks (MPI,
 * torch.distributed, a file: any host channel); every rank then calls bqx_comm_init once, after bqx_init.  The callback
 * table above remains for hosts that bring their own collectives (the gloo-backed CPU tests). */
int bqx_comm_unique_id(void* id128);
int bqx_comm_init(int world, int rank, const void* id128, int keep_sharded);
/* The same over a rendezvous file for hosts without any channel of their own (the C++ command line): rank 0 writes the
 * id to `path` (created atomically), the others wait for it.  world / rank default to $WORLD_SIZE / $RANK when < 0. */
int bqx_comm_init_file(const char* path, int world, int rank, int keep_sharded);
/* collectives issued (calls5: all_gather, all_gather_v, all_to_all_v, all_reduce_sum_u32, host_all_gather_i64) and payload
 * bytes sent by this rank since bqx_comm_init (bench.py's exchange record) */
int bqx_comm_stats(uint64_t* calls5, uint64_t* bytes_sent);
/* switch the installed exchange between "a shuffled GROUP BY leaves the groups with their owners" and "gathers them" */
int bqx_exchange_keep_sharded(int on);

/* LogicalOp::to_string of the planned statement (plan-shape tests, tests/test_logical.cpp of the reference) */
int bqx_explain(const char* sql, unsigned parse_flags, char* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* BOSQL_B200_EXEC_H */
