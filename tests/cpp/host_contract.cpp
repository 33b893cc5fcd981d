// host_contract.cpp — the reference's C++ unit tests that need no execution, restated against bo-sql_b200/host/ (the
// drop-in's own headers): tests/test_types.cpp:4-43 (Datum, ColumnType), tests/test_columnar.cpp:4-68 (ColumnVector,
// RecordBatch), tests/test_catalog.cpp:7-53 (load_csv -> Catalog round trip), and the construction-time facts of
// tests/test_execution.cpp (:187-200 output names of a join + GROUP BY, :210-220 the root of SELECT COUNT(*) is-a
// HashAggregate named "COUNT(*)", dictionary propagation :168-185).  Plain asserts instead of Catch2 (not in this image);
// built and run by tests/test_host_contract.py on the CPU: nothing here may create a CUDA context.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>

#include "bosql_operator.hpp"

using namespace bosql;

static int failures = 0;
#define REQUIRE(cond)                                                            \
    do {                                                                         \
        if (!(cond)) {                                                           \
            std::fprintf(stderr, "%s:%d: REQUIRE(%s) failed\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                          \
        }                                                                        \
    } while (0)
#define REQUIRE_THROWS(expr)                                                     \
    do {                                                                         \
        bool threw = false;                                                      \
        try { (void)(expr); } catch (const std::runtime_error&) { threw = true; } \
        if (!threw) { std::fprintf(stderr, "%s:%d: %s did not throw\n", __FILE__, __LINE__, #expr); ++failures; } \
    } while (0)

static void datum_and_column_type() {
    Datum d1 = Datum::from_i64(42);
    REQUIRE(d1.type == TypeId::INT64);
    REQUIRE(d1.as_i64() == 42);
    Datum d2 = Datum::from_f64(3.14);
    REQUIRE(d2.type == TypeId::DOUBLE);
    REQUIRE(d2.as_f64() == 3.14);
    Datum d3 = Datum::from_str(123);
    REQUIRE(d3.type == TypeId::STRING);
    REQUIRE(d3.as_str() == 123);
    Datum d4 = Datum::from_date32(20231225);
    REQUIRE(d4.type == TypeId::DATE32);
    REQUIRE(d4.as_date32() == 20231225);
    REQUIRE_THROWS(d1.as_f64());

    ColumnType ct1(TypeId::INT64, "id");
    REQUIRE(ct1.type_id == TypeId::INT64);
    REQUIRE(ct1.name == "id");
    ColumnType ct2(TypeId::DOUBLE);
    REQUIRE(ct2.type_id == TypeId::DOUBLE);
    REQUIRE(ct2.name.empty());
    ColumnType ct3(TypeId::INT64, "other_id");
    REQUIRE(ct1 == ct3);      // same type, different name
    REQUIRE(ct1 != ct2);
}

static void column_vector_and_record_batch() {
    ColumnVector<int64_t> col(10);
    for (int64_t i = 0; i < 5; ++i) col.append(i * 10);
    REQUIRE(col.size() == 5);
    REQUIRE(col.type() == TypeId::INT64);
    REQUIRE(col.data[0] == 0);
    REQUIRE(col.data[4] == 40);

    std::vector<ColumnType> schema = {ColumnType(TypeId::INT64, "id"), ColumnType(TypeId::DOUBLE, "value")};
    RecordBatch batch(schema);
    REQUIRE(batch.num_columns() == 0);
    REQUIRE(batch.num_rows() == 0);
    auto c1 = std::make_unique<ColumnVector<int64_t>>();
    auto c2 = std::make_unique<ColumnVector<double>>();
    for (int i = 1; i <= 3; ++i) {
        c1->append(i);
        c2->append(i * 1.1);
    }
    batch.add_column(std::move(c1));
    batch.add_column(std::move(c2));
    REQUIRE(batch.num_columns() == 2);
    REQUIRE(batch.num_rows() == 3);
    auto* g1 = dynamic_cast<ColumnVector<int64_t>*>(batch.get_column(0));
    auto* g2 = dynamic_cast<ColumnVector<double>*>(batch.get_column(1));
    REQUIRE(g1 != nullptr);
    REQUIRE(g2 != nullptr);
    REQUIRE(g1 && g1->data[0] == 1);
    REQUIRE(g2 && g2->data[1] == 2.2);
    REQUIRE(batch.get_column_type(0).name == "id");
    REQUIRE(batch.get_column_type(1).type_id == TypeId::DOUBLE);
}

static void catalog_round_trip(const std::string& dir) {
    const std::string path = dir + "/host_contract_catalog.csv";
    {
        std::ofstream csv(path);
        csv << "id,value\n10,1.1\n20,2.2\n";
    }
    auto loaded = load_csv(path);
    loaded.first.name = "mytable";
    loaded.second.name = "mytable";
    Catalog catalog;
    catalog.register_table(std::move(loaded.first), std::move(loaded.second));
    auto names = catalog.list_tables();
    REQUIRE(names.size() == 1);
    REQUIRE(names[0] == "mytable");
    auto meta = catalog.get_table_meta("mytable");
    REQUIRE(meta.has_value());
    REQUIRE(meta->name == "mytable");
    REQUIRE(meta->row_count == 2);
    REQUIRE(meta->columns.size() == 2);
    REQUIRE(meta->columns[0].name == "id");
    REQUIRE(meta->columns[0].type == TypeId::INT64);
    REQUIRE(meta->columns[1].name == "value");
    REQUIRE(meta->columns[1].type == TypeId::DOUBLE);
    REQUIRE(!catalog.get_table_meta("nope").has_value());
    std::remove(path.c_str());
}

// fixtures of tests/test_execution.cpp:13-63
static Catalog full_catalog(std::shared_ptr<Dictionary>& detail_dict) {
    Catalog catalog;
    {
        Table t;
        t.name = "orders";
        t.dict = std::make_shared<Dictionary>();
        auto id = std::make_unique<ColumnVector<int64_t>>();
        auto qty = std::make_unique<ColumnVector<int64_t>>();
        for (int i = 1; i <= 3; ++i) {
            id->append(i);
            qty->append(i * 10);
        }
        t.columns.push_back({"orders.id", std::move(id)});
        t.columns.push_back({"orders.qty", std::move(qty)});
        std::vector<ColumnMeta> cols;
        cols.emplace_back("orders.id", TypeId::INT64);
        cols.emplace_back("orders.qty", TypeId::INT64);
        catalog.register_table(std::move(t), TableMeta("orders", std::move(cols), 3));
    }
    {
        Table t;
        t.name = "detail";
        detail_dict = std::make_shared<Dictionary>();
        t.dict = detail_dict;
        auto id = std::make_unique<ColumnVector<int64_t>>();
        auto region = std::make_unique<ColumnVector<uint32_t>>();
        const int ids[3] = {1, 2, 4};
        const char* regions[3] = {"north", "south", "west"};
        for (int i = 0; i < 3; ++i) {
            id->append(ids[i]);
            region->append(t.dict->get_or_add(regions[i]));
        }
        t.columns.push_back({"detail.id", std::move(id)});
        t.columns.push_back({"detail.region", std::move(region)});
        std::vector<ColumnMeta> cols;
        cols.emplace_back("detail.id", TypeId::INT64);
        cols.emplace_back("detail.region", TypeId::STRING);
        catalog.register_table(std::move(t), TableMeta("detail", std::move(cols), 3));
    }
    return catalog;
}

static std::unique_ptr<Operator> plan(const std::string& sql, const Catalog& catalog) {
    SelectStmt stmt = parse_sql(sql);
    LogicalPlanner planner;
    auto logical = planner.build_logical_plan(stmt);
    return build_physical_plan(logical.get(), catalog);
}

static void plan_construction() {
    std::shared_ptr<Dictionary> detail_dict;
    Catalog catalog = full_catalog(detail_dict);
    REQUIRE(detail_dict->get_or_add("north") == 0);          // ids start at 0 in first-seen order (dictionary.cpp:5-12)
    REQUIRE(detail_dict->get_or_add("west") == 2);

    auto agg = plan("SELECT detail.region, SUM(orders.qty) AS total FROM orders INNER JOIN detail ON orders.id = detail.id GROUP BY detail.region", catalog);
    REQUIRE(agg->output_names().size() == 2);
    REQUIRE(agg->output_names()[0] == "detail.region");
    REQUIRE(agg->output_names()[1] == "total");
    REQUIRE(agg->output_types()[0] == TypeId::STRING);
    REQUIRE(agg->output_types()[1] == TypeId::INT64);          // SUM over an INT64 argument prints as an integer (:1040-1045)
    REQUIRE(agg->dictionary() != nullptr);
    REQUIRE(dynamic_cast<HashAggregate*>(agg.get()) != nullptr);   // Project is elided above an aggregate (physical_planner.cpp:50-52)

    auto count = plan("SELECT COUNT(*) FROM orders", catalog);
    REQUIRE(dynamic_cast<HashAggregate*>(count.get()) != nullptr);
    REQUIRE(count->output_names().size() == 1);
    REQUIRE(count->output_names()[0] == "COUNT(*)");

    auto join = plan("SELECT orders.id, detail.region FROM orders INNER JOIN detail ON orders.id = detail.id", catalog);
    REQUIRE(join->dictionary() != nullptr);                    // the join hands the right side's dictionary up (:694-704)
    REQUIRE(join->output_names().size() == 2);

    auto top = plan("SELECT orders.id, orders.qty FROM orders ORDER BY orders.qty DESC LIMIT 1", catalog);
    REQUIRE(dynamic_cast<Limit*>(top.get()) != nullptr);
    auto avg = plan("SELECT AVG(orders.qty) FROM orders", catalog);
    REQUIRE(avg->output_types()[0] == TypeId::DOUBLE);
    REQUIRE(avg->output_names()[0] == "AVG(orders.qty)");

    REQUIRE_THROWS(plan("SELECT x FROM nowhere", catalog));
    REQUIRE_THROWS(plan("SELECT nope FROM orders", catalog));
    REQUIRE_THROWS(plan("SELECT orders.id FROM orders INNER JOIN detail ON orders.nokey = detail.id", catalog));
}

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    datum_and_column_type();
    column_vector_and_record_batch();
    catalog_round_trip(dir);
    plan_construction();
    if (failures) {
        std::fprintf(stderr, "%d check(s) failed\n", failures);
        return 1;
    }
    std::puts("host contract ok");
    return 0;
}
