// bq_cli.cpp — `bq_b200 [file.csv] --sql "<query>" [--output-format markdown|csv]`: configuration 1's entry point.
//
// Mirrors the non-interactive path of the reference's CLI (src/cli/main.cpp:59-129): load the CSV (or stdin) as table
// "table", plan the statement, run it — here on the GPU operators — and print it with the reference's Markdown / CSV
// layout (src/exec/formatter.cpp, src/exec/execution.cpp:8-61: cells through std::to_string, DOUBLE with six decimals,
// dictionary ids decoded through the base table's dictionary).  Header names follow get_output_schema
// (src/logical/planner.cpp:167-270): alias, else the column name, else "expr".  The REPL is not reproduced.
#include <thread>
#include <cstdlib>
#include <chrono>
#include <algorithm>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "bosql_operator.hpp"
#include "gpu_device.hpp"
#include "bosql_types.hpp"

using namespace bosql;

namespace {

struct Schema {
    std::vector<std::string> names;
    const Dictionary* dict = nullptr;
};

Schema output_schema(const LogicalOp* plan, const Catalog& catalog) {
    Schema s;
    const LogicalOp* cur = plan;
    while (cur && cur->type != LogicalOpType::PROJECT) cur = cur->children.empty() ? nullptr : cur->children[0].get();
    if (!cur) {
        s.names = {"result"};
        return s;
    }
    const auto* project = dynamic_cast<const LogicalProject*>(cur);
    const LogicalOp* base = cur;
    while (base && base->type != LogicalOpType::SCAN) base = base->children.empty() ? nullptr : base->children[0].get();
    OptionalRef<const Table> table;
    if (base) {
        table = catalog.get_table_data(dynamic_cast<const LogicalScan*>(base)->table_name);
        if (table.has_value()) s.dict = table->dict.get();
    }
    if (project->select_list.empty()) {
        if (table.has_value())
            for (const auto& c : table->columns) s.names.push_back(c.name);
        if (s.names.empty()) s.names = {"col1"};
        return s;
    }
    for (size_t k = 0; k < project->select_list.size(); ++k) {
        const Expr* e = project->select_list[k].get();
        if (!project->aliases[k].empty()) s.names.push_back(project->aliases[k]);
        else if (e->type == ExprType::COLUMN_REF) s.names.push_back(e->str_val);
        else s.names.push_back("expr");
    }
    return s;
}

std::string csv_cell(const std::string& cell) {
    if (cell.find_first_of(",\"\n\r") == std::string::npos) return cell;
    std::string out = "\"";
    for (char ch : cell) {
        if (ch == '"') out.push_back('"');
        out.push_back(ch);
    }
    return out + "\"";
}

void print_markdown(const std::vector<std::string>& headers, const std::vector<std::vector<std::string>>& rows) {
    if (rows.empty()) {
        std::cout << "(no results)\n";
        return;
    }
    std::vector<size_t> w(headers.size(), 0);
    for (size_t i = 0; i < headers.size(); ++i) w[i] = headers[i].size();
    for (const auto& r : rows)
        for (size_t i = 0; i < r.size(); ++i) {
            if (i >= w.size()) w.resize(i + 1, 0);
            w[i] = std::max(w[i], r[i].size());
        }
    auto line = [&](const std::vector<std::string>& cells) {
        std::cout << "|";
        for (size_t i = 0; i < w.size(); ++i)
            std::cout << " " << std::left << std::setw(static_cast<int>(w[i])) << (i < cells.size() ? cells[i] : std::string()) << " |";
        std::cout << '\n';
    };
    line(headers);
    std::cout << "|";
    for (size_t i = 0; i < w.size(); ++i) std::cout << " " << std::string(w[i], '-') << " |";
    std::cout << '\n';
    for (const auto& r : rows) line(r);
}

int run(const std::string& sql, const Catalog& catalog, const std::string& format, const ParseOptions& parse_opts) {
    try {
        SelectStmt stmt = parse_sql(sql, parse_opts);
        LogicalPlanner planner;
        auto logical = planner.build_logical_plan(stmt);
        auto root = build_physical_plan(logical.get(), catalog);
        Schema schema = output_schema(logical.get(), catalog);
        std::vector<std::vector<std::string>> rows;
        root->open();
        ExecBatch batch;
        while (root->next(batch)) {
            for (size_t i = 0; i < batch.length; ++i) {
                std::vector<std::string> row;
                for (size_t j = 0; j < batch.columns.size(); ++j) {
                    switch (batch.columns[j].type) {
                        case TypeId::INT64: row.push_back(std::to_string(get_col<int64_t>(batch, j)[i])); break;
                        case TypeId::DOUBLE: row.push_back(std::to_string(get_col<double>(batch, j)[i])); break;
                        case TypeId::DATE32: row.push_back(std::to_string(get_col<int32_t>(batch, j)[i])); break;
                        case TypeId::STRING: {
                            uint32_t id = get_col<uint32_t>(batch, j)[i];
                            row.push_back(schema.dict ? schema.dict->get(id) : std::to_string(id));
                            break;
                        }
                    }
                }
                rows.push_back(std::move(row));
            }
        }
        root->close();
        if (format == "csv") {
            for (size_t i = 0; i < schema.names.size(); ++i) std::cout << (i ? "," : "") << csv_cell(schema.names[i]);
            if (!schema.names.empty()) std::cout << '\n';
            for (const auto& r : rows) {
                for (size_t i = 0; i < r.size(); ++i) std::cout << (i ? "," : "") << csv_cell(r[i]);
                std::cout << '\n';
            }
        } else {
            print_markdown(schema.names, rows);
        }
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << "\n";      // the reference prints the message and still exits 0 (main.cpp:54-56)
    }
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    std::vector<std::string> args(argv + 1, argv + argc);
    std::string csv_file, sql, format = "markdown";
    std::vector<std::pair<std::string, std::string>> named_tables;      // --table name=file.csv (repeatable): LOAD TABLE without a REPL
    ParseOptions parse_opts;                                             // --extended-sql: BETWEEN, 1.5, -5, keywords in any case
    bool have_sql = false;
    for (size_t i = 0; i < args.size(); ++i) {
        if (args[i] == "--extended-sql") {
            parse_opts.between = parse_opts.decimal_literals = parse_opts.negative_literals = parse_opts.keywords_any_case = true;
            continue;
        }
        if (args[i] == "--table") {
            if (i + 1 >= args.size() || args[i + 1].find('=') == std::string::npos) {
                std::cerr << "--table requires name=file.csv\n";
                return 1;
            }
            const std::string& spec = args[++i];
            named_tables.emplace_back(spec.substr(0, spec.find('=')), spec.substr(spec.find('=') + 1));
            continue;
        }
        if (args[i] == "--sql" || args[i] == "--output-format") {
            if (i + 1 >= args.size()) {
                std::cerr << args[i] << " requires an argument\n";
                return 1;
            }
            if (args[i] == "--sql") {
                sql = args[++i];
                have_sql = true;
            } else {
                format = args[++i];
                std::transform(format.begin(), format.end(), format.begin(), [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
            }
        } else if (args[i].rfind("--", 0) == 0) {
            std::cerr << "Unknown option: " << args[i] << "\n";
            return 1;
        } else if (csv_file.empty()) {
            csv_file = args[i];
        } else {
            std::cerr << "Too many positional arguments\n";
            return 1;
        }
    }
    if (format != "markdown" && format != "csv") {
        std::cerr << "Unsupported output format '" << format << "'. Use 'markdown' or 'csv'.\n";
        return 1;
    }
    if (!have_sql) {
        std::cerr << "bq_b200 runs one statement: bq_b200 [file.csv] [--table name=file.csv ...] --sql \"SELECT ...\" "
                     "[--extended-sql] [--output-format markdown|csv]\n";
        return 1;
    }
    // The CUDA context (driver initialisation, primary context, stream-ordered pool) takes longer than parsing a million-row
    // file: it is created on a second thread while the CSV loads, and the statement waits for whichever finishes last.
    // $BOSQL_TRACE=1 prints where the wall time went.
    const bool trace = std::getenv("BOSQL_TRACE") && *std::getenv("BOSQL_TRACE") == '1';
    const auto t0 = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    if (!std::getenv("CUDA_VISIBLE_DEVICES")) {
        // one process uses one GPU: initialising the driver for the other seven of a full box costs about as much again
        const char* dev = std::getenv("BOSQL_DEVICE") ? std::getenv("BOSQL_DEVICE") : (std::getenv("LOCAL_RANK") ? std::getenv("LOCAL_RANK") : "0");
        setenv("CUDA_VISIBLE_DEVICES", dev, 1);
        setenv("BOSQL_DEVICE", "0", 1);
        unsetenv("LOCAL_RANK");
    }
    std::string context_error;
    double context_s = 0.0;
    std::thread warm([&] {
        try {
            gpu::context();
        } catch (const std::exception& e) {
            context_error = e.what();          // reported by the first operator that needs the device, as before
        }
        context_s = since();
    });
    Catalog catalog;
    double load_s = 0.0;
    try {
        // All tables of one invocation share ONE dictionary, so string ids compare across tables.  With a
        // single table this is exactly the reference's per-table dictionary.
        std::shared_ptr<Dictionary> dict;
        if (!csv_file.empty() || named_tables.empty()) {
            auto [table, meta] = csv_file.empty() ? load_csv(std::cin) : load_csv(csv_file);
            dict = table.dict;
            table.name = "table";                 // src/cli/main.cpp:105
            meta.name = "table";
            catalog.register_table(std::move(table), std::move(meta));
        }
        for (const auto& [name, path] : named_tables) {
            auto [table, meta] = load_csv(path, dict);
            dict = table.dict;
            table.name = name;                    // LOAD TABLE <name> FROM '<file>' (src/cli/main.cpp:152-168)
            meta.name = name;
            catalog.register_table(std::move(table), std::move(meta));
        }
        load_s = since();
    } catch (const std::exception& e) {
        warm.join();
        std::cerr << "Error loading CSV" << (csv_file.empty() ? " from stdin" : "") << ": " << e.what() << "\n";
        return 1;
    }
    warm.join();
    const double ready_s = since();
    const int rc = run(sql, catalog, format, parse_opts);
    if (trace)
        std::cerr << "[bosql trace] csv load " << load_s << " s | cuda context (concurrent) " << context_s << " s | both ready " << ready_s
                  << " s | statement " << since() - ready_s << " s | total " << since() << " s\n";
    return rc;
}
