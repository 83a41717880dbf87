// example.cpp — the reference's README / test scenarios through the C++ host mirror (host/csvplus.hpp).
// Mirrors TestSimpleDataSource (csvplus_test.go:118-151), TestSimpleUniqueJoin (:368-452), TestErrors (:808-909)
// and TestWriteFile (:172-196) on fixtures generated like csvplus_test.go:1207-1333.  Exit code 0 = all checks hold.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <cstring>
#include <sstream>

#include "csvplus.hpp"

using namespace csvplus;

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
    const char* names[] = {"Amelia", "Olivia", "Emily", "Ava", "Isla", "Oliver", "Jack", "Harry", "Jacob", "Charlie"};
    const char* surnames[] = {"Smith", "Jones", "Taylor", "Williams", "Brown", "Davies", "Evans", "Wilson", "Thomas", "Roberts", "Johnson", "Lewis"};
    std::string people = "id,name,surname,born\n";
    for (int i = 0; i < 10; i++) for (int j = 0; j < 12; j++)
        people += std::to_string(i * 12 + j) + "," + names[i] + "," + surnames[j] + "," + std::to_string(1916 + (i * 37 + j * 11) % 90) + "\n";
    std::string orders = "order_id,cust_id,prod_id,qty,ts\n";
    std::vector<int> cust(10000), qty(10000);
    unsigned s = 12345;
    for (int i = 0; i < 10000; i++) {
        s = s * 1103515245u + 12345u; cust[i] = (s >> 8) % 120;
        s = s * 1103515245u + 12345u; qty[i] = 1 + (s >> 8) % 100;
        orders += std::to_string(i) + "," + std::to_string(cust[i]) + "," + std::to_string(i % 8) + "," + std::to_string(qty[i]) + ",2016-09-14T08:48:22+01:00\n";
    }
    try {
        // TestSimpleDataSource
        auto src = Take(FromString(people).SelectColumns({"born", "id", "name", "surname"}))
                       .Filter(Any({Like({{"name", "Jack"}}), Like({{"name", "Amelia"}})}));
        int n = 0;
        src([&](const Row& row) { if ((row.at("name") == "Jack" || row.at("name") == "Amelia") && row.size() == 4) n++; });
        CHECK(n == 24);
        // TestSimpleUniqueJoin
        auto idx = Take(FromString(people).SelectColumns({"id", "name", "surname"})).UniqueIndexOn({"id"});
        std::vector<long> qsum(120, 0), want(120, 0);
        for (int i = 0; i < 10000; i++) want[cust[i]] += qty[i];
        long joined = 0;
        Take(FromString(orders).SelectColumns({"order_id", "cust_id", "qty"})).Join(idx, {"cust_id"})([&](const Row& row) {
            if (row.size() == 6 && row.at("id") == row.at("cust_id")) { qsum[std::stoi(row.at("id"))] += std::stol(row.at("qty")); joined++; }
        });
        CHECK(joined == 10000);
        CHECK(qsum == want);
        // TestWriteFile: parse -> ToCsv round trip
        std::ostringstream out;
        Take(FromString(people).SelectColumns({"id", "name", "surname", "born"})).ToCsv(out, {"id", "name", "surname", "born"});
        CHECK(out.str() == people);
        // TestErrors
        try { Take(FromString(people).SelectColumns({"id", "name", "xxx"})).ToRows(); CHECK(false); }
        catch (const DataSourceError& e) { CHECK(std::string(e.what()) == "row 1: column not found: xxx"); }
        try { Take(FromString(people).SelectColumns({"id", "name", "surname"})).UniqueIndexOn({"name"}); CHECK(false); }
        catch (const Error& e) { CHECK(std::string(e.what()).find("duplicate value while creating unique index: { \"name\" : \"Amelia\" }") == 0); }
        try { Take(FromString(people).SelectColumns({"id", "name"})).IndexOn({"name", "xxx"}); CHECK(false); }
        catch (const DataSourceError& e) { CHECK(std::string(e.what()).find("missing column \"xxx\" while creating an index") != std::string::npos); }
        // ResolveDuplicates + Find (TestErrors :845-863, TestIndexImpl)
        auto byname = Take(FromString(people).SelectColumns({"id", "name", "surname"})).IndexOn({"name"});
        byname.ResolveDuplicates([](const std::vector<Row>& g) { return g.size() == 12 ? 0L : -1L; });
        CHECK(byname.size() == 10);
        CHECK(byname.Find({"Jack"}).ToRows().size() == 1);
        // host closure boundary: Filter(func) + Map then a device join again
        auto rows = Take(FromString(orders)).Filter(Func([](const Row& r) { return std::stoi(r.at("qty")) > 90; }))
                        .Map([](const Row& r) { Row x = r; x["tag"] = "big"; return x; }).Join(idx, {"cust_id"}).Top(5).ToRows();
        CHECK(rows.size() == 5 && rows[0].at("tag") == "big" && rows[0].count("surname") == 1);
    } catch (const std::exception& e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
    // ---- multi-GPU entry points of the C ABI, on as many devices as the box has (at most 2 here): the build
    // side is parsed in row-range shards, one per context, all-gathered (cpb_allgather_tables) and indexed on every rank;
    // every rank joins its own probe shard; the concatenation of the results equals the single-GPU join.
    {
        int ndev = 0;
        for (int d = 0; d < 2; d++) { cpb_ctx* probe = nullptr; if (cpb_init(d, &probe) == CPB_OK) { ndev++; cpb_shutdown(probe); } else break; }
        CHECK(ndev >= 1);
        std::vector<int> devs(ndev);
        for (int d = 0; d < ndev; d++) devs[d] = d;
        std::vector<cpb_ctx*> cs(ndev, nullptr);
        CHECK(cpb_init_multi(devs.data(), ndev, cs.data()) == CPB_OK);
        CHECK(cpb_comm_size(cs[0]) == ndev && cpb_comm_rank(cs[ndev - 1]) == ndev - 1);
        auto cs_ = [](const char* s) { return cpb_str{s, strlen(s)}; };
        cpb_reader_opts opts{',', 0, 0, 0, 0, 1, 0};
        cpb_header_col cust_spec[3] = {{cs_("id"), -1, 0}, {cs_("name"), -1, 0}, {cs_("surname"), -1, 0}};
        // row-range shards of the people file: the header + a contiguous run of lines per rank
        std::vector<std::string> lines;
        { std::istringstream in(people); std::string ln; while (std::getline(in, ln)) lines.push_back(ln + "\n"); }
        std::vector<cpb_table*> locals(ndev, nullptr), gathered(ndev, nullptr);
        cpb_error e;
        for (int r = 0; r < ndev; r++) {
            std::string shard = lines[0];
            const size_t lo = 1 + (lines.size() - 1) * r / ndev, hi = 1 + (lines.size() - 1) * (r + 1) / ndev;
            for (size_t i = lo; i < hi; i++) shard += lines[i];
            CHECK(cpb_parse_csv(cs[r], shard.data(), shard.size(), 0, &opts, cust_spec, 3, nullptr, &locals[r], &e) == CPB_OK);
        }
        CHECK(cpb_allgather_tables(cs.data(), locals.data(), ndev, gathered.data()) == CPB_OK);
        long total = 0;
        for (int r = 0; r < ndev; r++) {
            CHECK(cpb_table_num_rows(gathered[r]) == 120);
            cpb_index* ix = nullptr;
            cpb_str key = cs_("id");
            CHECK(cpb_index_build(cs[r], gathered[r], &key, 1, 1, &ix, &e) == CPB_OK);
            // this rank's probe shard: a row range of the orders file
            std::vector<std::string> ol;
            { std::istringstream in(orders); std::string ln; while (std::getline(in, ln)) ol.push_back(ln + "\n"); }
            std::string shard = ol[0];
            const size_t lo = 1 + (ol.size() - 1) * r / ndev, hi = 1 + (ol.size() - 1) * (r + 1) / ndev;
            for (size_t i = lo; i < hi; i++) shard += ol[i];
            cpb_header_col ord_spec[2] = {{cs_("cust_id"), -1, 0}, {cs_("qty"), -1, 0}};
            cpb_table *probe = nullptr, *joined = nullptr;
            CHECK(cpb_parse_csv(cs[r], shard.data(), shard.size(), 0, &opts, ord_spec, 2, nullptr, &probe, &e) == CPB_OK);
            cpb_str jc = cs_("cust_id");
            CHECK(cpb_join(cs[r], probe, ix, &jc, 1, &joined, &e) == CPB_OK);
            CHECK(cpb_table_num_rows(joined) == (int64_t)(hi - lo) && cpb_table_num_cols(joined) == 5);
            total += (long)cpb_table_num_rows(joined);
            cpb_table_free(joined); cpb_table_free(probe); cpb_index_free(ix);
        }
        CHECK(total == 10000);
        // the metadata exchange of byte-range shards
        uint64_t in2[2] = {7, 9}, out2[2 * 2] = {0, 0, 0, 0};
        if (ndev == 1) { CHECK(cpb_allgather_u64(cs[0], in2, 2, out2) == CPB_OK && out2[0] == 7 && out2[1] == 9); }
        for (int r = 0; r < ndev; r++) { cpb_table_free(gathered[r]); cpb_table_free(locals[r]); }
        for (int r = 0; r < ndev; r++) cpb_shutdown(cs[r]);
        printf("multi-GPU entry points ok on %d device(s)\n", ndev);
    }
    printf("host example ok\n");
    return 0;
}
