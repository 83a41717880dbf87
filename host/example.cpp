// example.cpp — the reference's README / test scenarios through the C++ host mirror (host/csvplus.hpp).
// Mirrors TestSimpleDataSource (csvplus_test.go:118-151), TestSimpleUniqueJoin (:368-452), TestErrors (:808-909)
// and TestWriteFile (:172-196) on fixtures generated like csvplus_test.go:1207-1333.  Exit code 0 = all checks hold.
#include <cstdio>
#include <cstdlib>
#include <iostream>
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
    printf("host example ok\n");
    return 0;
}
