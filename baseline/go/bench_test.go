// Benchmarks of the UNMODIFIED reference on the bench.py workloads (BASELINE.json configs 1-3, 5).
// Never compiled in this repository's image (no Go toolchain); kept for maintainers.
package csvplusbaseline

import (
	"os"
	"path/filepath"
	"testing"

	"github.com/maxim2266/csvplus"
)

func input(name string) string { return filepath.Join(os.Getenv("CSVPLUS_INPUTS"), name) }

// config 1/2: Take(FromFile(...).SelectColumns(name,surname,id)).Filter(Like{name: Amelia}).ToRows
func BenchmarkParseSelectFilter(b *testing.B) {
	for i := 0; i < b.N; i++ {
		rows, err := csvplus.Take(csvplus.FromFile(input("people.csv")).SelectColumns("name", "surname", "id")).
			Filter(csvplus.Like(csvplus.Row{"name": "Amelia"})).ToRows()
		if err != nil {
			b.Fatal(err)
		}
		b.ReportMetric(float64(len(rows)), "rows_out")
	}
}

// config 3: customers -> UniqueIndexOn(id); orders -> Join(idx, "cust_id")
func BenchmarkUniqueIndexJoin(b *testing.B) {
	for i := 0; i < b.N; i++ {
		idx, err := csvplus.Take(csvplus.FromFile(input("customers.csv")).SelectColumns("id", "name", "surname")).UniqueIndexOn("id")
		if err != nil {
			b.Fatal(err)
		}
		n := 0
		err = csvplus.Take(csvplus.FromFile(input("orders.csv")).SelectColumns("cust_id", "prod_id", "qty", "ts")).
			Join(idx, "cust_id")(func(csvplus.Row) error { n++; return nil })
		if err != nil {
			b.Fatal(err)
		}
		b.ReportMetric(float64(n), "rows_out")
	}
}

// config 5: IndexOn 2-column composite key + ResolveDuplicates (keep the smallest id)
func BenchmarkCompositeIndexDedup(b *testing.B) {
	for i := 0; i < b.N; i++ {
		idx, err := csvplus.Take(csvplus.FromFile(input("people.csv")).SelectColumns("id", "name", "surname")).IndexOn("surname", "name")
		if err != nil {
			b.Fatal(err)
		}
		err = idx.ResolveDuplicates(func(rows []csvplus.Row) (csvplus.Row, error) {
			best := rows[0]
			for _, r := range rows[1:] {
				if r["id"] < best["id"] {
					best = r
				}
			}
			return best, nil
		})
		if err != nil {
			b.Fatal(err)
		}
	}
}
