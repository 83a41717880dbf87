// Golden dumps from the UNMODIFIED reference: go run ./dump <inputs dir> <out dir>
// Writes ToCsv output of the pipelines tests/ compare bit-exactly.  Never compiled here (no Go toolchain).
package main

import (
	"log"
	"os"
	"path/filepath"

	"github.com/maxim2266/csvplus"
)

func main() {
	in, out := os.Args[1], os.Args[2]
	must := func(err error) {
		if err != nil {
			log.Fatal(err)
		}
	}
	must(os.MkdirAll(out, 0o755))
	must(csvplus.Take(csvplus.FromFile(filepath.Join(in, "people.csv")).SelectColumns("name", "surname", "id")).
		Filter(csvplus.Like(csvplus.Row{"name": "Amelia"})).ToCsvFile(filepath.Join(out, "people_amelia.csv"), "name", "surname", "id"))
	idx, err := csvplus.Take(csvplus.FromFile(filepath.Join(in, "customers.csv")).SelectColumns("id", "name", "surname")).UniqueIndexOn("id")
	must(err)
	must(csvplus.Take(csvplus.FromFile(filepath.Join(in, "orders.csv")).SelectColumns("cust_id", "prod_id", "qty", "ts")).
		Join(idx, "cust_id").ToCsvFile(filepath.Join(out, "orders_join_customers.csv"), "cust_id", "prod_id", "qty", "ts", "id", "name", "surname"))
	must(csvplus.Take(idx).ToCsvFile(filepath.Join(out, "customers_sorted_by_id.csv"), "id", "name", "surname"))
}
