package csvplus

// row.go: the Row type and its helpers (csvplus.go:59-205).  Host-side conveniences of the API surface, not part of
// the data-parallel path: they run on the rows a sink hands to user code.

import (
	"fmt"
	"sort"
	"strconv"
	"strings"
)

// Row is one record: column name -> value (csvplus.go:59).
type Row map[string]string

// RowFunc is called once per row by a DataSource (csvplus.go:208).
type RowFunc func(Row) error

// HasColumn tells whether the column is present.
func (row Row) HasColumn(col string) bool {
	_, ok := row[col]
	return ok
}

// SafeGetValue returns the column's value, or subst when the column is absent.
func (row Row) SafeGetValue(col, subst string) string {
	if v, ok := row[col]; ok {
		return v
	}
	return subst
}

// Header lists the column names in sort.Strings order.
func (row Row) Header() []string {
	names := make([]string, 0, len(row))
	for name := range row {
		names = append(names, name)
	}
	sort.Strings(names)
	return names
}

// String renders `{ "a" : "1", "b" : "2" }` with the columns sorted, `{}` for an empty row — the text the
// duplicate-key error of UniqueIndexOn embeds (csvplus.go:90-104, :751).
func (row Row) String() string {
	if len(row) == 0 {
		return "{}"
	}
	var sb strings.Builder
	sb.WriteString("{ ")
	for i, name := range row.Header() {
		if i > 0 {
			sb.WriteString(", ")
		}
		sb.WriteByte('"')
		sb.WriteString(name)
		sb.WriteString(`" : "`)
		sb.WriteString(row[name])
		sb.WriteByte('"')
	}
	sb.WriteString(" }")
	return sb.String()
}

// SelectExisting keeps those of cols that are present.
func (row Row) SelectExisting(cols ...string) Row {
	out := make(Row, len(cols))
	for _, c := range cols {
		if v, ok := row[c]; ok {
			out[c] = v
		}
	}
	return out
}

// Select keeps exactly cols; a missing one is an error (`missing column "x"`, csvplus.go:129).
func (row Row) Select(cols ...string) (Row, error) {
	out := make(Row, len(cols))
	for _, c := range cols {
		v, ok := row[c]
		if !ok {
			return nil, fmt.Errorf(`missing column %q`, c)
		}
		out[c] = v
	}
	return out, nil
}

// SelectValues returns the values of cols in order; a missing one is an error (csvplus.go:145).
func (row Row) SelectValues(cols ...string) ([]string, error) {
	out := make([]string, len(cols))
	for i, c := range cols {
		v, ok := row[c]
		if !ok {
			return nil, fmt.Errorf(`missing column %q`, c)
		}
		out[i] = v
	}
	return out, nil
}

// Clone copies the row.
func (row Row) Clone() Row {
	out := make(Row, len(row))
	for k, v := range row {
		out[k] = v
	}
	return out
}

// ValueAsInt converts the column's value with strconv.Atoi (error texts of csvplus.go:164-183).
func (row Row) ValueAsInt(column string) (int, error) {
	val, ok := row[column]
	if !ok {
		return 0, fmt.Errorf(`missing column %q`, column)
	}
	res, err := strconv.Atoi(val)
	if err != nil {
		if ne, isNum := err.(*strconv.NumError); isNum {
			return res, fmt.Errorf(`column %q: cannot convert %q to integer: %s`, column, val, ne.Err)
		}
		return res, fmt.Errorf(`column %q: %s`, column, err)
	}
	return res, nil
}

// ValueAsFloat64 converts the column's value with strconv.ParseFloat (error texts of csvplus.go:187-205).
func (row Row) ValueAsFloat64(column string) (float64, error) {
	val, ok := row[column]
	if !ok {
		return 0, fmt.Errorf(`missing column %q`, column)
	}
	res, err := strconv.ParseFloat(val, 64)
	if err != nil {
		if ne, isNum := err.(*strconv.NumError); isNum {
			return res, fmt.Errorf(`column %q: cannot convert %q to float: %s`, column, val, ne.Err)
		}
		return res, fmt.Errorf(`column %q: %s`, column, err.Error())
	}
	return res, nil
}
