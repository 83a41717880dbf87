// Package csvplus — cgo binding of the B200 hot path behind the csvplus API names.
//
// SOURCE ONLY: this image has no Go toolchain, so this file has never been compiled.  It shows the
// reference-side stub INTEGRATION.md describes; the same C ABI is exercised from Python (ctypes) and C++.
package csvplus

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../csvplus_b200 -lcsvplus_b200
#include <stdlib.h>
#include <string.h>
#include "csvplus_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"io"
	"os"
	"runtime"
	"unsafe"
)

// Row, RowFunc, DataSourceError keep the reference's definitions (csvplus.go:59, :208, :1230).
type Row map[string]string
type RowFunc func(Row) error

type DataSourceError struct {
	Line uint64
	Err  error
}

func (e *DataSourceError) Error() string { return fmt.Sprintf(`row %d: %s`, e.Line, e.Err) }

type context struct{ h *C.cpb_ctx }

var defaultCtx *context

func ctx() *context {
	if defaultCtx == nil {
		var h *C.cpb_ctx
		if st := C.cpb_init(0, &h); st != C.CPB_OK {
			panic("csvplus: no usable CUDA device (there is no CPU fallback)")
		}
		defaultCtx = &context{h}
	}
	return defaultCtx
}

func mapErr(st C.int, e *C.cpb_error) error {
	if st == C.CPB_OK {
		return nil
	}
	msg := C.GoString(&e.msg[0])
	if st == C.CPB_ERR_DATA && e.has_line != 0 {
		return &DataSourceError{Line: uint64(e.line), Err: errors.New(msg)}
	}
	return errors.New(msg)
}

// Table wraps cpb_table.
type Table struct{ h *C.cpb_table }

func newTable(h *C.cpb_table) *Table {
	t := &Table{h}
	runtime.SetFinalizer(t, func(t *Table) { C.cpb_table_free(t.h) })
	return t
}

// Reader mirrors csvplus.go:924-1076; only the fields the ABI needs.
type Reader struct {
	name                         string
	delimiter, comment           rune
	numFields                    int
	lazyQuotes, trimLeadingSpace bool
	header                       map[string]int
	headerFromFirstRow           bool
}

func FromFile(name string) *Reader {
	return &Reader{name: name, delimiter: ',', headerFromFirstRow: true}
}

func (r *Reader) SelectColumns(names ...string) *Reader {
	if len(names) == 0 {
		panic("empty header spec")
	}
	r.header = make(map[string]int, len(names))
	for _, n := range names {
		if _, found := r.header[n]; found {
			panic("header spec: duplicate column name: " + n)
		}
		r.header[n] = -1
	}
	r.headerFromFirstRow = true
	return r
}

// parse lowers Reader.Iterate (+ an optional recognised predicate) to cpb_parse_csv.
func (r *Reader) parse(pred *C.cpb_pred) (*Table, error) {
	data, err := os.ReadFile(r.name) // the production wrapper reads into cpb_host_alloc'ed pinned memory
	if err != nil {
		var pe *os.PathError
		if errors.As(err, &pe) {
			return nil, &DataSourceError{Line: 1, Err: errors.New(pe.Op + ": " + pe.Err.Error())}
		}
		return nil, &DataSourceError{Line: 1, Err: err}
	}
	opts := C.cpb_reader_opts{delimiter: C.uint32_t(r.delimiter), comment: C.uint32_t(r.comment),
		num_fields: C.int32_t(r.numFields)}
	if r.lazyQuotes {
		opts.lazy_quotes = 1
	}
	if r.trimLeadingSpace {
		opts.trim_leading_space = 1
	}
	if r.headerFromFirstRow {
		opts.header_from_first_row = 1
	}
	spec := make([]C.cpb_header_col, 0, len(r.header))
	var keep []unsafe.Pointer
	for name, idx := range r.header {
		p := C.CString(name)
		keep = append(keep, unsafe.Pointer(p))
		spec = append(spec, C.cpb_header_col{name: C.cpb_str{ptr: p, len: C.uint64_t(len(name))}, index: C.int32_t(idx)})
	}
	defer func() {
		for _, p := range keep {
			C.free(p)
		}
	}()
	var sp *C.cpb_header_col
	if len(spec) > 0 {
		sp = &spec[0]
	}
	var out *C.cpb_table
	var e C.cpb_error
	var dp unsafe.Pointer
	if len(data) > 0 {
		dp = unsafe.Pointer(&data[0])
	}
	st := C.cpb_parse_csv(ctx().h, dp, C.uint64_t(len(data)), 0, &opts, sp, C.int(len(spec)), pred, &out, &e)
	var t *Table
	if out != nil {
		t = newTable(out)
	}
	return t, mapErr(st, &e)
}

// rows materialises a table as []Row for opaque Go closures and the final RowFunc.
func (t *Table) rows() []Row {
	n := int64(C.cpb_table_num_rows(t.h))
	nc := int(C.cpb_table_num_cols(t.h))
	rows := make([]Row, n)
	for i := range rows {
		rows[i] = make(Row, nc)
	}
	for c := 0; c < nc; c++ {
		var nm C.cpb_str
		C.cpb_table_col_name(t.h, C.int(c), &nm)
		name := C.GoStringN(nm.ptr, C.int(nm.len))
		var nb C.uint64_t
		C.cpb_table_col_bytes(ctx().h, t.h, C.int(c), 0, C.int64_t(n), &nb)
		off := make([]int64, n+1)
		data := make([]byte, nb+1)
		C.cpb_table_fetch_column(ctx().h, t.h, C.int(c), 0, C.int64_t(n), (*C.int64_t)(unsafe.Pointer(&off[0])),
			(*C.uint8_t)(unsafe.Pointer(&data[0])), C.uint64_t(len(data)))
		for i := int64(0); i < n; i++ {
			rows[i][name] = string(data[off[i]:off[i+1]])
		}
	}
	return rows
}

// DataSource keeps the reference's shape: calling it pulls the rows (csvplus.go:215).
type DataSource func(RowFunc) error

// Take(reader): parse on the GPU, then hand the rows to fn; a DataSourceError is returned after the rows
// delivered before it, like the streaming reference.
func Take(r *Reader) DataSource {
	return func(fn RowFunc) error {
		t, perr := r.parse(nil)
		if t != nil {
			for _, row := range t.rows() {
				if err := fn(row); err != nil {
					if err == io.EOF {
						return nil
					}
					return err
				}
			}
		}
		return perr
	}
}
