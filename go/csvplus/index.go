package csvplus

// index.go: IndexOn / UniqueIndexOn / Index (csvplus.go:529-537, :610-767) over cpb_index.

/*
#include <stdlib.h>
#include "csvplus_b200.h"
*/
import "C"

import (
	"encoding/gob"
	"errors"
	"fmt"
	"os"
	"runtime"
	"unsafe"
)

// Index is a collection of rows sorted on the key columns, resident in HBM, with hash-probe lookups for joins and
// binary-search lookups for Find (csvplus.go:610-614).
type Index struct {
	c       *Context
	h       *C.cpb_index
	columns []string
}

func newIndex(c *Context, h *C.cpb_index, columns []string) *Index {
	ix := &Index{c, h, columns}
	runtime.SetFinalizer(ix, func(ix *Index) { ix.Close() })
	return ix
}

// Close releases the device memory of the index.
func (index *Index) Close() {
	if index.h != nil {
		C.cpb_index_free(index.h)
		index.h = nil
	}
}

func (index *Index) table() *Table {
	var t *C.cpb_table
	C.cpb_index_table(index.c.h, index.h, &t)
	return newTable(index.c, t)
}

func createIndex(src DataSource, columns []string, unique bool) (*Index, error) {
	switch len(columns) {
	case 0:
		panic("empty column list in CreateIndex()") // csvplus.go:709-710
	case 1:
	default:
		seen := make(map[string]struct{}, len(columns))
		for _, col := range columns {
			if _, dup := seen[col]; dup {
				panic("duplicate column name(s) in CreateIndex()") // csvplus.go:714-716
			}
			seen[col] = struct{}{}
		}
	}
	p := planOf(src)
	res := p.eval()
	if res.err != nil { // the reference builds nothing when the source fails (csvplus.go:729-731)
		return nil, res.err
	}
	c := p.context()
	t, err := res.deviceTable(c)
	if err != nil {
		return nil, err
	}
	if t == nil {
		if t, err = tableFromRows(c, nil); err != nil {
			return nil, err
		}
	}
	cn := newCstrs(columns)
	defer cn.free()
	var h *C.cpb_index
	var e C.cpb_error
	u := C.int(0)
	if unique {
		u = 1
	}
	if st := C.cpb_index_build(c.h, t.h, cn.ptr(), C.int(len(columns)), u, &h, &e); st != C.CPB_OK {
		return nil, mapErr(C.int(st), &e)
	}
	return newIndex(c, h, append([]string(nil), columns...)), nil
}

// IndexOn builds an index on the columns, left to right (csvplus.go:529).
func (src DataSource) IndexOn(columns ...string) (*Index, error) { return createIndex(src, columns, false) }

// UniqueIndexOn also requires the keys to be unique (csvplus.go:535): the error names the duplicate key like the
// reference does (`duplicate value while creating unique index: { "id" : "7" }`).
func (src DataSource) UniqueIndexOn(columns ...string) (*Index, error) { return createIndex(src, columns, true) }

// Iterate calls fn for every row in index order (csvplus.go:618).
func (index *Index) Iterate(fn RowFunc) error { return Take(index)(fn) }

// Find returns the rows whose leading key columns equal the values (csvplus.go:625).
func (index *Index) Find(values ...string) DataSource {
	if len(values) > len(index.columns) {
		panic("too many columns in indexImpl.find()") // csvplus.go:876-878
	}
	cv := newCstrs(values)
	defer cv.free()
	var t *C.cpb_table
	if st := C.cpb_index_find(index.c.h, index.h, cv.ptr(), C.int(len(values)), &t); st != C.CPB_OK {
		return TakeRows(nil)
	}
	return newSource(&plan{kind: opTable, table: newTable(index.c, t), lineBase: 0})
}

// SubIndex returns the index of the rows matching the values, keyed by the remaining columns (csvplus.go:632).
func (index *Index) SubIndex(values ...string) *Index {
	if len(values) >= len(index.columns) {
		panic("too many values in SubIndex()")
	}
	cv := newCstrs(values)
	defer cv.free()
	var h *C.cpb_index
	if st := C.cpb_index_sub(index.c.h, index.h, cv.ptr(), C.int(len(values)), &h); st != C.CPB_OK {
		panic(fmt.Sprintf("csvplus: cpb_index_sub failed with status %d", int(st)))
	}
	return newIndex(index.c, h, append([]string(nil), index.columns[len(values):]...))
}

// ResolveDuplicates calls resolve once per run of rows with equal keys (csvplus.go:651, dedup :810-867).  The groups are
// found on the GPU (cpb_index_dup_groups), resolve runs on the host like the Go closure of the reference, and the
// choices are applied on the GPU (cpb_index_dedup_apply2), including a returned row that is not one of the group's rows.
// The reference's loss of a trailing singleton (SURVEY §Q1) is reproduced.
func (index *Index) ResolveDuplicates(resolve func(rows []Row) (Row, error)) error {
	var ng C.int64_t
	var lo, hi *C.int64_t
	if st := C.cpb_index_dup_groups(index.c.h, index.h, &ng, &lo, &hi); st != C.CPB_OK {
		return errors.New("csvplus: cpb_index_dup_groups failed")
	}
	defer C.cpb_free(unsafe.Pointer(lo))
	defer C.cpb_free(unsafe.Pointer(hi))
	n := int(ng)
	if n == 0 {
		return nil
	}
	los, his := unsafe.Slice(lo, n), unsafe.Slice(hi, n)
	keep := make([]C.int64_t, n)
	var replacements []Row
	t := index.table()
	defer t.Close()
	for g := 0; g < n; g++ {
		rows, err := t.rows(int64(los[g]), int64(his[g]))
		if err != nil {
			return err
		}
		chosen, err := resolve(rows)
		if err != nil {
			return err // the index is left unchanged (csvplus.go:862-864)
		}
		if len(chosen) < len(index.columns) { // csvplus.go:845: an "empty" row drops the group
			keep[g] = -1
			continue
		}
		pick := -1
		for i, r := range rows {
			if reflect_same(r, chosen) {
				pick = i
				break
			}
		}
		if pick >= 0 {
			keep[g] = los[g] + C.int64_t(pick)
		} else {
			keep[g] = C.int64_t(-2 - len(replacements))
			replacements = append(replacements, chosen)
		}
	}
	var repl *C.cpb_table
	if len(replacements) > 0 {
		rt, err := tableFromRows(index.c, replacements)
		if err != nil {
			return err
		}
		defer rt.Close()
		repl = rt.h
	}
	var e C.cpb_error
	if st := C.cpb_index_dedup_apply2(index.c.h, index.h, C.int64_t(n), &keep[0], repl, 1, &e); st != C.CPB_OK {
		return mapErr(C.int(st), &e)
	}
	return nil
}

// reflect_same: is `b` the very map `a` (the resolver returned one of the rows it was given) or an equal row
func reflect_same(a, b Row) bool {
	if len(a) != len(b) {
		return false
	}
	for k, v := range a {
		if w, ok := b[k]; !ok || w != v {
			return false
		}
	}
	return true
}

// WriteTo writes the index to a file in the reference's on-disk format (csvplus.go:655-681): gob of the column list,
// then gob of the rows, so files are interchangeable with the reference's LoadIndex.
func (index *Index) WriteTo(fileName string) (err error) {
	t := index.table()
	defer t.Close()
	rows, err := t.rows(0, t.NumRows())
	if err != nil {
		return err
	}
	file, err := os.Create(fileName)
	if err != nil {
		return err
	}
	defer func() {
		if e := file.Close(); e != nil || err != nil {
			os.Remove(fileName)
			if err == nil {
				err = e
			}
		}
	}()
	enc := gob.NewEncoder(file)
	if err = enc.Encode(index.columns); err == nil {
		err = enc.Encode(rows)
	}
	return err
}

// LoadIndex reads an index written by WriteTo — this package's or the reference's (csvplus.go:683-705) — and uploads it.
// The rows are stored sorted, so the key image is rebuilt without a sort changing their order.
func LoadIndex(fileName string) (*Index, error) {
	file, err := os.Open(fileName)
	if err != nil {
		return nil, err
	}
	defer file.Close()
	var columns []string
	var rows []Row
	dec := gob.NewDecoder(file)
	if err = dec.Decode(&columns); err != nil {
		return nil, err
	}
	if err = dec.Decode(&rows); err != nil {
		return nil, err
	}
	return TakeRows(rows).IndexOn(columns...)
}
