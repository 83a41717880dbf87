package csvplus

// multi.go: the multi-GPU entry points of the C ABI (SURVEY §8e) — not part of the reference's API (csvplus is
// single-threaded, csvplus.go:33-46), offered next to it: the build side of a Join is parsed 1/N per GPU and all-gathered
// (cpb_allgather_tables), the probe side is sharded by row / byte range with no data-path collective.

/*
#include <stdlib.h>
#include <string.h>
#include "csvplus_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"unsafe"
)

type unsafePointer = unsafe.Pointer

func unsafePointerOf(b []byte) unsafe.Pointer { return unsafe.Pointer(&b[0]) }

// AllGatherTables concatenates, on every context of an InitMulti communicator, the rows of all the contexts' tables in
// context order (cpb_allgather_tables: grouped NCCL broadcasts straight into the final columns, one host round trip).
func AllGatherTables(ctxs []*Context, locals []*Table) ([]*Table, error) {
	if len(ctxs) == 0 || len(ctxs) != len(locals) {
		return nil, errors.New("csvplus: AllGatherTables needs one table per context")
	}
	hs := make([]*C.cpb_ctx, len(ctxs))
	ts := make([]*C.cpb_table, len(ctxs))
	outs := make([]*C.cpb_table, len(ctxs))
	for i := range ctxs {
		hs[i], ts[i] = ctxs[i].h, locals[i].h
	}
	if st := C.cpb_allgather_tables(&hs[0], &ts[0], C.int(len(ctxs)), &outs[0]); st != C.CPB_OK {
		return nil, fmt.Errorf("csvplus: cpb_allgather_tables failed with status %d: %s", int(st), C.GoString(C.cpb_last_error(hs[0])))
	}
	res := make([]*Table, len(ctxs))
	for i := range outs {
		res[i] = newTable(ctxs[i], outs[i])
	}
	return res, nil
}

// TakeTable turns a device table into a DataSource (rows are numbered from lineBase in errors raised by callbacks).
func TakeTable(t *Table, lineBase uint64) DataSource {
	return newSource(&plan{kind: opTable, table: t, lineBase: lineBase})
}

// ParseShard parses the byte range of one file that `rank` of `world` owns (cpb_parse_csv_shard).  buf holds the bytes
// [lo, hi + look-ahead) of the file, own = hi - lo; parity is the XOR of QuoteParity of the shards before it; fields /
// numFields are the resolved header (Table.ParsedFrom of the first shard or of the head of the file).  Returns the rows,
// the number of records the shard owns and, on a data error, a DataSourceError whose Line is LOCAL (0-based ordinal among
// the shard's records): the caller adds the records of the shards before it and the reader's base.
func ParseShard(c *Context, buf []byte, own uint64, rank int, isLast bool, parity uint32, r *Reader, numFields int) (*Table, uint64, error) {
	opts := C.cpb_reader_opts{delimiter: C.uint32_t(r.delimiter), num_fields: C.int32_t(numFields)}
	if r.headerFromFirstRow {
		opts.header_from_first_row = 1
	}
	names := make([]string, 0, len(r.header))
	for name := range r.header {
		names = append(names, name)
	}
	cn := newCstrs(names)
	defer cn.free()
	spec := make([]C.cpb_header_col, len(names)+1)
	for i, name := range names {
		spec[i] = C.cpb_header_col{name: cn.arr[i], index: C.int32_t(r.header[name])}
	}
	var staging unsafePointer
	if len(buf) > 0 {
		if st := C.cpb_host_alloc(c.h, C.uint64_t(len(buf)), &staging); st != C.CPB_OK {
			return nil, 0, errors.New("csvplus: cpb_host_alloc failed")
		}
		defer C.cpb_host_free(c.h, staging)
		C.memcpy(staging, unsafePointerOf(buf), C.size_t(len(buf)))
	}
	last := C.int(0)
	if isLast {
		last = 1
	}
	var out *C.cpb_table
	var recs C.uint64_t
	var e C.cpb_error
	st := C.cpb_parse_csv_shard(c.h, staging, C.uint64_t(len(buf)), 0, C.uint64_t(own), C.int(rank), last, C.uint32_t(parity), &opts,
		&spec[0], C.int(len(names)), nil, &out, &recs, &e)
	var t *Table
	if out != nil {
		t = newTable(c, out)
	}
	return t, uint64(recs), mapErr(C.int(st), &e)
}

// QuoteParity is the parity of the quote bytes of a shard's own byte range (cpb_csv_quote_parity).
func QuoteParity(c *Context, own []byte) (uint32, error) {
	var p C.uint32_t
	var ptr unsafePointer
	if len(own) > 0 {
		ptr = unsafePointerOf(own)
	}
	if st := C.cpb_csv_quote_parity(c.h, ptr, C.uint64_t(len(own)), 0, &p); st != C.CPB_OK {
		return 0, fmt.Errorf("csvplus: cpb_csv_quote_parity failed with status %d", int(st))
	}
	return uint32(p), nil
}
