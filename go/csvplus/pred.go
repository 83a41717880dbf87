package csvplus

// pred.go: Like / All / Any / Not (csvplus.go:1243-1293).  They return plain func(Row) bool values, as in the
// reference, so they also work on host rows; a predicate built only from these constructors is *recognisable*: when a
// Filter receives one it asks the closure for its description (a probe with a sentinel row) and lowers it to a
// cpb_pred tree evaluated inside the scan kernel.  Any other func(Row) bool is opaque and runs on the host.

/*
#include <stdlib.h>
#include "csvplus_b200.h"
*/
import "C"

import (
	"reflect"
	"sync"
	"unsafe"
)

const (
	opLike = C.CPB_PRED_LIKE
	opAll  = C.CPB_PRED_ALL
	opAny  = C.CPB_PRED_ANY
	opNot  = C.CPB_PRED_NOT
)

type predSpec struct {
	op    int
	keys  []string
	vals  []string
	kids  []*predSpec
}

var (
	probeMu      sync.Mutex
	predProbeRow = Row{} // identity sentinel: recognisable predicates answer it with their description
	predReply    *predSpec
)

func isPredProbe(row Row) bool {
	return reflect.ValueOf(row).Pointer() == reflect.ValueOf(predProbeRow).Pointer()
}

// describe returns the cpb_pred description of a predicate built by Like/All/Any/Not, nil for an opaque one.
// (A foreign predicate merely sees one call with an empty row; predicates must be pure in the reference too.)
func describe(pred func(Row) bool) *predSpec {
	probeMu.Lock()
	defer probeMu.Unlock()
	predReply = nil
	pred(predProbeRow)
	r := predReply
	predReply = nil
	return r
}

// Like is true for rows that have every column of match with the same value (csvplus.go:1279-1293).
func Like(match Row) func(Row) bool {
	if len(match) == 0 {
		panic("empty match row in Like() predicate")
	}
	spec := &predSpec{op: opLike}
	for _, k := range match.Header() {
		spec.keys = append(spec.keys, k)
		spec.vals = append(spec.vals, match[k])
	}
	return func(row Row) bool {
		if isPredProbe(row) {
			predReply = spec
			return false
		}
		for i, k := range spec.keys {
			if v, ok := row[k]; !ok || v != spec.vals[i] {
				return false
			}
		}
		return true
	}
}

func combine(op int, funcs []func(Row) bool, eval func(Row) bool) func(Row) bool {
	spec := &predSpec{op: op}
	for _, f := range funcs {
		k := describe(f)
		if k == nil {
			spec = nil // one opaque child makes the combination opaque
			break
		}
		spec.kids = append(spec.kids, k)
	}
	return func(row Row) bool {
		if isPredProbe(row) {
			predReply = spec
			return false
		}
		return eval(row)
	}
}

// All is the conjunction of the predicates (csvplus.go:1243-1253).
func All(funcs ...func(Row) bool) func(Row) bool {
	return combine(opAll, funcs, func(row Row) bool {
		for _, f := range funcs {
			if !f(row) {
				return false
			}
		}
		return true
	})
}

// Any is the disjunction of the predicates (csvplus.go:1256-1266).
func Any(funcs ...func(Row) bool) func(Row) bool {
	return combine(opAny, funcs, func(row Row) bool {
		for _, f := range funcs {
			if f(row) {
				return true
			}
		}
		return false
	})
}

// Not negates the predicate (csvplus.go:1269-1273).
func Not(pred func(Row) bool) func(Row) bool {
	return combine(opNot, []func(Row) bool{pred}, func(row Row) bool { return !pred(row) })
}

// cPred is a cpb_pred tree in C memory.
type cPred struct {
	root  *C.cpb_pred
	frees []unsafe.Pointer
}

func (p *cPred) alloc(n C.size_t) unsafe.Pointer {
	m := C.calloc(1, n)
	p.frees = append(p.frees, m)
	return m
}

func (p *cPred) str(s string) C.cpb_str {
	c := C.CString(s)
	p.frees = append(p.frees, unsafe.Pointer(c))
	return C.cpb_str{ptr: c, len: C.uint64_t(len(s))}
}

func (p *cPred) build(s *predSpec) *C.cpb_pred {
	node := (*C.cpb_pred)(p.alloc(C.size_t(unsafe.Sizeof(C.cpb_pred{}))))
	node.op = C.int32_t(s.op)
	if s.op == opLike {
		n := len(s.keys)
		node.n = C.int32_t(n)
		keys := (*C.cpb_str)(p.alloc(C.size_t(n) * C.size_t(unsafe.Sizeof(C.cpb_str{}))))
		vals := (*C.cpb_str)(p.alloc(C.size_t(n) * C.size_t(unsafe.Sizeof(C.cpb_str{}))))
		ks, vs := unsafe.Slice(keys, n), unsafe.Slice(vals, n)
		for i := range s.keys {
			ks[i], vs[i] = p.str(s.keys[i]), p.str(s.vals[i])
		}
		node.keys, node.values = keys, vals
		return node
	}
	n := len(s.kids)
	node.n = C.int32_t(n)
	var kid *C.cpb_pred
	kids := (**C.cpb_pred)(p.alloc(C.size_t(n+1) * C.size_t(unsafe.Sizeof(kid))))
	kslice := unsafe.Slice(kids, n+1)
	for i, k := range s.kids {
		kslice[i] = p.build(k)
	}
	node.children = kids
	return node
}

func (s *predSpec) toC() *cPred {
	p := &cPred{}
	p.root = p.build(s)
	return p
}

func (p *cPred) free() {
	for _, m := range p.frees {
		C.free(m)
	}
	p.frees = nil
}
