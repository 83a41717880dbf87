"""Host-side mirror of the csvplus Go API (csvplus.go) over the C ABI (include/csvplus_b200.h).

The reference's toolchain (Go) is absent from this image, so this module — together with
host/csvplus.hpp (C++) — plays the role of the thin Go package described in INTEGRATION.md:
same names, argument meaning and error text as the reference, so that tests read like
csvplus_test.go.  Nothing here parses, filters, sorts or joins on the CPU: a `DataSource` is a
*plan*; a sink (ToRows / ToCsv / IndexOn / UniqueIndexOn / calling the source with a RowFunc)
lowers the recognisable prefix (parse -> SelectColumns -> Filter(Like/All/Any/Not) -> Join ...) to
C-ABI calls running CUDA kernels.  Opaque Python callables (Map, Filter(func), Transform, ...)
run on the host at a materialisation boundary, exactly like opaque Go closures would,
and whatever follows them is uploaded again with TakeRows semantics (cpb_table_from_host).
"""
from __future__ import annotations

import ctypes as C
import io
import os
from typing import Callable, Iterable

import numpy as np

from . import _abi

Row = dict  # map[string]string, csvplus.go:59 (values are str; bytes survive via surrogateescape)


def _enc(s) -> bytes:
    return s if isinstance(s, (bytes, bytearray)) else str(s).encode("utf-8", "surrogateescape")


def _dec(b: bytes) -> str:
    return b.decode("utf-8", "surrogateescape")


def _go_json_string(s) -> bytes:
    """encoding/json's string encoding with HTML escaping off: `"` `\\` and control characters escaped (\\n \\r \\t, else
    \\u00xx), U+2028 / U+2029 escaped, every invalid UTF-8 byte replaced by \\ufffd, everything else verbatim"""
    b = _enc(s)
    out = bytearray(b'"')
    i, n = 0, len(b)
    while i < n:
        c = b[i]
        if c < 0x80:
            if c == 0x22: out += b'\\"'
            elif c == 0x5C: out += b"\\\\"
            elif c == 0x0A: out += b"\\n"
            elif c == 0x0D: out += b"\\r"
            elif c == 0x09: out += b"\\t"
            elif c < 0x20: out += b"\\u00" + b"%02x" % c
            else: out.append(c)
            i += 1
            continue
        width = 2 if 0xC2 <= c <= 0xDF else 3 if 0xE0 <= c <= 0xEF else 4 if 0xF0 <= c <= 0xF4 else 0
        chunk = b[i:i + width]
        try:
            ch = chunk.decode("utf-8") if width and len(chunk) == width else None
        except UnicodeDecodeError:
            ch = None
        if ch is None:
            out += b"\\ufffd"; i += 1
        elif ch in ("\u2028", "\u2029"):
            out += b"\\u2028" if ch == "\u2028" else b"\\u2029"; i += width
        else:
            out += chunk; i += width
    out += b'"'
    return bytes(out)


class DataSourceError(Exception):
    """csvplus.go:1230-1238: `row %d: %s`"""

    def __init__(self, line: int, err: str, kind: int = 0):
        self.Line, self.Err, self.kind = line, err, kind
        super().__init__(f"row {line}: {err}")


class CsvPlusError(Exception):
    """errors the reference returns without a row number (e.g. duplicate key) or panics with"""

    def __init__(self, msg, status=0, kind=0):
        super().__init__(msg)
        self.status, self.kind = status, kind


class StopIterationEOF(Exception):
    """a RowFunc raises this to stop the iteration cleanly (io.EOF, csvplus.go:238-239)"""


def _strs(items: Iterable) -> tuple:
    bs = [_enc(x) for x in items]
    arr = (_abi.Str * max(1, len(bs)))()
    for i, b in enumerate(bs):
        arr[i].ptr = b
        arr[i].len = len(b)
    return arr, bs  # keep bs alive


def _raise(st: int, err: _abi.Error, ctx=None):
    msg = err.msg.decode("utf-8", "replace") if err is not None else ""
    if st == 1:  # CPB_ERR_DATA
        if err.has_line:
            raise DataSourceError(err.line, msg, err.kind)
        raise CsvPlusError(msg, st, err.kind)
    if not msg and ctx is not None:
        msg = _abi.load().cpb_last_error(ctx.h).decode("utf-8", "replace")
    raise CsvPlusError(msg or f"csvplus_b200 call failed with status {st}", st, err.kind if err is not None else 0)


class Context:
    """cpb_ctx: one device, one stream.  `default()` gives a process-wide context on cuda:LOCAL_RANK."""

    _default = None

    def __init__(self, device: int = 0):
        self.lib = _abi.load()
        h = C.c_void_p()
        st = self.lib.cpb_init(device, C.byref(h))
        if st != 0:
            raise CsvPlusError(f"cpb_init(device={device}) failed with status {st}: no usable CUDA device "
                               "(csvplus_b200 has no CPU fallback)", st)
        self.h, self.device = h, device

    @classmethod
    def default(cls) -> "Context":
        if cls._default is None:
            cls._default = cls(int(os.environ.get("LOCAL_RANK", "0")))
        return cls._default

    def close(self):
        if self.h:
            self.lib.cpb_shutdown(self.h)
            self.h = None

    def sync(self):
        self.lib.cpb_sync(self.h)

    def reserve(self, nbytes: int) -> bool:
        """cpb_pool_reserve: map `nbytes` of device memory into this context's pool ahead of time (False: not available)"""
        return self.lib.cpb_pool_reserve(self.h, int(nbytes)) == 0

    @property
    def stream(self) -> int:
        return self.lib.cpb_ctx_stream(self.h) or 0

    # ---- measurement
    def stats(self, enable: bool | None = None, reset: bool = False):
        if enable is not None:
            self.lib.cpb_stats_enable(self.h, int(enable))
        if reset:
            self.lib.cpb_stats_reset(self.h)
            return {}
        arr = (_abi.KStat * 64)()
        n = C.c_int()
        self.lib.cpb_stats_get(self.h, arr, 64, C.byref(n))
        return {arr[i].name.decode(): {"launches": arr[i].launches, "ms": arr[i].ms, "algo_bytes": arr[i].algo_bytes}
                for i in range(min(n.value, 64))}

    def kernel_launches(self) -> int:
        return self.lib.cpb_kernel_launches(self.h)

    def host_syncs(self) -> int:
        return self.lib.cpb_host_syncs(self.h)

    # ---- staging memory
    def host_alloc(self, n: int) -> "HostBuffer":
        return HostBuffer(self, n)

    def device_alloc(self, n: int) -> "DeviceBuffer":
        return DeviceBuffer(self, n)

    def gen_csv(self, kind: str, rows: tuple, seed=0xC5B200, n_cust=1, n_prod=1, header=True, permute=False,
                chunk_rows=20_000_000) -> "DeviceBuffer":
        """synthetic CSV of SURVEY §8(d) generated on the GPU (bench / tests only)"""
        k = {"people": 0, "customers": 0, "orders": 1, "products": 2}[kind]
        lo, hi = rows
        sizes, total = [], 0
        r = lo
        while True:
            e = min(hi, r + chunk_rows)
            nb = C.c_uint64()
            st = self.lib.cpb_gen_csv(self.h, k, seed, r, e, n_cust, n_prod, int(header and r == lo), int(permute), None, 0, C.byref(nb))
            if st:
                _raise(st, None, self)
            sizes.append((r, e, nb.value)); total += nb.value
            r = e
            if r >= hi:
                break
        buf = self.device_alloc(total)
        off = 0
        for (r0, r1, nb) in sizes:
            got = C.c_uint64()
            st = self.lib.cpb_gen_csv(self.h, k, seed, r0, r1, n_cust, n_prod, int(header and r0 == lo), int(permute),
                                      C.c_void_p(buf.ptr + off), nb, C.byref(got))
            if st:
                _raise(st, None, self)
            off += got.value
        buf.nbytes = total
        return buf


class HostBuffer:
    def __init__(self, ctx: Context, n: int):
        self.ctx, self.nbytes = ctx, n
        p = C.c_void_p()
        st = ctx.lib.cpb_host_alloc(ctx.h, n, C.byref(p))
        if st:
            _raise(st, None, ctx)
        self.ptr = p.value

    def array(self) -> np.ndarray:
        return np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.ctx.lib.cpb_host_free(self.ctx.h, C.c_void_p(self.ptr)); self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceBuffer:
    def __init__(self, ctx: Context, n: int):
        self.ctx, self.nbytes = ctx, n
        p = C.c_void_p()
        st = ctx.lib.cpb_device_alloc(ctx.h, n, C.byref(p))
        if st:
            _raise(st, None, ctx)
        self.ptr = p.value

    def to_host(self, n: int | None = None, offset: int = 0) -> np.ndarray:
        n = self.nbytes - offset if n is None else n
        out = np.empty(n, np.uint8)
        st = self.ctx.lib.cpb_memcpy_d2h(self.ctx.h, out.ctypes.data, C.c_void_p(self.ptr + offset), n)
        if st:
            _raise(st, None, self.ctx)
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.cpb_device_free(self.ctx.h, C.c_void_p(self.ptr)); self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ------------------------------------------------------------------ predicates (csvplus.go:1240-1293)
class Predicate:
    """A recognisable predicate: callable on a Row (host) and lowerable to cpb_pred (device)."""

    def __init__(self, op: int, match: dict | None = None, kids: list | None = None):
        self.op, self.match, self.kids = op, match or {}, kids or []

    def __call__(self, row: Row) -> bool:
        if self.op == 0:
            return all(k in row and row[k] == v for k, v in self.match.items())
        if self.op == 1:
            return all(k(row) for k in self.kids)
        if self.op == 2:
            return any(k(row) for k in self.kids)
        return not self.kids[0](row)

    def lowerable(self) -> bool:
        return all(isinstance(k, Predicate) and k.lowerable() for k in self.kids)

    def _c(self, keep: list) -> _abi.Pred:
        p = _abi.Pred()
        p.op = self.op
        if self.op == 0:
            ka, kb = _strs(self.match.keys()); va, vb = _strs(self.match.values())
            keep += [ka, kb, va, vb]
            p.n, p.keys, p.values = len(self.match), ka, va
        else:
            kids = [k._c(keep) for k in self.kids]
            arr = (C.POINTER(_abi.Pred) * max(1, len(kids)))(*[C.pointer(k) for k in kids])
            keep += [kids, arr]
            p.n, p.children = len(kids), arr
        return p


def Like(match: Row) -> Predicate:
    if len(match) == 0:
        raise ValueError("empty match row in Like() predicate")  # panic, csvplus.go:1280-1282
    return Predicate(0, {str(k): (v if isinstance(v, str) else _dec(_enc(v))) for k, v in match.items()})


def All(*funcs) -> Callable[[Row], bool]:
    return Predicate(1, kids=list(funcs)) if all(isinstance(f, Predicate) for f in funcs) else (lambda row: all(f(row) for f in funcs))


def Any(*funcs) -> Callable[[Row], bool]:
    return Predicate(2, kids=list(funcs)) if all(isinstance(f, Predicate) for f in funcs) else (lambda row: any(f(row) for f in funcs))


def Not(pred) -> Callable[[Row], bool]:
    return Predicate(3, kids=[pred]) if isinstance(pred, Predicate) else (lambda row: not pred(row))


# ------------------------------------------------------------------ device tables
class Table:
    """cpb_table handle: a columnar batch of rows in HBM."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h

    def __del__(self):
        try:
            if self.h:
                self.ctx.lib.cpb_table_free(self.h); self.h = None
        except Exception:
            pass

    def __len__(self):
        return self.ctx.lib.cpb_table_num_rows(self.h)

    @property
    def columns(self) -> list[str]:
        out = []
        for i in range(self.ctx.lib.cpb_table_num_cols(self.h)):
            s = _abi.Str()
            self.ctx.lib.cpb_table_col_name(self.h, i, C.byref(s))
            out.append(_dec(C.string_at(s.ptr, s.len)))
        return out

    def column(self, name: str, lo: int = 0, hi: int | None = None):
        """-> (offsets int64[n+1], data uint8[]) of rows [lo,hi) copied to the host"""
        hi = len(self) if hi is None else hi
        idx = self.columns.index(name)
        nb = C.c_uint64()
        st = self.ctx.lib.cpb_table_col_bytes(self.ctx.h, self.h, idx, lo, hi, C.byref(nb))
        if st:
            _raise(st, None, self.ctx)
        off = np.empty(hi - lo + 1, np.int64)
        data = np.empty(max(1, nb.value), np.uint8)
        st = self.ctx.lib.cpb_table_fetch_column(self.ctx.h, self.h, idx, lo, hi, off.ctypes.data, data.ctypes.data, data.size)
        if st:
            _raise(st, None, self.ctx)
        return off, data[: nb.value]

    def values(self, name: str, lo: int = 0, hi: int | None = None) -> list[bytes]:
        off, data = self.column(name, lo, hi)
        d = data.tobytes()
        return [d[off[i]:off[i + 1]] for i in range(len(off) - 1)]

    def rows(self, lo: int = 0, hi: int | None = None) -> list[Row]:
        cols = self.columns
        vals = {c: self.values(c, lo, hi) for c in cols}
        n = (len(self) if hi is None else hi) - lo
        return [{c: _dec(vals[c][i]) for c in cols} for i in range(n)]

    def parsed_from(self) -> tuple:
        """(name -> field index of every column, field count of the file's first record) of a parsed table: the resolved
        header that the shards of a file after the first are given (cpb_table_col_field / cpb_table_record_fields)"""
        cols = self.columns
        return ({c: self.ctx.lib.cpb_table_col_field(self.h, i) for i, c in enumerate(cols)}, self.ctx.lib.cpb_table_record_fields(self.h))

    def device_column(self, name: str):
        """raw device pointers (offsets uint32[n+1], data) for zero-copy interop (torch / NCCL)"""
        po, pd = C.c_void_p(), C.c_void_p()
        self.ctx.lib.cpb_table_column_device(self.h, self.columns.index(name), C.byref(po), C.byref(pd))
        return po.value, pd.value

    @staticmethod
    def from_rows(ctx: Context, rows: list[Row], columns: list[str] | None = None) -> "Table":
        """TakeRows (csvplus.go:218): rows must share one column set on the device path"""
        cols = columns if columns is not None else (sorted(rows[0].keys()) if rows else [])
        for r in rows:
            if set(r.keys()) != set(cols):
                raise CsvPlusError("TakeRows on the device requires rows with identical column sets")
        na, nb = _strs(cols)
        offs, datas = [], []
        for cname in cols:
            vals = [_enc(r[cname]) for r in rows]
            off = np.zeros(len(rows) + 1, np.int64)
            if vals:
                off[1:] = np.cumsum([len(v) for v in vals])
            datas.append(np.frombuffer(b"".join(vals) + b"\0", np.uint8).copy()); offs.append(off)
        oa = (C.c_void_p * max(1, len(cols)))(*[o.ctypes.data for o in offs])
        da = (C.c_void_p * max(1, len(cols)))(*[d.ctypes.data for d in datas])
        h = C.c_void_p()
        st = ctx.lib.cpb_table_from_host(ctx.h, len(cols), na, oa, da, len(rows), C.byref(h))
        if st:
            _raise(st, None, ctx)
        return Table(ctx, h)

    @staticmethod
    def from_device_columns(ctx: Context, names: list[str], columns: list[tuple], nrows: int) -> "Table":
        """cpb_table_from_device: columns[k] = (device pointer of uint32 offsets[nrows+1], device pointer of the bytes);
        the buffers are copied, so the caller may release them once the ctx has been synchronised"""
        na, keep = _strs(names)
        oa = (C.c_void_p * max(1, len(names)))(*[c[0] for c in columns])
        da = (C.c_void_p * max(1, len(names)))(*[c[1] or 0 for c in columns])
        h = C.c_void_p()
        st = ctx.lib.cpb_table_from_device(ctx.h, len(names), na, oa, da, nrows, C.byref(h))
        if st:
            _raise(st, None, ctx)
        return Table(ctx, h)

    # thin wrappers of the table-level ABI
    def select(self, *cols) -> "Table":
        a, keep = _strs(cols); h = C.c_void_p(); e = _abi.Error()
        st = self.ctx.lib.cpb_table_select(self.ctx.h, self.h, a, len(cols), C.byref(h), C.byref(e))
        if st:
            _raise(st, e, self.ctx)
        return Table(self.ctx, h)

    def drop(self, *cols) -> "Table":
        a, keep = _strs(cols); h = C.c_void_p()
        st = self.ctx.lib.cpb_table_drop(self.ctx.h, self.h, a, len(cols), C.byref(h))
        if st:
            _raise(st, None, self.ctx)
        return Table(self.ctx, h)

    def filter(self, pred: Predicate) -> "Table":
        keep = []; p = pred._c(keep); h = C.c_void_p()
        st = self.ctx.lib.cpb_table_filter(self.ctx.h, self.h, C.byref(p), C.byref(h))
        if st:
            _raise(st, None, self.ctx)
        return Table(self.ctx, h)

    def first_false(self, pred: Predicate) -> int:
        """index of the first row for which the recognisable predicate is false (len(self) if none): cpb_table_first_false"""
        keep = []; p = pred._c(keep); out = C.c_int64()
        st = self.ctx.lib.cpb_table_first_false(self.ctx.h, self.h, C.byref(p), C.byref(out))
        if st:
            _raise(st, None, self.ctx)
        return out.value

    def slice(self, lo: int, hi: int) -> "Table":
        h = C.c_void_p()
        st = self.ctx.lib.cpb_table_slice(self.ctx.h, self.h, lo, hi, C.byref(h))
        if st:
            _raise(st, None, self.ctx)
        return Table(self.ctx, h)

    def index_on(self, *cols, unique=False) -> "Index":
        a, keep = _strs(cols); h = C.c_void_p(); e = _abi.Error()
        st = self.ctx.lib.cpb_index_build(self.ctx.h, self.h, a, len(cols), int(unique), C.byref(h), C.byref(e))
        if st:
            _raise(st, e, self.ctx)
        return Index(self.ctx, h, list(cols))

    def join(self, index: "Index", *cols, anti=False) -> "Table":
        a, keep = _strs(cols); h = C.c_void_p(); e = _abi.Error()
        fn = self.ctx.lib.cpb_except if anti else self.ctx.lib.cpb_join
        st = fn(self.ctx.h, self.h, index.h, a, len(cols), C.byref(h), C.byref(e))
        if st:
            _raise(st, e, self.ctx)
        return Table(self.ctx, h)

    def to_csv(self, *cols) -> bytes:
        a, keep = _strs(cols); p = C.c_void_p(); n = C.c_uint64(); e = _abi.Error()
        st = self.ctx.lib.cpb_table_to_csv(self.ctx.h, self.h, a, len(cols), C.byref(p), C.byref(n), C.byref(e))
        if st:
            _raise(st, e, self.ctx)
        out = C.string_at(p.value, n.value)
        self.ctx.lib.cpb_host_free(self.ctx.h, p)
        return out

    def to_csv_into(self, dst: "HostBuffer", offset: int, *cols, header: bool = True) -> int:
        """one batch of a streamed ToCsv into pinned host memory (cpb_table_to_csv_into); returns the bytes written"""
        a, keep = _strs(cols); n = C.c_uint64(); e = _abi.Error()
        st = self.ctx.lib.cpb_table_to_csv_into(self.ctx.h, self.h, a, len(cols), int(header), C.c_void_p(dst.ptr + offset),
                                                dst.nbytes - offset, C.byref(n), C.byref(e))
        if st:
            _raise(st, e, self.ctx)
        return n.value

    def to_csv_device(self, *cols) -> "DeviceBuffer":
        """ToCsv with the serialised bytes left in HBM (cpb_table_to_csv_device)"""
        a, keep = _strs(cols); p = C.c_void_p(); n = C.c_uint64(); e = _abi.Error()
        st = self.ctx.lib.cpb_table_to_csv_device(self.ctx.h, self.h, a, len(cols), C.byref(p), C.byref(n), C.byref(e))
        if st:
            _raise(st, e, self.ctx)
        b = DeviceBuffer.__new__(DeviceBuffer)
        b.ctx, b.nbytes, b.ptr = self.ctx, n.value, p.value
        return b

    @staticmethod
    def concat(parts: list["Table"]) -> "Table":
        ctx = parts[0].ctx
        arr = (C.c_void_p * len(parts))(*[p.h for p in parts]); h = C.c_void_p()
        st = ctx.lib.cpb_table_concat(ctx.h, arr, len(parts), C.byref(h))
        if st:
            _raise(st, None, ctx)
        return Table(ctx, h)


def parse_csv(ctx: Context, data, *, on_device=False, nbytes=None, delimiter=",", comment="", num_fields=0, lazy_quotes=False,
              trim_leading_space=False, header_from_first_row=True, spec: list | None = None, pred: Predicate | None = None):
    """cpb_parse_csv.  data: bytes / numpy uint8 / HostBuffer (host) or DeviceBuffer / int pointer (device).
    Returns (Table, DataSourceError | None): the table holds the rows delivered before the error."""
    opts = _abi.ReaderOpts(ord(delimiter), ord(comment) if comment else 0, num_fields, int(lazy_quotes), int(trim_leading_space),
                           int(header_from_first_row), 0)
    keep = []
    if isinstance(data, DeviceBuffer):
        ptr, n, on_device = data.ptr, data.nbytes if nbytes is None else nbytes, True
    elif isinstance(data, HostBuffer):
        ptr, n = data.ptr, data.nbytes if nbytes is None else nbytes
    elif isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data, np.uint8); keep.append(a)
        ptr, n = a.ctypes.data, a.size if nbytes is None else nbytes
    elif isinstance(data, int):
        ptr, n = data, nbytes
    else:
        b = bytes(data); keep.append(b)
        ptr, n = C.cast(C.c_char_p(b), C.c_void_p).value, len(b)
    spec = spec or []
    sa = (_abi.HeaderCol * max(1, len(spec)))()
    for i, (name, idx) in enumerate(spec):
        nb = _enc(name); keep.append(nb)
        sa[i].name.ptr, sa[i].name.len, sa[i].index = nb, len(nb), idx
    pp = None
    if pred is not None:
        pc = pred._c(keep); keep.append(pc); pp = C.byref(pc)
    h = C.c_void_p(); e = _abi.Error()
    st = ctx.lib.cpb_parse_csv(ctx.h, C.c_void_p(ptr), n, int(on_device), C.byref(opts), sa, len(spec), pp, C.byref(h), C.byref(e))
    if st == 1 and h.value:
        return Table(ctx, h), DataSourceError(e.line, e.msg.decode("utf-8", "replace"), e.kind)
    if st:
        _raise(st, e, ctx)
    return Table(ctx, h), None


def _input_ptr(data, nbytes, on_device, keep):
    if isinstance(data, DeviceBuffer):
        return data.ptr, data.nbytes if nbytes is None else nbytes, True
    if isinstance(data, HostBuffer):
        return data.ptr, data.nbytes if nbytes is None else nbytes, on_device
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data, np.uint8); keep.append(a)
        return a.ctypes.data, a.size if nbytes is None else nbytes, on_device
    if isinstance(data, int):
        return data, nbytes, on_device
    b = bytes(data); keep.append(b)
    return C.cast(C.c_char_p(b), C.c_void_p).value, len(b) if nbytes is None else nbytes, on_device


def csv_quote_parity(ctx: Context, data, *, nbytes=None, on_device=False) -> int:
    """cpb_csv_quote_parity: parity of the quote bytes of a shard's own byte range"""
    keep = []
    ptr, n, dev = _input_ptr(data, nbytes, on_device, keep)
    out = C.c_uint32()
    st = ctx.lib.cpb_csv_quote_parity(ctx.h, C.c_void_p(ptr), n, int(dev), C.byref(out))
    if st:
        _raise(st, None, ctx)
    return out.value


def parse_csv_shard(ctx: Context, data, *, own_bytes: int, shard_index: int, is_last: bool, initial_parity: int, nbytes=None, on_device=False,
                    delimiter=",", num_fields=0, header_from_first_row=True, spec: list | None = None, pred: Predicate | None = None):
    """cpb_parse_csv_shard.  Returns (Table, records owned by the shard, DataSourceError | None) — the error's Line is the
    LOCAL 0-based ordinal of the failing record among the shard's records."""
    opts = _abi.ReaderOpts(ord(delimiter), 0, num_fields, 0, 0, int(header_from_first_row), 0)
    keep = []
    ptr, n, dev = _input_ptr(data, nbytes, on_device, keep)
    spec = spec or []
    sa = (_abi.HeaderCol * max(1, len(spec)))()
    for i, (name, idx) in enumerate(spec):
        nb = _enc(name); keep.append(nb)
        sa[i].name.ptr, sa[i].name.len, sa[i].index = nb, len(nb), idx
    pp = None
    if pred is not None:
        pc = pred._c(keep); keep.append(pc); pp = C.byref(pc)
    h = C.c_void_p(); e = _abi.Error(); recs = C.c_uint64()
    st = ctx.lib.cpb_parse_csv_shard(ctx.h, C.c_void_p(ptr), n, int(dev), own_bytes, shard_index, int(is_last), initial_parity, C.byref(opts), sa,
                                     len(spec), pp, C.byref(h), C.byref(recs), C.byref(e))
    if st == 1 and h.value:
        return Table(ctx, h), recs.value, DataSourceError(e.line, e.msg.decode("utf-8", "replace"), e.kind)
    if st:
        _raise(st, e, ctx)
    return Table(ctx, h), recs.value, None


# ------------------------------------------------------------------ Index (csvplus.go:610-705)
class Index:
    def __init__(self, ctx: Context, h, columns: list[str]):
        self.ctx, self.h, self.columns = ctx, h, columns

    def __del__(self):
        try:
            if self.h:
                self.ctx.lib.cpb_index_free(self.h); self.h = None
        except Exception:
            pass

    def __len__(self):
        return self.ctx.lib.cpb_index_num_rows(self.h)

    def table(self) -> Table:
        h = C.c_void_p()
        self.ctx.lib.cpb_index_table(self.ctx.h, self.h, C.byref(h))
        return Table(self.ctx, h)

    def Iterate(self, fn):  # csvplus.go:618-620
        return TakeTable(self.table(), line_base=0)(fn)

    def Find(self, *values) -> "DataSource":  # csvplus.go:625-627
        a, keep = _strs(values); h = C.c_void_p()
        st = self.ctx.lib.cpb_index_find(self.ctx.h, self.h, a, len(values), C.byref(h))
        if st:
            raise ValueError("too many columns in indexImpl.find()") if st == 2 else CsvPlusError("find failed", st)
        return TakeTable(Table(self.ctx, h), line_base=0)

    def SubIndex(self, *values) -> "Index":  # csvplus.go:632-641
        if len(values) >= len(self.columns):
            raise ValueError("too many values in SubIndex()")
        a, keep = _strs(values); h = C.c_void_p()
        st = self.ctx.lib.cpb_index_sub(self.ctx.h, self.h, a, len(values), C.byref(h))
        if st:
            _raise(st, None, self.ctx)
        return Index(self.ctx, h, self.columns[len(values):])

    def dup_groups(self):
        """[lo, hi) sorted-row ranges of every run of >= 2 rows with equal keys (cpb_index_dup_groups) as two int64 arrays"""
        ng = C.c_int64(); lo = C.POINTER(C.c_int64)(); hi = C.POINTER(C.c_int64)()
        st = self.ctx.lib.cpb_index_dup_groups(self.ctx.h, self.h, C.byref(ng), C.byref(lo), C.byref(hi))
        if st:
            _raise(st, None, self.ctx)
        n = ng.value
        try:
            a = np.ctypeslib.as_array(lo, shape=(max(n, 1),))[:n].copy()
            b = np.ctypeslib.as_array(hi, shape=(max(n, 1),))[:n].copy()
        finally:
            self.ctx.lib.cpb_free(lo); self.ctx.lib.cpb_free(hi)
        return a, b

    def dedup_apply(self, keep, bug_compatible: bool = True):
        """keep[g] = sorted position of the row kept for group g, or -1 to drop the group (cpb_index_dedup_apply)"""
        k = np.ascontiguousarray(keep, dtype=np.int64)
        st = self.ctx.lib.cpb_index_dedup_apply(self.ctx.h, self.h, len(k), k.ctypes.data_as(C.POINTER(C.c_int64)), int(bug_compatible))
        if st:
            _raise(st, None, self.ctx)

    def ResolveDuplicates(self, resolve: Callable[[list[Row]], Row | None], bug_compatible: bool = True):
        """csvplus.go:651-653 / dedup :810-867.  `resolve` gets each group of rows with equal keys and returns the row to
        keep — one of the group's rows or any other row with the index's columns (:846 stores whatever comes back,
        without re-sorting) — or an empty row / None to drop the group, or raises (the index is left unchanged)."""
        lo, hi = self.dup_groups()
        n = len(lo)
        keep = np.full(max(1, n), -1, np.int64)
        replacements: list[Row] = []
        t = self.table()
        for g in range(n):
            rows = t.rows(int(lo[g]), int(hi[g]))
            chosen = resolve(rows)
            if not chosen or len(chosen) < len(self.columns):  # csvplus.go:845
                keep[g] = -1
                continue
            pick = next((i for i, r in enumerate(rows) if r is chosen), None)
            if pick is None:
                pick = next((i for i, r in enumerate(rows) if r == chosen), None)
            if pick is None:
                keep[g] = -2 - len(replacements)
                replacements.append(chosen)
            else:
                keep[g] = lo[g] + pick
        rt = Table.from_rows(self.ctx, replacements, columns=t.columns) if replacements else None
        e = _abi.Error()
        st = self.ctx.lib.cpb_index_dedup_apply2(self.ctx.h, self.h, n, keep.ctypes.data_as(C.POINTER(C.c_int64)),
                                                 rt.h if rt is not None else None, int(bug_compatible), C.byref(e))
        if st:
            _raise(st, e, self.ctx)


# ------------------------------------------------------------------ Reader (csvplus.go:922-1146)
class Reader:
    def __init__(self, source: Callable[[], bytes | np.ndarray | DeviceBuffer | HostBuffer], ctx: Context | None = None):
        self._source, self._ctx = source, ctx
        self.delimiter, self.comment = ",", ""
        self.numFields, self.lazyQuotes, self.trimLeadingSpace = 0, False, False
        self.header: dict | None = None
        self.headerFromFirstRow = True
        self._pred: Predicate | None = None

    @property
    def ctx(self) -> Context:
        return self._ctx or Context.default()

    def Delimiter(self, c): self.delimiter = c; return self
    def CommentChar(self, c): self.comment = c; return self
    def LazyQuotes(self): self.lazyQuotes = True; return self
    def TrimLeadingSpace(self): self.trimLeadingSpace = True; return self

    def AssumeHeader(self, spec: dict):  # csvplus.go:998-1012
        if len(spec) == 0:
            raise ValueError("Empty header spec")
        for name, col in spec.items():
            if col < 0:
                raise ValueError("header spec: negative index for column " + name)
        self.header, self.headerFromFirstRow = dict(spec), False
        return self

    def ExpectHeader(self, spec: dict):  # csvplus.go:1020-1033
        if len(spec) == 0:
            raise ValueError("empty header spec")
        self.header, self.headerFromFirstRow = dict(spec), True
        return self

    def SelectColumns(self, *names):  # csvplus.go:1039-1056
        if len(names) == 0:
            raise ValueError("empty header spec")
        h = {}
        for n in names:
            if n in h:
                raise ValueError("header spec: duplicate column name: " + n)
            h[n] = -1
        self.header, self.headerFromFirstRow = h, True
        return self

    def NumFields(self, n): self.numFields = n; return self
    def NumFieldsAuto(self): return self.NumFields(0)
    def NumFieldsAny(self): return self.NumFields(-1)

    def _parse(self, pred: Predicate | None = None):
        data = self._source()
        spec = list(self.header.items()) if self.header else []
        return parse_csv(self.ctx, data, delimiter=self.delimiter, comment=self.comment, num_fields=self.numFields,
                         lazy_quotes=self.lazyQuotes, trim_leading_space=self.trimLeadingSpace,
                         header_from_first_row=self.headerFromFirstRow, spec=spec, pred=pred)

    def Iterate(self, fn):  # csvplus.go:1080
        return Take(self)(fn)


def FromFile(name: str, ctx: Context | None = None) -> Reader:  # csvplus.go:950
    def src():
        try:
            return np.fromfile(name, dtype=np.uint8)
        except OSError as e:  # mapError's *os.PathError branch, csvplus.go:1216-1220
            raise DataSourceError(1, f"open: {e.strerror.lower() if e.strerror else e}") from None
    return Reader(src, ctx)


def FromReader(inp: io.IOBase, ctx: Context | None = None) -> Reader:  # csvplus.go:936
    return Reader(lambda: inp.read(), ctx)


FromReadCloser = FromReader


def FromBytes(data, ctx: Context | None = None) -> Reader:
    """bytes / numpy / HostBuffer / DeviceBuffer already in memory"""
    return Reader(lambda: data, ctx)


# ------------------------------------------------------------------ DataSource (csvplus.go:207-608)
class DataSource:
    """A lazily evaluated plan: source + operations.  Calling it with a RowFunc pulls the rows."""

    def __init__(self, source, ops: tuple = ()):
        self._src, self._ops = source, ops  # source: Reader | ("table", Table, line_base) | ("rows", list[Row])

    def _with(self, op) -> "DataSource":
        return DataSource(self._src, self._ops + (op,))

    # ---- combinators
    def Transform(self, trans): return self._with(("transform", trans))
    def Filter(self, pred): return self._with(("filter", pred))
    def Map(self, mf): return self._with(("map", mf))
    def Validate(self, vf): return self._with(("validate", vf))
    def Top(self, n: int): return self._with(("top", n))
    def Drop(self, n: int): return self._with(("drop", n))
    def TakeWhile(self, pred): return self._with(("takewhile", pred))
    def DropWhile(self, pred): return self._with(("dropwhile", pred))

    def DropColumns(self, *columns):
        if len(columns) == 0:
            raise ValueError("no columns specified in DropColumns()")
        return self._with(("dropcols", columns))

    def SelectColumns(self, *columns):
        if len(columns) == 0:
            raise ValueError("no columns specified in SelectColumns()")
        return self._with(("select", columns))

    def Join(self, index: Index, *columns):
        if len(columns) > len(index.columns):
            raise ValueError("too many source columns in Join()")
        return self._with(("join", index, columns))

    def Except(self, index: Index, *columns):
        if len(columns) > len(index.columns):
            raise ValueError("too many source columns in Except()")
        return self._with(("except", index, columns))

    # ---- evaluation
    def _ctx(self) -> Context:
        s = self._src
        if isinstance(s, Reader):
            return s.ctx
        if s[0] == "table":
            return s[1].ctx
        return Context.default()

    def _evaluate(self):
        """-> (state, pending error).  state: ("table", Table) or ("rows", list[Row])."""
        ops = list(self._ops)
        err = None
        s = self._src
        if isinstance(s, Reader):
            pred = None
            if ops and ops[0][0] == "filter" and isinstance(ops[0][1], Predicate) and ops[0][1].lowerable():
                pred = ops.pop(0)[1]  # fused into the parse kernel
            t, err = s._parse(pred)
            state = ("table", t)
        elif s[0] == "table":
            state = ("table", s[1])
        else:
            state = ("rows", [dict(r) for r in s[1]])  # iterate() clones, csvplus.go:230
        ctx = self._ctx()
        for op in ops:
            kind = op[0]
            device_ok = (kind in ("select", "dropcols", "join", "except", "top", "drop")
                         or (kind in ("filter", "takewhile", "dropwhile") and isinstance(op[1], Predicate) and op[1].lowerable()))
            if device_ok:
                if state[0] == "rows":
                    state = ("table", Table.from_rows(ctx, state[1]) if state[1] else None)
                t = state[1]
                if t is None:
                    continue
                try:
                    if kind == "select": t = t.select(*op[1])
                    elif kind == "dropcols": t = t.drop(*op[1])
                    elif kind == "filter": t = t.filter(op[1])
                    elif kind == "join": t = t.join(op[1], *op[2])
                    elif kind == "except": t = t.join(op[1], *op[2], anti=True)
                    elif kind == "top": t = t.slice(0, op[1])  # note: the reference pulls one extra row
                    elif kind == "drop": t = t.slice(op[1], len(t))
                    elif kind == "takewhile":  # csvplus.go:346-358: io.EOF at the first failing row — a later error is never met
                        cut = t.first_false(op[1])
                        if cut < len(t):
                            err = None
                        t = t.slice(0, cut)
                    elif kind == "dropwhile": t = t.slice(t.first_false(op[1]), len(t))  # :361-374
                except DataSourceError as e:
                    if err is None or kind in ("select", "join", "except"):
                        return ("rows", []), e
                    raise
                if kind == "top" and err is not None and len(state[1]) > op[1]:
                    err = None  # Top(n) pulls n+1 rows then stops (io.EOF): the failing record is never reached
                state = ("table", t)
            else:
                rows = state[1].rows() if state[0] == "table" and state[1] is not None else (state[1] or [])
                out, stop = [], False
                fn = op[1]
                dropping = True
                for r in rows:
                    if kind == "filter":
                        if fn(r): out.append(r)
                    elif kind == "map":
                        out.append(fn(r))
                    elif kind == "transform":
                        r2 = fn(r)
                        if r2: out.append(r2)
                    elif kind == "validate":
                        fn(r); out.append(r)
                    elif kind == "takewhile":
                        if not fn(r):
                            stop = True; err = None
                            break
                        out.append(r)
                    elif kind == "dropwhile":
                        dropping = dropping and fn(r)
                        if not dropping: out.append(r)
                state = ("rows", out)
        return state, err

    def _table(self) -> tuple:
        state, err = self._evaluate()
        if state[0] == "rows":
            t = Table.from_rows(self._ctx(), state[1]) if state[1] else None
            return t, err
        return state[1], err

    def __call__(self, fn: Callable[[Row], None]):
        state, err = self._evaluate()
        rows = state[1].rows() if state[0] == "table" and state[1] is not None else (state[1] or [])
        for r in rows:
            try:
                fn(r)
            except StopIterationEOF:
                return
        if err is not None:
            raise err

    # ---- sinks
    def ToRows(self) -> list[Row]:  # csvplus.go:483-490
        out = []
        self(out.append)
        return out

    def ToCsv(self, out: io.IOBase, *columns):  # csvplus.go:379-406
        if len(columns) == 0:
            raise ValueError("empty column list in ToCsv() function")
        t, err = self._table()
        if t is None:
            out.write((",".join(columns) + "\n").encode())
        else:
            out.write(t.to_csv(*columns))
        if err is not None:
            raise err

    def ToCsvFile(self, name: str, *columns):  # csvplus.go:411-443: the file is removed on error
        try:
            with open(name, "wb") as f:
                self.ToCsv(f, *columns)
        except BaseException:
            if os.path.exists(name):
                os.remove(name)
            raise

    def ToJSON(self, out: io.IOBase):  # csvplus.go:446-474
        """all rows as a JSON array, formatted the way the reference's json.Encoder (SetEscapeHTML(false), no indent) writes
        it: `[`, then per row the object with its keys sorted and a newline, a `,` before every row but the first, `]`."""
        parts = [b"["]
        first = [True]

        def fn(row):
            if not first[0]:
                parts.append(b",")
            first[0] = False
            parts.append(b"{" + b",".join(_go_json_string(k) + b":" + _go_json_string(row[k]) for k in sorted(row, key=_enc)) + b"}\n")
        self(fn)
        parts.append(b"]")
        out.write(b"".join(parts))

    def ToJSONFile(self, name: str):  # csvplus.go:477-479
        try:
            with open(name, "wb") as f:
                self.ToJSON(f)
        except BaseException:
            if os.path.exists(name):
                os.remove(name)
            raise

    def IndexOn(self, *columns) -> Index:  # csvplus.go:529-531
        return self._index(columns, False)

    def UniqueIndexOn(self, *columns) -> Index:  # csvplus.go:535-537
        return self._index(columns, True)

    def _index(self, columns, unique) -> Index:
        if len(columns) == 0:
            raise ValueError("empty column list in CreateIndex()")
        if len(set(columns)) != len(columns):
            raise ValueError("duplicate column name(s) in CreateIndex()")
        t, err = self._table()
        if err is not None:
            raise err
        if t is None:
            t = Table.from_rows(self._ctx(), [], list(columns))
        return t.index_on(*columns, unique=unique)


def Take(src) -> DataSource:  # csvplus.go:252-256
    if isinstance(src, Reader):
        return DataSource(src)
    if isinstance(src, Index):
        return TakeTable(src.table(), line_base=0)
    raise TypeError("Take() needs a Reader or an Index")


def TakeRows(rows: list[Row]) -> DataSource:  # csvplus.go:218-222
    return DataSource(("rows", rows))


def TakeTable(t: Table, line_base: int = 0) -> DataSource:
    return DataSource(("table", t, line_base))
