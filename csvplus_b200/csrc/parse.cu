// parse.cu — K1..K4 of SURVEY §2: single-pass CSV scan -> column offsets -> per-column gather into
// string columns, fused with the Like/All/Any/Not filter and row compaction.
//
// Replaces (reference): Reader.Iterate csvplus.go:1080-1146 (hot loop :1117-1138), the stdlib
// encoding/csv.Reader it drives (readRecord/readLine), Row construction
// csvplus.go:1118-1122, Filter csvplus.go:276-286 and Like csvplus.go:1279-1293.
//
// Design: the input is read from HBM exactly once.  A persistent grid walks
// 32 KiB tiles in ticket order; each tile (+16 B look-behind, +2 KiB look-ahead) is staged into
// shared memory with one bulk-async copy (TMA 1-D, cp.async.bulk + mbarrier).  Bytes are classified
// with SWAR compares into newline / delimiter bitmaps; quote parity and the running
// (records, rows, bytes-per-column) totals are carried across tiles by two decoupled look-back
// chains.  Every thread owns the records that *start* in its 128-byte slice: pass 1 counts,
// a block scan + look-back turns counts into global output positions, pass 2 writes offsets and
// gathers field bytes.  Records containing quotes (or running past the staged window) take an exact
// sequential state machine that restates Go's readRecord byte for byte.
#include <algorithm>

#include "core.hpp"
#include "pred.cuh"
#include "util.cuh"

namespace cpb {

constexpr int TILE = 32768;
constexpr int PRE = 16;
constexpr int HALO = 2048;
constexpr int WIN = TILE + HALO;
constexpr int THREADS = 256;
constexpr int WIN_WORDS = WIN / 32;    // 1088
constexpr int TILE_WORDS = TILE / 32;  // 1024
constexpr int WPT = TILE_WORDS / THREADS;  // bitmap words per thread (4)
constexpr int MAXSEL = CPB_MAX_PARSE_COLS;
constexpr int HDR_MAX_FIELDS = 4096;
constexpr int HDR_MAX_BYTES = 1 << 16;

enum { K_OK = 0, K_BARE = CPB_E_BARE_QUOTE, K_QUOTE = CPB_E_QUOTE, K_FIELDS = CPB_E_FIELD_COUNT, K_COLIDX = CPB_E_COLUMN_INDEX };

struct ParseResult {  // device -> host
    uint64_t totals[2 + MAXSEL];  // records, rows, bytes per slot
    unsigned long long err_key;   // (record ordinal << 16) | (kind << 8) | slot ; ~0 = none
    unsigned long long err_rows;  // rows delivered before the failing record
};

struct ParseParams {
    const uint8_t* in;
    uint64_t n;
    uint64_t data_start;  // records starting before this byte are not data (header row)
    uint32_t ntiles;
    uint32_t delim;
    int32_t expect_fields;  // >0: every record must have exactly this many fields
    int32_t pad_missing;    // numFields < 0: short records pad selected columns with ""
    int32_t nsel;
    int32_t sel_field[MAXSEL];  // ascending, distinct
    uint32_t* out_off[MAXSEL];
    uint8_t* out_data[MAXSEL];
    uint64_t data_cap[MAXSEL];
    uint64_t row_cap;
    const uint8_t* lits;
    PredProg pred;
    uint32_t slot_terms[MAXSEL];  // per slot: mask of Like terms comparing that slot
    // look-back state
    uint32_t* st1;   // [ntiles] quote-parity chain: bits1:0 status, bit2 value
    uint32_t* st2;   // [ntiles] totals chain status
    uint64_t* agg;   // [ntiles][2+nsel]
    uint64_t* inc;   // [ntiles][2+nsel]
    uint32_t* ticket;
    ParseResult* result;
};

// ------------------------------------------------------------------ byte source: staged window or HBM
struct ByteSrc {
    const uint8_t* g; uint64_t n;
    const uint8_t* s; uint64_t s_lo, s_hi;  // smem copy of absolute [s_lo, s_hi)
    __device__ __forceinline__ int get(uint64_t pos) const {
        if (pos >= n) return -1;
        if (pos >= s_lo && pos < s_hi) return s[pos - s_lo];
        return g[pos];
    }
};

struct SeqResult { int err; int nfields; uint64_t next; };

// Exact sequential restatement of encoding/csv readRecord (+readLine's \r\n and trailing-\r rules)
// for one record starting at `start` (a non-empty line start), default options, single-byte comma.
// Sink: begin_field(f) / put(byte) / end_field().
template <class Sink>
__device__ SeqResult seq_parse_record(const ByteSrc& src, uint64_t start, int delim, Sink& sink) {
    uint64_t pos = start;
    int f = 0;
    for (;;) {  // parseField
        sink.begin_field(f);
        int c = src.get(pos);
        if (c != '"') {
            // non-quoted field: up to the next comma or end of line
            uint64_t fb = pos;
            for (;;) {
                c = src.get(pos);
                if (c == delim) { sink.end_field(); pos++; f++; break; }
                if (c == '\n' || c < 0) {
                    // line[n-2]=='\r' normalisation / trailing \r before EOF: the sink has already seen the \r;
                    // it is retracted here (only one, only if inside this field)
                    if (pos > fb && src.get(pos - 1) == '\r') sink.unput();
                    sink.end_field();
                    return {K_OK, f + 1, c < 0 ? src.n : pos + 1};
                }
                if (c == '"') return {K_BARE, f + 1, pos};
                sink.put(c);
                pos++;
            }
        } else {
            pos++;  // opening quote
            for (;;) {
                c = src.get(pos);
                if (c < 0) return {K_QUOTE, f + 1, pos};  // EOF inside quotes (non-lazy)
                if (c == '"') {
                    int c2 = src.get(pos + 1);
                    if (c2 == '"') { sink.put('"'); pos += 2; continue; }
                    if (c2 == delim) { sink.end_field(); pos += 2; f++; break; }
                    if (c2 == '\n') { sink.end_field(); return {K_OK, f + 1, pos + 2}; }
                    if (c2 < 0) { sink.end_field(); return {K_OK, f + 1, src.n}; }
                    if (c2 == '\r') {
                        int c3 = src.get(pos + 2);
                        if (c3 == '\n') { sink.end_field(); return {K_OK, f + 1, pos + 3}; }
                        if (c3 < 0) { sink.end_field(); return {K_OK, f + 1, src.n}; }
                    }
                    return {K_QUOTE, f + 1, pos};
                }
                if (c == '\r') {
                    int c2 = src.get(pos + 1);
                    if (c2 == '\n') { sink.put('\n'); pos += 2; continue; }  // \r\n -> \n on every physical line
                    if (c2 < 0) return {K_QUOTE, f + 1, pos};               // trailing \r dropped, then EOF in quotes
                }
                sink.put(c);
                pos++;
            }
        }
    }
}

// ------------------------------------------------------------------ header kernel
struct HeaderOut {
    int32_t err;        // K_* of the first record (0 ok)
    int32_t nfields;
    int32_t eof;        // 1: no record at all
    int32_t truncated;  // names did not fit
    uint64_t rec_start, data_start;
    uint64_t sample_bytes, sample_newlines;
    uint32_t field_len[HDR_MAX_FIELDS];
    uint8_t bytes[HDR_MAX_BYTES];
};
struct HeaderSink {
    HeaderOut* o; int f = 0; uint32_t len = 0; uint32_t used = 0;
    __device__ void begin_field(int fi) { f = fi; len = 0; }
    __device__ void put(int c) { if (used < HDR_MAX_BYTES) o->bytes[used] = (uint8_t)c; else o->truncated = 1; used++; len++; }
    __device__ void unput() { used--; len--; }
    __device__ void end_field() { if (f < HDR_MAX_FIELDS) o->field_len[f] = len; else o->truncated = 1; }
};

// Parses the first record (makeHeader's reader.Read(), csvplus.go:1150) and samples newline density
// in three 64 KiB windows for the row-capacity estimate.
__global__ void csv_header_kernel(const uint8_t* in, uint64_t n, int delim, HeaderOut* out) {
    __shared__ unsigned long long s_nl;
    if (threadIdx.x == 0) {
        s_nl = 0;
        ByteSrc src{in, n, nullptr, 0, 0};
        uint64_t pos = 0;
        // skip empty lines: "\n", "\r\n", and a lone trailing "\r" before EOF
        for (;;) {
            int c = src.get(pos);
            if (c == '\n') { pos++; continue; }
            if (c == '\r') {
                int c2 = src.get(pos + 1);
                if (c2 == '\n') { pos += 2; continue; }
                if (c2 < 0) { pos += 1; continue; }
            }
            break;
        }
        out->truncated = 0;
        out->rec_start = pos;
        if (pos >= n) { out->eof = 1; out->err = 0; out->nfields = 0; out->data_start = n; }
        else {
            HeaderSink sink{out};
            SeqResult r = seq_parse_record(src, pos, delim, sink);
            out->eof = 0; out->err = r.err; out->nfields = r.nfields; out->data_start = r.next;
        }
    }
    __syncthreads();
    const uint64_t S = 65536;
    unsigned long long cnt = 0, tot = 0;
    for (int w = 0; w < 3; w++) {
        uint64_t lo = w == 0 ? 0 : (w == 1 ? (n / 2) : (n > S ? n - S : 0));
        uint64_t hi = lo + S < n ? lo + S : n;
        for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) cnt += in[i] == '\n';
        tot += hi - lo;
    }
    atomicAdd(&s_nl, cnt);
    __syncthreads();
    if (threadIdx.x == 0) { out->sample_bytes = tot; out->sample_newlines = s_nl; }
}

// ------------------------------------------------------------------ main kernel
struct __align__(16) ParseSmem {
    uint8_t data[PRE + WIN + 16];
    uint32_t Tb[WIN_WORDS + 4];  // record terminators: '\n' outside quotes
    uint32_t Db[WIN_WORDS + 4];  // delimiter bytes
    uint32_t Qb[WIN_WORDS + 4];  // quote bytes (exact; only built for tiles that contain quotes)
    uint64_t mbar;
    uint64_t tile_prefix[2 + MAXSEL];
    uint32_t wtot[1 + MAXSEL][THREADS / 32];
    uint32_t wpar[THREADS / 32];
    uint32_t ticket;
    uint32_t pin;
};

__device__ __forceinline__ int next_set(const uint32_t* bm, int from, int lim) {
    int wi = from >> 5;
    uint32_t m = bm[wi] & (0xffffffffu << (from & 31));
    for (;;) {
        if (m) { int i = (wi << 5) + __ffs(m) - 1; return i < lim ? i : lim; }
        wi++;
        if ((wi << 5) >= lim) return lim;
        m = bm[wi];
    }
}
__device__ __forceinline__ int count_bits(const uint32_t* bm, int a, int b) {  // bits set in [a,b)
    if (a >= b) return 0;
    int wa = a >> 5, wb = b >> 5;
    uint32_t ma = 0xffffffffu << (a & 31);
    uint32_t mb = (b & 31) ? (0xffffffffu >> (32 - (b & 31))) : 0u;
    if (wa == wb) return __popc(bm[wa] & ma & mb);
    int c = __popc(bm[wa] & ma);
    for (int w = wa + 1; w < wb; w++) c += __popc(bm[w]);
    if (mb) c += __popc(bm[wb] & mb);
    return c;
}
__device__ __forceinline__ bool bytes_eq(const uint8_t* a, const uint8_t* lit, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) if (a[i] != __ldg(lit + i)) return false;
    return true;
}
// Sink of the slow path inside the main kernel: tracks the selected slots of one record.
struct SlowSink {
    const ParseParams& P;
    bool emit;
    int next_slot = 0, cur = -1;
    uint32_t len = 0, present = 0, alive = 0, eq = 0;
    uint32_t ulen[MAXSEL];
    uint8_t* dst[MAXSEL];
    uint32_t maxlen[MAXSEL];  // emit: value lengths known from pass 1 (a retracted '\r' must never be stored)
    __device__ SlowSink(const ParseParams& p, bool e) : P(p), emit(e) {}
    __device__ void begin_field(int f) {
        cur = -1;
        if (next_slot < P.nsel && P.sel_field[next_slot] == f) { cur = next_slot++; len = 0; alive = P.slot_terms[cur]; }
    }
    __device__ void put(int c) {
        if (cur < 0) return;
        if (emit) { if (dst[cur] && len < maxlen[cur]) dst[cur][len] = (uint8_t)c; }
        else {
            uint32_t m = alive;
            while (m) {
                int t = __ffs(m) - 1; m &= m - 1;
                if (len >= P.pred.term_len[t] || __ldg(P.lits + P.pred.term_off[t] + len) != (uint8_t)c) alive &= ~(1u << t);
            }
        }
        len++;
    }
    __device__ void unput() { if (cur >= 0) len--; }
    __device__ void end_field() {
        if (cur < 0) return;
        ulen[cur] = len; present |= 1u << cur;
        uint32_t m = alive;
        while (m) { int t = __ffs(m) - 1; m &= m - 1; if (len == P.pred.term_len[t]) eq |= 1u << t; }
        // note: a retracted '\r' (unput) can only shorten the value; `alive` was computed on a prefix, still exact
    }
};

template <int KMAX>
struct Rec {
    uint32_t f[KMAX];  // fast: beg | len << 16 (window-relative); slow: unescaped length
    uint32_t present, eq;
    int nf, err, err_slot;
    bool slow;
};

struct SlowOut { uint32_t ulen[MAXSEL]; uint32_t present, eq; int nf, err; };

// count mode (emit=false): lengths / predicate terms / error of one record; emit mode: store the unescaped values.
__device__ __noinline__ void slow_record(const ParseParams& P, const ByteSrc& src, uint64_t start, bool emit,
                                         const uint64_t* dst_off, const uint32_t* maxlen, SlowOut* o) {
    SlowSink sink(P, emit);
    if (emit) {
        for (int k = 0; k < P.nsel; k++) {
            bool fits = dst_off[k] + maxlen[k] <= P.data_cap[k];
            sink.dst[k] = fits ? P.out_data[k] + dst_off[k] : nullptr;
            sink.maxlen[k] = maxlen[k];
        }
    }
    SeqResult s = seq_parse_record(src, start, (int)P.delim, sink);
    o->err = s.err; o->nf = s.nfields; o->present = sink.present; o->eq = sink.eq;
    for (int k = 0; k < P.nsel; k++) o->ulen[k] = ((sink.present >> k) & 1) ? sink.ulen[k] : 0;
}

// Extracts the record starting at window offset ws.  Returns false for an empty line (no record).
template <int KMAX>
__device__ __forceinline__ bool scan_record(const ParseParams& P, const ParseSmem& sm, const ByteSrc& src, uint64_t tile_base,
                                            int ws, int lim, bool eof_in_win, bool tile_has_q, Rec<KMAX>& r) {
    int e_nl = next_set(sm.Tb, ws, lim);
    r.present = 0; r.eq = 0; r.err = K_OK; r.err_slot = 0; r.slow = false;
    bool to_slow = false;
    if (e_nl >= lim && !eof_in_win) to_slow = true;  // runs past the staged window
    int e = e_nl;
    if (!to_slow) {
        if (e > ws && sm.data[PRE + e - 1] == '\r') e--;  // \r\n -> \n ; trailing \r before EOF
        if (e == ws) return false;                          // empty line: not a record
        if (tile_has_q && count_bits(sm.Qb, ws, e_nl) != 0) to_slow = true;
    }
    if (to_slow) {
        SlowOut so;
        slow_record(P, src, tile_base + ws, false, nullptr, nullptr, &so);
        r.err = so.err; r.nf = so.nf; r.present = so.present; r.eq = so.eq; r.slow = true;
#pragma unroll
        for (int k = 0; k < KMAX; k++) r.f[k] = k < P.nsel ? so.ulen[k] : 0;
    } else {
        int fb = ws, f = 0;
        bool ended = false;
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
            r.f[k] = 0;
            if (k < P.nsel) {
                int target = P.sel_field[k];
                while (!ended && f < target) {
                    int d = next_set(sm.Db, fb, e);
                    if (d >= e) ended = true; else { fb = d + 1; f++; }
                }
                if (!ended) {
                    int d = next_set(sm.Db, fb, e);
                    int fe = d < e ? d : e;
                    r.f[k] = (uint32_t)fb | ((uint32_t)(fe - fb) << 16);
                    r.present |= 1u << k;
                    uint32_t tm = P.slot_terms[k];
                    while (tm) {
                        int t = __ffs(tm) - 1; tm &= tm - 1;
                        if ((uint32_t)(fe - fb) == P.pred.term_len[t] &&
                            bytes_eq(sm.data + PRE + fb, P.lits + P.pred.term_off[t], fe - fb)) r.eq |= 1u << t;
                    }
                    if (d < e) { fb = d + 1; f++; } else ended = true;
                }
            }
        }
        r.nf = 1 + count_bits(sm.Db, ws, e);
    }
    // record-level checks in the reference's order: parse error (already set) > field count > missing column
    if (r.err == K_OK) {
        if (P.expect_fields > 0 && r.nf != P.expect_fields) r.err = K_FIELDS;
        else {
            uint32_t want = P.nsel >= 32 ? 0xffffffffu : ((1u << P.nsel) - 1);
            uint32_t missing = want & ~r.present;
            if (missing) {
                if (P.pad_missing) {
                    // padded "" values still take part in Like comparisons against empty literals
                    uint32_t m = missing;
                    while (m) {
                        int k = __ffs(m) - 1; m &= m - 1;
                        uint32_t tm = P.slot_terms[k];
                        while (tm) { int t = __ffs(tm) - 1; tm &= tm - 1; if (P.pred.term_len[t] == 0) r.eq |= 1u << t; }
                    }
                } else { r.err = K_COLIDX; r.err_slot = __ffs(missing) - 1; }
            }
        }
    }
    return true;
}

template <int KMAX>
__global__ void __launch_bounds__(THREADS) csv_scan_kernel(const __grid_constant__ ParseParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ParseSmem& sm = *reinterpret_cast<ParseSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NP = 2 + P.nsel;
    const uint32_t NL4 = 0x0a0a0a0au, Q4 = 0x22222222u, D4 = P.delim * 0x01010101u;

    if (tid == 0) { mbar_init(&sm.mbar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t phase = 0;

    for (;;) {
        if (tid == 0) sm.ticket = atomicAdd(P.ticket, 1u);
        __syncthreads();
        const uint32_t tile = sm.ticket;
        if (tile >= P.ntiles) break;
        const uint64_t tile_base = (uint64_t)tile * TILE;
        // ---- stage the window [tile_base-PRE, tile_base+WIN) with one bulk copy
        const uint64_t w_lo = tile_base >= PRE ? tile_base - PRE : 0;
        const uint64_t n16 = (P.n + 15) & ~15ull;
        const uint64_t w_hi = tile_base + WIN < n16 ? tile_base + WIN : n16;
        const uint32_t lead = (uint32_t)(PRE - (tile_base - w_lo));  // 0, or PRE for tile 0
        const uint32_t nbytes = (uint32_t)(w_hi - w_lo);
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(&sm.mbar, nbytes);
            bulk_g2s(sm.data + lead, P.in + w_lo, nbytes, &sm.mbar);
        }
        mbar_wait(&sm.mbar, phase);
        phase ^= 1;
        // bytes at absolute positions >= n are zeroed so that they classify as nothing
        const int64_t rel_n = (int64_t)(P.n - tile_base);  // > 0
        if (rel_n < WIN) {
            for (int i = (int)rel_n + tid; i < WIN + 16; i += THREADS) sm.data[PRE + i] = 0;
            __syncthreads();
        }
        const int lim = rel_n < WIN ? (int)rel_n : WIN;
        const bool eof_in_win = rel_n <= WIN;

        // ---- classify: newline / delimiter bitmaps, quote presence
        const uint4* d4 = reinterpret_cast<const uint4*>(sm.data + PRE);
        uint16_t* T16 = reinterpret_cast<uint16_t*>(sm.Tb);
        uint16_t* D16 = reinterpret_cast<uint16_t*>(sm.Db);
        uint32_t anyq = 0;
        for (int v = tid; v < WIN / 16; v += THREADS) {
            uint4 x = d4[v];
            T16[v] = (uint16_t)flags16(eq_flags(x.x, NL4), eq_flags(x.y, NL4), eq_flags(x.z, NL4), eq_flags(x.w, NL4));
            D16[v] = (uint16_t)flags16(eq_flags(x.x, D4), eq_flags(x.y, D4), eq_flags(x.z, D4), eq_flags(x.w, D4));
            anyq |= eq_any(x.x, Q4) | eq_any(x.y, Q4) | eq_any(x.z, Q4) | eq_any(x.w, Q4);
        }
        if (tid < 4) { sm.Tb[WIN_WORDS + tid] = 0; sm.Db[WIN_WORDS + tid] = 0; sm.Qb[WIN_WORDS + tid] = 0; }
        const bool hasq = __syncthreads_or(anyq != 0);
        uint32_t tile_par = 0;
        if (hasq) {
            uint16_t* Q16 = reinterpret_cast<uint16_t*>(sm.Qb);
            for (int v = tid; v < WIN / 16; v += THREADS) {
                uint4 x = d4[v];
                Q16[v] = (uint16_t)flags16(eq_flags(x.x, Q4), eq_flags(x.y, Q4), eq_flags(x.z, Q4), eq_flags(x.w, Q4));
            }
            __syncthreads();
            uint32_t par = 0;
            for (int w = tid; w < TILE_WORDS; w += THREADS) par ^= __popc(sm.Qb[w]);
            tile_par = __syncthreads_count(par & 1) & 1;
        }
        // ---- chain 1: quote parity at the tile start
        if (warp == 0) {
            uint32_t pin = 0;
            if (tile == 0) { if (lane == 0) st_release_u32(&P.st1[0], 2u | (tile_par << 2)); }
            else {
                if (lane == 0) st_release_u32(&P.st1[tile], 1u | (tile_par << 2));
                int64_t base = (int64_t)tile - 1;
                for (;;) {
                    int64_t p = base - lane;
                    uint32_t s = 2u;
                    if (p >= 0) { do { s = ld_acquire_u32(&P.st1[p]); } while ((s & 3u) == 0); }
                    uint32_t incl = __ballot_sync(0xffffffffu, (s & 3u) == 2u);
                    uint32_t vals = __ballot_sync(0xffffffffu, (s >> 2) & 1u);
                    int f = __ffs(incl) - 1;
                    uint32_t mask = f >= 0 ? (0xffffffffu >> (31 - f)) : 0xffffffffu;
                    pin ^= __popc(vals & mask) & 1;
                    if (f >= 0) break;
                    base -= 32;
                }
                if (lane == 0) st_release_u32(&P.st1[tile], 2u | ((pin ^ tile_par) << 2));
            }
            if (lane == 0) sm.pin = pin;
        }
        __syncthreads();
        const uint32_t pin = sm.pin;
        const bool tile_has_q = hasq;
        if (hasq || pin) {
            // in-quote mask by prefix-XOR of the quote bitmap; terminators are newlines outside quotes
            uint32_t carry = pin;
            for (int r0 = 0; r0 < WIN_WORDS; r0 += THREADS) {
                int w = r0 + tid;
                uint32_t q = (w < WIN_WORDS && hasq) ? sm.Qb[w] : 0;
                uint32_t px = q; px ^= px << 1; px ^= px << 2; px ^= px << 4; px ^= px << 8; px ^= px << 16;
                uint32_t b = __ballot_sync(0xffffffffu, px >> 31);
                uint32_t before = __popc(b & lanemask_lt()) & 1;
                if (lane == 0) sm.wpar[warp] = __popc(b) & 1;
                __syncthreads();
                uint32_t c = carry, tot = 0;
                for (int i = 0; i < THREADS / 32; i++) { if (i < warp) c ^= sm.wpar[i]; tot ^= sm.wpar[i]; }
                uint32_t cin = c ^ before;
                uint32_t iq = (px ^ q) ^ (0u - cin);
                if (w < WIN_WORDS) sm.Tb[w] &= ~iq;
                carry ^= tot;
                __syncthreads();
            }
        }

        // ---- record-start bits of this thread's 4 words
        uint32_t rs[WPT];
        {
            const uint4 tw = reinterpret_cast<const uint4*>(sm.Tb)[tid];
            uint32_t prev;
            if (tid == 0) prev = (tile_base > 0 && sm.data[PRE - 1] == '\n' && pin == 0) ? 0x80000000u : 0u;
            else prev = sm.Tb[tid * WPT - 1];
            rs[0] = (tw.x << 1) | (prev >> 31);
            rs[1] = (tw.y << 1) | (tw.x >> 31);
            rs[2] = (tw.z << 1) | (tw.y >> 31);
            rs[3] = (tw.w << 1) | (tw.z >> 31);
            // only positions in [data_start, n); data_start itself always starts a record
            const int64_t rel_ds = (int64_t)P.data_start - (int64_t)tile_base;
#pragma unroll
            for (int j = 0; j < WPT; j++) {
                const int64_t b0 = (int64_t)(tid * WPT + j) * 32;
                uint32_t keep = 0xffffffffu;
                if (rel_ds > b0) keep = rel_ds >= b0 + 32 ? 0u : (0xffffffffu << (rel_ds - b0));
                if (rel_n < b0 + 32) keep &= rel_n <= b0 ? 0u : (0xffffffffu >> (32 - (rel_n - b0)));
                rs[j] &= keep;
                if (rel_ds >= b0 && rel_ds < b0 + 32 && rel_ds < rel_n) rs[j] |= 1u << (rel_ds - b0);
            }
        }
        ByteSrc src{P.in, P.n, sm.data + PRE, tile_base, tile_base + (uint64_t)lim};

        // ---- pass 1: count records / surviving rows / bytes per column
        uint32_t nrec = 0, nrow = 0, cb[KMAX];
        uint32_t my_err = 0xffffffffu;  // (local record idx << 16) | kind << 8 | slot
        uint32_t err_rows_local = 0;
#pragma unroll
        for (int k = 0; k < KMAX; k++) cb[k] = 0;
#pragma unroll
        for (int j = 0; j < WPT; j++) {
            uint32_t m = rs[j];
            while (m) {
                int b = __ffs(m) - 1; m &= m - 1;
                int ws = (tid * WPT + j) * 32 + b;
                Rec<KMAX> r;
                if (!scan_record<KMAX>(P, sm, src, tile_base, ws, lim, eof_in_win, tile_has_q, r)) continue;
                if (r.err != K_OK) {
                    if (my_err == 0xffffffffu) { my_err = (nrec << 16) | ((uint32_t)r.err << 8) | (uint32_t)r.err_slot; err_rows_local = nrow; }
                } else if (eval_pred(P.pred, r.eq)) {
                    nrow++;
#pragma unroll
                    for (int k = 0; k < KMAX; k++) if (k < P.nsel) cb[k] += r.slow ? r.f[k] : (r.f[k] >> 16);
                }
                nrec++;
            }
        }
        // ---- block scan of (records | rows << 16, bytes[k])
        uint32_t v0 = nrec | (nrow << 16);
        uint32_t i0 = warp_incl_scan(v0);
        uint32_t ik[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; k++) if (k < P.nsel) ik[k] = warp_incl_scan(cb[k]);
        if (lane == 31) {
            sm.wtot[0][warp] = i0;
#pragma unroll
            for (int k = 0; k < KMAX; k++) if (k < P.nsel) sm.wtot[1 + k][warp] = ik[k];
        }
        __syncthreads();
        uint32_t ex0 = i0 - v0, tot0 = 0;
        uint32_t exk[KMAX], totk[KMAX];
#pragma unroll
        for (int i = 0; i < THREADS / 32; i++) { uint32_t t = sm.wtot[0][i]; if (i < warp) ex0 += t; tot0 += t; }
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
            exk[k] = 0; totk[k] = 0;
            if (k < P.nsel) {
                exk[k] = ik[k] - cb[k];
#pragma unroll
                for (int i = 0; i < THREADS / 32; i++) { uint32_t t = sm.wtot[1 + k][i]; if (i < warp) exk[k] += t; totk[k] += t; }
            }
        }
        // ---- chain 2: global prefix of (records, rows, bytes[k])
        if (warp == 0) {
            uint64_t mine = 0;
            if (lane == 0) mine = tot0 & 0xffffu;
            else if (lane == 1) mine = tot0 >> 16;
#pragma unroll
            for (int k = 0; k < KMAX; k++) if (lane == 2 + k) mine = totk[k];
            uint64_t excl = 0;
            if (tile == 0) {
                if (lane < NP) st_relaxed_u64(&P.inc[lane], mine);
                __syncwarp();
                if (lane == 0) { __threadfence(); st_release_u32(&P.st2[0], 2u); }
            } else {
                if (lane < NP) st_relaxed_u64(&P.agg[(uint64_t)tile * NP + lane], mine);
                __syncwarp();
                if (lane == 0) { __threadfence(); st_release_u32(&P.st2[tile], 1u); }
                int64_t base = (int64_t)tile - 1;
                for (;;) {
                    int64_t p = base - lane;
                    uint32_t s = 3u;
                    if (p >= 0) { do { s = ld_acquire_u32(&P.st2[p]); } while (s == 0); }
                    uint32_t incl = __ballot_sync(0xffffffffu, s >= 2u);
                    int f = __ffs(incl) - 1;
                    bool take = (f < 0 || lane <= f) && p >= 0;
                    for (int c = 0; c < NP; c++) {
                        uint64_t v = 0;
                        if (take) v = ld_relaxed_u64(s == 2u ? &P.inc[(uint64_t)p * NP + c] : &P.agg[(uint64_t)p * NP + c]);
                        v = warp_sum_u64(v);
                        if (lane == c) excl += v;
                    }
                    if (f >= 0) break;
                    base -= 32;
                }
                if (lane < NP) st_relaxed_u64(&P.inc[(uint64_t)tile * NP + lane], excl + mine);
                __syncwarp();
                if (lane == 0) { __threadfence(); st_release_u32(&P.st2[tile], 2u); }
            }
            if (lane < NP) sm.tile_prefix[lane] = excl;
            if (tile == P.ntiles - 1) {  // totals + end-of-column sentinels
                uint64_t total = excl + mine;
                if (lane < NP) P.result->totals[lane] = total;
                uint64_t rows = __shfl_sync(0xffffffffu, total, 1);
                if (lane >= 2 && lane < NP && rows <= P.row_cap) P.out_off[lane - 2][rows] = (uint32_t)total;
            }
        }
        __syncthreads();

        // ---- pass 2: write offsets, gather field bytes
        {
            const uint64_t rec0 = sm.tile_prefix[0] + (ex0 & 0xffffu);
            uint64_t row = sm.tile_prefix[1] + (ex0 >> 16);
            if (my_err != 0xffffffffu) {
                unsigned long long key = ((rec0 + (my_err >> 16)) << 16) | (my_err & 0xffffu);
                atomicMin(&P.result->err_key, key);
                atomicMin(&P.result->err_rows, (unsigned long long)(row + err_rows_local));
            }
            if (nrow != 0) {
                uint64_t off[KMAX];
#pragma unroll
                for (int k = 0; k < KMAX; k++) off[k] = k < P.nsel ? sm.tile_prefix[2 + k] + exk[k] : 0;
#pragma unroll
                for (int j = 0; j < WPT; j++) {
                    uint32_t m = rs[j];
                    while (m) {
                        int b = __ffs(m) - 1; m &= m - 1;
                        int ws = (tid * WPT + j) * 32 + b;
                        Rec<KMAX> r;
                        if (!scan_record<KMAX>(P, sm, src, tile_base, ws, lim, eof_in_win, tile_has_q, r)) continue;
                        if (r.err != K_OK || !eval_pred(P.pred, r.eq)) continue;
                        const bool row_ok = row < P.row_cap;
                        if (r.slow) {
                            uint64_t dst_off[MAXSEL];
                            uint32_t maxlen[MAXSEL];
                            SlowOut so;
#pragma unroll
                            for (int k = 0; k < KMAX; k++) { dst_off[k] = off[k]; maxlen[k] = r.f[k]; }
                            slow_record(P, src, tile_base + ws, true, dst_off, maxlen, &so);
                        }
#pragma unroll
                        for (int k = 0; k < KMAX; k++) {
                            if (k < P.nsel) {
                                uint32_t len = r.slow ? r.f[k] : (r.f[k] >> 16);
                                if (row_ok) P.out_off[k][row] = (uint32_t)off[k];
                                if (!r.slow && off[k] + len <= P.data_cap[k]) {
                                    const uint8_t* s = sm.data + PRE + (r.f[k] & 0xffffu);
                                    uint8_t* d = P.out_data[k] + off[k];
                                    for (uint32_t i = 0; i < len; i++) d[i] = s[i];
                                }
                                off[k] += len;
                            }
                        }
                        row++;
                    }
                }
            }
        }
        __syncthreads();  // smem is reused by the next tile
    }
}

// ------------------------------------------------------------------ host driver
namespace {

const char* kind_text(int k) {
    switch (k) {
        case CPB_E_BARE_QUOTE: return "bare \" in non-quoted-field";
        case CPB_E_QUOTE: return "extraneous or missing \" in quoted-field";
        case CPB_E_FIELD_COUNT: return "wrong number of fields";
        default: return "";
    }
}

template <int KMAX>
void launch_scan(Ctx* c, const ParseParams& P, uint64_t algo_bytes) {
    static bool configured = false;
    const size_t smem = sizeof(ParseSmem);
    if (!configured) {
        CPB_CUDA(cudaFuncSetAttribute(csv_scan_kernel<KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int occ = 0;
    CPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, csv_scan_kernel<KMAX>, THREADS, smem));
    if (occ < 1) occ = 1;
    uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)c->sm_count * occ, P.ntiles);
    KernelTimer kt(c, "csv_scan", algo_bytes);
    csv_scan_kernel<KMAX><<<grid, THREADS, smem, c->stream>>>(P);
    CPB_CUDA(cudaGetLastError());
}

}  // namespace

bool valid_delim(uint32_t r) {
    return r != 0 && r != '"' && r != '\r' && r != '\n' && r != 0xFFFD && r <= 0x10FFFF && !(r >= 0xD800 && r <= 0xDFFF);
}

std::shared_ptr<Table> parse_csv(Ctx* c, const uint8_t* in, uint64_t n, const cpb_reader_opts& o,
                                 const std::vector<std::pair<std::string, int>>& spec, const cpb_pred* filter,
                                 bool* had_error, DataError* derr) {
    *had_error = false;
    auto fail = [&](int kind, int col, uint64_t line, const std::string& msg) {
        *had_error = true; *derr = DataError{kind, col, line, true, msg};
    };
    auto empty_table = [&](const std::vector<std::string>& names) {
        auto t = std::make_shared<Table>(); t->ctx = c; t->nrows = 0;
        for (auto& nm : names) {
            Column col; col.name = nm; col.offsets = dev_alloc(c, 4); col.data = dev_alloc(c, 1);
            CPB_CUDA(cudaMemsetAsync(col.offsets->p, 0, 4, c->stream));
            t->cols.push_back(col);
        }
        return t;
    };
    std::vector<std::string> spec_names;
    for (auto& s : spec) spec_names.push_back(s.first);

    // encoding/csv readRecord's option validation
    if (o.delimiter == o.comment || !valid_delim(o.delimiter) || (o.comment != 0 && !valid_delim(o.comment))) {
        fail(CPB_E_INVALID_DELIM, -1, 1, "csv: invalid field or comment delimiter");
        return empty_table(spec_names);
    }
    if (o.comment != 0 || o.lazy_quotes || o.trim_leading_space || o.delimiter >= 0x80)
        throw ArgError{CPB_ERR_UNSUPPORTED,
                       "CommentChar / LazyQuotes / TrimLeadingSpace / multi-byte Delimiter are not lowered to kernels yet"};
    if ((reinterpret_cast<uintptr_t>(in) & 15) != 0) throw ArgError{CPB_ERR_ARG, "device input must be 16-byte aligned"};

    // ---- header kernel (first record + sampling)
    Buf hbuf = dev_alloc(c, sizeof(HeaderOut));
    {
        KernelTimer kt(c, "csv_header", 0);
        csv_header_kernel<<<1, 256, 0, c->stream>>>(in, n, (int)o.delimiter, hbuf->as<HeaderOut>());
        CPB_CUDA(cudaGetLastError());
    }
    HeaderOut* h = (HeaderOut*)c->pinned_scratch(sizeof(HeaderOut));
    // only the fixed part + what is needed: copy the whole struct (≈80 KB) — small next to the input
    CPB_CUDA(cudaMemcpyAsync(h, hbuf->p, sizeof(HeaderOut), cudaMemcpyDeviceToHost, c->stream));
    CPB_CUDA(cudaStreamSynchronize(c->stream));
    if (h->truncated) throw ArgError{CPB_ERR_UNSUPPORTED, "header row larger than 64 KiB / 4096 fields"};

    // ---- header resolution: makeHeader, csvplus.go:1149-1206
    std::vector<std::pair<std::string, int>> cols;  // output columns (name, field index)
    uint64_t data_start = 0;
    const bool hdr = o.header_from_first_row != 0;
    int first_nfields = h->eof ? 0 : h->nfields;
    if (hdr) {
        if (h->eof) { fail(CPB_E_EOF, -1, 1, "EOF"); return empty_table(spec_names); }
        if (h->err) { fail(h->err, -1, 1, kind_text(h->err)); return empty_table(spec_names); }
        if (o.num_fields > 0 && h->nfields != o.num_fields) { fail(CPB_E_FIELD_COUNT, -1, 1, kind_text(CPB_E_FIELD_COUNT)); return empty_table(spec_names); }
        std::vector<std::string> line;
        size_t offb = 0;
        for (int i = 0; i < h->nfields; i++) { line.emplace_back((const char*)h->bytes + offb, h->field_len[i]); offb += h->field_len[i]; }
        if (spec.empty()) {  // :1160-1168: later duplicates overwrite earlier ones
            for (size_t i = 0; i < line.size(); i++) {
                bool found = false;
                for (auto& cc : cols) if (cc.first == line[i]) { cc.second = (int)i; found = true; }
                if (!found) cols.emplace_back(line[i], (int)i);
            }
        } else {
            std::vector<int> got(spec.size(), -1);
            for (size_t i = 0; i < line.size(); i++) {  // :1174-1183
                for (size_t s = 0; s < spec.size(); s++) {
                    if (spec[s].first != line[i]) continue;
                    if (spec[s].second == -1 || spec[s].second == (int)i) got[s] = (int)i;
                    else {
                        fail(CPB_E_MISPLACED_COLUMN, (int)s, 1,
                             "misplaced column " + go_quote(line[i]) + ": expected at pos. " + std::to_string(spec[s].second) +
                                 ", but found at pos. " + std::to_string(i));
                        return empty_table(spec_names);
                    }
                }
            }
            std::vector<std::string> missing; int first_missing = -1;
            for (size_t s = 0; s < spec.size(); s++) if (got[s] < 0) { missing.push_back(spec[s].first); if (first_missing < 0) first_missing = (int)s; }
            if (!missing.empty()) {  // :1186-1202
                std::string m = missing.size() > 1 ? "columns not found: " : "column not found: ";
                for (size_t i = 0; i < missing.size(); i++) { if (i) m += ", "; m += missing[i]; }
                fail(CPB_E_COLUMN_NOT_FOUND, first_missing, 1, m);
                return empty_table(spec_names);
            }
            for (size_t s = 0; s < spec.size(); s++) cols.emplace_back(spec[s].first, got[s]);
        }
        data_start = h->data_start;
    } else {
        if (spec.empty()) throw ArgError{CPB_ERR_ARG, "Empty header spec"};  // csvplus.go:999-1001
        for (auto& s : spec) { if (s.second < 0) throw ArgError{CPB_ERR_ARG, "header spec: negative index for column " + s.first}; cols.push_back(s); }
        data_start = 0;
    }
    std::vector<std::string> names;
    for (auto& cc : cols) names.push_back(cc.first);
    const uint64_t line_base = hdr ? 2 : 1;  // DataSourceError.Line of data record 0 (csvplus.go:1102-1109)

    // ---- slots: distinct field indices, ascending
    std::vector<int> fields;
    for (auto& cc : cols) fields.push_back(cc.second);
    std::sort(fields.begin(), fields.end());
    fields.erase(std::unique(fields.begin(), fields.end()), fields.end());
    if ((int)fields.size() > MAXSEL)
        throw ArgError{CPB_ERR_UNSUPPORTED, "more than 16 columns in one fused parse; select fewer columns"};
    std::vector<int> col_slot;
    for (auto& cc : cols) col_slot.push_back((int)(std::lower_bound(fields.begin(), fields.end(), cc.second) - fields.begin()));
    const int nsel = (int)fields.size();

    Compiled comp;
    compile_pred(filter, [&](const std::string& key) {
        int col = -1;
        for (size_t i = 0; i < names.size(); i++) if (names[i] == key) col = (int)i;
        return col < 0 ? -1 : col_slot[col];
    }, comp);

    ParseParams P{};
    P.in = in; P.n = n; P.data_start = data_start; P.delim = o.delimiter;
    P.expect_fields = o.num_fields > 0 ? o.num_fields : (o.num_fields == 0 ? first_nfields : 0);
    P.pad_missing = o.num_fields < 0;
    P.nsel = nsel;
    for (int k = 0; k < nsel; k++) P.sel_field[k] = fields[k];
    P.pred = comp.prog;
    for (int t = 0; t < comp.prog.nterms; t++) P.slot_terms[comp.prog.term_col[t]] |= 1u << t;
    P.ntiles = data_start >= n ? 0 : (uint32_t)((n + TILE - 1) / TILE);

    if (P.ntiles == 0) return empty_table(names);

    Buf lits = dev_alloc(c, comp.lits.size() + 16);
    if (!comp.lits.empty())  // pageable source: the runtime stages it before returning
        CPB_CUDA(cudaMemcpyAsync(lits->p, comp.lits.data(), comp.lits.size(), cudaMemcpyHostToDevice, c->stream));
    P.lits = lits->as<uint8_t>();

    // ---- capacities (exact totals always come back; overflow => one exact rerun)
    const int NP = 2 + nsel;
    const uint64_t sample_bytes = h->sample_bytes, sample_newlines = h->sample_newlines;  // h aliases pinned scratch
    double avg = sample_newlines ? (double)sample_bytes / (double)sample_newlines : (double)n;
    if (avg < 2) avg = 2;
    uint64_t row_cap = (uint64_t)((double)(n - data_start) / avg * 1.10) + 4096;
    if (row_cap > (n - data_start) / 2 + 2) row_cap = (n - data_start) / 2 + 2;
    std::vector<uint64_t> data_cap(nsel, std::min<uint64_t>(n - data_start, 0xffffffffull));

    Buf state = dev_alloc(c, (size_t)P.ntiles * (8 + 16ull * NP) + 64 + sizeof(ParseResult));
    std::vector<Buf> offs(nsel), datas(nsel);
    ParseResult res{};
    for (int attempt = 0; attempt < 2; attempt++) {
        for (int k = 0; k < nsel; k++) {
            offs[k] = dev_alloc(c, (row_cap + 1) * 4);
            datas[k] = dev_alloc(c, data_cap[k] + 16);
            P.out_off[k] = offs[k]->as<uint32_t>(); P.out_data[k] = datas[k]->as<uint8_t>(); P.data_cap[k] = data_cap[k];
        }
        P.row_cap = row_cap;
        uint8_t* sp = state->as<uint8_t>();
        P.result = reinterpret_cast<ParseResult*>(sp); sp += sizeof(ParseResult);
        P.ticket = reinterpret_cast<uint32_t*>(sp); sp += 64;
        P.agg = reinterpret_cast<uint64_t*>(sp); sp += (size_t)P.ntiles * NP * 8;
        P.inc = reinterpret_cast<uint64_t*>(sp); sp += (size_t)P.ntiles * NP * 8;
        P.st1 = reinterpret_cast<uint32_t*>(sp); sp += (size_t)P.ntiles * 4;
        P.st2 = reinterpret_cast<uint32_t*>(sp);
        // zero: result totals, ticket, status words; err fields = ~0
        CPB_CUDA(cudaMemsetAsync(state->p, 0, sizeof(ParseResult) + 64, c->stream));
        CPB_CUDA(cudaMemsetAsync(&P.result->err_key, 0xff, 16, c->stream));
        CPB_CUDA(cudaMemsetAsync(P.st1, 0, (size_t)P.ntiles * 8, c->stream));
        uint64_t algo = n;  // S_in; S_out added by the caller of stats from the totals
        if (nsel <= 4) launch_scan<4>(c, P, algo);
        else if (nsel <= 8) launch_scan<8>(c, P, algo);
        else launch_scan<16>(c, P, algo);
        ParseResult* hr = (ParseResult*)c->pinned_scratch(sizeof(ParseResult));
        CPB_CUDA(cudaMemcpyAsync(hr, P.result, sizeof(ParseResult), cudaMemcpyDeviceToHost, c->stream));
        CPB_CUDA(cudaStreamSynchronize(c->stream));
        res = *hr;
        bool overflow = res.totals[1] > row_cap;
        for (int k = 0; k < nsel; k++) {
            if (res.totals[2 + k] > 0xffffffffull)
                throw DataError{CPB_E_TOO_LARGE, k, 0, false, "a column of this batch exceeds 4 GiB; parse the input in smaller batches"};
            if (res.totals[2 + k] > data_cap[k]) overflow = true;
        }
        if (!overflow) break;
        if (attempt == 1) throw ArgError{CPB_ERR_CUDA, "parse capacity overflow after exact resize"};
        row_cap = res.totals[1] + 1;
        for (int k = 0; k < nsel; k++) data_cap[k] = res.totals[2 + k];
    }
    // fold S_out into the algorithmic bytes of the scan record
    if (c->stats_on) {
        uint64_t s_out = 0;
        for (int k = 0; k < nsel; k++) s_out += res.totals[2 + k] + 4 * (res.totals[1] + 1);
        c->drain_events();
        c->stats["csv_scan"].bytes += s_out;
    }

    auto t = std::make_shared<Table>();
    t->ctx = c;
    t->nrows = (int64_t)res.totals[1];
    if (res.err_key != ~0ull) {
        uint64_t ordinal = res.err_key >> 16;
        int kind = (int)((res.err_key >> 8) & 0xff), slot = (int)(res.err_key & 0xff);
        t->nrows = (int64_t)res.err_rows;
        if (kind == CPB_E_COLUMN_INDEX) {
            int ci = 0;
            for (size_t i = 0; i < cols.size(); i++) if (col_slot[i] == slot) { ci = (int)i; break; }
            fail(kind, ci, line_base + ordinal,
                 "column not found: " + go_quote(cols[ci].first) + " (" + std::to_string(cols[ci].second) + ")");  // csvplus.go:1128
        } else fail(kind, -1, line_base + ordinal, kind_text(kind));
    }
    for (size_t i = 0; i < cols.size(); i++) {
        Column col; col.name = cols[i].first; col.offsets = offs[col_slot[i]]; col.data = datas[col_slot[i]];
        t->cols.push_back(col);
    }
    return t;
}

}  // namespace cpb
