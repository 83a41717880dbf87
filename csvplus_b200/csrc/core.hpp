// core.hpp — host-side internals shared by the C-ABI translation units:
// context (device, stream, stream-ordered pool), refcounted device buffers, string columns,
// tables, indices, error helpers and per-kernel timing records.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/csvplus_b200.h"

namespace cpb {

struct Ctx;

// ------------------------------------------------------------------ errors
struct CudaFail { cudaError_t e; const char* what; const char* file; int line; };
#define CPB_CUDA(x)                                                        \
    do {                                                                   \
        cudaError_t _e = (x);                                              \
        if (_e != cudaSuccess) throw ::cpb::CudaFail{_e, #x, __FILE__, __LINE__}; \
    } while (0)

struct DataError {  // reference-visible error (cpb_error payload)
    int kind; int column_index; uint64_t line; bool has_line; std::string msg;
};
struct ArgError { int status; std::string msg; };

inline void fill_error(cpb_error* e, int kind, int col, uint64_t line, bool has_line, const std::string& msg) {
    if (!e) return;
    e->kind = kind; e->column_index = col; e->line = line; e->has_line = has_line ? 1 : 0; e->_pad = 0;
    size_t n = msg.size() < sizeof(e->msg) - 1 ? msg.size() : sizeof(e->msg) - 1;
    memcpy(e->msg, msg.data(), n); e->msg[n] = 0;
}
inline void clear_error(cpb_error* e) { if (e) { memset(e, 0, sizeof(*e)); e->column_index = -1; } }

// Go's %q for the plain names csvplus prints (csvplus.go:129, :725, :1128, :1179)
std::string go_quote(const std::string& s);

// ------------------------------------------------------------------ device memory
struct DevBuf {  // stream-ordered allocation owned by a ctx
    Ctx* ctx = nullptr; void* p = nullptr; size_t n = 0;
    DevBuf(Ctx* c, size_t bytes);
    // memory of `owner`'s pool, allocated in the stream order of `user` (usable by user's kernels at once) and
    // released on owner's stream: for structures another context builds lazily inside an object owner owns
    DevBuf(Ctx* owner, Ctx* user, size_t bytes);
    struct Borrow {};
    DevBuf(Borrow, const void* ptr, size_t bytes) : ctx(nullptr), p(const_cast<void*>(ptr)), n(bytes) {}  // caller-owned memory: never freed here
    ~DevBuf();
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
using Buf = std::shared_ptr<DevBuf>;

// Arrow-style string column: value i = data[offsets[i] .. offsets[i+1]); offsets are uint32 (a
// column of one table/batch holds < 4 GiB of bytes — DESIGN.md "batches").  `row0` lets row-range
// views (Top/Drop/Find/SubIndex) share the buffers.
struct Column {
    std::string name;
    Buf offsets;      // uint32[nrows_alloc+1]
    Buf data;         // uint8[]
    int64_t row0 = 0; // first row of this view inside `offsets`
    const uint32_t* off() const { return offsets->as<uint32_t>() + row0; }
    const uint8_t* bytes() const { return data ? data->as<uint8_t>() : nullptr; }
};

struct Table {
    Ctx* ctx = nullptr;
    int64_t nrows = 0;
    uint64_t first_line = 0;  // DataSourceError.Line the source reports for row 0 (csvplus.go:1137 / :243)
    std::vector<Column> cols;
    // set by the parser: the field index every column was read from and the field count of the file's first record —
    // what the shards of a file after the first need instead of the header row (cpb_parse_csv_shard)
    std::vector<int> src_field;
    int record_fields = 0;
    int find(const std::string& name) const {
        for (size_t i = 0; i < cols.size(); i++) if (cols[i].name == name) return (int)i;
        return -1;
    }
};

// csvplus Index (csvplus.go:610-614, :785-788): rows sorted on `key_cols`.  `table` holds the
// rows physically in sorted order.  `image` is the order-preserving fixed-width key image
// (sort.cu) of the sorted rows; `hash*` is the probe table over distinct full keys / key prefixes.
// "Built on some context's stream": what several contexts may use (the lazily built structures of an index) carries the
// event recorded after its last kernel; a user makes its own stream wait for it — no host synchronisation.
struct ReadyEvent {
    cudaEvent_t ev = nullptr;
    ~ReadyEvent() { if (ev) cudaEventDestroy(ev); }
};
using Ready = std::shared_ptr<ReadyEvent>;

struct HashTable {
    Ready ready;
    int nkeys = 0;           // number of leading key columns hashed
    uint32_t pbytes = 0;     // image bytes those columns cover
    uint64_t nslots = 0;     // power of two
    uint64_t nheads = 0;     // distinct key prefixes
    Buf slots;               // uint32 head ordinal or 0xFFFFFFFF (open addressing, linear probing)
    Buf heads;               // uint32[nheads+1]: sorted position of each distinct prefix; heads[nheads] = nrows
    uint64_t nslots32 = 0;   // embedded-key table (prefix <= 24 bytes, table too large for shared memory)
    Buf slots32;             // 32-byte slots: 3 key words + (first row | run length << 32)
    uint64_t nslots16 = 0;   // unique keys with prefix <= 12 bytes: 16-byte slots (12 key bytes + row + 1)
    Buf slots16;
};
// Row slots (gather.cu): the output columns of the sorted rows re-laid as one fixed-size slot per row (their
// bytes back to back) + one packed word of lengths per row, so that a join fetches an index row with ONE random
// access instead of two per column.  Built lazily per set of output columns when every value is short.
struct RowSlots {
    Ready ready;
    bool usable = false;
    uint32_t S = 0;   // slot bytes (multiple of 16, <= 64)
    Buf slots;        // uint8[nrows][S]
    Buf lens;         // uint32[nrows]: length of column c in bits 8c..8c+7
};
struct Index {
    Ctx* ctx = nullptr;
    // The sorted rows exist in two forms.  A join needs neither the sorted columns nor their offsets: it probes the
    // hash table (built from the sorted key image) and fetches payload through row slots laid out straight from the
    // source rows + the sort permutation.  So the index build keeps (src, perm) and the physically sorted table is
    // materialised only when something iterates / looks up / dedups the index (sorted_table()).
    std::shared_ptr<Table> src;        // the rows as they were given to IndexOn (unsorted); null after a dedup
    Buf perm;                          // uint32[nrows]: sorted position -> row of src
    int64_t nrows = 0;
    std::shared_ptr<Table> table;      // sorted rows (lazily materialised, under `mu`)
    // A UNIQUE index whose duplicate check passed is not even sorted until something needs the order (iteration, Find,
    // SubIndex, a prefix join, dedup): a full-key join only needs "key -> row", which the hash table built over the rows
    // in SOURCE order gives (and that build doubles as the duplicate check: every key must find itself).  `uimage` is the
    // key image in source order; `sorted` tells whether perm / image (sorted order) exist yet (ensure_sorted()).
    Buf uimage;
    bool sorted = true;
    std::map<int, HashTable> hash_src;                   // probe tables over the source order (payload = source row)
    std::map<std::vector<int>, RowSlots> row_slots_src;  // row slots in source order (filled sequentially, no permutation)
    const Table& schema() const { return table ? *table : *src; }  // column names / count only
    bool unique = false;               // built by UniqueIndexOn and verified: every full key occurs once
    std::vector<std::string> key_cols; // index.impl.columns
    std::vector<int> key_col_idx;      // positions in table->cols
    std::vector<uint32_t> key_width;   // per key column: max value length (bytes) in this index
    uint32_t image_words = 0;          // uint64 words per row of the key image
    Buf image;                         // uint64[image_words][nrows] (word-major / SoA)
    std::map<int, HashTable> hash;     // by number of leading key columns (built lazily, under `mu`)
    std::map<std::vector<int>, RowSlots> row_slots;  // by output column set (built lazily, under `mu`)
    std::mutex mu;                     // several contexts may probe one index concurrently
};

// ------------------------------------------------------------------ kernel stats
struct KStat { uint64_t launches = 0; double ms = 0; uint64_t bytes = 0; };
struct PendingEvent { std::string name; cudaEvent_t a, b; uint64_t bytes; };

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;
    bool own_pool = false;
    std::mutex mu;              // one call in flight per ctx
    std::string last_error;
    int sm_count = 148;
    size_t smem_optin = 0;
    // stats
    bool stats_on = false;
    std::map<std::string, KStat> stats;
    std::vector<PendingEvent> pending;
    std::vector<cudaEvent_t> event_pool;  // recycled timing events
    uint64_t launches = 0;
    // small pinned scratch for D2H of results
    void* pinned = nullptr; size_t pinned_n = 0;
    // multi-GPU (comm.cu): ncclComm_t of this rank, null for a single-GPU ctx
    void* comm = nullptr; int nranks = 1, rank = 0; bool comm_owned = false;
    Ctx* alloc_for = nullptr;  // see dev_alloc
    // host synchronisations this ctx has issued (cudaStreamSynchronize on its stream): a pipeline step should need few
    uint64_t host_syncs = 0;

    void* pinned_scratch(size_t n);
    void drain_events();
    void harvest_events();
};

// RAII: time a kernel (or a group of launches) on the ctx stream when stats are enabled
struct KernelTimer {
    Ctx* c; bool on; cudaEvent_t a{}, b{}; std::string name; uint64_t bytes; int nlaunch;
    KernelTimer(Ctx* ctx, const char* nm, uint64_t algo_bytes, int launches = 1);
    ~KernelTimer();
};

// every blocking wait of the host on a ctx stream goes through here and is counted (cpb_host_syncs): a pipeline step
// should need few of them
inline void sync_stream(Ctx* c) {
    c->host_syncs++;
    CPB_CUDA(cudaStreamSynchronize(c->stream));
}

inline Ready record_ready(Ctx* c) {
    auto r = std::make_shared<ReadyEvent>();
    CPB_CUDA(cudaEventCreateWithFlags(&r->ev, cudaEventDisableTiming));
    CPB_CUDA(cudaEventRecord(r->ev, c->stream));
    return r;
}
inline void wait_ready(Ctx* c, const Ready& r) { if (r && r->ev) CPB_CUDA(cudaStreamWaitEvent(c->stream, r->ev, 0)); }

struct DeviceGuard {  // every entry point: select device, serialise on the ctx
    std::unique_lock<std::mutex> lk;
    explicit DeviceGuard(Ctx* c) : lk(c->mu) { cudaSetDevice(c->device); }
};

inline Buf dev_alloc_owned(Ctx* owner, Ctx* user, size_t bytes) { return std::make_shared<DevBuf>(owner, user, bytes ? bytes : 1); }
// (while a ctx builds something that stays inside an object another ctx owns, its allocations come from that owner's pool)
inline Buf dev_alloc(Ctx* c, size_t bytes) {
    if (c->alloc_for && c->alloc_for != c) return dev_alloc_owned(c->alloc_for, c, bytes);
    return std::make_shared<DevBuf>(c, bytes ? bytes : 1);
}

inline std::string to_string(cpb_str s) { return std::string(s.ptr ? s.ptr : "", (size_t)s.len); }

// exception -> status translation used by every extern "C" body
int translate_exception(Ctx* c, cpb_error* err);
#define CPB_TRY(ctx, err) try {
#define CPB_CATCH(ctx, err) } catch (...) { return ::cpb::translate_exception(ctx, err); }

// ------------------------------------------------------------------ cross-TU operations
// parse.cu
struct ShardArgs {  // byte-range shard of one file (cpb_parse_csv_shard)
    uint64_t own_bytes;  // records starting at buffer positions (0, own_bytes] are this shard's ([0, own_bytes] for index 0)
    int index;           // 0 = the shard holding the start of the file
    bool is_last;        // the buffer ends where the file ends
    uint32_t pin0;       // parity of the quote bytes of the file before this buffer
    uint64_t* records;   // out: records owned by this shard (before filtering)
};
std::shared_ptr<Table> parse_csv(Ctx* c, const uint8_t* dev_bytes, uint64_t n, const cpb_reader_opts& o,
                                 const std::vector<std::pair<std::string, int>>& spec, const cpb_pred* filter,
                                 bool* had_error, DataError* derr, const ShardArgs* sh = nullptr);
uint32_t quote_parity(Ctx* c, const uint8_t* dev_bytes, uint64_t n);
// gather.cu
void exclusive_scan_u32(Ctx* c, const uint32_t* in, uint32_t* out, uint64_t n, uint64_t* total_dev);  // out[n] = total
Column gather_column(Ctx* c, const Column& src, const uint32_t* row_ids, int64_t nout);
std::shared_ptr<Table> gather_rows(Ctx* c, const Table& t, const uint32_t* row_ids, int64_t nout);
std::shared_ptr<Table> filter_table(Ctx* c, const Table& t, const cpb_pred* pred);
int64_t first_false_row(Ctx* c, const Table& t, const cpb_pred* pred);
Column materialize(Ctx* c, const Column& col, int64_t nrows);  // view -> own compact buffers
// rows `ids` of the index's sorted table restricted to columns `cols` (row-slot path when it pays, else gather_rows)
std::shared_ptr<Table> gather_index_rows(Ctx* c, Index& ix, const std::vector<int>& cols, const uint32_t* ids, int64_t nout,
                                         bool by_src = false);
std::shared_ptr<Table> concat_tables(Ctx* c, const std::vector<const Table*>& parts);
// sort.cu
std::shared_ptr<Index> build_index(Ctx* c, std::shared_ptr<Table> t, const std::vector<std::string>& keys, bool unique,
                                   DataError* derr, bool* failed);
std::shared_ptr<Table> sorted_table(Ctx* c, Index& ix);  // materialises ix.table on first use
void ensure_sorted(Ctx* c, Index& ix);                    // sorts a lazily-sorted unique index on first need of the order
constexpr int MAXKEYS = 16;
struct KeyDesc {  // key columns + the widths of the order-preserving image (sort.cu)
    int nkeys;
    const uint32_t* off[MAXKEYS];
    const uint8_t* data[MAXKEYS];
    uint32_t width[MAXKEYS];     // bytes of value kept
    uint32_t lenbytes[MAXKEYS];  // bytes of the big-endian length field
    uint32_t words;              // image words per row
};
void describe_keys(Ctx* c, const Table& t, const std::vector<int>& kidx, const std::vector<uint32_t>& width, KeyDesc& kd);
uint32_t prefix_bytes(const Index& ix, int nk);  // image bytes covered by the first nk key columns
Buf pack_with_widths(Ctx* c, const Table& t, const std::vector<int>& kidx, const std::vector<uint32_t>& width, uint32_t* words_out);
// join.cu
std::shared_ptr<Table> join_tables(Ctx* c, const Table& probe, Index& ix, const std::vector<std::string>& cols,
                                   bool anti, DataError* derr, bool* failed);
void find_range(Ctx* c, Index& ix, const std::vector<std::string>& values, int64_t* lo, int64_t* hi);
bool index_has_duplicates(Ctx* c, Index& ix);  // full-key probe table over the source order + self probe
void index_dup_groups(Ctx* c, Index& ix, std::vector<int64_t>& lo, std::vector<int64_t>& hi);
void index_dedup_apply(Ctx* c, Index& ix, const std::vector<int64_t>& keep, bool bug_compatible, const Table* repl = nullptr);
// write.cu
Buf table_to_csv(Ctx* c, const Table& t, const std::vector<int>& cols, const std::vector<std::string>& names,
                 uint64_t* nbytes, uint64_t* header_bytes);

}  // namespace cpb

extern "C" void cpb_comm_release(cpb::Ctx* c);  // comm.cu

// opaque handle layouts of the C ABI
struct cpb_ctx { cpb::Ctx c; };
struct cpb_table { std::shared_ptr<cpb::Table> t; };
struct cpb_index { std::shared_ptr<cpb::Index> ix; };
