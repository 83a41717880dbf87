// abi.cu — the extern "C" boundary (include/csvplus_b200.h): context, staging memory, table
// accessors and the thin argument marshalling around the kernels in parse/gather/sort/join/write.cu.
#include <algorithm>
#include <new>

#include "core.hpp"

using namespace cpb;


namespace cpb {

std::string go_quote(const std::string& s) {
    std::string r = "\"";
    for (unsigned char ch : s) {
        if (ch == '"' || ch == '\\') { r += '\\'; r += (char)ch; }
        else if (ch == '\n') r += "\\n";
        else if (ch == '\r') r += "\\r";
        else if (ch == '\t') r += "\\t";
        else if (ch < 0x20 || ch == 0x7f) { char b[8]; snprintf(b, sizeof b, "\\x%02x", ch); r += b; }
        else r += (char)ch;
    }
    return r + "\"";
}

// Requests are rounded up to size classes (1/8 steps of the enclosing power of two, <= 12.5 % slack): batches of
// slightly different sizes then ask the stream-ordered pool for identical blocks and reuse them instead of
// mapping fresh physical memory (which costs tens of milliseconds per GB).
static size_t size_class(size_t b) {
    if (b < (1u << 20)) return (b + 511) & ~(size_t)511;
    size_t p = 1;
    while ((p << 1) <= b) p <<= 1;
    const size_t step = p >> 3;
    return (b + step - 1) / step * step;
}
DevBuf::DevBuf(Ctx* c, size_t bytes) : ctx(c), n(bytes) {
    cudaError_t e = cudaMallocAsync(&p, size_class(bytes), c->pool, c->stream);
    if (e != cudaSuccess) { cudaGetLastError(); throw CudaFail{e, "cudaMallocAsync", __FILE__, __LINE__}; }
}
DevBuf::DevBuf(Ctx* owner, Ctx* user, size_t bytes) : ctx(owner), n(bytes) {
    cudaError_t e = cudaMallocAsync(&p, size_class(bytes), owner->pool, user->stream);
    if (e != cudaSuccess) { cudaGetLastError(); throw CudaFail{e, "cudaMallocAsync", __FILE__, __LINE__}; }
}
DevBuf::~DevBuf() {
    if (p && ctx) { cudaSetDevice(ctx->device); cudaFreeAsync(p, ctx->stream); }
}

void* Ctx::pinned_scratch(size_t n) {
    if (n > pinned_n) {
        if (pinned) { cudaStreamSynchronize(stream); cudaFreeHost(pinned); pinned = nullptr; }
        size_t want = std::max<size_t>(n, 1 << 20);
        CPB_CUDA(cudaMallocHost(&pinned, want));
        pinned_n = want;
    }
    return pinned;
}

// Timing events are recycled: a step launches ~60 timed kernels, and creating / keeping thousands of live events
// over many steps is not free.  harvest() folds the records whose kernels have already finished into the stats
// without blocking; drain_events() waits for the rest.
static void fold_event(Ctx* c, PendingEvent& pe) {
    float ms = 0;
    cudaEventElapsedTime(&ms, pe.a, pe.b);
    KStat& s = c->stats[pe.name];
    s.ms += ms; s.bytes += pe.bytes;
    c->event_pool.push_back(pe.a); c->event_pool.push_back(pe.b);
}
void Ctx::harvest_events() {
    size_t done = 0;
    while (done < pending.size() && cudaEventQuery(pending[done].b) == cudaSuccess) { fold_event(this, pending[done]); done++; }
    cudaGetLastError();  // cudaErrorNotReady of the first unfinished record
    if (done) pending.erase(pending.begin(), pending.begin() + (long)done);
}
void Ctx::drain_events() {
    for (auto& pe : pending) {
        cudaEventSynchronize(pe.b);
        fold_event(this, pe);
    }
    pending.clear();
}
static cudaEvent_t take_event(Ctx* c) {
    if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

KernelTimer::KernelTimer(Ctx* ctx, const char* nm, uint64_t algo_bytes, int launches)
    : c(ctx), on(ctx->stats_on), name(nm), bytes(algo_bytes), nlaunch(launches) {
    c->launches += launches;
    if (on) {
        if (c->pending.size() >= 64) c->harvest_events();
        a = take_event(c); b = take_event(c);
        cudaEventRecord(a, c->stream);
    }
}
KernelTimer::~KernelTimer() {
    if (on) {
        cudaEventRecord(b, c->stream);
        c->stats[name].launches += nlaunch;
        c->pending.push_back(PendingEvent{name, a, b, bytes});
    }
}

int translate_exception(Ctx* c, cpb_error* err) {
    try { throw; }
    catch (const CudaFail& f) {
        char buf[400];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)f.e, cudaGetErrorString(f.e), f.file, f.line, f.what);
        if (c) c->last_error = buf;
        fill_error(err, CPB_E_CUDA, -1, 0, false, buf);
        cudaGetLastError();
        return CPB_ERR_CUDA;
    } catch (const DataError& d) {
        fill_error(err, d.kind, d.column_index, d.line, d.has_line, d.msg);
        if (c) c->last_error = d.msg;
        return CPB_ERR_DATA;
    } catch (const ArgError& a) {
        if (c) c->last_error = a.msg;
        fill_error(err, a.status == CPB_ERR_UNSUPPORTED ? CPB_E_UNSUPPORTED : CPB_E_NONE, -1, 0, false, a.msg);
        return a.status;
    } catch (const std::bad_alloc&) {
        if (c) c->last_error = "out of host memory";
        return CPB_ERR_NOMEM;
    } catch (const std::exception& e) {
        if (c) c->last_error = e.what();
        fill_error(err, CPB_E_NONE, -1, 0, false, e.what());
        return CPB_ERR_ARG;
    } catch (...) {
        if (c) c->last_error = "unknown error";
        return CPB_ERR_ARG;
    }
}

}  // namespace cpb

static std::vector<std::string> str_list(const cpb_str* s, int n) {
    std::vector<std::string> v;
    for (int i = 0; i < n; i++) v.push_back(to_string(s[i]));
    return v;
}
static cpb_table* wrap(std::shared_ptr<Table> t) { return new cpb_table{std::move(t)}; }

extern "C" {

int cpb_abi_version(void) { return CPB_ABI_VERSION; }

int cpb_init(int device, cpb_ctx** out) {
    if (!out) return CPB_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return CPB_ERR_CUDA; }  // no CPU fallback
    cpb_ctx* h = new (std::nothrow) cpb_ctx();
    if (!h) return CPB_ERR_NOMEM;
    Ctx* c = &h->c;
    try {
        c->device = device;
        CPB_CUDA(cudaSetDevice(device));
        // CPB_BLOCKING_SYNC=1: host waits sleep on an interrupt instead of spinning.  One process per GPU on a shared host:
        // N spinning ranks (+ NCCL's proxy threads) can run a cgroup into its CPU quota, and a throttled rank stalls all
        // the others at the next collective (measured: steps of 55 ms with sporadic 150-750 ms ones at N = 2).
        if (const char* bs = getenv("CPB_BLOCKING_SYNC")) {
            if (bs[0] == '1') { cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync); cudaGetLastError(); }
        }
        cudaDeviceProp prop;
        CPB_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) throw ArgError{CPB_ERR_UNSUPPORTED, "csvplus_b200 kernels are built for sm_100a only"};
        c->sm_count = prop.multiProcessorCount;
        c->smem_optin = prop.sharedMemPerBlockOptin;
        CPB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        // a private stream-ordered pool per ctx: contexts running on different streams never wait on each other's
        // frees (the shared default pool may insert cross-stream dependencies to recycle memory)
        if (getenv("CPB_SHARED_POOL")) {
            CPB_CUDA(cudaDeviceGetDefaultMemPool(&c->pool, device));
        } else {
            cudaMemPoolProps props{};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = device;
            CPB_CUDA(cudaMemPoolCreate(&c->pool, &props));
            c->own_pool = true;
        }
        uint64_t thr = UINT64_MAX;  // keep freed blocks cached in the pool
        CPB_CUDA(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr));
    } catch (...) {
        int st = translate_exception(c, nullptr);
        delete h;
        return st;
    }
    *out = h;
    return CPB_OK;
}

void cpb_shutdown(cpb_ctx* h) {
    if (!h) return;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->drain_events();
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    c->event_pool.clear();
    cpb_comm_release(c);
    if (c->pinned) cudaFreeHost(c->pinned);
    cudaStreamDestroy(c->stream);
    if (c->own_pool) cudaMemPoolDestroy(c->pool);
    delete h;
}

void* cpb_ctx_stream(cpb_ctx* h) { return h ? (void*)h->c.stream : nullptr; }
const char* cpb_last_error(cpb_ctx* h) { return h ? h->c.last_error.c_str() : ""; }

int cpb_sync(cpb_ctx* h) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    sync_stream(c);
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

int cpb_pool_reserve(cpb_ctx* h, uint64_t nbytes) {
    if (!h) return CPB_ERR_ARG;
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    if (nbytes == 0) return CPB_OK;
    void* p = nullptr;
    CPB_CUDA(cudaMallocAsync(&p, nbytes, c->pool, c->stream));
    CPB_CUDA(cudaFreeAsync(p, c->stream));
    sync_stream(c);
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

int cpb_host_alloc(cpb_ctx* h, uint64_t n, void** out) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    CPB_CUDA(cudaMallocHost(out, n ? n : 1));
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}
int cpb_host_free(cpb_ctx* h, void* p) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    sync_stream(c);
    CPB_CUDA(cudaFreeHost(p));
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}
int cpb_device_alloc(cpb_ctx* h, uint64_t n, void** out) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    uint64_t padded = ((n + 15) & ~15ull) + 256;
    CPB_CUDA(cudaMallocAsync(out, padded, c->pool, c->stream));
    CPB_CUDA(cudaMemsetAsync((uint8_t*)*out + (n & ~15ull), 0, padded - (n & ~15ull), c->stream));
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}
int cpb_device_free(cpb_ctx* h, void* p) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    if (p) CPB_CUDA(cudaFreeAsync(p, c->stream));
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}
int cpb_memcpy_h2d(cpb_ctx* h, void* dst, const void* src, uint64_t n) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    CPB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, c->stream));
    sync_stream(c);
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}
int cpb_memcpy_d2h(cpb_ctx* h, void* dst, const void* src, uint64_t n) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    CPB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, c->stream));
    sync_stream(c);
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

// ------------------------------------------------------------------ parse
int cpb_parse_csv(cpb_ctx* h, const void* bytes, uint64_t nbytes, int on_device, const cpb_reader_opts* opts,
                  const cpb_header_col* spec, int nspec, const cpb_pred* filter, cpb_table** out, cpb_error* err) {
    if (!h || !out || !opts || (nbytes && !bytes) || nspec < 0) return CPB_ERR_ARG;
    Ctx* c = &h->c; DeviceGuard g(c);
    clear_error(err);
    *out = nullptr;
    CPB_TRY(c, err)
    std::vector<std::pair<std::string, int>> sp;
    for (int i = 0; i < nspec; i++) {
        std::string nm = to_string(spec[i].name);
        for (auto& s : sp) if (s.first == nm) throw ArgError{CPB_ERR_ARG, "header spec: duplicate column name: " + nm};  // csvplus.go:1047-1049
        sp.emplace_back(nm, spec[i].index);
    }
    Buf staged;
    const uint8_t* dev = (const uint8_t*)bytes;
    if (!on_device) {
        staged = dev_alloc(c, (nbytes + (32u << 20)) / (32u << 20) * (32u << 20));  // size class: batches reuse pool blocks
        if (nbytes) {
            KernelTimer kt(c, "h2d_input", nbytes, 0);
            CPB_CUDA(cudaMemcpyAsync(staged->p, bytes, nbytes, cudaMemcpyHostToDevice, c->stream));
        }
        dev = staged->as<uint8_t>();
    }
    bool had_error = false; DataError de{};
    auto t = parse_csv(c, dev, nbytes, *opts, sp, filter, &had_error, &de);
    *out = wrap(t);
    if (had_error) { fill_error(err, de.kind, de.column_index, de.line, de.has_line, de.msg); return CPB_ERR_DATA; }
    return CPB_OK;
    CPB_CATCH(c, err)
}

// ------------------------------------------------------------------ byte-range shards of one file
static const uint8_t* stage_input(Ctx* c, const void* bytes, uint64_t nbytes, int on_device, Buf& staged) {
    if (on_device) return (const uint8_t*)bytes;
    staged = dev_alloc(c, (nbytes + (32u << 20)) / (32u << 20) * (32u << 20));
    CPB_CUDA(cudaMemsetAsync(staged->as<uint8_t>() + (nbytes & ~15ull), 0, 32, c->stream));
    if (nbytes) {
        KernelTimer kt(c, "h2d_input", nbytes, 0);
        CPB_CUDA(cudaMemcpyAsync(staged->p, bytes, nbytes, cudaMemcpyHostToDevice, c->stream));
    }
    return staged->as<uint8_t>();
}
int cpb_csv_quote_parity(cpb_ctx* h, const void* bytes, uint64_t nbytes, int on_device, uint32_t* parity) {
    if (!h || !parity || (nbytes && !bytes)) return CPB_ERR_ARG;
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    Buf staged;
    const uint8_t* dev = stage_input(c, bytes, nbytes, on_device, staged);
    if ((reinterpret_cast<uintptr_t>(dev) & 15) != 0) throw ArgError{CPB_ERR_ARG, "device input must be 16-byte aligned"};
    *parity = quote_parity(c, dev, nbytes);
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}
int cpb_parse_csv_shard(cpb_ctx* h, const void* bytes, uint64_t nbytes, int on_device, uint64_t own_bytes, int shard_index, int is_last,
                        uint32_t initial_parity, const cpb_reader_opts* opts, const cpb_header_col* spec, int nspec, const cpb_pred* filter,
                        cpb_table** out, uint64_t* records, cpb_error* err) {
    if (!h || !out || !opts || (nbytes && !bytes) || nspec < 0 || shard_index < 0) return CPB_ERR_ARG;
    Ctx* c = &h->c; DeviceGuard g(c);
    clear_error(err);
    *out = nullptr;
    CPB_TRY(c, err)
    std::vector<std::pair<std::string, int>> sp;
    for (int i = 0; i < nspec; i++) sp.emplace_back(to_string(spec[i].name), spec[i].index);
    Buf staged;
    const uint8_t* dev = stage_input(c, bytes, nbytes, on_device, staged);
    uint64_t recs = 0;
    ShardArgs sh{own_bytes, shard_index, is_last != 0, initial_parity, &recs};
    bool had_error = false; DataError de{};
    auto t = parse_csv(c, dev, nbytes, *opts, sp, filter, &had_error, &de, &sh);
    if (records) *records = recs;
    *out = wrap(t);
    if (had_error) {
        // the line is LOCAL: (0-based) ordinal of the failing record among this shard's records; the host adds the records
        // of the shards before it and the reader's base (csvplus.go:1102-1137)
        const uint64_t base = opts->header_from_first_row ? 2 : 1;
        fill_error(err, de.kind, de.column_index, de.line >= base ? de.line - base : 0, de.has_line, de.msg);
        return CPB_ERR_DATA;
    }
    return CPB_OK;
    CPB_CATCH(c, err)
}

// ------------------------------------------------------------------ tables
int64_t cpb_table_num_rows(const cpb_table* t) { return t ? t->t->nrows : -1; }
int cpb_table_num_cols(const cpb_table* t) { return t ? (int)t->t->cols.size() : -1; }
int cpb_table_col_name(const cpb_table* t, int col, cpb_str* out) {
    if (!t || col < 0 || col >= (int)t->t->cols.size()) return CPB_ERR_ARG;
    out->ptr = t->t->cols[col].name.data(); out->len = t->t->cols[col].name.size();
    return CPB_OK;
}
int cpb_table_find_col(const cpb_table* t, cpb_str name) { return t ? t->t->find(to_string(name)) : -1; }
int cpb_table_col_field(const cpb_table* t, int col) {
    if (!t || col < 0 || col >= (int)t->t->src_field.size()) return -1;
    return t->t->src_field[col];
}
int cpb_table_record_fields(const cpb_table* t) { return t ? t->t->record_fields : -1; }

int cpb_table_col_bytes(cpb_ctx* h, const cpb_table* t, int col, int64_t lo, int64_t hi, uint64_t* nbytes) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    const Table& T = *t->t;
    if (col < 0 || col >= (int)T.cols.size() || lo < 0 || hi < lo || hi > T.nrows) throw ArgError{CPB_ERR_ARG, "bad column/row range"};
    uint32_t e[2];
    CPB_CUDA(cudaMemcpyAsync(&e[0], T.cols[col].off() + lo, 4, cudaMemcpyDeviceToHost, c->stream));
    CPB_CUDA(cudaMemcpyAsync(&e[1], T.cols[col].off() + hi, 4, cudaMemcpyDeviceToHost, c->stream));
    sync_stream(c);
    *nbytes = e[1] - e[0];
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

int cpb_table_fetch_column(cpb_ctx* h, const cpb_table* t, int col, int64_t lo, int64_t hi, int64_t* offsets_out,
                           uint8_t* data_out, uint64_t cap) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    const Table& T = *t->t;
    if (col < 0 || col >= (int)T.cols.size() || lo < 0 || hi < lo || hi > T.nrows) throw ArgError{CPB_ERR_ARG, "bad column/row range"};
    size_t cnt = (size_t)(hi - lo) + 1;
    std::vector<uint32_t> tmp(cnt);
    CPB_CUDA(cudaMemcpyAsync(tmp.data(), T.cols[col].off() + lo, cnt * 4, cudaMemcpyDeviceToHost, c->stream));
    sync_stream(c);
    uint64_t total = tmp[cnt - 1] - tmp[0];
    if (offsets_out) for (size_t i = 0; i < cnt; i++) offsets_out[i] = (int64_t)(tmp[i] - tmp[0]);
    if (data_out) {
        if (total > cap) throw ArgError{CPB_ERR_ARG, "data_out too small"};
        if (total) {
            CPB_CUDA(cudaMemcpyAsync(data_out, T.cols[col].bytes() + tmp[0], total, cudaMemcpyDeviceToHost, c->stream));
            sync_stream(c);
        }
    }
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

int cpb_table_column_device(const cpb_table* t, int col, const uint32_t** offsets, const uint8_t** data) {
    if (!t || col < 0 || col >= (int)t->t->cols.size()) return CPB_ERR_ARG;
    *offsets = t->t->cols[col].off(); *data = t->t->cols[col].bytes();
    return CPB_OK;
}

int cpb_table_from_host(cpb_ctx* h, int ncols, const cpb_str* names, const int64_t* const* offsets,
                        const uint8_t* const* data, int64_t nrows, cpb_table** out) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    auto t = std::make_shared<Table>(); t->ctx = c; t->nrows = nrows;
    for (int i = 0; i < ncols; i++) {
        Column col; col.name = to_string(names[i]);
        std::vector<uint32_t> off((size_t)nrows + 1);
        for (int64_t r = 0; r <= nrows; r++) {
            int64_t v = offsets[i][r] - offsets[i][0];
            if (v < 0 || v > 0xffffffffll) throw DataError{CPB_E_TOO_LARGE, i, 0, false, "column exceeds 4 GiB"};
            off[(size_t)r] = (uint32_t)v;
        }
        col.offsets = dev_alloc(c, off.size() * 4);
        CPB_CUDA(cudaMemcpyAsync(col.offsets->p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, c->stream));
        uint64_t total = off.back();
        col.data = dev_alloc(c, total + 16);
        if (total) CPB_CUDA(cudaMemcpyAsync(col.data->p, data[i] + offsets[i][0], total, cudaMemcpyHostToDevice, c->stream));
        sync_stream(c);
        t->cols.push_back(col);
    }
    *out = wrap(t);
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

int cpb_table_from_device(cpb_ctx* h, int ncols, const cpb_str* names, const uint32_t* const* offsets,
                          const uint8_t* const* data, int64_t nrows, cpb_table** out) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    // the caller's arrays are wrapped without ownership and copied by the concat path, which rebases offsets that do not
    // start at 0 (row-range views exported by cpb_table_column_device) and reads all extents with one host round trip
    Table view; view.ctx = c; view.nrows = nrows;
    for (int i = 0; i < ncols; i++) {
        Column col; col.name = to_string(names[i]);
        col.offsets = std::make_shared<DevBuf>(DevBuf::Borrow{}, offsets[i], ((size_t)nrows + 1) * 4);
        col.data = std::make_shared<DevBuf>(DevBuf::Borrow{}, data[i], 0);
        view.cols.push_back(col);
    }
    *out = wrap(concat_tables(c, {&view}));
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

int cpb_table_select(cpb_ctx* h, const cpb_table* t, const cpb_str* cols, int n, cpb_table** out, cpb_error* err) {
    Ctx* c = &h->c; DeviceGuard g(c);
    clear_error(err);
    CPB_TRY(c, err)
    if (n <= 0) throw ArgError{CPB_ERR_ARG, "no columns specified in SelectColumns()"};  // csvplus.go:512-514
    auto r = std::make_shared<Table>(); r->ctx = c; r->nrows = t->t->nrows; r->first_line = t->t->first_line;
    for (int i = 0; i < n; i++) {
        std::string nm = to_string(cols[i]);
        int k = t->t->find(nm);
        if (k < 0) {
            if (t->t->nrows == 0) { *out = wrap(r); r->cols.clear(); return CPB_OK; }  // no rows => Row.Select never runs
            throw DataError{CPB_E_MISSING_COLUMN, i, t->t->first_line, true, "missing column " + go_quote(nm)};  // csvplus.go:129
        }
        bool dup = false;
        for (auto& cc : r->cols) if (cc.name == nm) dup = true;  // a Go map holds a name once
        if (!dup) r->cols.push_back(t->t->cols[k]);
    }
    *out = wrap(r);
    return CPB_OK;
    CPB_CATCH(c, err)
}

int cpb_table_drop(cpb_ctx* h, const cpb_table* t, const cpb_str* cols, int n, cpb_table** out) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    if (n <= 0) throw ArgError{CPB_ERR_ARG, "no columns specified in DropColumns()"};  // csvplus.go:494-496
    auto names = str_list(cols, n);
    auto r = std::make_shared<Table>(); r->ctx = c; r->nrows = t->t->nrows; r->first_line = t->t->first_line;
    for (auto& col : t->t->cols)
        if (std::find(names.begin(), names.end(), col.name) == names.end()) r->cols.push_back(col);
    *out = wrap(r);
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

int cpb_table_slice(cpb_ctx* h, const cpb_table* t, int64_t lo, int64_t hi, cpb_table** out) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    const Table& T = *t->t;
    if (lo < 0) lo = 0;
    if (hi > T.nrows) hi = T.nrows;
    if (hi < lo) hi = lo;
    auto r = std::make_shared<Table>(); r->ctx = c; r->nrows = hi - lo; r->first_line = T.first_line + (uint64_t)lo;
    for (auto col : T.cols) { col.row0 += lo; r->cols.push_back(col); }
    *out = wrap(r);
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

int cpb_table_filter(cpb_ctx* h, const cpb_table* t, const cpb_pred* pred, cpb_table** out) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    *out = wrap(filter_table(c, *t->t, pred));
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

int cpb_table_first_false(cpb_ctx* h, const cpb_table* t, const cpb_pred* pred, int64_t* row) {
    if (!h || !t || !row) return CPB_ERR_ARG;
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    *row = first_false_row(c, *t->t, pred);
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

int cpb_table_concat(cpb_ctx* h, const cpb_table* const* parts, int nparts, cpb_table** out) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    if (nparts <= 0) throw ArgError{CPB_ERR_ARG, "no parts"};
    std::vector<const Table*> ps;
    for (int i = 0; i < nparts; i++) ps.push_back(parts[i]->t.get());
    *out = wrap(concat_tables(c, ps));
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}

void cpb_table_free(cpb_table* t) {
    if (!t) return;
    if (t->t && t->t->ctx) { std::unique_lock<std::mutex> lk(t->t->ctx->mu); cudaSetDevice(t->t->ctx->device); t->t.reset(); }
    delete t;
}

// ------------------------------------------------------------------ index / join
int cpb_index_build(cpb_ctx* h, const cpb_table* t, const cpb_str* key_cols, int nkeys, int unique, cpb_index** out,
                    cpb_error* err) {
    Ctx* c = &h->c; DeviceGuard g(c);
    clear_error(err);
    *out = nullptr;
    CPB_TRY(c, err)
    if (nkeys <= 0) throw ArgError{CPB_ERR_ARG, "empty column list in CreateIndex()"};  // csvplus.go:709-710
    auto keys = str_list(key_cols, nkeys);
    for (int i = 0; i < nkeys; i++)
        for (int j = i + 1; j < nkeys; j++)
            if (keys[i] == keys[j]) throw ArgError{CPB_ERR_ARG, "duplicate column name(s) in CreateIndex()"};  // :714-716
    DataError de{}; bool failed = false;
    auto ix = build_index(c, t->t, keys, unique != 0, &de, &failed);
    if (failed) { fill_error(err, de.kind, de.column_index, de.line, de.has_line, de.msg); return CPB_ERR_DATA; }
    *out = new cpb_index{ix};
    return CPB_OK;
    CPB_CATCH(c, err)
}
int64_t cpb_index_num_rows(const cpb_index* ix) { return ix ? ix->ix->nrows : -1; }
int cpb_index_num_keys(const cpb_index* ix) { return ix ? (int)ix->ix->key_cols.size() : -1; }
int cpb_index_table(cpb_ctx* h, const cpb_index* ix, cpb_table** out) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    *out = wrap(sorted_table(c, *ix->ix));
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}
int cpb_index_find(cpb_ctx* h, const cpb_index* ix, const cpb_str* values, int n, cpb_table** out) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    if (n > (int)ix->ix->key_cols.size()) throw ArgError{CPB_ERR_ARG, "too many columns in indexImpl.find()"};  // csvplus.go:876-878
    int64_t lo = 0, hi = ix->ix->nrows;
    if (n > 0) find_range(c, *ix->ix, str_list(values, n), &lo, &hi);
    const Table& T = *sorted_table(c, *ix->ix);
    auto r = std::make_shared<Table>(); r->ctx = c; r->nrows = hi - lo;
    for (auto col : T.cols) { col.row0 += lo; r->cols.push_back(col); }
    *out = wrap(r);
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}
int cpb_index_sub(cpb_ctx* h, const cpb_index* ix, const cpb_str* values, int n, cpb_index** out) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    if (n >= (int)ix->ix->key_cols.size()) throw ArgError{CPB_ERR_ARG, "too many values in SubIndex()"};  // csvplus.go:633-635
    int64_t lo = 0, hi = ix->ix->nrows;
    if (n > 0) find_range(c, *ix->ix, str_list(values, n), &lo, &hi);
    // rows are already sorted on the remaining key columns inside the range: rebuild the index image on the slice
    auto view = std::make_shared<Table>(); view->ctx = c; view->nrows = hi - lo;
    for (auto col : sorted_table(c, *ix->ix)->cols) { col.row0 += lo; view->cols.push_back(col); }
    std::vector<std::string> keys(ix->ix->key_cols.begin() + n, ix->ix->key_cols.end());
    DataError de{}; bool failed = false;
    auto sub = build_index(c, view, keys, false, &de, &failed);
    if (failed) throw de;
    *out = new cpb_index{sub};
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}
int cpb_index_dup_groups(cpb_ctx* h, const cpb_index* ix, int64_t* ngroups, int64_t** lo, int64_t** hi) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    std::vector<int64_t> a, b;
    index_dup_groups(c, *ix->ix, a, b);
    *ngroups = (int64_t)a.size();
    *lo = (int64_t*)malloc((a.size() + 1) * 8); *hi = (int64_t*)malloc((a.size() + 1) * 8);
    if (!*lo || !*hi) return CPB_ERR_NOMEM;
    if (!a.empty()) { memcpy(*lo, a.data(), a.size() * 8); memcpy(*hi, b.data(), b.size() * 8); }
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}
int cpb_index_dedup_apply(cpb_ctx* h, cpb_index* ix, int64_t ngroups, const int64_t* keep, int bug_compatible) {
    Ctx* c = &h->c; DeviceGuard g(c);
    CPB_TRY(c, nullptr)
    std::vector<int64_t> k(keep, keep + (ngroups > 0 ? ngroups : 0));
    index_dedup_apply(c, *ix->ix, k, bug_compatible != 0);
    return CPB_OK;
    CPB_CATCH(c, nullptr)
}
int cpb_index_dedup_apply2(cpb_ctx* h, cpb_index* ix, int64_t ngroups, const int64_t* keep, const cpb_table* replacements,
                           int bug_compatible, cpb_error* err) {
    Ctx* c = &h->c; DeviceGuard g(c);
    clear_error(err);
    CPB_TRY(c, err)
    std::vector<int64_t> k(keep, keep + (ngroups > 0 ? ngroups : 0));
    index_dedup_apply(c, *ix->ix, k, bug_compatible != 0, replacements ? replacements->t.get() : nullptr);
    return CPB_OK;
    CPB_CATCH(c, err)
}
void cpb_index_free(cpb_index* ix) {
    if (!ix) return;
    if (ix->ix && ix->ix->ctx) { std::unique_lock<std::mutex> lk(ix->ix->ctx->mu); cudaSetDevice(ix->ix->ctx->device); ix->ix.reset(); }
    delete ix;
}
void cpb_free(void* p) { free(p); }

static int join_common(cpb_ctx* h, const cpb_table* probe, const cpb_index* ix, const cpb_str* cols, int n, bool anti,
                       cpb_table** out, cpb_error* err) {
    Ctx* c = &h->c; DeviceGuard g(c);
    clear_error(err);
    *out = nullptr;
    CPB_TRY(c, err)
    std::vector<std::string> names = n ? str_list(cols, n) : ix->ix->key_cols;  // natural join, csvplus.go:546-547
    if (names.size() > ix->ix->key_cols.size())
        throw ArgError{CPB_ERR_ARG, anti ? "too many source columns in Except()" : "too many source columns in Join()"};  // :548-549, :591-592
    DataError de{}; bool failed = false;
    auto t = join_tables(c, *probe->t, *ix->ix, names, anti, &de, &failed);
    if (failed) { fill_error(err, de.kind, de.column_index, de.line, de.has_line, de.msg); return CPB_ERR_DATA; }
    *out = wrap(t);
    return CPB_OK;
    CPB_CATCH(c, err)
}
int cpb_join(cpb_ctx* h, const cpb_table* probe, const cpb_index* ix, const cpb_str* cols, int n, cpb_table** out, cpb_error* err) {
    return join_common(h, probe, ix, cols, n, false, out, err);
}
int cpb_except(cpb_ctx* h, const cpb_table* probe, const cpb_index* ix, const cpb_str* cols, int n, cpb_table** out, cpb_error* err) {
    return join_common(h, probe, ix, cols, n, true, out, err);
}

// ------------------------------------------------------------------ ToCsv
static int to_csv_common(cpb_ctx* h, const cpb_table* t, const cpb_str* cols, int n, bool to_host, void** bytes, uint64_t* nbytes,
                         cpb_error* err) {
    Ctx* c = &h->c; DeviceGuard g(c);
    clear_error(err);
    CPB_TRY(c, err)
    if (n <= 0) throw ArgError{CPB_ERR_ARG, "empty column list in ToCsv() function"};  // csvplus.go:380-382
    auto names = str_list(cols, n);
    std::vector<int> idx;
    for (int i = 0; i < n; i++) {
        int k = t->t->find(names[i]);
        if (k < 0 && t->t->nrows > 0)
            throw DataError{CPB_E_MISSING_COLUMN, i, t->t->first_line, true, "missing column " + go_quote(names[i])};  // :392, :145
        idx.push_back(k);
    }
    uint64_t total = 0;
    Buf out = table_to_csv(c, *t->t, idx, names, &total, nullptr);
    *nbytes = total;
    if (to_host) {
        void* hp = nullptr;
        CPB_CUDA(cudaMallocHost(&hp, total ? total : 1));
        if (total) CPB_CUDA(cudaMemcpyAsync(hp, out->p, total, cudaMemcpyDeviceToHost, c->stream));
        sync_stream(c);
        *bytes = hp;
    } else {
        void* dp = nullptr;
        CPB_CUDA(cudaMallocAsync(&dp, total + 16, c->pool, c->stream));
        if (total) CPB_CUDA(cudaMemcpyAsync(dp, out->p, total, cudaMemcpyDeviceToDevice, c->stream));
        *bytes = dp;
    }
    return CPB_OK;
    CPB_CATCH(c, err)
}
// ToCsv of one batch of a streamed result into caller-owned (ideally pinned) host memory: no allocation per call,
// the header line only when asked for (a csv.Writer writes it once, before the first batch: csvplus.go:387).
int cpb_table_to_csv_into(cpb_ctx* h, const cpb_table* t, const cpb_str* cols, int n, int with_header, void* host_dst,
                          uint64_t cap, uint64_t* nbytes, cpb_error* err) {
    if (!h || !t || !nbytes || (cap && !host_dst)) return CPB_ERR_ARG;
    Ctx* c = &h->c; DeviceGuard g(c);
    clear_error(err);
    CPB_TRY(c, err)
    if (n <= 0) throw ArgError{CPB_ERR_ARG, "empty column list in ToCsv() function"};  // csvplus.go:380-382
    auto names = str_list(cols, n);
    std::vector<int> idx;
    for (int i = 0; i < n; i++) {
        int k = t->t->find(names[i]);
        if (k < 0 && t->t->nrows > 0)
            throw DataError{CPB_E_MISSING_COLUMN, i, t->t->first_line, true, "missing column " + go_quote(names[i])};  // :392, :145
        idx.push_back(k);
    }
    uint64_t total = 0, hdr = 0;
    Buf out = table_to_csv(c, *t->t, idx, names, &total, &hdr);
    const uint64_t skip = with_header ? 0 : hdr;
    *nbytes = total - skip;
    if (*nbytes > cap) throw ArgError{CPB_ERR_ARG, "ToCsv destination too small"};
    if (*nbytes) {
        KernelTimer kt(c, "d2h_output", *nbytes, 0);
        CPB_CUDA(cudaMemcpyAsync(host_dst, out->as<uint8_t>() + skip, *nbytes, cudaMemcpyDeviceToHost, c->stream));
    }
    sync_stream(c);
    return CPB_OK;
    CPB_CATCH(c, err)
}
int cpb_table_to_csv(cpb_ctx* h, const cpb_table* t, const cpb_str* cols, int n, void** bytes, uint64_t* nbytes, cpb_error* err) {
    return to_csv_common(h, t, cols, n, true, bytes, nbytes, err);
}
int cpb_table_to_csv_device(cpb_ctx* h, const cpb_table* t, const cpb_str* cols, int n, void** bytes, uint64_t* nbytes, cpb_error* err) {
    return to_csv_common(h, t, cols, n, false, bytes, nbytes, err);
}

// ------------------------------------------------------------------ stats
int cpb_stats_enable(cpb_ctx* h, int on) { DeviceGuard g(&h->c); h->c.stats_on = on != 0; return CPB_OK; }
int cpb_stats_reset(cpb_ctx* h) {
    DeviceGuard g(&h->c);
    h->c.drain_events(); h->c.stats.clear();
    return CPB_OK;
}
int cpb_stats_get(cpb_ctx* h, cpb_kstat* out, int cap, int* n) {
    DeviceGuard g(&h->c);
    h->c.drain_events();
    int i = 0;
    for (auto& kv : h->c.stats) {
        if (i < cap) {
            memset(&out[i], 0, sizeof(cpb_kstat));
            strncpy(out[i].name, kv.first.c_str(), sizeof(out[i].name) - 1);
            out[i].launches = kv.second.launches; out[i].ms = kv.second.ms; out[i].algo_bytes = kv.second.bytes;
        }
        i++;
    }
    *n = i;
    return CPB_OK;
}
uint64_t cpb_kernel_launches(cpb_ctx* h) { return h ? h->c.launches : 0; }
uint64_t cpb_host_syncs(cpb_ctx* h) { return h ? h->c.host_syncs : 0; }

}  // extern "C"
