// gather.cu — column-level primitives shared by Filter / IndexOn / Join:
//   * device-wide exclusive scan (single pass, decoupled look-back);
//   * string-column gather by row ids (lengths -> scan -> byte copy) = the columnar form of the
//     per-row map copies the reference makes (Row.Select csvplus.go:122-134, mergeRows :571-583);
//   * Filter(Like/All/Any/Not) over a materialised table (csvplus.go:276-286, :1243-1293);
//   * row-wise concatenation of tables (assembling all-gathered shards).
#include <algorithm>

#include "core.hpp"
#include "pred.cuh"
#include "util.cuh"

namespace cpb {

// ------------------------------------------------------------------ exclusive scan (uint32 in, uint32 out, uint64 total)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// tile state word: bits 63:62 status (0 none, 1 aggregate, 2 inclusive), bits 61:0 value
__global__ void __launch_bounds__(SCAN_THREADS) scan_u32_kernel(const uint32_t* in, uint32_t* out, uint64_t n,
                                                                unsigned long long* state, uint32_t* ticket,
                                                                unsigned long long* total) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    for (;;) {
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint64_t tile = s_tile;
        if (tile >= ntiles) break;
        const uint64_t base = tile * SCAN_TILE + (uint64_t)tid * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS];
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = base + i < n ? in[base + i] : 0; sum += v[i]; }
        uint32_t inc = warp_incl_scan(sum);
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t woff = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < SCAN_THREADS / 32; i++) { uint32_t t = s_warp[i]; if (i < warp) woff += t; tot += t; }
        if (warp == 0) {
            unsigned long long excl = 0;
            if (tile == 0) {
                if (lane == 0) atomicExch(&state[0], (2ull << 62) | tot);
            } else {
                if (lane == 0) atomicExch(&state[tile], (1ull << 62) | tot);
                int64_t b = (int64_t)tile - 1;
                for (;;) {
                    int64_t p = b - lane;
                    unsigned long long s = 2ull << 62;
                    if (p >= 0) { do { s = ld_relaxed_u64((const uint64_t*)&state[p]); } while ((s >> 62) == 0); }
                    uint32_t incl = __ballot_sync(0xffffffffu, (s >> 62) == 2);
                    int f = __ffs(incl) - 1;
                    unsigned long long val = (f < 0 || lane <= f) ? (s & ((1ull << 62) - 1)) : 0ull;
                    excl += warp_sum_u64(val);
                    if (f >= 0) break;
                    b -= 32;
                }
                if (lane == 0) atomicExch(&state[tile], (2ull << 62) | (excl + tot));
            }
            if (lane == 0) { s_prefix = excl; if (tile == ntiles - 1) { *total = excl + tot; out[n] = (uint32_t)(excl + tot); } }
        }
        __syncthreads();
        uint32_t run = (uint32_t)s_prefix + woff + (inc - sum);
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out[base + i] = run; run += v[i]; }
        __syncthreads();
    }
}

__global__ void scan_empty_kernel(uint32_t* out, unsigned long long* total) { out[0] = 0; *total = 0; }

// out may alias in; out has n+1 entries (out[n] = total, truncated to 32 bits); *total_dev holds the 64-bit total
void exclusive_scan_u32(Ctx* c, const uint32_t* in, uint32_t* out, uint64_t n, uint64_t* total_dev) {
    if (n == 0) {
        KernelTimer kt(c, "scan_u32", 0);
        scan_empty_kernel<<<1, 1, 0, c->stream>>>(out, (unsigned long long*)total_dev);
        return;
    }
    uint64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    Buf st = dev_alloc(c, ntiles * 8 + 64);
    CPB_CUDA(cudaMemsetAsync(st->p, 0, ntiles * 8 + 64, c->stream));
    unsigned long long* state = st->as<unsigned long long>() + 8;
    uint32_t* ticket = st->as<uint32_t>();
    uint32_t grid = (uint32_t)std::min<uint64_t>(ntiles, (uint64_t)c->sm_count * 8);
    KernelTimer kt(c, "scan_u32", n * 8);
    scan_u32_kernel<<<grid, SCAN_THREADS, 0, c->stream>>>(in, out, n, state, ticket, (unsigned long long*)total_dev);
    CPB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ gather by row ids
__global__ void gather_len_kernel(const uint32_t* __restrict__ off, const uint32_t* __restrict__ ids, uint32_t* out_len, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { uint32_t r = ids ? ids[i] : (uint32_t)i; out_len[i] = off[r + 1] - off[r]; }
}
// One warp gathers 32 consecutive output values.  Their destination bytes are contiguous, so the lanes first
// copy their (randomly placed) source strings into a per-warp shared-memory stage laid out like the
// destination, then the warp writes the stage with aligned 16-byte stores: HBM/L2 see full sectors instead of one
// scattered byte store per lane.  ids==nullptr means identity (compaction of a view).
constexpr int GW_WARPS = 8;
constexpr int GW_STAGE = 2048;  // bytes staged per warp; longer groups take the direct path
__global__ void __launch_bounds__(GW_WARPS * 32) gather_copy_kernel(const uint32_t* __restrict__ src_off, const uint8_t* __restrict__ src,
                                                                    const uint32_t* __restrict__ ids, const uint32_t* __restrict__ dst_off,
                                                                    uint8_t* __restrict__ dst, uint64_t n) {
    __shared__ __align__(16) uint8_t stage_all[GW_WARPS][GW_STAGE + 16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* stage = stage_all[warp];
    const uint64_t nwarps = (uint64_t)gridDim.x * GW_WARPS;
    for (uint64_t g = (uint64_t)blockIdx.x * GW_WARPS + warp; g * 32 < n; g += nwarps) {
        const uint64_t i = g * 32 + lane;
        uint32_t s = 0, len = 0, d = 0;
        if (i < n) {
            const uint32_t r = ids ? ids[i] : (uint32_t)i;
            s = src_off[r]; len = src_off[r + 1] - s; d = dst_off[i];
        }
        const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
        const uint64_t last = (g * 32 + 31 < n ? g * 32 + 31 : n - 1) - g * 32;
        const uint32_t dl = __shfl_sync(0xffffffffu, d + len, (int)last);
        const uint32_t total = dl - d0, sh = d0 & 15u;
        const uint8_t* sp = src + s;
        if (sh + total <= GW_STAGE) {
            uint8_t* q = stage + sh + (d - d0);
            for (uint32_t k = 0; k < len; k++) q[k] = sp[k];
            __syncwarp();
            uint8_t* gb = dst + (d0 - sh);
            for (uint32_t x = lane * 16; x < sh + total; x += 32 * 16) {
                if (x >= sh && x + 16 <= sh + total) *reinterpret_cast<uint4*>(gb + x) = *reinterpret_cast<const uint4*>(stage + x);
                else for (uint32_t y = x; y < x + 16; y++) if (y >= sh && y < sh + total) gb[y] = stage[y];
            }
            __syncwarp();
        } else {
            uint8_t* dp = dst + d;
            for (uint32_t k = 0; k < len; k++) dp[k] = sp[k];
        }
    }
}

static inline uint32_t blocks_for(uint64_t n, int threads) { return (uint32_t)((n + threads - 1) / threads); }

struct PendingGather { Column col; const Column* src; Buf total; };

// gathers several columns with ONE host synchronisation (totals of all columns read back together)
static std::vector<Column> gather_columns(Ctx* c, const std::vector<const Column*>& srcs, const uint32_t* ids, int64_t nout) {
    std::vector<Column> out(srcs.size());
    if (srcs.empty()) return out;
    Buf totals = dev_alloc(c, srcs.size() * 8);
    for (size_t k = 0; k < srcs.size(); k++) {
        out[k].name = srcs[k]->name;
        out[k].offsets = dev_alloc(c, ((size_t)nout + 1) * 4);
        uint32_t* o = out[k].offsets->as<uint32_t>();
        if (nout) {
            KernelTimer kt(c, "gather_len", (uint64_t)nout * 12);
            gather_len_kernel<<<blocks_for(nout, 256), 256, 0, c->stream>>>(srcs[k]->off(), ids, o, (uint64_t)nout);
            CPB_CUDA(cudaGetLastError());
        }
        exclusive_scan_u32(c, o, o, (uint64_t)nout, totals->as<uint64_t>() + k);
    }
    uint64_t* ht = (uint64_t*)c->pinned_scratch(srcs.size() * 8);
    CPB_CUDA(cudaMemcpyAsync(ht, totals->p, srcs.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    CPB_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<uint64_t> tot(ht, ht + srcs.size());
    for (size_t k = 0; k < srcs.size(); k++) {
        if (tot[k] > 0xffffffffull) throw DataError{CPB_E_TOO_LARGE, (int)k, 0, false, "a result column exceeds 4 GiB; process in smaller batches"};
        out[k].data = dev_alloc(c, tot[k] + 16);
        if (nout && tot[k]) {
            KernelTimer kt(c, "gather_copy", 2 * tot[k] + (uint64_t)nout * 12);
            const uint32_t gblocks = (uint32_t)std::min<uint64_t>(((uint64_t)nout + GW_WARPS * 32 - 1) / (GW_WARPS * 32), (uint64_t)c->sm_count * 16);
            gather_copy_kernel<<<gblocks, GW_WARPS * 32, 0, c->stream>>>(srcs[k]->off(), srcs[k]->bytes(), ids,
                                                                              out[k].offsets->as<uint32_t>(), out[k].data->as<uint8_t>(), (uint64_t)nout);
            CPB_CUDA(cudaGetLastError());
        }
    }
    return out;
}

Column gather_column(Ctx* c, const Column& src, const uint32_t* row_ids, int64_t nout) {
    return gather_columns(c, {&src}, row_ids, nout)[0];
}

std::shared_ptr<Table> gather_rows(Ctx* c, const Table& t, const uint32_t* row_ids, int64_t nout) {
    auto r = std::make_shared<Table>(); r->ctx = c; r->nrows = nout; r->first_line = t.first_line;
    // columns sharing buffers (AssumeHeader aliases) are gathered once
    std::vector<const Column*> uniq; std::vector<int> which(t.cols.size());
    for (size_t i = 0; i < t.cols.size(); i++) {
        int f = -1;
        for (size_t u = 0; u < uniq.size(); u++)
            if (uniq[u]->offsets == t.cols[i].offsets && uniq[u]->data == t.cols[i].data && uniq[u]->row0 == t.cols[i].row0) f = (int)u;
        if (f < 0) { f = (int)uniq.size(); uniq.push_back(&t.cols[i]); }
        which[i] = f;
    }
    auto g = gather_columns(c, uniq, row_ids, nout);
    for (size_t i = 0; i < t.cols.size(); i++) { Column col = g[which[i]]; col.name = t.cols[i].name; r->cols.push_back(col); }
    return r;
}

Column materialize(Ctx* c, const Column& col, int64_t nrows) { return gather_columns(c, {&col}, nullptr, nrows)[0]; }

// ------------------------------------------------------------------ Filter over a table
struct FilterCols { const uint32_t* off[MAXTERMS]; const uint8_t* data[MAXTERMS]; };

__global__ void filter_flags_kernel(FilterCols cols, PredProg prog, const uint8_t* __restrict__ lits, uint32_t* flags, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t eq = 0;
    for (int t = 0; t < prog.nterms; t++) {
        const uint32_t* off = cols.off[t];
        uint32_t s = off[i], len = off[i + 1] - s;
        if (len != prog.term_len[t]) continue;
        const uint8_t* p = cols.data[t] + s;
        const uint8_t* l = lits + prog.term_off[t];
        bool ok = true;
        for (uint32_t k = 0; k < len; k++) if (p[k] != l[k]) { ok = false; break; }
        if (ok) eq |= 1u << t;
    }
    flags[i] = eval_pred(prog, eq) ? 1u : 0u;
}
__global__ void compact_ids_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, uint32_t* ids, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) ids[pos[i]] = (uint32_t)i;
}

std::shared_ptr<Table> filter_table(Ctx* c, const Table& t, const cpb_pred* pred) {
    if (!pred) throw ArgError{CPB_ERR_ARG, "nil predicate"};
    Compiled comp;
    compile_pred(pred, [&](const std::string& key) { return t.find(key); }, comp);
    const uint64_t n = (uint64_t)t.nrows;
    if (n == 0) return gather_rows(c, t, nullptr, 0);
    FilterCols fc{};
    for (int i = 0; i < comp.prog.nterms; i++) { fc.off[i] = t.cols[comp.prog.term_col[i]].off(); fc.data[i] = t.cols[comp.prog.term_col[i]].bytes(); }
    Buf lits = dev_alloc(c, comp.lits.size() + 16);
    if (!comp.lits.empty()) CPB_CUDA(cudaMemcpyAsync(lits->p, comp.lits.data(), comp.lits.size(), cudaMemcpyHostToDevice, c->stream));
    Buf flags = dev_alloc(c, n * 4), pos = dev_alloc(c, (n + 1) * 4), total = dev_alloc(c, 8);
    {
        KernelTimer kt(c, "filter_like", n * 12);
        filter_flags_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(fc, comp.prog, lits->as<uint8_t>(), flags->as<uint32_t>(), n);
        CPB_CUDA(cudaGetLastError());
    }
    exclusive_scan_u32(c, flags->as<uint32_t>(), pos->as<uint32_t>(), n, total->as<uint64_t>());
    uint64_t* ht = (uint64_t*)c->pinned_scratch(8);
    CPB_CUDA(cudaMemcpyAsync(ht, total->p, 8, cudaMemcpyDeviceToHost, c->stream));
    CPB_CUDA(cudaStreamSynchronize(c->stream));
    const uint64_t m = *ht;
    Buf ids = dev_alloc(c, (m + 1) * 4);
    {
        KernelTimer kt(c, "compact_ids", n * 8 + m * 4);
        compact_ids_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(flags->as<uint32_t>(), pos->as<uint32_t>(), ids->as<uint32_t>(), n);
        CPB_CUDA(cudaGetLastError());
    }
    return gather_rows(c, t, ids->as<uint32_t>(), (int64_t)m);
}

// ------------------------------------------------------------------ concat
__global__ void rebase_offsets_kernel(const uint32_t* __restrict__ in, uint32_t* out, uint64_t n, uint32_t add) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = in[i] - in[0] + add;  // also writes the sentinel (overwritten by the next part's first entry: same value)
}

std::shared_ptr<Table> concat_tables(Ctx* c, const std::vector<const Table*>& parts) {
    const Table& first = *parts[0];
    for (auto* p : parts) {
        if (p->cols.size() != first.cols.size()) throw ArgError{CPB_ERR_ARG, "concat: column count mismatch"};
        for (size_t k = 0; k < first.cols.size(); k++) if (p->cols[k].name != first.cols[k].name) throw ArgError{CPB_ERR_ARG, "concat: column name mismatch"};
    }
    int64_t nrows = 0;
    for (auto* p : parts) nrows += p->nrows;
    const size_t K = first.cols.size(), NPARTS = parts.size();
    // byte extents of every (part, column): one D2H round trip
    std::vector<uint32_t> ext(2 * K * NPARTS);
    uint32_t* hp = (uint32_t*)c->pinned_scratch(ext.size() * 4);
    for (size_t p = 0; p < NPARTS; p++)
        for (size_t k = 0; k < K; k++) {
            const uint32_t* off = parts[p]->cols[k].off();
            CPB_CUDA(cudaMemcpyAsync(hp + 2 * (p * K + k), off, 4, cudaMemcpyDeviceToHost, c->stream));
            CPB_CUDA(cudaMemcpyAsync(hp + 2 * (p * K + k) + 1, off + parts[p]->nrows, 4, cudaMemcpyDeviceToHost, c->stream));
        }
    CPB_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(ext.data(), hp, ext.size() * 4);
    auto r = std::make_shared<Table>(); r->ctx = c; r->nrows = nrows; r->first_line = first.first_line;
    for (size_t k = 0; k < K; k++) {
        uint64_t total = 0;
        for (size_t p = 0; p < NPARTS; p++) total += ext[2 * (p * K + k) + 1] - ext[2 * (p * K + k)];
        if (total > 0xffffffffull) throw DataError{CPB_E_TOO_LARGE, (int)k, 0, false, "concatenated column exceeds 4 GiB"};
        Column col; col.name = first.cols[k].name;
        col.offsets = dev_alloc(c, ((size_t)nrows + 1) * 4);
        col.data = dev_alloc(c, total + 16);
        uint64_t row = 0, byte = 0;
        for (size_t p = 0; p < NPARTS; p++) {
            uint32_t b0 = ext[2 * (p * K + k)], b1 = ext[2 * (p * K + k) + 1];
            uint64_t pr = (uint64_t)parts[p]->nrows;
            KernelTimer kt(c, "concat", (uint64_t)(b1 - b0) * 2 + pr * 8);
            rebase_offsets_kernel<<<blocks_for(pr + 1, 256), 256, 0, c->stream>>>(parts[p]->cols[k].off(), col.offsets->as<uint32_t>() + row, pr, (uint32_t)byte);
            CPB_CUDA(cudaGetLastError());
            if (b1 > b0) CPB_CUDA(cudaMemcpyAsync(col.data->as<uint8_t>() + byte, parts[p]->cols[k].bytes() + b0, b1 - b0, cudaMemcpyDeviceToDevice, c->stream));
            row += pr; byte += b1 - b0;
        }
        r->cols.push_back(col);
    }
    return r;
}

}  // namespace cpb
