// write.cu — K10 of SURVEY §2: DataSource.ToCsv (csvplus.go:379-406) with the default
// encoding/csv.Writer (comma ',', UseCRLF=false): header line, then one line per row
// with the named columns in caller order.  A field is quoted iff it is non-empty and (it is `\.` or
// contains , " \r \n or its first rune is unicode.IsSpace); inside quotes " is doubled.
// Two kernels: per-row output length -> exclusive scan -> per-row serialisation.
#include <algorithm>

#include "core.hpp"
#include "util.cuh"

namespace cpb {

constexpr int MAXW = 64;  // columns per ToCsv call
struct WriteCols { int n; const uint32_t* off[MAXW]; const uint8_t* data[MAXW]; };

__host__ __device__ inline bool first_rune_is_space(const uint8_t* p, uint32_t n) {
    uint8_t c = p[0];
    if (c < 0x80) return c == ' ' || (c >= '\t' && c <= '\r');
    if (c == 0xC2 && n >= 2) return p[1] == 0x85 || p[1] == 0xA0;
    if (c == 0xE1 && n >= 3) return p[1] == 0x9A && p[2] == 0x80;
    if (c == 0xE2 && n >= 3) {
        if (p[1] == 0x80) return (p[2] >= 0x80 && p[2] <= 0x8A) || p[2] == 0xA8 || p[2] == 0xA9 || p[2] == 0xAF;
        return p[1] == 0x81 && p[2] == 0x9F;
    }
    if (c == 0xE3 && n >= 3) return p[1] == 0x80 && p[2] == 0x80;
    return false;
}
// returns output length of one field; *quoted tells whether it is written in quotes
__host__ __device__ inline uint32_t field_out_len(const uint8_t* p, uint32_t n, bool* quoted) {
    *quoted = false;
    if (n == 0) return 0;
    uint32_t nq = 0; bool need = (n == 2 && p[0] == '\\' && p[1] == '.');
    for (uint32_t i = 0; i < n; i++) {
        uint8_t c = p[i];
        if (c == '"') { nq++; need = true; }
        else if (c == ',' || c == '\r' || c == '\n') need = true;
    }
    if (!need) need = first_rune_is_space(p, n);
    *quoted = need;
    return need ? n + nq + 2 : n;
}
__host__ __device__ inline uint32_t field_write(const uint8_t* p, uint32_t n, uint8_t* d) {
    bool q; uint32_t len = field_out_len(p, n, &q);
    if (!q) { for (uint32_t i = 0; i < n; i++) d[i] = p[i]; return len; }
    uint32_t k = 0;
    d[k++] = '"';
    for (uint32_t i = 0; i < n; i++) { if (p[i] == '"') d[k++] = '"'; d[k++] = p[i]; }
    d[k++] = '"';
    return k;
}

__global__ void csv_row_len_kernel(WriteCols wc, uint64_t n, uint32_t* len) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint32_t total = (uint32_t)wc.n;  // n-1 commas + newline
    for (int k = 0; k < wc.n; k++) {
        uint32_t s = wc.off[k][r], l = wc.off[k][r + 1] - s;
        bool q;
        total += field_out_len(wc.data[k] + s, l, &q);
    }
    len[r] = total;
}
// One warp serialises 32 consecutive rows.  Their output bytes are contiguous, so every lane writes its line into a
// per-warp shared-memory stage laid out like the destination and the warp stores the stage with aligned 16-byte
// vectors (a thread-per-row kernel storing bytes straight to HBM touches ~20 sectors per store instruction).
// Groups longer than the stage (very long rows) are written directly.
constexpr int CW_WARPS = 8;
constexpr int CW_STAGE = 4096;
__global__ void __launch_bounds__(CW_WARPS * 32) csv_row_write_kernel(WriteCols wc, uint64_t n, const uint32_t* __restrict__ pos, uint8_t* out) {
    __shared__ __align__(16) uint8_t stage_all[CW_WARPS][CW_STAGE + 16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* stage = stage_all[warp];
    const uint64_t nwarps = (uint64_t)gridDim.x * CW_WARPS;
    for (uint64_t g = (uint64_t)blockIdx.x * CW_WARPS + warp; g * 32 < n; g += nwarps) {
        const uint64_t r = g * 32 + lane;
        const bool valid = r < n;
        const uint32_t d = valid ? pos[r] : 0u;
        const uint64_t lastr = g * 32 + 31 < n ? g * 32 + 31 : n - 1;
        const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
        const uint32_t dend = pos[lastr + 1];  // pos has n+1 entries
        // the absolute destination is out + d; alignment is that of the pointer, not of the offset alone
        const uint32_t sh = (uint32_t)((reinterpret_cast<uintptr_t>(out) + d0) & 15u), total = dend - d0;
        const bool staged = sh + total <= CW_STAGE;
        if (valid) {
            uint8_t* q = staged ? stage + sh + (d - d0) : out + d;
            for (int k = 0; k < wc.n; k++) {
                if (k) *q++ = ',';
                const uint32_t s = wc.off[k][r], l = wc.off[k][r + 1] - s;
                q += field_write(wc.data[k] + s, l, q);
            }
            *q = '\n';
        }
        if (staged) {
            __syncwarp();
            uint8_t* gb = out + d0 - sh;
            for (uint32_t x = lane * 16; x < sh + total; x += 32 * 16) {
                if (x >= sh && x + 16 <= sh + total) *reinterpret_cast<uint4*>(gb + x) = *reinterpret_cast<const uint4*>(stage + x);
                else for (uint32_t y = x; y < x + 16; y++) if (y >= sh && y < sh + total) gb[y] = stage[y];
            }
            __syncwarp();
        }
    }
}

Buf table_to_csv(Ctx* c, const Table& t, const std::vector<int>& cols, const std::vector<std::string>& names, uint64_t* nbytes,
                 uint64_t* header_bytes) {
    if ((int)cols.size() > MAXW) throw ArgError{CPB_ERR_UNSUPPORTED, "more than 64 columns in ToCsv"};
    // header line (csvplus.go:387): the caller's column names through the same quoting rule
    std::string header;
    for (size_t i = 0; i < names.size(); i++) {
        if (i) header += ',';
        std::vector<uint8_t> tmp(names[i].size() * 2 + 2);
        uint32_t k = field_write((const uint8_t*)names[i].data(), (uint32_t)names[i].size(), tmp.data());
        header.append((const char*)tmp.data(), k);
    }
    header += '\n';
    const uint64_t n = (uint64_t)t.nrows;
    WriteCols wc{};
    wc.n = (int)cols.size();
    for (int k = 0; k < wc.n && n; k++) { wc.off[k] = t.cols[cols[k]].off(); wc.data[k] = t.cols[cols[k]].bytes(); }
    Buf len = dev_alloc(c, (n + 1) * 4), tot = dev_alloc(c, 8);
    uint64_t in_bytes = 0;
    if (n) {
        KernelTimer kt(c, "csv_row_len", n * 4 * (wc.n + 1));
        csv_row_len_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, c->stream>>>(wc, n, len->as<uint32_t>());
        CPB_CUDA(cudaGetLastError());
    }
    exclusive_scan_u32(c, len->as<uint32_t>(), len->as<uint32_t>(), n, tot->as<uint64_t>());
    uint64_t* ht = (uint64_t*)c->pinned_scratch(8);
    CPB_CUDA(cudaMemcpyAsync(ht, tot->p, 8, cudaMemcpyDeviceToHost, c->stream));
    sync_stream(c);
    const uint64_t body = *ht;
    if (body > 0xffffffffull) throw DataError{CPB_E_TOO_LARGE, -1, 0, false, "ToCsv output exceeds 4 GiB; write in smaller batches"};
    Buf out = dev_alloc(c, header.size() + body + 16);
    CPB_CUDA(cudaMemcpyAsync(out->p, header.data(), header.size(), cudaMemcpyHostToDevice, c->stream));
    if (n) {
        (void)in_bytes;
        KernelTimer kt(c, "csv_write", 2 * body + n * 4 * (wc.n + 1));
        const uint32_t wblocks = (uint32_t)std::min<uint64_t>((n + CW_WARPS * 32 - 1) / (CW_WARPS * 32), (uint64_t)c->sm_count * 16);
        csv_row_write_kernel<<<wblocks, CW_WARPS * 32, 0, c->stream>>>(wc, n, len->as<uint32_t>(), out->as<uint8_t>() + header.size());
        CPB_CUDA(cudaGetLastError());
    }
    *nbytes = header.size() + body;
    if (header_bytes) *header_bytes = header.size();
    return out;
}

}  // namespace cpb
