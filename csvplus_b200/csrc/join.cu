// join.cu — K8/K9 of SURVEY §2: the probe side of DataSource.Join / Except and Index.Find.
//
// Replaces (reference): Join csvplus.go:545-569 (per probe row: SelectValues, indexImpl.first =
// sort.Search lower bound :893-897 with cmp :907-920, then a forward scan while equal), mergeRows
// :571-583, Except :588-608 (indexImpl.has :899-905) and indexImpl.find :870-891.
//
// The index keeps its sorted key image (sort.cu).  For a join on the first nk key columns the
// distinct nk-column prefixes ("heads") of the sorted image are inserted into an open-addressing
// hash table that maps a prefix to its run [heads[j], heads[j+1]) of sorted rows — exactly the
// [lower_bound, upper_bound) range the reference finds by binary search, so output order (probe
// order, then index order) is unchanged.  Small tables are staged into shared memory per CTA; large
// ones are probed in L2/HBM.  Probe keys are packed with the index's column widths so equality of
// images is equality of keys.
#include <algorithm>

#include "core.hpp"
#include "util.cuh"

namespace cpb {

static inline uint32_t nblk(uint64_t n, int t) { return (uint32_t)((n + t - 1) / t); }
constexpr uint32_t EMPTY = 0xffffffffu;

__device__ __forceinline__ uint64_t mix64(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}
__device__ __forceinline__ uint64_t hash_prefix(const uint64_t* img, uint64_t n, uint64_t r, uint32_t pbytes) {
    uint32_t full = pbytes >> 3, rem = pbytes & 7;
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint32_t w = 0; w < full; w++) h = mix64(h ^ img[(uint64_t)w * n + r]);
    if (rem) h = mix64(h ^ (img[(uint64_t)full * n + r] & (~0ull << (8 * (8 - rem)))));
    return h;
}
__device__ __forceinline__ bool prefix_equal2(const uint64_t* a, uint64_t na, uint64_t ra, const uint64_t* b, uint64_t nb, uint64_t rb,
                                              uint32_t pbytes) {
    uint32_t full = pbytes >> 3, rem = pbytes & 7;
    for (uint32_t w = 0; w < full; w++) if (a[(uint64_t)w * na + ra] != b[(uint64_t)w * nb + rb]) return false;
    if (rem) {
        uint64_t m = ~0ull << (8 * (8 - rem));
        if ((a[(uint64_t)full * na + ra] & m) != (b[(uint64_t)full * nb + rb] & m)) return false;
    }
    return true;
}

__global__ void head_flags2_kernel(const uint64_t* __restrict__ image, uint64_t n, uint32_t pbytes, uint32_t* head) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || !prefix_equal2(image, n, i, image, n, i - 1, pbytes)) ? 1u : 0u;
}
__global__ void compact_heads_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, uint32_t* out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) out[pos[i]] = (uint32_t)i;
    if (i == 0) out[pos[n]] = (uint32_t)n;  // sentinel: heads[nheads] = nrows
}
__global__ void hash_insert_kernel(const uint64_t* __restrict__ image, uint64_t n, uint32_t pbytes, const uint32_t* __restrict__ heads,
                                   uint64_t nheads, uint32_t* slots, uint64_t mask) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nheads) return;
    uint64_t s = hash_prefix(image, n, heads[j], pbytes) & mask;
    for (;;) {
        uint32_t old = atomicCAS(&slots[s], EMPTY, (uint32_t)j);
        if (old == EMPTY) return;
        s = (s + 1) & mask;
    }
}

// total number of matches (= join result rows), one atomic per warp: flags layout [0..1] scan total, [2] not_one,
// [4..5] matches.  With it a join whose probe rows all match exactly once needs no scan of the counts at all.
__device__ __forceinline__ void add_matches(uint32_t* not_one, unsigned long long matches) {
    matches = warp_sum_u64(matches);
    if ((threadIdx.x & 31) == 0 && matches) atomicAdd(reinterpret_cast<unsigned long long*>(not_one + 2), matches);
}
// probe: per probe row the matching run of sorted index rows: lo[i], cnt[i] (cnt 0 = no match)
template <bool SMEM>
__global__ void __launch_bounds__(256) join_probe_kernel(const uint64_t* __restrict__ pimg, uint64_t np, const uint64_t* __restrict__ iimg,
                                                         uint64_t ni, uint32_t pbytes, const uint32_t* __restrict__ slots_g, uint64_t nslots,
                                                         const uint32_t* __restrict__ heads_g, uint64_t nheads, uint32_t* lo, uint32_t* cnt,
                                                         uint32_t* not_one) {
    extern __shared__ uint32_t sh[];
    const uint32_t* slots = slots_g;
    const uint32_t* heads = heads_g;
    if (SMEM) {  // stage the build-side table (slots + heads) into shared memory once per CTA
        for (uint64_t i = threadIdx.x; i < nslots; i += blockDim.x) sh[i] = slots_g[i];
        for (uint64_t i = threadIdx.x; i <= nheads; i += blockDim.x) sh[nslots + i] = heads_g[i];
        __syncthreads();
        slots = sh; heads = sh + nslots;
    }
    const uint64_t mask = nslots - 1;
    unsigned long long matches = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = hash_prefix(pimg, np, i, pbytes) & mask;
        uint32_t l = 0, c = 0;
        for (;;) {
            uint32_t j = slots[s];
            if (j == EMPTY) break;
            uint32_t r = heads[j];
            if (prefix_equal2(pimg, np, i, iimg, ni, r, pbytes)) { l = r; c = heads[j + 1] - r; break; }
            s = (s + 1) & mask;
        }
        lo[i] = l; cnt[i] = c; matches += c;
        if (__any_sync(__activemask(), c != 1) && c != 1) *not_one = 1u;  // benign race: every writer stores 1
    }
    add_matches(not_one, matches);
}

// ------------------------------------------------------------------ embedded-key table (key prefixes <= 24 bytes)
// One 32-byte slot = one DRAM sector: the whole key prefix (3 image words) + (first sorted row, run length).
// A probe is resolved by the sector it hashes to — no dependent reads of heads[] / the key image — and the probe
// key is packed in registers straight from the probe column (no materialised probe image).
struct __align__(32) Slot32 { unsigned long long k[3]; unsigned long long payload; };  // payload = lo | cnt << 32; 0 = empty

__device__ __forceinline__ uint64_t hash3(unsigned long long a, unsigned long long b, unsigned long long c) {
    return mix64(mix64(mix64(0x9E3779B97F4A7C15ull ^ a) ^ b) ^ c);
}
__device__ __forceinline__ uint64_t slot_of(uint64_t h, uint64_t nslots) { return __umul64hi(h, nslots); }

__global__ void hash32_insert_kernel(const uint64_t* __restrict__ image, uint64_t n, uint32_t pbytes, const uint32_t* __restrict__ heads,
                                     uint64_t nheads, Slot32* slots, uint64_t nslots) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nheads) return;
    const uint32_t r = heads[j], full = pbytes >> 3, rem = pbytes & 7;
    unsigned long long w[3] = {0, 0, 0};
    for (uint32_t i = 0; i < full; i++) w[i] = image[(uint64_t)i * n + r];
    if (rem) w[full] = image[(uint64_t)full * n + r] & (~0ull << (8 * (8 - rem)));
    const unsigned long long payload = (unsigned long long)r | ((unsigned long long)(heads[j + 1] - r) << 32);
    uint64_t s = slot_of(hash3(w[0], w[1], w[2]), nslots);
    for (;;) {
        if (atomicCAS(&slots[s].payload, 0ull, payload) == 0ull) { slots[s].k[0] = w[0]; slots[s].k[1] = w[1]; slots[s].k[2] = w[2]; return; }
        if (++s == nslots) s = 0;
    }
}

__device__ __forceinline__ unsigned long long bswap64(unsigned long long x) {
    return ((unsigned long long)__byte_perm((uint32_t)x, 0, 0x0123) << 32) | __byte_perm((uint32_t)(x >> 32), 0, 0x0123);
}
// the first min(len,16) bytes of a string, little-endian in (a, b), the rest zero: aligned 8-byte loads only (the
// probe rows of a warp are ~one sector apart, a byte loop costs one L1 request per byte)
__device__ __forceinline__ void load16(const uint8_t* p, uint32_t len, unsigned long long& a, unsigned long long& b) {
    a = b = 0;
    if (len == 0) return;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 7u), sh = mis * 8;
    const unsigned long long* wp = reinterpret_cast<const unsigned long long*>(p - mis);
    const uint32_t need = mis + (len < 16u ? len : 16u);
    const unsigned long long w0 = __ldg(wp), w1 = need > 8 ? __ldg(wp + 1) : 0ull, w2 = need > 16 ? __ldg(wp + 2) : 0ull;
    a = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
    b = sh ? (w1 >> sh) | (w2 << (64 - sh)) : w1;
    if (len < 8) { a &= (1ull << (8 * len)) - 1; b = 0; }
    else if (len < 16) b &= (1ull << (8 * (len - 8))) - 1;
}
// packs the probe key exactly like key_pack_kernel (sort.cu), but into three registers
__device__ __forceinline__ void pack3(const KeyDesc& kd, uint64_t r, unsigned long long& w0, unsigned long long& w1, unsigned long long& w2) {
    unsigned long long cur = 0; int cnt = 0, word = 0;
    w0 = w1 = w2 = 0;
    auto flush = [&](unsigned long long v) { if (word == 0) w0 = v; else if (word == 1) w1 = v; else w2 = v; word++; };
    auto push = [&](uint32_t b) { cur = (cur << 8) | b; if (++cnt == 8) { flush(cur); cnt = 0; cur = 0; } };
    for (int k = 0; k < kd.nkeys; k++) {
        const uint32_t s = kd.off[k][r], len = kd.off[k][r + 1] - s, wd = kd.width[k];
        const uint8_t* p = kd.data[k] + s;
        unsigned long long a, b;
        load16(p, len < wd ? len : wd, a, b);
        uint32_t i = 0;
        if (cnt == 0 && wd >= 8) { flush(bswap64(a)); i = 8; if (wd >= 16) { flush(bswap64(b)); i = 16; } }
        for (; i < wd; i++)
            push(i < 8 ? (uint32_t)(a >> (8 * i)) & 255u : i < 16 ? (uint32_t)(b >> (8 * (i - 8))) & 255u : (i < len ? p[i] : 0u));
        const uint32_t lf = len > wd ? 0xffffffffu : len;
        for (int i2 = (int)kd.lenbytes[k] - 1; i2 >= 0; i2--) push((lf >> (8 * i2)) & 0xffu);
    }
    if (cnt) flush(cur << (8 * (8 - cnt)));
}
__global__ void __launch_bounds__(256) join_probe32_kernel(KeyDesc kd, uint64_t np, const Slot32* __restrict__ slots, uint64_t nslots,
                                                           uint32_t* lo, uint32_t* cnt, uint32_t* not_one) {
    unsigned long long matches = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long w0, w1, w2;
        pack3(kd, i, w0, w1, w2);
        uint64_t s = slot_of(hash3(w0, w1, w2), nslots);
        uint32_t l = 0, c = 0;
        for (;;) {
            const ulonglong4 sl = *reinterpret_cast<const ulonglong4*>(&slots[s]);  // one 32-byte sector
            if (sl.w == 0ull) break;
            if (sl.x == w0 && sl.y == w1 && sl.z == w2) { l = (uint32_t)sl.w; c = (uint32_t)(sl.w >> 32); break; }
            if (++s == nslots) s = 0;
        }
        lo[i] = l; cnt[i] = c; matches += c;
        if (c != 1) *not_one = 1u;  // benign race: every writer stores 1
    }
    add_matches(not_one, matches);
}

// 16-byte slots for key prefixes <= 12 bytes whose runs all have length 1 (unique keys): four slots per 64-byte
// DRAM granule at load factor 1/2, so a probe costs about one granule (a 32-byte slot at 2/3 costs about two).
struct __align__(16) Slot16 { unsigned long long k0; uint32_t k1; uint32_t row1; };  // row1 = sorted row + 1; 0 = empty
// (a slot is claimed and filled by ONE 128-bit compare-and-swap — SASS ATOMG.E.CAS.128 — so a slot is never seen half
// written and an insert that meets its own key knows it: `dup`, when given, is set if any key is inserted twice.  That
// is the whole duplicate check of UniqueIndexOn for keys <= 12 bytes: the second inserter of a key walks the same probe
// sequence as the first and reaches its slot.)
__device__ __forceinline__ void cas128(void* addr, unsigned long long vl, unsigned long long vh, unsigned long long& ol, unsigned long long& oh) {
    asm volatile("{\n.reg .b128 c, v, o;\nmov.b128 c, {%3, %4};\nmov.b128 v, {%5, %6};\natom.global.cas.b128 o, [%2], c, v;\nmov.b128 {%0, %1}, o;\n}"
                 : "=l"(ol), "=l"(oh) : "l"(addr), "l"(0ull), "l"(0ull), "l"(vl), "l"(vh) : "memory");
}
__global__ void hash16_insert_kernel(const uint64_t* __restrict__ image, uint64_t n, uint32_t pbytes, const uint32_t* __restrict__ heads,
                                     uint64_t nheads, Slot16* slots, uint64_t nslots, uint32_t* dup) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nheads) return;
    const uint32_t r = heads[j], full = pbytes >> 3, rem = pbytes & 7;
    unsigned long long w[2] = {0, 0};
    for (uint32_t i = 0; i < full; i++) w[i] = image[(uint64_t)i * n + r];
    if (rem) w[full] = image[(uint64_t)full * n + r] & (~0ull << (8 * (8 - rem)));
    const unsigned long long lo = w[0], hi = (w[1] >> 32) | ((unsigned long long)(r + 1u) << 32);  // {k0, k1 | row1 << 32}
    uint64_t s = slot_of(hash3(w[0], w[1], 0ull), nslots);
    for (;;) {
        unsigned long long ol, oh;
        cas128(&slots[s], lo, hi, ol, oh);
        if ((ol | oh) == 0ull) return;
        if (dup && ol == lo && (uint32_t)oh == (uint32_t)hi) { *dup = 1u; return; }  // benign race: every writer stores 1
        if (++s == nslots) s = 0;
    }
}
__global__ void __launch_bounds__(256) join_probe16_kernel(KeyDesc kd, uint64_t np, const Slot16* __restrict__ slots, uint64_t nslots,
                                                           uint32_t* lo, uint32_t* cnt, uint32_t* not_one) {
    unsigned long long matches = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long w0, w1, w2;
        pack3(kd, i, w0, w1, w2);
        uint64_t s = slot_of(hash3(w0, w1, 0ull), nslots);
        uint32_t l = 0, c = 0;
        for (;;) {
            const uint4 sl = __ldg(reinterpret_cast<const uint4*>(&slots[s]));
            if (sl.w == 0u) break;
            if (sl.x == (uint32_t)w0 && sl.y == (uint32_t)(w0 >> 32) && sl.z == (uint32_t)(w1 >> 32)) { l = sl.w - 1u; c = 1u; break; }
            if (++s == nslots) s = 0;
        }
        lo[i] = l; cnt[i] = c; matches += c;
        if (c != 1) *not_one = 1u;  // benign race: every writer stores 1
    }
    add_matches(not_one, matches);
}

__global__ void expand_pairs_kernel(const uint32_t* __restrict__ lo, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ pos,
                                    uint32_t* pid, uint32_t* iid, uint64_t np) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    uint32_t c = cnt[i], p = pos[i], l = lo[i];
    for (uint32_t j = 0; j < c; j++) { pid[p + j] = (uint32_t)i; iid[p + j] = l + j; }
}
__global__ void zero_flag_kernel(const uint32_t* __restrict__ cnt, uint32_t* flag, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = cnt[i] == 0 ? 1u : 0u;
}
__global__ void compact_ids2_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, uint32_t* ids, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) ids[pos[i]] = (uint32_t)i;
}

static uint64_t read_u64(Ctx* c, const void* dev) {
    uint64_t* h = (uint64_t*)c->pinned_scratch(8);
    CPB_CUDA(cudaMemcpyAsync(h, dev, 8, cudaMemcpyDeviceToHost, c->stream));
    sync_stream(c);
    return *h;
}

__global__ void iota_heads_kernel(uint32_t* heads, uint64_t n) {  // heads[i] = i for i in [0, n]
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) heads[i] = (uint32_t)i;
}

// by_src: the table is built over the rows in SOURCE order (a lazily sorted unique index: payload = source row, key
// image = ix.uimage); otherwise over the sorted order (payload = sorted position, key image = ix.image).
static HashTable& ensure_hash(Ctx* c, Index& ix, int nk, bool by_src, uint32_t* dup_flag = nullptr) {
    std::lock_guard<std::mutex> lk(ix.mu);
    auto& hmap = by_src ? ix.hash_src : ix.hash;
    const Buf& image = by_src ? ix.uimage : ix.image;
    auto it = hmap.find(nk);
    if (it != hmap.end()) { wait_ready(c, it->second.ready); return it->second; }
    HashTable ht;
    ht.nkeys = nk; ht.pbytes = prefix_bytes(ix, nk);
    const uint64_t n = (uint64_t)ix.nrows;
    if (ix.unique && nk == (int)ix.key_cols.size()) {
        // a verified unique index probed on its full key: every sorted row is its own run — no adjacent compare, no scan,
        // no host round trip for the number of distinct keys
        ht.nheads = n;
        ht.heads = dev_alloc_owned(ix.ctx, c, (n + 1) * 4);
        iota_heads_kernel<<<nblk(n + 1, 256), 256, 0, c->stream>>>(ht.heads->as<uint32_t>(), n);
        CPB_CUDA(cudaGetLastError());
    } else {
        Buf head = dev_alloc(c, (n + 1) * 4), pos = dev_alloc(c, (n + 1) * 4), tot = dev_alloc(c, 8);
        {
            KernelTimer kt(c, "hash_heads", (uint64_t)ix.image_words * n * 8 + n * 4);
            head_flags2_kernel<<<nblk(n, 256), 256, 0, c->stream>>>(image->as<uint64_t>(), n, ht.pbytes, head->as<uint32_t>());
            CPB_CUDA(cudaGetLastError());
        }
        exclusive_scan_u32(c, head->as<uint32_t>(), pos->as<uint32_t>(), n, tot->as<uint64_t>());
        ht.nheads = read_u64(c, tot->p);
        // what stays in the index comes from the index owner's pool: the probing context may be shut down first
        ht.heads = dev_alloc_owned(ix.ctx, c, (ht.nheads + 1) * 4);
        compact_heads_kernel<<<nblk(n, 256), 256, 0, c->stream>>>(head->as<uint32_t>(), pos->as<uint32_t>(), ht.heads->as<uint32_t>(), n);
    }
    uint64_t want = std::max<uint64_t>(16, ht.nheads * 2);
    ht.nslots = 1; while (ht.nslots < want) ht.nslots <<= 1;
    // one of three probe tables: ordinal slots (+ heads + key image; shared-memory sized or wide keys), or the
    // embedded-key 16 / 32 byte slots
    const bool large = (ht.nslots + ht.nheads + 1) * 4 > 200 * 1024;
    if (!(large && ht.pbytes <= 24)) {
        ht.slots = dev_alloc_owned(ix.ctx, c, ht.nslots * 4);
        CPB_CUDA(cudaMemsetAsync(ht.slots->p, 0xff, ht.nslots * 4, c->stream));
        KernelTimer kt(c, "hash_build", ht.nheads * (ht.pbytes + 8));
        hash_insert_kernel<<<nblk(ht.nheads, 256), 256, 0, c->stream>>>(image->as<uint64_t>(), n, ht.pbytes, ht.heads->as<uint32_t>(), ht.nheads,
                                                                        ht.slots->as<uint32_t>(), ht.nslots - 1);
        CPB_CUDA(cudaGetLastError());
    }
    const bool too_large_for_smem = (ht.nslots + ht.nheads + 1) * 4 > 200 * 1024;
    if (ht.pbytes <= 12 && ht.nheads == n && too_large_for_smem) {  // unique short keys: 16-byte slots
        ht.nslots16 = ht.nheads * 2 + 16;
        ht.slots16 = dev_alloc_owned(ix.ctx, c, ht.nslots16 * sizeof(Slot16));
        CPB_CUDA(cudaMemsetAsync(ht.slots16->p, 0, ht.nslots16 * sizeof(Slot16), c->stream));
        KernelTimer kt(c, "hash_build", ht.nheads * (ht.pbytes + 8) + ht.nslots16 * 16);
        hash16_insert_kernel<<<nblk(ht.nheads, 256), 256, 0, c->stream>>>(image->as<uint64_t>(), n, ht.pbytes, ht.heads->as<uint32_t>(), ht.nheads,
                                                                          ht.slots16->as<Slot16>(), ht.nslots16, dup_flag);
        CPB_CUDA(cudaGetLastError());
    } else if (ht.pbytes <= 24 && too_large_for_smem) {  // key fits a sector
        ht.nslots32 = ht.nheads + ht.nheads / 2 + 16;
        ht.slots32 = dev_alloc_owned(ix.ctx, c, ht.nslots32 * sizeof(Slot32));
        CPB_CUDA(cudaMemsetAsync(ht.slots32->p, 0, ht.nslots32 * sizeof(Slot32), c->stream));
        KernelTimer kt(c, "hash_build", ht.nheads * (ht.pbytes + 8) + ht.nslots32 * 32);
        hash32_insert_kernel<<<nblk(ht.nheads, 256), 256, 0, c->stream>>>(image->as<uint64_t>(), n, ht.pbytes, ht.heads->as<uint32_t>(), ht.nheads,
                                                                          ht.slots32->as<Slot32>(), ht.nslots32);
        CPB_CUDA(cudaGetLastError());
    }
    ht.ready = record_ready(c);  // another context (stream) that finds it in the map waits for this event, not the host
    return hmap.emplace(nk, std::move(ht)).first->second;
}

// per probe row the matching run [lo, lo+cnt) of index rows, through whichever probe table the index has for nk columns
static void probe_dispatch(Ctx* c, const Table& probe, const std::vector<int>& pidx, Index& ix, HashTable& ht, const Buf& iimage, int nk,
                           uint32_t* lo, uint32_t* cnt, uint32_t* not_one) {
    const uint64_t np = (uint64_t)probe.nrows, ni = (uint64_t)ix.nrows;
    std::vector<uint32_t> widths(ix.key_width.begin(), ix.key_width.begin() + nk);
    uint32_t pwords = 0;
    Buf pimg;
    if (!ht.slots32 && !ht.slots16) pimg = pack_with_widths(c, probe, pidx, widths, &pwords);
    else pwords = (ht.pbytes + 7) / 8;
    // algorithmic bytes: probe keys once + build table once
    uint64_t algo = np * ((uint64_t)pwords * 8 + 8) + ht.nslots * 4 + ht.nheads * ((uint64_t)ht.pbytes + 4);
    size_t smem = (ht.nslots + ht.nheads + 1) * 4;
    KernelTimer kt(c, "join_probe", algo);
    if (ht.slots16) {
        KeyDesc kd{};
        describe_keys(c, probe, pidx, widths, kd);
        uint32_t grid = (uint32_t)std::min<uint64_t>(nblk(np, 256), (uint64_t)c->sm_count * 16);
        join_probe16_kernel<<<grid, 256, 0, c->stream>>>(kd, np, ht.slots16->as<Slot16>(), ht.nslots16, lo, cnt, not_one);
    } else if (ht.slots32) {
        KeyDesc kd{};
        describe_keys(c, probe, pidx, widths, kd);
        uint32_t grid = (uint32_t)std::min<uint64_t>(nblk(np, 256), (uint64_t)c->sm_count * 16);
        join_probe32_kernel<<<grid, 256, 0, c->stream>>>(kd, np, ht.slots32->as<Slot32>(), ht.nslots32, lo, cnt, not_one);
    } else if (smem <= 200 * 1024) {
        // (per device, and cheap: set on every launch rather than cached per process)
        CPB_CUDA(cudaFuncSetAttribute(join_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        uint32_t grid = (uint32_t)std::min<uint64_t>(nblk(np, 256), (uint64_t)c->sm_count);
        join_probe_kernel<true><<<grid, 256, smem, c->stream>>>(pimg->as<uint64_t>(), np, iimage->as<uint64_t>(), ni, ht.pbytes,
                                                                ht.slots->as<uint32_t>(), ht.nslots, ht.heads->as<uint32_t>(), ht.nheads, lo, cnt, not_one);
    } else {
        uint32_t grid = (uint32_t)std::min<uint64_t>(nblk(np, 256), (uint64_t)c->sm_count * 16);
        join_probe_kernel<false><<<grid, 256, 0, c->stream>>>(pimg->as<uint64_t>(), np, iimage->as<uint64_t>(), ni, ht.pbytes,
                                                              ht.slots->as<uint32_t>(), ht.nslots, ht.heads->as<uint32_t>(), ht.nheads, lo, cnt, not_one);
    }
    CPB_CUDA(cudaGetLastError());
}

std::shared_ptr<Table> join_tables(Ctx* c, const Table& probe, Index& ix, const std::vector<std::string>& cols, bool anti,
                                   DataError* derr, bool* failed) {
    *failed = false;
    const uint64_t np = (uint64_t)probe.nrows, ni = (uint64_t)ix.nrows;
    const int nk = (int)cols.size();
    std::vector<int> pidx;
    for (int k = 0; k < nk; k++) {
        int ci = probe.find(cols[k]);
        if (ci < 0 && np > 0) {  // row.SelectValues(columns...) fails on the first probe row (csvplus.go:556, :145)
            *failed = true;
            *derr = DataError{CPB_E_MISSING_COLUMN, k, probe.first_line, true, "missing column " + go_quote(cols[k])};
            return nullptr;
        }
        pidx.push_back(ci);
    }
    // output schema: mergeRows(indexRow, probeRow) — probe wins name collisions (csvplus.go:571-583)
    std::vector<const Column*> icols, pcols;
    std::vector<int> icol_idx;
    const Table& ischema = ix.schema();  // names only: the sorted columns may not be materialised
    if (!anti) for (size_t q = 0; q < ischema.cols.size(); q++)
        if (probe.find(ischema.cols[q].name) < 0) { icols.push_back(&ischema.cols[q]); icol_idx.push_back((int)q); }
    for (auto& col : probe.cols) pcols.push_back(&col);

    auto out = std::make_shared<Table>(); out->ctx = c; out->first_line = probe.first_line;
    auto empty_result = [&]() {
        Table e; e.ctx = c; e.nrows = 0;
        for (auto* p : icols) e.cols.push_back(*p);
        for (auto* p : pcols) e.cols.push_back(*p);
        return gather_rows(c, e, nullptr, 0);
    };
    if (np == 0) return empty_result();
    Buf lo = dev_alloc(c, np * 4), cnt = dev_alloc(c, (np + 1) * 4), flags = dev_alloc(c, 32);
    CPB_CUDA(cudaMemsetAsync(flags->p, 0, 32, c->stream));
    uint32_t* not_one = flags->as<uint32_t>() + 2;  // [0..1] = scan total
    // a lazily sorted unique index joined on its full key: probe table and row slots over the source order
    const bool by_src = ix.unique && ix.uimage && nk == (int)ix.key_cols.size();
    if (ni != 0 && !by_src) ensure_sorted(c, ix);
    const Buf& iimage = by_src ? ix.uimage : ix.image;
    if (ni == 0) {
        CPB_CUDA(cudaMemsetAsync(cnt->p, 0, (np + 1) * 4, c->stream));
    } else {
        HashTable& ht = ensure_hash(c, ix, nk, by_src);
        probe_dispatch(c, probe, pidx, ix, ht, iimage, nk, lo->as<uint32_t>(), cnt->as<uint32_t>(), not_one);
    }
    Buf tot = flags;
    if (anti) {
        Buf flag = dev_alloc(c, np * 4), pos = dev_alloc(c, (np + 1) * 4);
        zero_flag_kernel<<<nblk(np, 256), 256, 0, c->stream>>>(cnt->as<uint32_t>(), flag->as<uint32_t>(), np);
        exclusive_scan_u32(c, flag->as<uint32_t>(), pos->as<uint32_t>(), np, tot->as<uint64_t>());
        uint64_t m = read_u64(c, tot->p);
        Buf ids = dev_alloc(c, (m + 1) * 4);
        compact_ids2_kernel<<<nblk(np, 256), 256, 0, c->stream>>>(flag->as<uint32_t>(), pos->as<uint32_t>(), ids->as<uint32_t>(), np);
        CPB_CUDA(cudaGetLastError());
        auto r = gather_rows(c, probe, ids->as<uint32_t>(), (int64_t)m);
        r->first_line = probe.first_line;
        return r;
    }
    uint32_t* hf = (uint32_t*)c->pinned_scratch(32);
    CPB_CUDA(cudaMemcpyAsync(hf, flags->p, 32, cudaMemcpyDeviceToHost, c->stream));
    sync_stream(c);
    const uint64_t m = (uint64_t)hf[4] | ((uint64_t)hf[5] << 32);  // counted by the probe kernel (0 when ni == 0)
    // every probe row matched exactly one index row (the usual foreign-key join): the probe-side columns of the
    // result ARE the probe columns, in order — share their buffers instead of copying them; no scan of the counts
    const bool probe_identity = ni != 0 && hf[2] == 0 && m == np;
    if (m > 0xfffffffeull) throw DataError{CPB_E_TOO_LARGE, -1, 0, false, "join result exceeds 2^32-2 rows; probe in smaller batches"};
    if (m == 0) return empty_result();
    if (probe_identity) {
        auto gi = gather_index_rows(c, ix, icol_idx, lo->as<uint32_t>(), (int64_t)m, by_src);  // lo[i] is the single matching index row
        out->nrows = (int64_t)m;
        for (auto& col : gi->cols) out->cols.push_back(col);
        for (auto* p : pcols) out->cols.push_back(*p);
        return out;
    }
    Buf pos = dev_alloc(c, (np + 1) * 4);
    exclusive_scan_u32(c, cnt->as<uint32_t>(), pos->as<uint32_t>(), np, tot->as<uint64_t>());
    Buf pid = dev_alloc(c, m * 4), iid = dev_alloc(c, m * 4);
    {
        KernelTimer kt(c, "join_pairs", np * 12 + m * 8);
        expand_pairs_kernel<<<nblk(np, 256), 256, 0, c->stream>>>(lo->as<uint32_t>(), cnt->as<uint32_t>(), pos->as<uint32_t>(), pid->as<uint32_t>(),
                                                                  iid->as<uint32_t>(), np);
        CPB_CUDA(cudaGetLastError());
    }
    Table pt; pt.ctx = c; pt.nrows = (int64_t)np;
    for (auto* p : pcols) pt.cols.push_back(*p);
    auto gi = gather_index_rows(c, ix, icol_idx, iid->as<uint32_t>(), (int64_t)m, by_src);
    auto gp = gather_rows(c, pt, pid->as<uint32_t>(), (int64_t)m);
    out->nrows = (int64_t)m;
    for (auto& col : gi->cols) out->cols.push_back(col);
    for (auto& col : gp->cols) out->cols.push_back(col);
    return out;
}

// ------------------------------------------------------------------ UniqueIndexOn without sorting
// Every row probes the full-key table built over the rows themselves: with unique keys row r finds r.  Two rows with
// the same key occupy two slots of one probe sequence, so at least one of them finds the other: a duplicate exists
// iff some row does not find itself.
__global__ void self_check_kernel(const uint32_t* __restrict__ lo, const uint32_t* __restrict__ cnt, uint64_t n, uint32_t* flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (cnt[i] != 1u || lo[i] != (uint32_t)i)) *flag = 1u;  // benign race: every writer stores 1
}
bool index_has_duplicates(Ctx* c, Index& ix) {
    const uint64_t n = (uint64_t)ix.nrows;
    const int nk = (int)ix.key_cols.size();
    Buf flags = dev_alloc(c, 64);
    CPB_CUDA(cudaMemsetAsync(flags->p, 0, 64, c->stream));
    HashTable& ht = ensure_hash(c, ix, nk, true, flags->as<uint32_t>() + 8);
    if (ht.slots16) {  // the 16-byte-slot insert saw every key it met twice: no self probe needed
        uint32_t* hf = (uint32_t*)c->pinned_scratch(64);
        CPB_CUDA(cudaMemcpyAsync(hf, flags->p, 64, cudaMemcpyDeviceToHost, c->stream));
        sync_stream(c);
        return hf[8] != 0;
    }
    Buf lo = dev_alloc(c, n * 4), cnt = dev_alloc(c, (n + 1) * 4);
    probe_dispatch(c, *ix.src, ix.key_col_idx, ix, ht, ix.uimage, nk, lo->as<uint32_t>(), cnt->as<uint32_t>(), flags->as<uint32_t>() + 2);
    {
        KernelTimer kt(c, "unique_check", n * 8);
        self_check_kernel<<<nblk(n, 256), 256, 0, c->stream>>>(lo->as<uint32_t>(), cnt->as<uint32_t>(), n, flags->as<uint32_t>() + 8);
        CPB_CUDA(cudaGetLastError());
    }
    uint32_t* hf = (uint32_t*)c->pinned_scratch(64);
    CPB_CUDA(cudaMemcpyAsync(hf, flags->p, 64, cudaMemcpyDeviceToHost, c->stream));
    sync_stream(c);
    return hf[8] != 0;
}

// ------------------------------------------------------------------ Index.Find: [lower, upper) of a key prefix
__global__ void find_range_kernel(const uint64_t* __restrict__ image, uint64_t n, const uint64_t* __restrict__ val, uint32_t pbytes,
                                  unsigned long long* out) {
    // compare row r's first pbytes with val: <0, 0, >0
    auto cmp = [&](uint64_t r) {
        uint32_t full = pbytes >> 3, rem = pbytes & 7;
        for (uint32_t w = 0; w < full; w++) {
            uint64_t a = image[(uint64_t)w * n + r], b = val[w];
            if (a != b) return a < b ? -1 : 1;
        }
        if (rem) {
            uint64_t m = ~0ull << (8 * (8 - rem));
            uint64_t a = image[(uint64_t)full * n + r] & m, b = val[full] & m;
            if (a != b) return a < b ? -1 : 1;
        }
        return 0;
    };
    uint64_t i = 0, j = n;  // lower bound: first row >= val
    while (i < j) { uint64_t h = i + (j - i) / 2; if (cmp(h) < 0) i = h + 1; else j = h; }
    out[0] = i;
    j = n;                  // upper bound: first row > val
    while (i < j) { uint64_t h = i + (j - i) / 2; if (cmp(h) <= 0) i = h + 1; else j = h; }
    out[1] = i;
}

void find_range(Ctx* c, Index& ix, const std::vector<std::string>& values, int64_t* lo, int64_t* hi) {
    const uint64_t n = (uint64_t)ix.nrows;
    *lo = 0; *hi = 0;
    if (n == 0) return;
    ensure_sorted(c, ix);
    // pack the lookup values on the host with the index's widths (a handful of bytes)
    std::vector<uint8_t> img((size_t)ix.image_words * 8 + 8, 0);
    size_t b = 0;
    for (size_t k = 0; k < values.size(); k++) {
        uint32_t w = ix.key_width[k], lb = w < 255 ? 1 : (w < 65535 ? 2 : 4);
        for (uint32_t i = 0; i < w; i++) img[b++] = i < values[k].size() ? (uint8_t)values[k][i] : 0;
        uint32_t lf = values[k].size() > w ? 0xffffffffu : (uint32_t)values[k].size();
        for (int i = (int)lb - 1; i >= 0; i--) img[b++] = (lf >> (8 * i)) & 0xff;
    }
    std::vector<uint64_t> words(ix.image_words + 1, 0);
    for (size_t w = 0; w < words.size() && w * 8 < img.size(); w++)
        for (int i = 0; i < 8 && w * 8 + i < img.size(); i++) words[w] |= (uint64_t)img[w * 8 + i] << (8 * (7 - i));
    Buf dv = dev_alloc(c, words.size() * 8), out = dev_alloc(c, 16);
    CPB_CUDA(cudaMemcpyAsync(dv->p, words.data(), words.size() * 8, cudaMemcpyHostToDevice, c->stream));
    {
        KernelTimer kt(c, "index_find", 0);
        find_range_kernel<<<1, 1, 0, c->stream>>>(ix.image->as<uint64_t>(), n, dv->as<uint64_t>(), (uint32_t)b, (unsigned long long*)out->p);
        CPB_CUDA(cudaGetLastError());
    }
    uint64_t* h = (uint64_t*)c->pinned_scratch(16);
    CPB_CUDA(cudaMemcpyAsync(h, out->p, 16, cudaMemcpyDeviceToHost, c->stream));
    sync_stream(c);
    *lo = (int64_t)h[0]; *hi = (int64_t)h[1];
}

}  // namespace cpb
