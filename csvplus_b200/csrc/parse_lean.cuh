// parse_lean.cuh — csv_scan_lean_kernel: the same fused scan as csv_scan_kernel (same tiles, same staging, same two
// look-back chains, same outputs), with a lean path for *regular* tiles in front of the general tile body.
//
// Replaces (reference): the hot loop of Reader.Iterate csvplus.go:1117-1138 with the field-count policy of
// csvplus.go:1058-1076 in force (every record has exactly expect_fields fields — the default) and encoding/csv's
// readRecord on lines that hold no quote.
//
// Why a second kernel: csv_scan_kernel is instruction-bound (≈ 0.9 warp-instructions per input byte, ≈ 25 block barriers
// per tile).  A tile is *regular* when its window holds no quote, every line that starts in it has exactly NF fields and
// ends inside the window.  For such a tile
//   * each WARP owns 4 KiB of the tile (+16 B look-behind, +496 B look-ahead) and works on it alone: classification,
//     structural index, field extents, Like terms and the staged output all synchronise with __syncwarp only;
//   * the newline bitmap never reaches shared memory (three warp reductions give the line count, the first line start and
//     the terminator the last line must have), there is no terminator-ordinal array: field j of line i is structural
//     s0 + NF*i + j of the flat index by arithmetic;
//   * regularity is *verified*, not assumed: every line's NF-th structural must be a newline and the last line must end
//     at the first newline of the look-ahead, which together leave no room for a line with another field count;
//   * the CTA meets at three barriers per tile (ticket, totals, prefix) instead of ≈ 25.
// Anything else — a quote anywhere in the window, a line that is too long, the header tile, the tile that sees the end of
// the input, a tile entered inside a quoted field, any error — makes the whole tile take general_tile(), out of line,
// exactly as csv_scan_kernel runs it, so results (rows, offsets, bytes, first error, totals) are those of that kernel.
#pragma once
#include "parse_kernels.cuh"

namespace cpb {

constexpr int LN_NW = THREADS / 32;                 // warps per CTA (8)
constexpr int LN_SUB = TILE / LN_NW;                // bytes of a tile one warp owns (4096)
constexpr int LN_ROUNDS = (PRE + LN_SUB) / 512 + 1; // classification rounds of 32 lanes x 16 bytes (9)
constexpr int LN_WLEN = LN_ROUNDS * 512;            // bytes of a warp's window (4608)
constexpr int LN_WWORDS = LN_WLEN / 32;             // 144
constexpr int LN_MAINW = 128;                       // bitmap words expanded 4 per lane; the rest one per lane
constexpr int LN_LOOK = LN_WLEN - PRE - LN_SUB;     // look-ahead bytes (496)
static_assert(LN_SUB == 4096 && LN_ROUNDS == 9 && LN_WWORDS - LN_MAINW <= 32, "lean path layout");
static_assert(LN_LOOK <= HALO, "the last warp's window stays inside the staged window");
constexpr uint32_t LN_REGION = 3 * (WIN_WORDS + 4) * 4 + SCAP * 2 + LCAP * 2;  // Tb|Sb|Qb|sidx|ord of ParseSmem, contiguous
constexpr uint32_t LN_WS = (LN_REGION / LN_NW) & ~15u;                          // scratch bytes per warp (4192)
constexpr uint32_t LN_SB_BYTES = (LN_WWORDS + 4) * 4;                           // structural bitmap of a warp window
constexpr int LN_SCAP = (int)((LN_WS - LN_SB_BYTES) / 2);                       // structural index capacity per warp (1800)
constexpr int LN_PK = (1 + 8 + 1) / 2;                                          // packed 16-bit counters: rows, bytes[k<=8]

struct LeanExtra {
    uint32_t wtot[LN_NW][LN_PK];  // per warp: counter 2j | counter 2j+1 << 16 (0 = rows, 1+k = bytes of slot k)
    uint32_t wpre[LN_NW][LN_PK];  // the same, exclusive over the warps of the tile
    uint32_t wrec[LN_NW], wrecpre[LN_NW];  // records (lines) per warp
    uint32_t irregular;  // some warp declined the lean path
    uint32_t mode;       // 0 lean emit, 1 general tile
    uint32_t store_ok;   // the tile's rows / bytes fit the output capacities
    uint32_t pad;
};

// quote presence accumulates without the final mask: (x - 0x01..) & ~x has bit 7 of a byte set iff ... (exact as an any-test
// once masked with 0x80808080)
__device__ __forceinline__ uint32_t quote_acc(uint32_t acc, uint32_t w) {
    const uint32_t x = w ^ 0x22222222u;
    return acc | ((x - 0x01010101u) & ~x);
}

#ifdef CPB_LEAN_DP4A
// 16 flag bits via four byte dot products: flags are 0x80 per byte, weights 1,2,4,8 (16,32,64,128): sum = nibble << 7
__device__ __forceinline__ uint32_t lean_flags16(uint32_t z0, uint32_t z1, uint32_t z2, uint32_t z3) {
    const uint32_t lo = __dp4a(z1, 0x80402010u, __dp4a(z0, 0x08040201u, 0u));
    const uint32_t hi = __dp4a(z3, 0x80402010u, __dp4a(z2, 0x08040201u, 0u));
    return (lo >> 7) | (hi << 1);
}
#else
__device__ __forceinline__ uint32_t lean_flags16(uint32_t z0, uint32_t z1, uint32_t z2, uint32_t z3) { return flags16(z0, z1, z2, z3); }
#endif

template <int KMAX, bool HP>
static __device__ __noinline__ void general_tile_cold(const ParseParams& P, ParseSmem& sm, uint32_t tile, const uint8_t* lits, bool lits_in_smem) {
    general_tile<KMAX, true, HP>(P, sm, tile, lits, lits_in_smem);
}

// Warp 0, once per tile, after every warp's totals are in shared memory: exclusive prefixes over the warps, both
// look-back chains, the capacity check.  Out of line so that its registers (the look-back holds 2 + KMAX 64-bit words per
// lane) are not the lean path's.
template <int KMAX>
static __device__ __noinline__ void lean_chain(const ParseParams& P, ParseSmem& sm, LeanExtra& lx, const uint32_t tile) {
    constexpr int NP = 2 + KMAX;
    constexpr int PK = (1 + KMAX + 1) / 2;
    const int lane = threadIdx.x & 31;
    if (lx.irregular) { if (lane == 0) lx.mode = 1; return; }
    // exclusive prefixes over the warps; lane j < PK walks packed word j, lane PK the record counts
    uint32_t run = 0;
    if (lane < PK) {
#pragma unroll
        for (int w = 0; w < LN_NW; w++) { const uint32_t t = lx.wtot[w][lane]; lx.wpre[w][lane] = run; run += t; }
    } else if (lane == PK) {
#pragma unroll
        for (int w = 0; w < LN_NW; w++) { const uint32_t t = lx.wrec[w]; lx.wrecpre[w] = run; run += t; }
    }
    // component c of the tile's totals on lane c: 0 records, 1 rows, 2+k bytes of slot k
    const int cv = lane >= 1 ? lane - 1 : 0;
    const uint32_t tsrc = __shfl_sync(0xffffffffu, run, lane == 0 ? PK : (cv >> 1));
    const unsigned long long mine = lane == 0 ? tsrc : ((cv & 1) ? (tsrc >> 16) : (tsrc & 0xffffu));
    // chain 1: this tile holds no quote (parity 0); it must also *start* outside quotes
    if (lane == 0) st_release_u32(&P.st1[tile], 1u);
    const uint32_t pin = lookback_parity_w0(P.st1, tile);
    if (pin) { if (lane == 0) lx.mode = 1; return; }
    if (lane == 0) st_release_u32(&P.st1[tile], 2u);
    // chain 2: totals
    unsigned long long* wt = P.words + (uint64_t)tile * NP;
    if (lane < NP) st_relaxed_u64((uint64_t*)(wt + lane), LB_AGG | mine);
    lookback_totals_w0<KMAX>(P.words, tile, NP, sm);
    bool ok = true;
    if (lane < NP) {
        const unsigned long long incl = sm.tile_prefix[lane] + mine;
        st_relaxed_u64((uint64_t*)(wt + lane), LB_INCL | incl);
        if (lane == 1) ok = incl <= P.row_cap;
#pragma unroll
        for (int k = 0; k < KMAX; k++) if (lane == 2 + k) ok = incl <= P.data_cap[k];
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) { lx.mode = 0; lx.store_ok = ok ? 1u : 0u; }
}

#ifndef CPB_LEAN_CTAS
#define CPB_LEAN_CTAS (768 / CPB_THREADS)
#endif
template <int KMAX, bool HP>
__global__ void __launch_bounds__(THREADS, CPB_LEAN_CTAS) csv_scan_lean_kernel(const __grid_constant__ ParseParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ParseSmem& sm = *reinterpret_cast<ParseSmem*>(smem_raw);
    LeanExtra& lx = *reinterpret_cast<LeanExtra*>(smem_raw + ((sizeof(ParseSmem) + 15) & ~size_t(15)));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NP = 2 + KMAX;
    constexpr int PK = (1 + KMAX + 1) / 2;  // packed counter words in use
    constexpr int RC = KMAX <= 2 ? 6 : (KMAX <= 4 ? 4 : 3);  // lines per lane whose field extents stay in registers
    const uint32_t NL4 = 0x0a0a0a0au, D4 = P.delim * 0x01010101u;
    const int NF = P.expect_fields;

    if (tid == 0) { mbar_init(&sm.mbar, 1); fence_mbar_init(); }
    const bool lits_in_smem = P.lits_len <= LITS_SMEM;
    if (lits_in_smem) for (uint32_t i = tid; i < P.lits_len; i += THREADS) sm.lits[i] = P.lits[i];
    const uint8_t* lits = lits_in_smem ? sm.lits : P.lits;
    __syncthreads();
    uint32_t phase = 0;

    // this warp's window: byte b of it is wdata[b]; b = PRE is the first byte of the warp's 4 KiB
    const uint8_t* const wdata = sm.data + warp * LN_SUB;
    uint8_t* const ws = reinterpret_cast<uint8_t*>(sm.Tb) + (uint32_t)warp * LN_WS;
    uint32_t* const Sbw = reinterpret_cast<uint32_t*>(ws);
    uint16_t* const S16 = reinterpret_cast<uint16_t*>(ws);
    uint16_t* const sidx = reinterpret_cast<uint16_t*>(ws + LN_SB_BYTES);
    uint8_t* const stage = ws;  // the bitmap and the index are dead once the field extents sit in registers

    for (;;) {
        if (tid == 0) { sm.ticket = atomicAdd(P.ticket, 1u); lx.irregular = 0; }
        __syncthreads();
        const uint32_t tile = sm.ticket;
        if (tile >= P.ntiles) break;
        const uint64_t tile_base = (uint64_t)tile * TILE;
        // ---- stage the window [tile_base-PRE, tile_base+WIN) with one bulk copy (as csv_scan_kernel)
        const uint64_t w_lo = tile_base >= PRE ? tile_base - PRE : 0;
        const uint64_t n16 = (P.n + 15) & ~15ull;
        const uint64_t w_hi = tile_base + WIN < n16 ? tile_base + WIN : n16;
        const uint32_t lead = (uint32_t)(PRE - (tile_base - w_lo));
        const uint32_t nbytes = (uint32_t)(w_hi - w_lo);
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(&sm.mbar, nbytes);
            bulk_g2s(sm.data + lead, P.in + w_lo, nbytes, &sm.mbar);
            if (P.l2_ahead) {
                const uint64_t pf = tile_base + (uint64_t)gridDim.x * TILE;
                if (pf + TILE <= (P.n & ~15ull)) bulk_prefetch_l2(P.in + pf, TILE);
            }
        }
        mbar_wait(&sm.mbar, phase);
        phase ^= 1;

        // the lean path needs a full window of data lines: not the first tile, not a tile that sees the end of the input,
        // not a tile at or before the end of the header row
        const bool try_lean = tile != 0 && tile_base + WIN <= P.n && P.data_start <= tile_base;
        if (!try_lean) {
            if (tid == 0) atomicAdd(&P.result->general_tiles, 1u);
            general_tile_cold<KMAX, HP>(P, sm, tile, lits, lits_in_smem);
            __syncthreads();
            continue;
        }

        // ================================================================== lean path, one warp per 4 KiB
        bool irr = false;
        uint32_t nlines, q0, qh, s0, nstruct;
        {
            // ---- classify: structural bitmap to shared memory; newlines only counted / located
            const uint4* w4 = reinterpret_cast<const uint4*>(wdata);
            uint32_t anyq = 0, nlcnt = 0, nl_first = 0, nl_last = 0, s_first = 0;
#pragma unroll
            for (int r = 0; r < LN_ROUNDS; r++) {
                const int c = r * 32 + lane;
                const uint4 x = w4[c];
                const uint32_t nl = lean_flags16(eq_flags(x.x, NL4), eq_flags(x.y, NL4), eq_flags(x.z, NL4), eq_flags(x.w, NL4));
                const uint32_t dl = lean_flags16(eq_flags(x.x, D4), eq_flags(x.y, D4), eq_flags(x.z, D4), eq_flags(x.w, D4));
                S16[c] = (uint16_t)(dl | nl);
                anyq = quote_acc(quote_acc(quote_acc(quote_acc(anyq, x.x), x.y), x.z), x.w);
                if (r == 0) { nl_first = nl; s_first = dl | nl; }
                else if (r == LN_ROUNDS - 1) nl_last = nl;
                else nlcnt += __popc(nl);
            }
            if (lane < 4) Sbw[LN_WWORDS + lane] = 0;
            // newlines that start a line of this warp: window bits [PRE-1, PRE-1+SUB) = [15, 4111)
            const uint32_t mf = nl_first & (lane == 0 ? 0x8000u : 0xffffu);
            const uint32_t ml = nl_last & (lane == 0 ? 0x8000u : 0xffffu);   // newlines of the look-ahead: bits >= 4111
            nlcnt += __popc(mf) + (lane == 0 ? __popc(nl_last & 0x7fffu) : 0u);
            nlines = __reduce_add_sync(0xffffffffu, nlcnt);
            q0 = __reduce_min_sync(0xffffffffu, mf ? (uint32_t)(lane * 16 + __ffs(mf) - 1) : 0xffffu);
            qh = __reduce_min_sync(0xffffffffu, ml ? (uint32_t)((LN_ROUNDS - 1) * 512 + lane * 16 + __ffs(ml) - 1) : 0xffffu);
            const bool hasq = __any_sync(0xffffffffu, (anyq & 0x80808080u) != 0);
            // structurals below the first line start: ordinal of that newline + 1
            const uint32_t cq = q0 >> 4;
            const uint32_t below = (uint32_t)lane < cq ? __popc(s_first) : ((uint32_t)lane == cq ? __popc(s_first & ((1u << (q0 & 15u)) - 1u)) : 0u);
            s0 = 1u + __reduce_add_sync(0xffffffffu, below);
            irr = hasq || q0 == 0xffffu || qh == 0xffffu;
            __syncwarp();
        }
        const uint32_t Lpl = (nlines + 31u) >> 5;  // lines per lane (blocked: lane l owns lines [l*Lpl, (l+1)*Lpl))
        if (Lpl > (uint32_t)RC) irr = true;
        {
            // ---- flat structural index of the window
            const uint4 sw = reinterpret_cast<const uint4*>(Sbw)[lane];
            const uint32_t hs = lane < LN_WWORDS - LN_MAINW ? Sbw[LN_MAINW + lane] : 0u;
            const uint32_t v = (uint32_t)(__popc(sw.x) + __popc(sw.y) + __popc(sw.z) + __popc(sw.w)) | ((uint32_t)__popc(hs) << 16);
            const uint32_t inc = warp_incl_scan(v);
            const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
            nstruct = (tot & 0xffffu) + (tot >> 16);
            if (nstruct > (uint32_t)LN_SCAP) irr = true;
            if (!irr) {
                uint32_t o = (inc - v) & 0xffffu;
                const uint32_t sws[4] = {sw.x, sw.y, sw.z, sw.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t m = sws[j];
                    const uint32_t pos0 = (uint32_t)(lane * 128 + j * 32);
                    while (m) {
                        const int b = __ffs(m) - 1; m &= m - 1;
                        sidx[o++] = (uint16_t)(pos0 + b);
                    }
                }
                uint32_t o2 = (tot & 0xffffu) + ((inc - v) >> 16);
                uint32_t m = hs;
                const uint32_t pos0 = (uint32_t)(LN_MAINW * 32 + lane * 32);
                while (m) {
                    const int b = __ffs(m) - 1; m &= m - 1;
                    sidx[o2++] = (uint16_t)(pos0 + b);
                }
            }
            __syncwarp();
        }

        // ---- lines: field extents, Like terms, counts
        uint32_t cf[RC][KMAX];
        uint32_t cmask = 0, nrow = 0, cb[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; k++) cb[k] = 0;
        if (!irr) {
#pragma unroll
            for (int q = 0; q < RC; q++) {
#pragma unroll
                for (int k = 0; k < KMAX; k++) cf[q][k] = 0;
                const uint32_t i = (uint32_t)lane * Lpl + q;
                if ((uint32_t)q < Lpl && i < nlines) {
                    const uint32_t sA = s0 + (uint32_t)NF * i;  // ordinal of the line's first structural
                    const uint32_t sE = sA + (uint32_t)NF - 1;  // its terminator, if the line is regular
                    if (sE >= nstruct) { irr = true; continue; }
                    const uint32_t start = (uint32_t)sidx[sA - 1] + 1u;
                    const uint32_t e_nl = sidx[sE];
                    if (wdata[e_nl] != '\n' || (i == nlines - 1 && e_nl != qh)) { irr = true; continue; }
                    uint32_t e = e_nl;
                    if (wdata[e - 1] == '\r') e--;           // "\r\n" -> "\n"; e_nl > start - 1 >= PRE - 1 so the byte exists
                    if (e <= start && NF == 1) { irr = true; continue; }  // an empty line is not a record
                    uint32_t eq = 0;
#pragma unroll
                    for (int k = 0; k < KMAX; k++) {
                        const uint32_t t = (uint32_t)P.sel_field[k];
                        const uint32_t fb = t == 0 ? start : (uint32_t)sidx[sA + t - 1] + 1u;
                        const uint32_t fe = t + 1 == (uint32_t)NF ? e : (uint32_t)sidx[sA + t];
                        const uint32_t len = fe - fb;
                        cf[q][k] = fb | (len << 16);
                        uint32_t tm = HP ? P.slot_terms[k] : 0u;
                        while (tm) {
                            const int tt = __ffs(tm) - 1; tm &= tm - 1;
                            if (len == P.pred.term_len[tt]) {
                                const uint32_t doff = (uint32_t)(wdata - sm.data) + fb;
                                const bool same = lits_in_smem ? field_eq_smem(sm.data, doff, sm.lits, P.pred.term_off[tt], len)
                                                               : bytes_eq(wdata + fb, lits + P.pred.term_off[tt], len);
                                if (same) eq |= 1u << tt;
                            }
                        }
                    }
                    if (!HP || eval_pred(P.pred, eq)) {
                        cmask |= 1u << q;
                        nrow++;
#pragma unroll
                        for (int k = 0; k < KMAX; k++) cb[k] += cf[q][k] >> 16;
                    }
                }
            }
        }
        // ---- warp scan of the packed counters; totals to shared memory
        uint32_t pv[PK], pinc[PK];
#pragma unroll
        for (int j = 0; j < PK; j++) {
            const uint32_t lo = j == 0 ? nrow : cb[2 * j - 1];
            const uint32_t hi = 2 * j < KMAX ? cb[2 * j] : 0u;
            pv[j] = lo | (hi << 16);
            pinc[j] = warp_incl_scan(pv[j]);
        }
        if (__any_sync(0xffffffffu, irr)) { if (lane == 0) lx.irregular = 1; }
        if (lane == 31) {
#pragma unroll
            for (int j = 0; j < PK; j++) lx.wtot[warp][j] = pinc[j];
            lx.wrec[warp] = nlines;
        }
        __syncthreads();  // (1) every warp's totals

        if (warp == 0) lean_chain<KMAX>(P, sm, lx, tile);
        __syncthreads();  // (2) the tile's prefix

        if (lx.mode != 0) {
            if (tid == 0) atomicAdd(&P.result->general_tiles, 1u);
            general_tile_cold<KMAX, HP>(P, sm, tile, lits, lits_in_smem);
        } else if (lx.store_ok) {
            // ================================================================== lean emit
            uint32_t ex[PK];
#pragma unroll
            for (int j = 0; j < PK; j++) ex[j] = pinc[j] - pv[j] + lx.wpre[warp][j];  // exclusive inside the tile, packed
            const uint64_t row0 = sm.tile_prefix[1] + (ex[0] & 0xffffu);
            if (nrow != 0 && row0 == 0)
                P.result->first_row_ordinal = sm.tile_prefix[0] + lx.wrecpre[warp] + (uint32_t)lane * Lpl + (uint32_t)(__ffs(cmask) - 1);
#pragma unroll
            for (int k = 0; k < KMAX; k++) {
                const int cj = (1 + k) >> 1, ch = (1 + k) & 1;
                const uint32_t lane_ex = ch ? (ex[cj] >> 16) : (ex[cj] & 0xffffu);               // bytes of slot k before this lane, in the tile
                const uint32_t wpre_k = ch ? (lx.wpre[warp][cj] >> 16) : (lx.wpre[warp][cj] & 0xffffu);  // ... before this warp
                const uint32_t wt_k = __shfl_sync(0xffffffffu, pinc[cj], 31);
                const uint32_t B = ch ? (wt_k >> 16) : (wt_k & 0xffffu);                         // bytes of slot k of this warp
                const uint64_t dtile = sm.tile_prefix[2 + k];
                // ---- offsets of this lane's rows
                {
                    uint32_t run = (uint32_t)(dtile + lane_ex);
                    uint32_t* po = P.out_off[k] + row0;
#pragma unroll
                    for (int q = 0; q < RC; q++)
                        if ((cmask >> q) & 1) { *po++ = run; run += cf[q][k] >> 16; }
                }
                // ---- field bytes: rows -> stage (shifted so that stage byte x belongs at gbase[x]) -> aligned 16-byte stores
                const uint64_t dw = dtile + wpre_k;
                const uint32_t r16 = (uint32_t)(dw & 15);
                uint8_t* gbase = P.out_data[k] + (dw - r16);
                const uint32_t hi_ok = r16 + B;
                for (uint32_t c0 = 0; c0 < hi_ok; c0 += LN_WS) {
                    uint32_t st = r16 + (lane_ex - wpre_k);
#pragma unroll
                    for (int q = 0; q < RC; q++) {
                        if ((cmask >> q) & 1) {
                            const uint32_t f = cf[q][k], len = f >> 16;
                            const uint32_t lo = st > c0 ? st : c0, hi = st + len < c0 + LN_WS ? st + len : c0 + LN_WS;
                            const uint8_t* sp = wdata + (f & 0xffffu) - st;
                            for (uint32_t x = lo; x < hi; x++) stage[x - c0] = sp[x];
                            st += len;
                        }
                    }
                    __syncwarp();
                    const uint32_t cend = c0 + LN_WS < hi_ok ? c0 + LN_WS : hi_ok;
                    for (uint32_t x0 = c0 + lane * 16; x0 < cend; x0 += 32 * 16)
                        if (x0 >= r16 && x0 + 16 <= hi_ok) *reinterpret_cast<uint4*>(gbase + x0) = *reinterpret_cast<const uint4*>(stage + (x0 - c0));
                    {   // the (at most two) partial vectors at the first and last byte of the warp's range: one byte per lane
                        const uint32_t tv0 = hi_ok & ~15u;
                        const uint32_t x = lane < 16 ? (uint32_t)lane : tv0 + (uint32_t)(lane - 16);
                        const bool head = lane < 16 && c0 == 0 && r16 != 0;
                        const bool tail = lane >= 16 && (hi_ok & 15u) != 0 && tv0 >= c0 && tv0 < cend && !(tv0 == 0 && r16 != 0);
                        if ((head || tail) && x >= r16 && x < hi_ok) gbase[x] = stage[x - c0];
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();  // (3) smem is reused by the next tile
    }
}

}  // namespace cpb
