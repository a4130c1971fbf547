// plain.cu -- the .plain text path on the device (compressPlain :1246-1297, decompressPlain
// :1299-1335, convertBinToPlain :1414-1465, convertPlainToBin :1467-1533 of compress_file.cpp).
//
// Text in:   k_find_records<0/1>   byte-parallel search for the "e" terminator lines, listed in
//                                  file order (count pass, scan, write pass)
//            k_parse_records       one thread per record walks its lines backwards (the nearest
//                                  line of each key wins, which is exactly "fields persist across
//                                  records", :1254), parses FEN / UCI / integers into an Entry
//            k_entries_link_encode the same link + encode step as the .bin compressor, fed from
//                                  Entries; the payload scan / chunk orbit / emission are shared
//            k_entries_to_bin      SfenPacker::pack of every Entry (.plain -> .bin)
// Text out:  k_bin_text<0/1>       per .bin record: decode, then size pass / write pass of
//                                  emitPlainEntry (:1216-1237)
//            k_chain_text          per chain of a binpack (see decompress.cu for how chains are found)
//
// Supported layout of .plain input: what the reference tokeniser (`>> key`, `>> std::ws`, getline)
// reads the same way line by line -- every non-blank line is either `e` or `<key> <value>`; keys in
// any order, unknown keys ignored, missing keys inherited from earlier records. Layouts where the
// stream semantics differ from the line structure (a key whose value is on the next line, tokens
// after `e` on its line) are rejected with NNP_ERR_BAD_TEXT.
#include "common.cuh"
#include "kernels.h"
#include "link.cuh"
#include "text.cuh"

namespace nnp {

// ------------------------------------------------------------------ large exclusive sum (u32 -> u64)

constexpr int LSUM_THREADS = 256;
constexpr int LSUM_ITEMS = 16;
constexpr int LSUM_TILE = LSUM_THREADS * LSUM_ITEMS;

__device__ __forceinline__ u32 warp_inclusive_sum(u32 v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}
template <int THREADS>
__device__ __forceinline__ u32 block_excl_sum(u32 v, u32& total, u32* warp_tot)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u32 inc = warp_inclusive_sum(v);
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    u32 wprefix = 0, all = 0;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) {
        if (i == wid) wprefix = all;
        all += warp_tot[i];
    }
    total = all;
    __syncthreads();
    return wprefix + inc - v;
}

__global__ void __launch_bounds__(LSUM_THREADS) k_tile_sums(const u32* __restrict__ in, u64 n, u32* __restrict__ tile_sum)
{
    __shared__ u32 warp_tot[LSUM_THREADS / 32];
    const u64 base = (u64)blockIdx.x * LSUM_TILE;
    u32 s = 0;
#pragma unroll
    for (int j = 0; j < LSUM_ITEMS; ++j) {
        const u64 i = base + (u64)j * LSUM_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    u32 total;
    block_excl_sum<LSUM_THREADS>(s, total, warp_tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(LSUM_THREADS)
k_tile_scan(const u32* __restrict__ in, u64 n, const u64* __restrict__ tile_prefix, u64* __restrict__ out)
{
    __shared__ u32 warp_tot[LSUM_THREADS / 32];
    const u64 base = (u64)blockIdx.x * LSUM_TILE + (u64)threadIdx.x * LSUM_ITEMS;
    u32 v[LSUM_ITEMS];
    u32 s = 0;
#pragma unroll
    for (int j = 0; j < LSUM_ITEMS; ++j) {
        v[j] = base + j < n ? in[base + j] : 0u;
        s += v[j];
    }
    u32 total;
    u64 run = tile_prefix[blockIdx.x] + block_excl_sum<LSUM_THREADS>(s, total, warp_tot);
#pragma unroll
    for (int j = 0; j < LSUM_ITEMS; ++j) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == LSUM_THREADS - 1) out[n] = run;
}

u64 large_sum_tiles(u64 n) { return (n + LSUM_TILE - 1) / LSUM_TILE; }
// out[i] = sum in[0..i), out[n] = total; tile_sum [tiles] u32 and tile_prefix [tiles + 1] u64 are scratch
void launch_exclusive_sum_large(const u32* in, u64 n, u64* out, u32* tile_sum, u64* tile_prefix, cudaStream_t s)
{
    if (n == 0) {
        launch_exclusive_sum(in, 0, out, s);
        return;
    }
    const u64 tiles = large_sum_tiles(n);
    k_tile_sums<<<(unsigned)tiles, LSUM_THREADS, 0, s>>>(in, n, tile_sum);
    launch_exclusive_sum(tile_sum, tiles, tile_prefix, s);
    k_tile_scan<<<(unsigned)tiles, LSUM_THREADS, 0, s>>>(in, n, tile_prefix, out);
}

// ------------------------------------------------------------------ text in: record terminators

constexpr int FIND_THREADS = 256;
constexpr int FIND_BYTES = 16;
constexpr int FIND_TILE = FIND_THREADS * FIND_BYTES;

__device__ __forceinline__ bool is_blank(unsigned char c) { return is_ws(c) && c != '\n'; }

// Is text[q] the token "e" alone on its line? Sets `bad` when the line holds more tokens after it.
__device__ __forceinline__ bool is_terminator(const unsigned char* text, u64 n, u64 q, bool& bad)
{
    if (text[q] != 'e') return false;
    if (q + 1 < n && !is_ws(text[q + 1])) return false;
    for (u64 i = q; i > 0;) {  // first token of its line?
        --i;
        if (text[i] == '\n') break;
        if (!is_blank(text[i])) return false;
    }
    for (u64 i = q + 1; i < n; ++i) {  // nothing but blanks up to the end of the line
        if (text[i] == '\n') break;
        if (!is_blank(text[i])) { bad = true; return false; }
    }
    return true;
}

template <bool WRITE>
__global__ void __launch_bounds__(FIND_THREADS)
k_find_records(const unsigned char* __restrict__ text, u64 n, u32* __restrict__ tile_count,
               const u64* __restrict__ tile_prefix, u64* __restrict__ rec_pos, PlainTotals* tot)
{
    __shared__ u32 warp_tot[FIND_THREADS / 32];
    const u64 base = (u64)blockIdx.x * FIND_TILE + (u64)threadIdx.x * FIND_BYTES;
    u32 mask = 0;
    bool bad = false;
    if (base < n) {
        const int m = (int)(n - base < (u64)FIND_BYTES ? n - base : (u64)FIND_BYTES);
        for (int j = 0; j < m; ++j)
            if (text[base + j] == 'e' && is_terminator(text, n, base + j, bad)) mask |= 1u << j;
    }
    if (bad) atomicMin(&tot->error_pos, base);
    u32 total;
    const u32 rank0 = block_excl_sum<FIND_THREADS>(__popc(mask), total, warp_tot);
    if (!WRITE) {
        if (threadIdx.x == 0) tile_count[blockIdx.x] = total;
        return;
    }
    u64 slot = tile_prefix[blockIdx.x] + rank0;
    while (mask) {
        const int b = __ffs((int)mask) - 1;
        mask &= mask - 1;
        rec_pos[slot++] = base + b;
    }
}

// ------------------------------------------------------------------ text in: per-record parse

struct Span {
    const unsigned char* p;
    int n;
};

// Splits the line [ls, le) into key token and value (rest of the line after the blanks that
// follow the key). Returns 0 for a blank line, 1 for a single token, 2 for key + value.
__device__ __forceinline__ int split_line(const unsigned char* ls, const unsigned char* le, Span& key, Span& val)
{
    while (ls < le && is_blank(*ls)) ++ls;
    if (ls >= le) return 0;
    const unsigned char* k = ls;
    while (ls < le && !is_blank(*ls)) ++ls;
    key.p = k;
    key.n = (int)(ls - k);
    while (ls < le && is_blank(*ls)) ++ls;
    val.p = ls;
    val.n = (int)(le - ls);
    return val.n > 0 ? 2 : 1;
}
__device__ __forceinline__ bool key_is(const Span& k, const char* lit, int n)
{
    if (k.n != n) return false;
    for (int i = 0; i < n; ++i)
        if (k.p[i] != (unsigned char)lit[i]) return false;
    return true;
}

// The five values of a record (nullptr = never defined: the default of `TrainingDataEntry e;`) -> Entry.
// Returns false where the reference would throw or misbehave (std::stoi on a non-number, no move).
__device__ __forceinline__ bool parse_entry(const Span& fen, const Span& move, const Span& score, const Span& ply,
                                            const Span& result, Entry& out)
{
    bool ok = true;
    Pos p;
    pos_clear(p);  // TrainingDataEntry e; -> Position()
    Move mv;
    mv.from = mv.to = 0; mv.type = MT_NORMAL; mv.promo = NO_PIECE;
    long long v;
    int sc = 0, pl = 0, res = 0;
    if (fen.p && !parse_fen(fen.p, fen.p + fen.n, p)) ok = false;
    if (!move.p || !parse_uci(p, move.p, move.n, mv)) ok = false;
    if (score.p) { if (parse_int(score.p, score.p + score.n, v)) sc = (int)(short)v; else ok = false; }
    if (ply.p) { if (parse_int(ply.p, ply.p + ply.n, v)) pl = (int)(v & 0xFFFF); else ok = false; }
    if (result.p) { if (parse_int(result.p, result.p + result.n, v)) res = (int)(short)v; else ok = false; }
    Entry e;
    e.occ0 = p.occ[0]; e.occ1 = p.occ[1]; e.t0 = p.t0; e.t1 = p.t1; e.t2 = p.t2;
    e.meta = (u32)p.stm | ((u32)p.ep << 1) | ((u32)p.cr << 8) | (((u32)p.rule50 & 0xFF) << 12);
    e.pos_ply = (u32)p.ply & 0xFFFF;
    e.mv = (u32)mv.from | ((u32)mv.to << 6) | ((u32)mv.type << 12) | ((u32)mv.promo << 14);
    e.score_ply = ((u32)sc & 0xFFFF) | ((u32)pl << 16);
    e.result = (u32)res & 0xFFFF;
    e.pad = 0;
    out = e;
    return ok;
}

// The block's records are contiguous text (about 105 bytes each): it is staged in shared memory with
// 16-byte loads and parsed there, because byte-wise walking of global memory costs one sector
// request per character and lane. Lines in front of the staged region (a record that inherits a
// field from far back) and blocks whose text does not fit are read from global memory.
constexpr int PARSE_THREADS = 128;
constexpr int PARSE_STAGE = 32768;
__global__ void __launch_bounds__(PARSE_THREADS)
k_parse_records(const unsigned char* __restrict__ text, u64 n, const u64* __restrict__ rec_pos, u64 nrec,
                Entry* __restrict__ entries, PlainTotals* tot)
{
    __shared__ __align__(16) unsigned char stage[PARSE_STAGE + 32];
    __shared__ u64 reg_lo_s, reg_hi_s;
    const u64 r0 = (u64)blockIdx.x * PARSE_THREADS;
    if (threadIdx.x == 0) {
        const u64 r_last = r0 + PARSE_THREADS <= nrec ? r0 + PARSE_THREADS - 1 : nrec - 1;
        u64 lo = r0 > 0 ? rec_pos[r0 - 1] : 0;
        while (lo > 0 && text[lo - 1] != '\n') --lo;  // start of the previous terminator's line
        u64 hi = rec_pos[r_last] + 1;
        if (hi - lo > PARSE_STAGE) hi = lo;  // does not fit: everything from global memory
        reg_lo_s = lo;
        reg_hi_s = hi;
    }
    __syncthreads();
    const u64 reg_lo = reg_lo_s, reg_hi = reg_hi_s;
    const unsigned char* gbase = reinterpret_cast<const unsigned char*>(reinterpret_cast<uintptr_t>(text + reg_lo) & ~(uintptr_t)15);
    const int delta = (int)((text + reg_lo) - gbase);
    const int nvec = reg_hi > reg_lo ? (int)((delta + (reg_hi - reg_lo) + 15) >> 4) : 0;
    for (int i = threadIdx.x; i < nvec; i += PARSE_THREADS)
        reinterpret_cast<uint4*>(stage)[i] = load16_clipped(gbase + 16 * i, text, text + n);
    __syncthreads();
    // byte i of the text: staged copy if inside the region
    const unsigned char* sbase = stage + delta - reg_lo;  // sbase + i is valid for reg_lo <= i < reg_hi
    auto at = [&](u64 i) -> const unsigned char* { return (i >= reg_lo && i < reg_hi) ? sbase + i : text + i; };

    const u64 r = r0 + threadIdx.x;
    if (r >= nrec) return;
    const u64 epos = rec_pos[r];
    const u64 stop_validate = r > 0 ? rec_pos[r - 1] : 0;  // the record's own lines end here
    Span fen{nullptr, 0}, move{nullptr, 0}, score{nullptr, 0}, ply{nullptr, 0}, result{nullptr, 0};
    int found = 0;
    bool bad = false;
    // start of the terminator's line
    u64 cur = epos;
    while (cur > 0 && *at(cur - 1) != '\n') --cur;
    // Only the record's own lines are walked: a key it does not define is inherited from an earlier
    // record ("fields persist", :1254), which the inheritance passes below resolve in linear time
    // (k_record_defs, a running maximum over the records, k_parse_inherited).
    while (cur > 0 && cur > stop_validate) {
        const u64 le = cur - 1;  // the '\n' that ends the previous line
        u64 ls = le;
        while (ls > 0 && *at(ls - 1) != '\n') --ls;
        Span key, val;
        // a line lies entirely inside or entirely in front of the staged region (it starts at a line start)
        const unsigned char* lp = at(ls);
        const int kind = split_line(lp, lp + (le - ls), key, val);
        if (kind == 1 && !key_is(key, "e", 1)) bad = true;  // "key\nvalue": stream and line semantics differ
        if (kind == 2) {
            if (key_is(key, "fen", 3)) { if (!(found & 1)) { fen = val; found |= 1; } }
            else if (key_is(key, "move", 4)) { if (!(found & 2)) { move = val; found |= 2; } }
            else if (key_is(key, "score", 5)) { if (!(found & 4)) { score = val; found |= 4; } }
            else if (key_is(key, "ply", 3)) { if (!(found & 8)) { ply = val; found |= 8; } }
            else if (key_is(key, "result", 6)) { if (!(found & 16)) { result = val; found |= 16; } }
            else if (key_is(key, "e", 1)) bad = true;  // tokens after "e" on its line
        }
        cur = ls;
    }
    if (found != 31 && r > 0) {  // inherits a field from an earlier record: parsed by k_parse_inherited
        tot->inherit = 1;
        if (bad) atomicMin(&tot->error_pos, epos);
        return;
    }
    if (!parse_entry(fen, move, score, ply, result, entries[r])) bad = true;
    if (bad) atomicMin(&tot->error_pos, epos);
}

// ---- fields that persist across records (:1254), in linear time
//
// k_record_defs: for every record and each of the five keys, the start of the line that defines it inside
// the record (+ 1; 0 = the record does not define it; the line nearest to the terminator wins, as in the
// stream). An inclusive running maximum over the records turns that into "the last definition at or
// before this record" (line offsets grow with the record index): three small kernels over 40 bytes
// per record. k_parse_inherited then parses the records that do not define all five keys themselves from
// those lines. The passes only run for files in which some record inherits a field.
constexpr int DEF_KEYS = 5;
__global__ void __launch_bounds__(PARSE_THREADS)
k_record_defs(const unsigned char* __restrict__ text, u64 n, const u64* __restrict__ rec_pos, u64 nrec, u64* __restrict__ defs)
{
    const u64 r = (u64)blockIdx.x * PARSE_THREADS + threadIdx.x;
    if (r >= nrec) return;
    const u64 stop = r > 0 ? rec_pos[r - 1] : 0;
    u64 d[DEF_KEYS] = {0, 0, 0, 0, 0};
    u64 cur = rec_pos[r];
    while (cur > 0 && text[cur - 1] != '\n') --cur;
    while (cur > 0 && cur > stop) {
        const u64 le = cur - 1;
        u64 ls = le;
        while (ls > 0 && text[ls - 1] != '\n') --ls;
        Span key, val;
        if (split_line(text + ls, text + le, key, val) == 2) {
            const int k = key_is(key, "fen", 3) ? 0 : key_is(key, "move", 4) ? 1 : key_is(key, "score", 5) ? 2
                        : key_is(key, "ply", 3) ? 3 : key_is(key, "result", 6) ? 4 : -1;
            if (k >= 0 && d[k] == 0) d[k] = ls + 1;
        }
        cur = ls;
    }
#pragma unroll
    for (int k = 0; k < DEF_KEYS; ++k) defs[r * DEF_KEYS + k] = d[k];
}

constexpr int MAXSCAN_THREADS = 256;
// tile_max[t][k] = maximum of defs[.][k] over the MAXSCAN_THREADS records of tile t
__global__ void __launch_bounds__(MAXSCAN_THREADS) k_defs_tile_max(const u64* __restrict__ defs, u64 nrec, u64* __restrict__ tile_max)
{
    __shared__ u64 sm[MAXSCAN_THREADS / 32][DEF_KEYS];
    const u64 r = (u64)blockIdx.x * MAXSCAN_THREADS + threadIdx.x;
    u64 v[DEF_KEYS];
#pragma unroll
    for (int k = 0; k < DEF_KEYS; ++k) v[k] = r < nrec ? defs[r * DEF_KEYS + k] : 0;
#pragma unroll
    for (int k = 0; k < DEF_KEYS; ++k)
        for (int d = 16; d > 0; d >>= 1) {
            const u64 o = __shfl_xor_sync(0xffffffffu, v[k], d);
            v[k] = o > v[k] ? o : v[k];
        }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < DEF_KEYS; ++k) sm[threadIdx.x >> 5][k] = v[k];
    __syncthreads();
    if (threadIdx.x < DEF_KEYS) {
        u64 m = 0;
        for (int w = 0; w < MAXSCAN_THREADS / 32; ++w) m = sm[w][threadIdx.x] > m ? sm[w][threadIdx.x] : m;
        tile_max[(u64)blockIdx.x * DEF_KEYS + threadIdx.x] = m;
    }
}
// exclusive running maximum over the tiles, in place (one thread per key: a few thousand tiles)
__global__ void k_defs_tile_scan(u64* __restrict__ tile_max, u64 tiles)
{
    if (threadIdx.x >= DEF_KEYS || blockIdx.x != 0) return;
    u64 run = 0;
    for (u64 t = 0; t < tiles; ++t) {
        const u64 v = tile_max[t * DEF_KEYS + threadIdx.x];
        tile_max[t * DEF_KEYS + threadIdx.x] = run;
        run = v > run ? v : run;
    }
}
// defs[r][k] = max(defs[0..r][k]): inclusive, in place
__global__ void __launch_bounds__(MAXSCAN_THREADS) k_defs_apply(u64* __restrict__ defs, u64 nrec, const u64* __restrict__ tile_excl)
{
    __shared__ u64 sm[MAXSCAN_THREADS][DEF_KEYS + 1];
    const u64 r = (u64)blockIdx.x * MAXSCAN_THREADS + threadIdx.x;
#pragma unroll
    for (int k = 0; k < DEF_KEYS; ++k) sm[threadIdx.x][k] = r < nrec ? defs[r * DEF_KEYS + k] : 0;
    __syncthreads();
    for (int d = 1; d < MAXSCAN_THREADS; d <<= 1) {
        u64 o[DEF_KEYS];
#pragma unroll
        for (int k = 0; k < DEF_KEYS; ++k) o[k] = (int)threadIdx.x >= d ? sm[threadIdx.x - d][k] : 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < DEF_KEYS; ++k)
            if (o[k] > sm[threadIdx.x][k]) sm[threadIdx.x][k] = o[k];
        __syncthreads();
    }
    if (r < nrec) {
#pragma unroll
        for (int k = 0; k < DEF_KEYS; ++k) {
            const u64 e = tile_excl[(u64)blockIdx.x * DEF_KEYS + k], v = sm[threadIdx.x][k];
            defs[r * DEF_KEYS + k] = e > v ? e : v;
        }
    }
}

__global__ void __launch_bounds__(PARSE_THREADS)
k_parse_inherited(const unsigned char* __restrict__ text, u64 n, const u64* __restrict__ rec_pos, u64 nrec,
                  const u64* __restrict__ defs, Entry* __restrict__ entries, PlainTotals* tot)
{
    const u64 r = (u64)blockIdx.x * PARSE_THREADS + threadIdx.x;
    if (r >= nrec || r == 0) return;
    const u64 stop = rec_pos[r - 1];
    Span sp[DEF_KEYS];
    bool own = true;
#pragma unroll
    for (int k = 0; k < DEF_KEYS; ++k) {
        const u64 d = defs[r * DEF_KEYS + k];
        sp[k].p = nullptr;
        sp[k].n = 0;
        if (d == 0 || d - 1 < stop) own = false;  // not defined by the record itself
        if (d != 0) {
            const u64 ls = d - 1;
            u64 le = ls;
            while (le < n && text[le] != '\n') ++le;
            Span key;
            split_line(text + ls, text + le, key, sp[k]);
        }
    }
    if (own) return;  // parsed by k_parse_records
    if (!parse_entry(sp[0], sp[1], sp[2], sp[3], sp[4], entries[r])) atomicMin(&tot->error_pos, rec_pos[r]);
}

__device__ __forceinline__ void entry_unpack(const Entry& e, Pos& p, RecordFields& f)
{
    p.occ[0] = e.occ0; p.occ[1] = e.occ1; p.t0 = e.t0; p.t1 = e.t1; p.t2 = e.t2;
    p.stm = e.meta & 1; p.ep = (e.meta >> 1) & 127; p.cr = (e.meta >> 8) & 15; p.rule50 = (e.meta >> 12) & 0xFF;
    p.ply = e.pos_ply & 0xFFFF;
    f.mv.from = e.mv & 63; f.mv.to = (e.mv >> 6) & 63; f.mv.type = (e.mv >> 12) & 3; f.mv.promo = (e.mv >> 14) & 15;
    f.score = (int)(short)(e.score_ply & 0xFFFF);
    f.ply = (int)(e.score_ply >> 16);
    f.result = (int)(short)(e.result & 0xFFFF);
}

__global__ void __launch_bounds__(128)
k_entries_link_encode(const Entry* __restrict__ entries, u64 n, u32* __restrict__ codes, u32* __restrict__ stems,
                      CompressTotals* tot, u64* __restrict__ bleed_list)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pos cur, prev;
    RecordFields cf, pf;
    entry_unpack(entries[i], cur, cf);
    pos_clear(prev);
    pf = cf;
    if (i > 0) entry_unpack(entries[i - 1], prev, pf);
    u32 bleed = 0;
    codes[i] = link_and_encode(i > 0, prev, pf, cur, cf, stems + i * 8, &bleed);
    if (bleed) bleed_report(BleedLog{bleed_list, &tot->bleeds}, i, bleed);
}

// trainingDataEntryToPackedSfenValue (:570-585) for parsed text records
__global__ void __launch_bounds__(128)
k_entries_to_bin(const Entry* __restrict__ entries, u64 n, unsigned char* __restrict__ out)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pos p;
    RecordFields f;
    entry_unpack(entries[i], p, f);
    u32 w[10];
    sfen_encode(p, w);
    w[8] = ((u32)f.score & 0xFFFFu) | (move_to_sfmove(f.mv) << 16);
    w[9] = ((u32)f.ply & 0xFFFFu) | (((u32)f.result & 0xFFu) << 16) | 0xFF000000u;
    uint2* d = reinterpret_cast<uint2*>(out + i * 40);
#pragma unroll
    for (int k = 0; k < 5; ++k) d[k] = make_uint2(w[2 * k], w[2 * k + 1]);
}

// ------------------------------------------------------------------ text out: .bin -> .plain

constexpr int BIN_TEXT_THREADS = 128;
template <bool WRITE>
__global__ void __launch_bounds__(BIN_TEXT_THREADS)
k_bin_text(const unsigned char* __restrict__ bin, u64 n, u32* __restrict__ lens, const u64* __restrict__ offs,
           unsigned char* __restrict__ out, CompressTotals* tot)
{
    __shared__ uint4 windows[WRITE ? 4 * BIN_TEXT_THREADS : 1];
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32* w = reinterpret_cast<const u32*>(bin + i * 40);
    Pos p;
    const bool ok = sfen_decode([&](int j) { return w[j]; }, p);
    if (!ok) {
        if (!WRITE) { atomicMin(&tot->error_index, i); lens[i] = 0; }
        return;
    }
    const u32 w8 = w[8], w9 = w[9];
    const Move mv = sfmove_to_move(w8 >> 16);
    const int score = (int)(short)(w8 & 0xFFFF), ply = (int)(w9 & 0xFFFF), result = (int)(signed char)((w9 >> 16) & 0xFF);
    if (WRITE) {
        BufferedSink s;
        s.init(out + offs[i], windows + threadIdx.x, BIN_TEXT_THREADS);
        put_plain_entry(s, p, mv, score, ply, result);
        s.finish();
    } else {
        CountSink s;
        put_plain_entry(s, p, mv, score, ply, result);
        lens[i] = s.n;
    }
}

// The reference hands its text buffer to the file only once it exceeds 1 MiB (:1318-1325,
// :1448-1455); after an exception the unflushed tail is lost. Given the text of all records that
// were emitted before the error ([0, limit)), follow the flush boundaries: b' = end of the first
// record that makes the buffer exceed 1 MiB. Record ends are found in the text itself ("\ne\n").
__global__ void k_text_flush_orbit(const unsigned char* __restrict__ text, u64 limit, int drop_last, u64* __restrict__ committed)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (drop_last) {
        // the record whose construction threw never reached the buffer: cut the text back to the
        // end of the record before the last one
        u64 q = limit >= 3 ? limit - 3 : 0;
        while (q >= 3 && !(text[q - 3] == '\n' && text[q - 2] == 'e' && text[q - 1] == '\n')) --q;
        limit = q >= 3 ? q : 0;
    }
    u64 b = 0;
    for (;;) {
        u64 q = b + (1u << 20) + 1;  // the flush happens at the first record end >= q
        if (q > limit) break;
        while (q <= limit && !(text[q - 3] == '\n' && text[q - 2] == 'e' && text[q - 1] == '\n')) ++q;
        if (q > limit) break;
        b = q;
    }
    *committed = b;
}

// ------------------------------------------------------------------ host launchers

u64 find_tiles(u64 n) { return (n + FIND_TILE - 1) / FIND_TILE; }
void launch_find_records(bool write, const void* text, u64 n, u32* tile_count, const u64* tile_prefix, u64* rec_pos,
                         PlainTotals* tot, cudaStream_t s)
{
    if (n == 0) return;
    const unsigned blocks = (unsigned)find_tiles(n);
    if (write) k_find_records<true><<<blocks, FIND_THREADS, 0, s>>>((const unsigned char*)text, n, tile_count, tile_prefix, rec_pos, tot);
    else k_find_records<false><<<blocks, FIND_THREADS, 0, s>>>((const unsigned char*)text, n, tile_count, tile_prefix, rec_pos, tot);
}
void launch_parse_records(const void* text, u64 n, const u64* rec_pos, u64 nrec, Entry* entries, PlainTotals* tot,
                          cudaStream_t s)
{
    if (nrec == 0) return;
    k_parse_records<<<(unsigned)((nrec + PARSE_THREADS - 1) / PARSE_THREADS), PARSE_THREADS, 0, s>>>((const unsigned char*)text, n, rec_pos, nrec, entries, tot);
}
u64 defs_tiles(u64 nrec) { return (nrec + MAXSCAN_THREADS - 1) / MAXSCAN_THREADS; }
// the inheritance passes: defs [nrec * 5] u64 and tile_max [defs_tiles(nrec) * 5] u64 are scratch
void launch_parse_inherited(const void* text, u64 n, const u64* rec_pos, u64 nrec, u64* defs, u64* tile_max, Entry* entries,
                            PlainTotals* tot, cudaStream_t s)
{
    if (nrec == 0) return;
    const unsigned pb = (unsigned)((nrec + PARSE_THREADS - 1) / PARSE_THREADS), tb = (unsigned)defs_tiles(nrec);
    k_record_defs<<<pb, PARSE_THREADS, 0, s>>>((const unsigned char*)text, n, rec_pos, nrec, defs);
    k_defs_tile_max<<<tb, MAXSCAN_THREADS, 0, s>>>(defs, nrec, tile_max);
    k_defs_tile_scan<<<1, 32, 0, s>>>(tile_max, tb);
    k_defs_apply<<<tb, MAXSCAN_THREADS, 0, s>>>(defs, nrec, tile_max);
    k_parse_inherited<<<pb, PARSE_THREADS, 0, s>>>((const unsigned char*)text, n, rec_pos, nrec, defs, entries, tot);
}
void launch_entries_link_encode(const Entry* entries, u64 n, u32* codes, u32* stems, CompressTotals* tot, u64* bleed_list,
                                cudaStream_t s)
{
    if (n == 0) return;
    k_entries_link_encode<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(entries, n, codes, stems, tot, bleed_list);
}
void launch_entries_to_bin(const Entry* entries, u64 n, void* out, cudaStream_t s)
{
    if (n == 0) return;
    k_entries_to_bin<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(entries, n, (unsigned char*)out);
}
void launch_bin_text(bool write, const void* bin, u64 n, u32* lens, const u64* offs, void* out, CompressTotals* tot,
                     cudaStream_t s)
{
    if (n == 0) return;
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (write) k_bin_text<true><<<blocks, 128, 0, s>>>((const unsigned char*)bin, n, lens, offs, (unsigned char*)out, tot);
    else k_bin_text<false><<<blocks, 128, 0, s>>>((const unsigned char*)bin, n, lens, offs, (unsigned char*)out, tot);
}
void launch_text_flush_orbit(const void* text, u64 limit, int drop_last, u64* committed, cudaStream_t s)
{
    k_text_flush_orbit<<<1, 32, 0, s>>>((const unsigned char*)text, limit, drop_last, committed);
}

}  // namespace nnp
