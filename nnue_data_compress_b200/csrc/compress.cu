// compress.cu -- .bin -> .binpack kernels (compressBin, compress_file.cpp:1338-1374, and the
// CompressedTrainingDataEntryWriter it drives, :1045-1126), restructured for a B200:
//
//   k_walk_runs, k_walk_items
//                         K1, chain-walking form (walk.cuh): every thread walks a run of records,
//                         splicing the predecessor's Huffman stream by its move and comparing it
//                         with the record's own bits instead of decoding every sfen (:364-446);
//                         continuation plies get their move/score bit string (:877-989), chain
//                         heads their 32-byte stem (:997-1020). Heads met inside a run are parked
//                         and worked off in dense rounds.
//   k_walk_chains         K1, chain-owning form: a thread owns the chains whose heads lie in its range and
//                         walks each from head to end, the warp opening and walking in lock-step
//   k_decode_link_encode  K1, record-parallel form: one thread per record decodes it and tests it
//                         against its predecessor (:587-593). Chosen when more than a third of the
//                         records start a chain (k_sample_heads) and kept as a cross-check of the walk.
//   k_heads_transcode     K1 for files of (nearly) single positions: a chain head's stem is its record's
//                         token stream in another order (heads.cuh), nothing is decoded
//   k_heads_direct        the whole conversion in one kernel when EVERY record starts a chain
//   k_tile_aggregate      per-tile summary of the segmented payload scan
//   k_scan_aggregates_*   exclusive scan of the tile summaries (+ totals), three small launches
//   k_write_payload       re-scans each tile with its carry-in and writes stems, numPlies
//                         fields and movetext bits at their final *payload* offsets
//   k_head_next, k_chunk_orbit
//                         replays the writer's greedy chunk-flush rule (:1076-1080) over the
//                         chain-head offsets; k_orbit_table / k_orbit_resolve do the same across
//                         the ranks of the sharded compressor
//   k_emit_chunks         inserts the 8-byte BINP headers (:486-498)
//
// Each record's bit string depends only on its own position/move/score and the previous
// record's score, so the encode is record-parallel; only bit offsets need a scan.
#include "common.cuh"
#include "kernels.h"
#include "link.cuh"
#include "walk.cuh"
#include "heads.cuh"

namespace nnp {

// ------------------------------------------------------------------ K1

#ifndef K1_THREADS_N
#define K1_THREADS_N 256
#endif
constexpr int K1_THREADS = K1_THREADS_N;
constexpr int K1_TILE = K1_THREADS - 1;  // thread 0 decodes the halo record (predecessor of the tile)

struct K1Shared {
    __align__(16) u32 raw[K1_THREADS * 10 + 4];
    u64 occ0[K1_THREADS], occ1[K1_THREADS], t0[K1_THREADS], t1[K1_THREADS], t2[K1_THREADS];
    u32 meta[K1_THREADS];   // stm | ep << 1 | cr << 8 | ok << 16 | rule50 << 17
    u32 w8[K1_THREADS];     // score | move << 16
    u32 w9[K1_THREADS];     // gamePly | result << 16
    u32 heads[K1_THREADS];  // threads whose record starts a chain
    u32 nheads;
};

__device__ __forceinline__ void k1_load(const K1Shared& sh, int t, Pos& p, RecordFields& f, bool& ok)
{
    const u32 w8 = sh.w8[t], w9 = sh.w9[t], meta = sh.meta[t];
    p.occ[0] = sh.occ0[t]; p.occ[1] = sh.occ1[t];
    p.t0 = sh.t0[t]; p.t1 = sh.t1[t]; p.t2 = sh.t2[t];
    p.stm = meta & 1; p.ep = (meta >> 1) & 127; p.cr = (meta >> 8) & 15;
    p.rule50 = (meta >> 17) & 255; p.ply = 0;
    ok = (meta >> 16) & 1;
    f.score = (int)(short)(w8 & 0xFFFF);
    f.mv = sfmove_to_move(w8 >> 16);
    f.ply = (int)(w9 & 0xFFFF);
    f.result = (int)(signed char)((w9 >> 16) & 0xFF);
}

#ifndef K1_MIN_BLOCKS
#define K1_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(K1_THREADS, K1_MIN_BLOCKS)
k_decode_link_encode(const unsigned char* __restrict__ bin, u64 n, u32* __restrict__ codes,
                     u32* __restrict__ stems, CompressTotals* tot, u64* __restrict__ bleed_list)
{
    __shared__ K1Shared sh;
    const int t = threadIdx.x;
    const long long first = (long long)blockIdx.x * K1_TILE - 1;  // record handled by thread 0
    const long long rec = first + t;
    if (t == 0) sh.nheads = 0;

    // coalesced 8-byte loads of the tile (+halo) into shared memory
    {
        const u64 lo = first < 0 ? 0ull : (u64)first;
        u64 hi = (u64)(first + K1_THREADS);
        if (hi > n) hi = n;
        const uint2* src = reinterpret_cast<const uint2*>(bin + lo * 40);
        const int nvec = (int)(hi - lo) * 5;
        uint2* dst = reinterpret_cast<uint2*>(sh.raw) + (int)((long long)lo - first) * 5;
        for (int i = t; i < nvec; i += K1_THREADS) dst[i] = src[i];
    }
    __syncthreads();

    const bool valid = rec >= 0 && (u64)rec < n;
    Pos p;
    pos_clear(p);
    bool ok = true;
    const u32* w = sh.raw + t * 10;
    u32 w8 = 0, w9 = 0;
    if (valid) {
        ok = sfen_decode([&](int j) { return w[j]; }, p);
        w8 = w[8];
        w9 = w[9];
        if (!ok) atomicMin(&tot->error_index, (u64)rec);
    }
    sh.occ0[t] = p.occ[0]; sh.occ1[t] = p.occ[1];
    sh.t0[t] = p.t0; sh.t1[t] = p.t1; sh.t2[t] = p.t2;
    sh.meta[t] = (u32)p.stm | ((u32)p.ep << 1) | ((u32)p.cr << 8) | ((u32)ok << 16) | (((u32)p.rule50 & 255u) << 17);
    sh.w8[t] = w8;
    sh.w9[t] = w9;
    __syncthreads();

    if (valid && t > 0) {
        RecordFields cf, pf;
        cf.score = (int)(short)(w8 & 0xFFFF);
        cf.mv = sfmove_to_move(w8 >> 16);
        cf.ply = (int)(w9 & 0xFFFF);
        cf.result = (int)(signed char)((w9 >> 16) & 0xFF);
        Pos a;
        bool prev_ok;
        k1_load(sh, t - 1, a, pf, prev_ok);
        const bool has_prev = rec > 0 && ok && prev_ok;
        u32 bleed = 0;
        const u32 code = link_code(has_prev, a, pf, p, cf, &bleed);
        codes[rec] = code;
        if (bleed) bleed_report(BleedLog{bleed_list, &tot->bleeds}, (u64)rec, bleed);
        if (code == 0u) sh.heads[atomicAdd(&sh.nheads, 1u)] = (u32)t;
    }
    __syncthreads();
    // chain heads are rare (about one record in a hundred): pack their stems with full warps
    // instead of letting one lane per warp diverge into stem_pack
    const u32 nheads = sh.nheads;
    for (u32 h = t; h < nheads; h += K1_THREADS) {
        const int th = (int)sh.heads[h];
        Pos hp;
        RecordFields hf;
        bool hok;
        k1_load(sh, th, hp, hf, hok);
        store_stem(hp, hf, stems + (u64)(first + th) * 8);
    }
}

// ------------------------------------------------------------------ K1 for files of single positions
//
// k_heads_transcode: when (nearly) every record starts a chain -- shuffled training data, the L = 1 end of
// the sweep -- K1 only has to produce stems, and a stem is the record's own token stream in another order
// (heads.cuh). One thread per record: the ply / result fields of the record in front decide whether it
// can continue a chain at all (isContinuation :589-590); if not, it is transcoded without building a
// position. The few records that pass the field test, and the streams the transcoder leaves alone, take
// the reference's route (decode both positions, compare, encode) out of line.

constexpr int KH_THREADS = 128;

static __device__ __noinline__ void heads_other(const unsigned char* __restrict__ bin, u64 rec, bool linked,
                                                u32* __restrict__ codes, u32* __restrict__ stems, CompressTotals* tot,
                                                u64* __restrict__ bleed_list)
{
    const u32* w = reinterpret_cast<const u32*>(bin + rec * 40);
    Pos cur, prev;
    const bool ok = decode_record(bin, rec, cur);
    if (!ok) atomicMin(&tot->error_index, rec);
    bool prev_ok = false;
    RecordFields pf = record_fields(0u, 0u);
    if (linked) {
        prev_ok = decode_record(bin, rec - 1, prev);
        pf = record_fields(w[-2], w[-1]);
    }
    u32 bleed = 0;
    codes[rec] = link_and_encode(linked && ok && prev_ok, prev, pf, cur, record_fields(w[8], w[9]), stems + rec * 8, &bleed);
    if (bleed) bleed_report(BleedLog{bleed_list, &tot->bleeds}, rec, bleed);
}

// occupancy of record `rec`, for the first thread of a block (its predecessor is in another block's tile)
static __device__ __noinline__ bool heads_occupancy(const unsigned char* __restrict__ bin, u64 rec, u64& occ)
{
    const u32* w = reinterpret_cast<const u32*>(bin + rec * 40);
    u32 s[8];
    if (record_to_stem([&](int j) { return w[j]; }, s) != HEADS_OK) return false;
    occ = bswap64((u64)s[0] | ((u64)s[1] << 32));
    return true;
}

// Can record `rec` (fields linked to its predecessor) continue it? A move changes at most four squares of the
// occupancy (castling), so two positions that differ in more cannot be a move apart -- which settles nearly
// every accidental field link between unrelated positions without decoding either of them. `prev_occ` is the
// predecessor's occupancy if a neighbouring thread has it.
__device__ __forceinline__ bool heads_may_continue(const unsigned char* __restrict__ bin, u64 rec, u64 occ, bool have_prev,
                                                   u64 prev_occ)
{
    if (!have_prev && !heads_occupancy(bin, rec - 1, prev_occ)) return true;
    return popc64(prev_occ ^ occ) <= 4;
}

__global__ void __launch_bounds__(KH_THREADS, 8)
k_heads_transcode(const unsigned char* __restrict__ bin, u64 n, u32* __restrict__ codes, u32* __restrict__ stems,
                  CompressTotals* tot, u64* __restrict__ bleed_list)
{
    __shared__ __align__(16) u32 raw[KH_THREADS * 10];
    __shared__ u64 occs[KH_THREADS];
    __shared__ unsigned char sts[KH_THREADS];
    const int t = threadIdx.x;
    const u64 first = (u64)blockIdx.x * KH_THREADS;
    const int count = (int)(n - first < (u64)KH_THREADS ? n - first : (u64)KH_THREADS);
    {
        const uint2* src = reinterpret_cast<const uint2*>(bin + first * 40);
        uint2* dst = reinterpret_cast<uint2*>(raw);
        for (int i = t; i < count * 5; i += KH_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const u64 rec = first + t;
    const u32* w = raw + t * 10;
    u32 s[8] = {};
    int st = HEADS_OTHER;
    if (t < count) {
        st = record_to_stem([&](int j) { return w[j]; }, s);
        occs[t] = bswap64((u64)s[0] | ((u64)s[1] << 32));
        sts[t] = (unsigned char)st;
    }
    __syncthreads();
    if (t >= count) return;
    bool linked = false;
    if (rec > 0) {
        const u32 prev9 = t > 0 ? w[-1] : reinterpret_cast<const u32*>(bin + rec * 40)[-1];
        linked = fields_link(prev9, w[9]);
    }
    if (st == HEADS_OK && linked && !heads_may_continue(bin, rec, occs[t], t > 0 && sts[t - 1] == HEADS_OK, t > 0 ? occs[t - 1] : 0ull))
        linked = false;
    if (st == HEADS_OK && !linked) {
        codes[rec] = 0u;
        uint4* d = reinterpret_cast<uint4*>(stems + rec * 8);
        d[0] = make_uint4(s[0], s[1], s[2], s[3]);
        d[1] = make_uint4(s[4], s[5], s[6], s[7]);
    } else {
        heads_other(bin, rec, linked, codes, stems, tot, bleed_list);  // also reports malformed streams
    }
}

// k_heads_direct: the whole conversion in one kernel for a file in which EVERY record starts a chain. Then every
// chain is 34 bytes (stem + numPlies 0), the writer's flush rule (:1076-1080) closes a chunk after exactly
// HEADS_PER_CHUNK = ceil(2^20 / 34) chains, and record r's bytes lie at a position that depends on r alone:
//   (r / HEADS_PER_CHUNK) * (8 + 34 * HEADS_PER_CHUNK) + 8 + 34 * (r % HEADS_PER_CHUNK).
// No codes, no stems array, no payload stream, no scan, no orbit, no chunk copy: a block transcodes 128
// records (record_to_stem), lays their 34-byte chains (and a chunk header where one falls among them) out in
// shared memory and writes the 4.3 KB image with whole-word stores. The premise is checked on the way: a
// record whose ply / result fields link to its predecessor AND whose occupancy differs from it in at most four
// squares (what a move can change) might continue a chain; that, a malformed stream or a stream the
// transcoder leaves alone raises *fallback and the host runs the general pipeline instead.
constexpr u32 HEADS_PER_CHUNK = (CHUNK_THRESHOLD + 33) / 34;
constexpr u64 HEADS_CHUNK_BYTES = (u64)HEADS_PER_CHUNK * 34;
constexpr int HD_THREADS = 128;

__host__ __device__ __forceinline__ u64 heads_direct_offset(u64 rec)
{
    return (rec / HEADS_PER_CHUNK) * (HEADS_CHUNK_BYTES + 8) + 8 + (rec % HEADS_PER_CHUNK) * 34;
}

__global__ void __launch_bounds__(HD_THREADS, 8)
k_heads_direct(const unsigned char* __restrict__ bin, u64 n, unsigned char* __restrict__ out, u32* __restrict__ fallback)
{
    __shared__ __align__(16) u32 raw[HD_THREADS * 10];
    __shared__ __align__(16) u32 image[(HD_THREADS * 34 + 8 + 8) / 4 + 2];
    __shared__ u64 occs[HD_THREADS];
    const int t = threadIdx.x;
    const u64 first = (u64)blockIdx.x * HD_THREADS;
    const int count = (int)(n - first < (u64)HD_THREADS ? n - first : (u64)HD_THREADS);
    {
        const uint2* src = reinterpret_cast<const uint2*>(bin + first * 40);
        uint2* dst = reinterpret_cast<uint2*>(raw);
        for (int i = t; i < count * 5; i += HD_THREADS) dst[i] = src[i];
    }
    // the block's bytes of the file: its chains, and in front of them the header of a chunk its first record opens
    const u64 g_lo = heads_direct_offset(first) - (first % HEADS_PER_CHUNK == 0 ? 8 : 0);
    const u64 g_hi = heads_direct_offset(first + count - 1) + 34;
    const u64 img_base = g_lo & ~3ull;
    __syncthreads();
    const u64 rec = first + t;
    const u32* w = raw + t * 10;
    u32 s[8] = {};
    int st = HEADS_OK;
    if (t < count) {
        st = record_to_stem([&](int j) { return w[j]; }, s);
        occs[t] = bswap64((u64)s[0] | ((u64)s[1] << 32));
    }
    __syncthreads();
    if (t < count) {
        bool doubt = st != HEADS_OK;
        if (!doubt && rec > 0) {
            const u32 prev9 = t > 0 ? w[-1] : reinterpret_cast<const u32*>(bin + rec * 40)[-1];
            // (a predecessor the transcoder refused raises the fallback itself)
            if (fields_link(prev9, w[9])) doubt = heads_may_continue(bin, rec, occs[t], t > 0, t > 0 ? occs[t - 1] : 0ull);
        }
        if (doubt) atomicOr(fallback, 1u);
        unsigned short* img = reinterpret_cast<unsigned short*>(image) + ((heads_direct_offset(rec) - img_base) >> 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            img[2 * k] = (unsigned short)(s[k] & 0xFFFFu);
            img[2 * k + 1] = (unsigned short)(s[k] >> 16);
        }
        img[16] = 0;  // numPlies, big-endian
        if (rec % HEADS_PER_CHUNK == 0) {
            // "BINP" + LE32 payload size of the chunk this record opens (:486-498)
            const u64 left = n - rec;
            const u32 size = (u32)((left < HEADS_PER_CHUNK ? left : (u64)HEADS_PER_CHUNK) * 34);
            img[-4] = 0x4942; img[-3] = 0x504E;  // 'B','I' / 'N','P'
            img[-2] = (unsigned short)(size & 0xFFFFu);
            img[-1] = (unsigned short)(size >> 16);
        }
    }
    __syncthreads();
    // whole words of [g_lo, g_hi); a first / last halfword when the range starts / ends in the middle of a word
    const u64 w_lo = (g_lo + 3) & ~3ull, w_hi = g_hi & ~3ull;
    u32* gw = reinterpret_cast<u32*>(out + w_lo);
    const u32* iw = image + ((w_lo - img_base) >> 2);
    const int words = (int)((w_hi - w_lo) >> 2);
    for (int i = t; i < words; i += HD_THREADS) gw[i] = iw[i];
    if (t == 0 && (g_lo & 2)) *reinterpret_cast<unsigned short*>(out + g_lo) = reinterpret_cast<const unsigned short*>(image)[1];
    if (t == 32 && (g_hi & 2))
        *reinterpret_cast<unsigned short*>(out + w_hi) = reinterpret_cast<const unsigned short*>(image)[(w_hi - img_base) >> 1];
}

// ------------------------------------------------------------------ K1, chain-walking form
//
// k_walk_runs: every thread owns a run of KW_RUN consecutive records and walks it with walk_item
// (walk.cuh): one from-scratch decode of the record before the run, then splice-and-compare per
// record. A chain head met on the way (the ply / result fields say so without decoding anything)
// ends the thread's walk and is appended to a global list; k_walk_items then works the list off
// with one thread per parked head -- decode it, emit its stem, walk the rest of its run -- which
// may park again (a second head in the same run), so the host repeats k_walk_items until a round
// parks nothing. Every round runs dense warps; heads never make a warp diverge into the decoder.

#ifndef KW_THREADS_N
#define KW_THREADS_N 128
#endif
#ifndef KW_RUN_N
#define KW_RUN_N 16
#endif
#ifndef KW_MIN_BLOCKS
#define KW_MIN_BLOCKS 5
#endif
constexpr int KW_THREADS = KW_THREADS_N;
constexpr int KW_RUN = KW_RUN_N;
constexpr u32 KW_NONE = 0xFFFFFFFFu;

// Fraction of chain heads, estimated from evenly spaced record pairs with the field tests of
// isContinuation alone: files of (nearly) single positions are better served by the record-parallel
// kernel, which does not pay for walking state it never uses.
__global__ void __launch_bounds__(256)
k_sample_heads(const unsigned char* __restrict__ bin, u64 n, u64 stride, u64 samples, u64* __restrict__ heads)
{
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    bool head = false;
    if (i < samples) {
        const u64 rec = 1 + i * stride;
        if (rec < n) {
            const u32* w = reinterpret_cast<const u32*>(bin + (rec - 1) * 40);
            head = !fields_link(w[9], w[19]);
        }
    }
    const u32 m = __ballot_sync(0xffffffffu, head);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(heads, (u64)__popc(m));
}

// appends the parked records of a warp with one atomic
__device__ __forceinline__ void park_append(u32 parked, u32* __restrict__ list, u64* count)
{
    const u32 m = __ballot_sync(0xffffffffu, parked != KW_NONE);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    u64 base = 0;
    if (lane == __ffs((int)m) - 1) base = atomicAdd(count, (u64)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs((int)m) - 1);
    if (parked != KW_NONE) list[base + __popc(m & ((1u << lane) - 1u))] = parked;
}

// Persistent form: the grid is sized to the machine (launch_walk_runs) and every block loops over tiles of
// KW_THREADS runs, so the lookup tables are staged in shared memory once per block, not once per tile.
__global__ void __launch_bounds__(KW_THREADS, KW_MIN_BLOCKS)
k_walk_runs(const unsigned char* __restrict__ bin, u64 n, u64 run_lo, u64 run_hi, u32* __restrict__ codes,
            u32* __restrict__ stems, CompressTotals* tot, u32* __restrict__ park_list, u64* park_count,
            u64* __restrict__ bleed_list)
{
    __shared__ StepTables T;
    step_tables_fill(T);
    const BleedLog B{bleed_list, &tot->bleeds};
    for (u64 tile = blockIdx.x; run_lo + tile * KW_THREADS < run_hi; tile += gridDim.x) {
        const u64 run = run_lo + tile * KW_THREADS + threadIdx.x;
        const u64 r0 = run * KW_RUN;
        u32 parked = KW_NONE;
        if (run < run_hi && r0 < n) {
            const u64 e = r0 + KW_RUN < n ? r0 + KW_RUN : n;
            bool head = r0 == 0;
            if (!head) {
                const u32* w = reinterpret_cast<const u32*>(bin + (r0 - 1) * 40);
                head = !fields_link(w[9], w[19]);
            }
            // a head right behind the anchor (files of single positions) is decoded on the spot, where all
            // lanes of the warp do the same thing anyway; any other head is parked for a dense round
            walk_item(bin, head ? r0 : r0 - 1, head, e, codes, stems, [&](u64 rec) { atomicMin(&tot->error_index, rec); },
                      [&](u64 rec, u64 a) { if (rec == a + 1) return true; parked = (u32)rec; return false; },
                      [](u64) {}, &T, &B);
        }
        park_append(parked, park_list, park_count);
    }
}

__global__ void __launch_bounds__(KW_THREADS, KW_MIN_BLOCKS)
k_walk_items(const unsigned char* __restrict__ bin, u64 n, u32* __restrict__ codes, u32* __restrict__ stems,
             CompressTotals* tot, const u32* __restrict__ items, u64 n_items, u32* __restrict__ park_list, u64* park_count,
             u64* __restrict__ bleed_list)
{
    __shared__ StepTables T;
    step_tables_fill(T);
    const BleedLog B{bleed_list, &tot->bleeds};
    const u64 i = (u64)blockIdx.x * KW_THREADS + threadIdx.x;
    u32 parked = KW_NONE;
    if (i < n_items) {
        const u64 rec = items[i];
        u64 e = (rec / KW_RUN + 1) * KW_RUN;
        if (e > n) e = n;
        walk_item(bin, rec, true, e, codes, stems, [&](u64 r) { atomicMin(&tot->error_index, r); },
                  [&](u64 r, u64 a) { if (r == a + 1) return true; parked = (u32)r; return false; }, [](u64) {}, &T, &B);
    }
    park_append(parked, park_list, park_count);
}

// ------------------------------------------------------------------ K1, chain-owning form
//
// k_walk_chains: every thread owns the chains whose HEADS lie in its range of `range` consecutive records
// and walks each of them from its head to its end, wherever that is (the tail of its last chain runs
// into the next thread's range; that thread starts at the first head of its own range, which it finds
// from the ply / result fields alone). Compared with k_walk_runs: the only from-scratch decodes are the
// chain heads, which need one anyway for their stem (one per ~100 records instead of one per 16), no
// head is parked and no item rounds follow. The warp stays dense because decode-then-walk is a nested
// loop: lanes that finish a chain early wait at the loop exit, so the warp decodes the next heads of
// all its lanes together and walks the next chains together -- efficient as long as the chains of a
// file are of similar length, which the density sample checks (launch site). A thread that would
// walk more than WALK_CHAINS_CAP records (a file with giant chains) gives up and raises
// tot->parked[1]: the host then runs the file through k_walk_runs, whose cost does not depend on
// chain length.
constexpr u64 WALK_CHAINS_CAP = 16384;
__global__ void __launch_bounds__(KW_THREADS, KW_MIN_BLOCKS)
k_walk_chains(const unsigned char* __restrict__ bin, u64 n, u64 range, u64 n_ranges, u32* __restrict__ codes,
              u32* __restrict__ stems, CompressTotals* tot, u64* __restrict__ bleed_list)
{
    __shared__ StepTables T;
    step_tables_fill(T);
    const BleedLog B{bleed_list, &tot->bleeds};
    for (u64 tile = blockIdx.x; tile * KW_THREADS < n_ranges; tile += gridDim.x) {
        const u64 g = tile * KW_THREADS + threadIdx.x;
        const u64 r0 = g * range, r1 = g < n_ranges ? (r0 + range < n ? r0 + range : n) : 0;
        // the first chain head at or behind r0, from the field tests of isContinuation (:589-590) alone
        u64 h = r0;
        bool active = g < n_ranges;
        if (active && r0 > 0) {
            u32 prev = reinterpret_cast<const u32*>(bin + (r0 - 1) * 40)[9];
            bool found = false;
            while (!found && h < r1) {  // eight records per step: the loads of a step are independent
                u32 w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = h + j < n ? reinterpret_cast<const u32*>(bin + (h + j) * 40)[9] : 0u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (!found && h + j < r1) {
                        if (!fields_link(prev, w[j])) { found = true; h += j; }
                        prev = w[j];
                    }
                }
                if (!found) h += 8;
            }
            active = found;  // else the whole range continues a chain of an earlier thread
        }
        const u64 stop = h + WALK_CHAINS_CAP < n ? h + WALK_CHAINS_CAP : n;
        auto on_error = [&](u64 rec) { atomicMin(&tot->error_index, rec); };
        // Both loops are warp-uniform (every lane takes part in the votes, lanes without work idle inside):
        // the warp opens the next chains of all its lanes together and walks them together. Left to the
        // compiler's reconvergence, a lane that finishes its chain early runs ahead into the next head's
        // from-scratch decode alone, and the warp pays that decode once per lane.
        WalkState S;
        while (__any_sync(0xffffffffu, active)) {
            if (active) walk_open(S, bin, h, true, stop, codes, stems, on_error);
            bool walking = active;
            while (__any_sync(0xffffffffu, walking)) {
                if (walking) walking = S.rec < stop && walk_step(S, bin, stop, codes, stems, on_error, &T, &B);
            }
            if (active) {
                if (S.rec >= stop) {
                    if (S.rec < n) tot->parked[1] = 1;  // gave up inside a chain: the file goes to k_walk_runs
                    active = false;
                } else if (S.rec < r1) {
                    h = S.rec;  // the next chain of this range
                } else {
                    active = false;  // the next head belongs to the next thread
                }
            }
        }
    }
}

// ------------------------------------------------------------------ payload scan

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int code_bits(u32 code) { return 32 - __ffs((int)code); }  // sentinel position

__device__ __forceinline__ Agg agg_identity()
{
    Agg a;
    a.bytes = 0; a.pre_bits = a.pre_plies = a.post_bits = a.post_plies = a.heads = a.pad = 0;
    return a;
}
__device__ __forceinline__ Agg agg_of_code(u32 code)
{
    Agg a = agg_identity();
    if (code == 0) { a.bytes = 34; a.heads = 1; }
    else { a.pre_bits = (u32)code_bits(code); a.pre_plies = 1; }
    return a;
}
__device__ __forceinline__ Agg agg_shfl_up(const Agg& a, int delta)
{
    Agg r;
    r.bytes = __shfl_up_sync(0xffffffffu, a.bytes, delta);
    r.pre_bits = __shfl_up_sync(0xffffffffu, a.pre_bits, delta);
    r.pre_plies = __shfl_up_sync(0xffffffffu, a.pre_plies, delta);
    r.post_bits = __shfl_up_sync(0xffffffffu, a.post_bits, delta);
    r.post_plies = __shfl_up_sync(0xffffffffu, a.post_plies, delta);
    r.heads = __shfl_up_sync(0xffffffffu, a.heads, delta);
    r.pad = 0;
    return r;
}

// Block-wide exclusive scan of one Agg per thread with the (non-commutative) agg_combine.
// Returns the exclusive prefix of the calling thread; `total` receives the block aggregate.
template <int THREADS>
__device__ __forceinline__ Agg block_exclusive_scan(const Agg& local, Agg& total, Agg* warp_tot /* [THREADS/32] shared */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Agg inc = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Agg o = agg_shfl_up(inc, d);
        if (lane >= d) inc = agg_combine(o, inc);
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    Agg wprefix, all;
    if (THREADS / 32 == 32) {
        // a full warp of warp totals: scan them with shuffles instead of a 32-step loop per thread
        Agg w = warp_tot[lane];
        Agg winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Agg o = agg_shfl_up(winc, d);
            if (lane >= d) winc = agg_combine(o, winc);
        }
        Agg wexc = agg_shfl_up(winc, 1);
        if (lane == 0) wexc = agg_identity();
        // every warp computed the same scan; pick this warp's entries
        wprefix.bytes = __shfl_sync(0xffffffffu, wexc.bytes, wid);
        wprefix.pre_bits = __shfl_sync(0xffffffffu, wexc.pre_bits, wid);
        wprefix.pre_plies = __shfl_sync(0xffffffffu, wexc.pre_plies, wid);
        wprefix.post_bits = __shfl_sync(0xffffffffu, wexc.post_bits, wid);
        wprefix.post_plies = __shfl_sync(0xffffffffu, wexc.post_plies, wid);
        wprefix.heads = __shfl_sync(0xffffffffu, wexc.heads, wid);
        wprefix.pad = 0;
        all.bytes = __shfl_sync(0xffffffffu, winc.bytes, 31);
        all.pre_bits = __shfl_sync(0xffffffffu, winc.pre_bits, 31);
        all.pre_plies = __shfl_sync(0xffffffffu, winc.pre_plies, 31);
        all.post_bits = __shfl_sync(0xffffffffu, winc.post_bits, 31);
        all.post_plies = __shfl_sync(0xffffffffu, winc.post_plies, 31);
        all.heads = __shfl_sync(0xffffffffu, winc.heads, 31);
        all.pad = 0;
    } else {
        wprefix = agg_identity();
        all = agg_identity();
#pragma unroll
        for (int i = 0; i < THREADS / 32; ++i) {
            if (i == wid) wprefix = all;
            all = agg_combine(all, warp_tot[i]);
        }
    }
    total = all;
    Agg exc = agg_shfl_up(inc, 1);
    if (lane == 0) exc = agg_identity();
    __syncthreads();
    return agg_combine(wprefix, exc);
}

// appends one record to a run summary: agg_combine(a, agg_of_code(code)) without the generic case analysis
__device__ __forceinline__ void agg_push(Agg& a, u32 code)
{
    if (code == 0u) {
        a.bytes += (a.heads ? ceil8(a.post_bits) : 0ull) + 34;
        a.post_bits = 0;
        a.post_plies = 0;
        a.heads += 1;
    } else {
        const u32 nb = (u32)code_bits(code);
        if (a.heads) { a.post_bits += nb; a.post_plies += 1; }
        else { a.pre_bits += nb; a.pre_plies += 1; }
    }
}

// the SCAN_ITEMS codes of one thread (two 16-byte loads; a zero-bit continuation fills the tail)
static_assert(SCAN_ITEMS == 8, "load_codes reads two uint4");
__device__ __forceinline__ void load_codes(const u32* __restrict__ codes, u64 base, u64 n, u32 (&c)[SCAN_ITEMS])
{
    // base is a multiple of 8; the array itself starts mid-way for a shard's owned range
    if (base + SCAN_ITEMS <= n && (reinterpret_cast<uintptr_t>(codes + base) & 15) == 0) {
        const uint4* v = reinterpret_cast<const uint4*>(codes + base);
        const uint4 a = v[0], b = v[1];
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) c[i] = base + i < n ? codes[base + i] : (1u << 31);
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_tile_aggregate(const u32* __restrict__ codes, u64 n, Agg* __restrict__ tile_agg)
{
    __shared__ Agg warp_tot[SCAN_THREADS / 32];
    const u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    u32 c[SCAN_ITEMS];
    load_codes(codes, base, n, c);
    Agg local = agg_identity();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) agg_push(local, c[i]);
    }
    Agg total;
    block_exclusive_scan<SCAN_THREADS>(local, total, warp_tot);
    if (threadIdx.x == 0) tile_agg[blockIdx.x] = total;
}

// Exclusive scan of the tile aggregates in place + totals, in three small launches: every block
// scans AGGSCAN_THREADS aggregates and leaves its total; one block scans the block totals; every
// block adds its prefix.
constexpr int AGGSCAN_THREADS = 1024;
__global__ void __launch_bounds__(AGGSCAN_THREADS)
k_scan_aggregates_local(Agg* __restrict__ tile_agg, u64 ntiles, Agg* __restrict__ block_tot)
{
    __shared__ Agg warp_tot[AGGSCAN_THREADS / 32];
    const u64 i = (u64)blockIdx.x * AGGSCAN_THREADS + threadIdx.x;
    const Agg mine = i < ntiles ? tile_agg[i] : agg_identity();
    Agg total;
    const Agg exc = block_exclusive_scan<AGGSCAN_THREADS>(mine, total, warp_tot);
    if (i < ntiles) tile_agg[i] = exc;
    if (threadIdx.x == 0) block_tot[blockIdx.x] = total;
}
__global__ void __launch_bounds__(AGGSCAN_THREADS)
k_scan_aggregates_top(Agg* __restrict__ block_tot, u64 nblocks, CompressTotals* tot)
{
    __shared__ Agg warp_tot[AGGSCAN_THREADS / 32];
    Agg carry = agg_identity();
    for (u64 base = 0; base < nblocks; base += AGGSCAN_THREADS) {
        const u64 i = base + threadIdx.x;
        const Agg mine = i < nblocks ? block_tot[i] : agg_identity();
        Agg total;
        const Agg exc = block_exclusive_scan<AGGSCAN_THREADS>(mine, total, warp_tot);
        if (i < nblocks) block_tot[i] = agg_combine(carry, exc);
        carry = agg_combine(carry, total);
    }
    if (threadIdx.x == 0) {
        tot->payload_bytes = carry.heads ? carry.bytes + ceil8(carry.post_bits) : 0;
        tot->heads = carry.heads;
    }
}
__global__ void __launch_bounds__(AGGSCAN_THREADS)
k_scan_aggregates_apply(Agg* __restrict__ tile_agg, u64 ntiles, const Agg* __restrict__ block_tot)
{
    const u64 i = (u64)blockIdx.x * AGGSCAN_THREADS + threadIdx.x;
    if (i < ntiles) tile_agg[i] = agg_combine(block_tot[blockIdx.x], tile_agg[i]);
}

// ---- byte/bit writers into the zero-initialised payload (all writes are ORs, so ragged
// edges shared between threads, warps and blocks need no ownership protocol)

__device__ __forceinline__ u32 bswap32(u32 v) { return __byte_perm(v, 0, 0x0123); }

// `code` holds n bits MSB-first, left-aligned; bitpos counts MSB-first bits from payload byte 0
__device__ __forceinline__ void or_bits(u32* payload, u64 bitpos, u32 code_left, int n)
{
    const u64 word = bitpos >> 5;
    const int k = (int)(bitpos & 31);
    atomicOr(payload + word, bswap32(code_left >> k));
    if (k + n > 32) atomicOr(payload + word + 1, bswap32(code_left << (32 - k)));
}
__device__ __forceinline__ void or_byte(u32* payload, u64 bytepos, u32 v)
{
    atomicOr(payload + (bytepos >> 2), (v & 0xFF) << ((bytepos & 3) * 8));
}
// eight memory-order words (a stem) at an arbitrary byte offset. The 32 bytes are the stem's alone: payload
// words that lie entirely inside them are stored plainly, only the two ragged edge words, which the
// neighbouring movetext / numPlies bytes share, are OR-ed.
__device__ __forceinline__ void or_words8(u32* payload, u64 bytepos, const uint4& a, const uint4& b)
{
    const u32 w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    u32* dst = payload + (bytepos >> 2);
    const int sh = (int)(bytepos & 3) * 8;
    if (sh == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = w[i];
    } else {
        atomicOr(dst, w[0] << sh);
        u32 carry = w[0] >> (32 - sh);
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            dst[i] = (w[i] << sh) | carry;
            carry = w[i] >> (32 - sh);
        }
        atomicOr(dst + 8, carry);
    }
}

// first entry of the sorted bleed list with record index >= rec
__device__ __forceinline__ u64 bleed_lower_bound(const u64* __restrict__ list, u64 count, u64 rec)
{
    u64 lo = 0, hi = count;
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if ((list[mid] & 0xFFFFFFFFull) < rec) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// BLEED = false: the writer. BLEED = true: the same scan, writing nothing but the stray bits of the plies
// in `bleed_list` (sorted by record index, indices relative to `rec_base`): an unmasked id ORs
// (id >> width) << bitsLeft into the byte its field starts in when that byte is already in use
// (addBitsLE8 :840-862; a field that opens a new byte is truncated by the byte store).
// DENSE: compiled with the warp path for 256 consecutive chain heads (files in which most records start a chain; the
// host knows the head count from the aggregate scan); ordinary files run the instantiation without it (39 registers
// against 58, 0.49 against 0.59 ms per 100 M positions)
template <bool BLEED, bool DENSE>
__global__ void __launch_bounds__(SCAN_THREADS)
k_write_payload(const u32* __restrict__ codes, const u32* __restrict__ stems, u64 n,
                const Agg* __restrict__ tile_prefix, u32* __restrict__ payload, u64* __restrict__ head_off,
                const u64* __restrict__ bleed_list, u64 bleed_count, u64 rec_base)
{
    __shared__ Agg warp_tot[SCAN_THREADS / 32];
    const u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    u32 c[SCAN_ITEMS];
    load_codes(codes, base, n, c);
    Agg local = agg_identity();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) agg_push(local, c[i]);
    }
    Agg total;
    const Agg exc = block_exclusive_scan<SCAN_THREADS>(local, total, warp_tot);
    const Agg st = agg_combine(tile_prefix[blockIdx.x], exc);
    // running writer state before this thread's first record: movetext start of the open
    // chain (payload bytes), its bits and plies so far, and the number of heads seen
    u64 M = st.bytes;
    u32 ob = st.post_bits, op = st.post_plies;
    u64 h = st.heads;
    if (!BLEED && DENSE) {
        // A warp whose 256 records all start chains (files of single positions) writes 256 x 34 contiguous
        // bytes: the stems go through shared memory 64 at a time and leave as whole words, each lane
        // assembling the words it owns; only the two ragged edge words of a batch are OR-ed. Per record this
        // replaces two atomics and nine scattered stores of one thread by a share of coalesced ones.
        __shared__ __align__(16) u32 stage[SCAN_THREADS / 32][64 * 8];
        const bool mine = base + SCAN_ITEMS <= n && local.heads == (u32)SCAN_ITEMS;
        if (__all_sync(0xffffffffu, mine)) {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            u64 P0 = 0;
            if (lane == 0 && h > 0) {
                P0 = M + ceil8(ob);
                if (op != 0) {  // numPlies of the chain that ends in front of the warp's first record (:1118-1119)
                    or_byte(payload, M - 2, op >> 8);
                    or_byte(payload, M - 1, op);
                }
            }
            P0 = __shfl_sync(0xffffffffu, P0, 0);
            const u64 h0 = __shfl_sync(0xffffffffu, h, 0), rec0 = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) head_off[h0 + 32 * k + lane] = P0 + 34ull * (32 * k + lane);
            const unsigned char* sb = reinterpret_cast<const unsigned char*>(stage[warp]);
            for (int q = 0; q < 4; ++q) {
                const uint4* src = reinterpret_cast<const uint4*>(stems + (rec0 + 64 * q) * 8);
                uint4* dst = reinterpret_cast<uint4*>(stage[warp]);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; ++k) dst[32 * k + lane] = src[32 * k + lane];
                __syncwarp();
                const u64 lo = P0 + 2176ull * q, hi = lo + 2176;  // the batch's bytes of the payload
                const u64 w_lo = lo >> 2, w_hi = (hi + 3) >> 2;
                for (u64 wd = w_lo + lane; wd < w_hi; wd += 32) {
                    u32 v = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const long long rel = (long long)(4 * wd + i) - (long long)lo;
                        if (rel >= 0 && rel < 2176) {
                            const u32 r = (u32)rel / 34u, within = (u32)rel - 34u * r;
                            if (within < 32u) v |= (u32)sb[32u * r + within] << (8 * i);  // (numPlies stays 0)
                        }
                    }
                    if (4 * wd >= lo && 4 * wd + 4 <= hi) payload[wd] = v;
                    else if (v) atomicOr(payload + wd, v);
                }
            }
            // (should the file end with this warp, its last chain has no plies and its numPlies stays 0)
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const u64 rec = base + i;
        if (rec >= n) break;
        if (c[i] == 0) {
            u64 P = 0;
            if (h > 0) {
                P = M + ceil8(ob);
                // numPlies of the chain that just ended, big-endian (:1118-1119)
                if (!BLEED && op != 0) {
                    or_byte(payload, M - 2, op >> 8);
                    or_byte(payload, M - 1, op);
                }
            }
            if (!BLEED) {
                const uint4* s = reinterpret_cast<const uint4*>(stems + rec * 8);
                or_words8(payload, P, s[0], s[1]);
                head_off[h] = P;
            }
            M = P + 34;
            ob = 0;
            op = 0;
            ++h;
        } else {
            const int nb = code_bits(c[i]);
            if (!BLEED) {
                or_bits(payload, M * 8 + ob, c[i] & ~(1u << (31 - nb)), nb);
            } else {
                const u64 j = bleed_lower_bound(bleed_list, bleed_count, rec_base + rec);
                if (j < bleed_count && (bleed_list[j] & 0xFFFFFFFFull) == rec_base + rec) {
                    const u32 pk = (u32)(bleed_list[j] >> 32);
                    const u32 raw[2] = {pk & 0xFFu, (pk >> 12) & 0xFFu}, width[2] = {(pk >> 8) & 15u, (pk >> 20) & 15u};
                    u32 q = ob;  // bit offset of the field in the chain's movetext
                    for (int f = 0; f < 2; ++f) {
                        const u32 left = (8u - (q & 7u)) & 7u;  // m_bitsLeft when the field is added
                        const u32 stray = ((raw[f] >> width[f]) << left) & 0xFFu;
                        if (width[f] != 0 && left != 0 && stray != 0) or_byte(payload, M + (q >> 3), stray);
                        q += width[f];
                    }
                }
            }
            ob += nb;
            op = (op + 1) & 0xFFFF;  // std::uint16_t numPlies
        }
        if (!BLEED && rec == n - 1) {  // ~CompressedTrainingDataEntryWriter (:1094-1106): last movelist
            or_byte(payload, M - 2, op >> 8);
            or_byte(payload, M - 1, op);
        }
    }
}

// ------------------------------------------------------------------ chunk orbit
//
// The writer flushes a chunk when a new chain head arrives and the bytes gathered since the
// last flush reached 1 MiB (:1076-1080). With P[h] the payload offset of head h that is the
// orbit b0 = 0, b(k+1) = min{h : P[h] - P[b(k)] >= 2^20}; one warp follows it with a 32-ary
// search per hop.
//
// Sharded compression runs the same orbit over one shard's heads: `base` is the
// shard's offset in the global payload and `carry` the global offset of the last chunk start
// before it (NO_CARRY for the first shard, whose first head opens chunk 0).
//
// Output: seg_off[0] = 0, seg_off[1 + k] = local payload offset of the k-th chunk start,
// seg_off[1 + K] = payload size; segment 0 = [0, first chunk start) belongs to a chunk opened by an
// earlier shard (empty for a whole file) and gets no header.

// first h in [lo, H) with head_off[h] >= target, H if none (whole warp)
__device__ __forceinline__ u64 orbit_search(const u64* __restrict__ head_off, u64 lo, u64 H, u64 target, int lane)
{
    u64 hi = H;  // answer in [lo, hi]; hi == H means none
    while (lo < hi) {
        const u64 span = hi - lo;
        const u64 step = (span + 31) / 32;  // 32 probes; the last one reaches hi - 1 or beyond
        const u64 probe = lo + (u64)(lane + 1) * step - 1;
        const bool ge = probe < hi ? (head_off[probe] >= target) : true;
        const u32 m = __ballot_sync(0xffffffffu, ge);
        if (m == 0) { lo = hi; break; }  // every head below hi is too small
        const int f = __ffs((int)m) - 1;  // first lane whose probe is >= target
        const u64 new_hi = lo + (u64)(f + 1) * step - 1;
        lo = lo + (u64)f * step;
        hi = new_hi < hi ? new_hi : hi;
    }
    return lo;
}

// next[h] = first head h' > h with P[h'] - P[h] >= 2^20 (H if none): the successor of h on the orbit
// if h opens a chunk. One thread per head, binary search; neighbouring threads probe the same lines.
__global__ void __launch_bounds__(256)
k_head_next(const u64* __restrict__ head_off, const CompressTotals* __restrict__ tot, u32* __restrict__ next)
{
    const u64 H = tot->heads;
    const u64 h = (u64)blockIdx.x * 256 + threadIdx.x;
    if (h >= H) return;
    const u64 target = head_off[h] + CHUNK_THRESHOLD;
    u64 lo = h + 1, hi = H;  // first index in [lo, hi) with head_off >= target, hi if none
    if (target > tot->payload_bytes) lo = hi;
    if (lo < hi) {
        // gallop from where the successor lies at the file's mean chain size: a handful of probes instead
        // of log2(heads) (files of short chains have tens of millions of heads)
        // (heads per MiB = 2^20 * H / payload, rounded down: exact to within one head where chains are all alike)
        u64 g = h + (u64)CHUNK_THRESHOLD * H / tot->payload_bytes;
        g = g < lo ? lo : g >= hi ? hi - 1 : g;
        u64 step = 16;
        if (head_off[g] >= target) {
            hi = g;
            while (hi > lo) {
                const u64 p = hi - lo > step ? hi - step : lo;
                if (head_off[p] >= target) hi = p; else { lo = p + 1; break; }
                step <<= 1;
                if (p == lo) break;
            }
        } else {
            lo = g + 1;
            while (lo < hi) {
                const u64 p = hi - lo > step ? lo + step - 1 : hi - 1;
                if (head_off[p] >= target) { hi = p; break; }
                lo = p + 1;
                step <<= 1;
            }
        }
    }
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (head_off[mid] >= target) hi = mid; else lo = mid + 1;
    }
    next[h] = (u32)lo;
}

// follows the orbit through next[]: one dependent load per chunk
__global__ void __launch_bounds__(32)
k_chunk_orbit(const u64* __restrict__ head_off, const u32* __restrict__ next, CompressTotals* tot, u64* __restrict__ seg_off,
              u64 max_chunks, u64 base, u64 carry, int speculate)
{
    const int lane = threadIdx.x;
    const u64 H = tot->heads;
    const u64 total = tot->payload_bytes;
    u64 k = 0;
    if (lane == 0) seg_off[0] = 0;
    u64 cur = H;  // head index of the current chunk start; H = none
    if (H > 0) {
        if (carry == NO_CARRY) {
            cur = 0;
        } else {
            const u64 target_global = carry + CHUNK_THRESHOLD;
            const u64 target = target_global > base ? target_global - base : 0;
            if (target <= total) cur = orbit_search(head_off, 0, H, target, lane);
        }
    }
    // One hop is one dependent load (0.5 us from HBM). Where chains are all alike -- files of single positions:
    // 3242 chunks per 100 M of them -- the next chunk starts as many heads further on as the last one did, so the
    // lanes load the 32 heads that lie 0, 1, 2 ... strides ahead at once and the longest prefix whose links
    // confirm each other is taken in one round trip; anywhere else the prefix is one hop long, as before.
    // (`speculate`: the host asks for it when most records start a chain; elsewhere strides do not repeat and the plain
    // loop is 0.12 us per hop cheaper)
    if (!speculate) {
        if (lane == 0) {
            while (cur < H) {
                const u64 off = head_off[cur];
                const u64 nx = next[cur];
                if (k < max_chunks) seg_off[1 + k] = off;
                ++k;
                cur = nx;
            }
        }
        cur = H;
    }
    u64 stride = 0;
    while (cur < H) {
        const u64 idx = cur + (u64)lane * stride;
        const bool in = idx < H;
        const u64 nx = in ? (u64)next[idx] : H;
        const u64 off = in ? head_off[idx] : 0;
        const u64 idx_up = __shfl_down_sync(0xffffffffu, idx, 1);
        const bool linked = in && lane < 31 && nx == idx_up && nx < H;  // the lane above looked at the right head
        const u32 votes = __ballot_sync(0xffffffffu, linked);
        const int m = 1 + (__ffs((int)~votes) - 1);  // lanes 0 .. m-1 are on the orbit
        if (lane < m && k + lane < max_chunks) seg_off[1 + k + lane] = off;
        k += (u64)m;
        const u64 last_nx = __shfl_sync(0xffffffffu, nx, m - 1), last_idx = __shfl_sync(0xffffffffu, idx, m - 1);
        stride = last_nx - last_idx;
        cur = last_nx;
    }
    if (lane == 0) {
        seg_off[1 + (k < max_chunks ? k : max_chunks)] = total;
        tot->chunks = k;
    }
}

// Orbit table of a shard (sharded compression): a chunk can only be entered within the first MiB of a
// shard's payload (the carry lies before the shard), i.e. at a head whose predecessor starts below
// 2^20. For every such entry head the orbit is followed to the end of the shard, which turns the
// rank-order carry chain into table lookups: entry e -> (offset of e, offset of the last chunk start,
// number of chunk starts). Unused entries hold first_local = ~0.
__global__ void __launch_bounds__(128)
k_orbit_table(const u64* __restrict__ head_off, const u32* __restrict__ next, const CompressTotals* __restrict__ tot,
              u64* __restrict__ table, u64 entries)
{
    const u64 e = (u64)blockIdx.x * 128 + threadIdx.x;
    if (e >= entries) return;
    const u64 H = tot->heads;
    u64 first = ~0ull, last = 0, count = 0;
    if (e < H && (e == 0 || head_off[e - 1] < CHUNK_THRESHOLD)) {
        first = head_off[e];
        for (u64 cur = e; cur < H; cur = next[cur]) {
            last = head_off[cur];
            ++count;
        }
    }
    table[3 * e] = first;
    table[3 * e + 1] = last;
    table[3 * e + 2] = count;
}

// The carry chain over the gathered tables of all ranks, for rank `rank`: out[0] = carry into the
// rank (NO_CARRY if no chunk was opened before it), out[1] = chunks opened before it, out[2] = offset
// of the first chunk start behind the rank's own last one (the total payload size if none),
// out[3] = chunks of the whole file, out[4] = 1 if a chunk starts inside the rank's payload.
__global__ void k_orbit_resolve(const u64* __restrict__ tables, u64 entries, const u64* __restrict__ sizes, int world, int rank,
                                u64* __restrict__ out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    u64 base = 0, carry = NO_CARRY, chunks = 0, next_start = ~0ull;
    u64 my_carry = NO_CARRY, my_before = 0, my_has = 0;
    for (int r = 0; r < world; ++r) {
        if (r == rank) { my_carry = carry; my_before = chunks; }
        const u64* t = tables + (u64)r * entries * 3;
        if (sizes[r] != 0) {
            u64 target = 0;
            if (carry != NO_CARRY) {
                const u64 g = carry + CHUNK_THRESHOLD;
                target = g > base ? g - base : 0;
            }
            u64 lo = 0, hi = entries;  // first entry with first_local >= target (padding is ~0)
            while (lo < hi) {
                const u64 mid = (lo + hi) >> 1;
                if (t[3 * mid] >= target) hi = mid; else lo = mid + 1;
            }
            if (lo < entries && t[3 * lo] != ~0ull) {
                if (r > rank && next_start == ~0ull) next_start = base + t[3 * lo];
                if (r == rank) my_has = 1;
                carry = base + t[3 * lo + 1];
                chunks += t[3 * lo + 2];
            }
        }
        base += sizes[r];
    }
    out[0] = my_carry;
    out[1] = my_before;
    out[2] = next_start == ~0ull ? base : next_start;
    out[3] = chunks;
    out[4] = my_has;
}

// ------------------------------------------------------------------ chunk emission
// grid = EMIT_BLOCKS_PER_SEG * (1 + chunks) blocks, EMIT_BLOCKS_PER_SEG per segment (a flat grid: any
// number of chunks): segment 0 is copied as it is, segment 1 + k behind its
// 8-byte header 'B','I','N','P',LE32(size) (:486-498). Source and destination of a segment differ by
// a multiple of 8 bytes, so they share their alignment and the body moves as aligned vectors.
// `last_size`: size field of the last chunk when it continues in the next shard (else NATURAL_SIZE).
constexpr int EMIT_THREADS = 256;
constexpr unsigned EMIT_BLOCKS_PER_SEG = 8;
__global__ void __launch_bounds__(EMIT_THREADS)
k_emit_chunks(const unsigned char* __restrict__ payload, const u64* __restrict__ seg_off, unsigned char* __restrict__ out,
              u64 last_size, u64 n_seg)
{
    const u64 seg = blockIdx.x / EMIT_BLOCKS_PER_SEG;
    const unsigned part = blockIdx.x % EMIT_BLOCKS_PER_SEG;
    const u64 s0 = seg_off[seg], s1 = seg_off[seg + 1];
    u64 size = s1 - s0;
    unsigned char* dst = out + s0;
    if (seg > 0) {
        dst += 8 * (seg - 1);
        const u64 field = (seg == n_seg - 1 && last_size != NATURAL_SIZE) ? last_size : size;
        if (part == 0 && threadIdx.x < 8) {
            const unsigned char hdr[8] = {'B', 'I', 'N', 'P', (unsigned char)field, (unsigned char)(field >> 8),
                                          (unsigned char)(field >> 16), (unsigned char)(field >> 24)};
            dst[threadIdx.x] = hdr[threadIdx.x];
        }
        dst += 8;
    }
    const unsigned char* src = payload + s0;
    const u64 tid = (u64)part * EMIT_THREADS + threadIdx.x;
    const u64 nthreads = (u64)EMIT_BLOCKS_PER_SEG * EMIT_THREADS;
    u64 head = (8 - ((uintptr_t)src & 7)) & 7;
    if (head > size) head = size;
    for (u64 i = tid; i < head; i += nthreads) dst[i] = src[i];
    const u64 nvec = (size - head) >> 3;
    const uint2* s8 = reinterpret_cast<const uint2*>(src + head);
    uint2* d8 = reinterpret_cast<uint2*>(dst + head);
    for (u64 i = tid; i < nvec; i += nthreads) d8[i] = s8[i];
    const u64 tail0 = head + (nvec << 3);
    for (u64 i = tail0 + tid; i < size; i += nthreads) dst[i] = src[i];
}

// first index i in [start, n) with codes[i] == 0 (a chain head), n if none; one warp
__global__ void __launch_bounds__(32) k_find_head(const u32* __restrict__ codes, u64 n, u64 start, u64* __restrict__ out)
{
    const int lane = threadIdx.x;
    u64 found = n;
    for (u64 i = start; i < n; i += 32) {
        const u64 j = i + lane;
        const bool head = j < n && codes[j] == 0u;
        const u32 m = __ballot_sync(0xffffffffu, head);
        if (m) { found = i + (u64)(__ffs((int)m) - 1); break; }
    }
    if (lane == 0) *out = found;
}

// ------------------------------------------------------------------ host launchers

void init_tables_compress(cudaStream_t s) { k_step_tables_init<<<1, 256, 0, s>>>(); }


void launch_decode_link_encode(const void* d_bin, u64 n, u32* codes, u32* stems, CompressTotals* tot, u64* bleed_list,
                               cudaStream_t s)
{
    if (n == 0) return;
    const u64 blocks = (n + K1_TILE - 1) / K1_TILE;
    k_decode_link_encode<<<(unsigned)blocks, K1_THREADS, 0, s>>>((const unsigned char*)d_bin, n, codes, stems, tot, bleed_list);
}
void launch_heads_transcode(const void* d_bin, u64 n, u32* codes, u32* stems, CompressTotals* tot, u64* bleed_list,
                            cudaStream_t s)
{
    if (n == 0) return;
    k_heads_transcode<<<(unsigned)((n + KH_THREADS - 1) / KH_THREADS), KH_THREADS, 0, s>>>((const unsigned char*)d_bin, n, codes, stems,
                                                                                         tot, bleed_list);
}
u64 heads_direct_bytes(u64 n) { return n == 0 ? 0 : heads_direct_offset(n - 1) + 34; }
void launch_heads_direct(const void* d_bin, u64 n, void* d_out, u32* fallback, cudaStream_t s)
{
    if (n == 0) return;
    k_heads_direct<<<(unsigned)((n + HD_THREADS - 1) / HD_THREADS), HD_THREADS, 0, s>>>((const unsigned char*)d_bin, n,
                                                                                        (unsigned char*)d_out, fallback);
}
void launch_sample_heads(const void* d_bin, u64 n, u64 stride, u64 samples, u64* heads, cudaStream_t s)
{
    if (samples == 0) return;
    k_sample_heads<<<(unsigned)((samples + 255) / 256), 256, 0, s>>>((const unsigned char*)d_bin, n, stride, samples, heads);
}
// SMs of the current device (148 on a B200), asked once per device
static int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}
u64 walk_runs(u64 n) { return (n + KW_RUN - 1) / KW_RUN; }
// ranges of k_walk_chains: a few per resident thread, at least 128 and at most 1024 records each
void launch_walk_chains(const void* d_bin, u64 n, u32* codes, u32* stems, CompressTotals* tot, u64* bleed_list, cudaStream_t s)
{
    if (n == 0) return;
    const u64 resident = (u64)sm_count() * KW_MIN_BLOCKS;
    u64 range = (n + resident * KW_THREADS * 2 - 1) / (resident * KW_THREADS * 2);
    range = range < 128 ? 128 : range > 1024 ? 1024 : range;
    range = (range + 7) & ~(u64)7;
    const u64 n_ranges = (n + range - 1) / range;
    u64 blocks = (n_ranges + KW_THREADS - 1) / KW_THREADS;
    if (blocks > resident) blocks = resident;
    k_walk_chains<<<(unsigned)blocks, KW_THREADS, 0, s>>>((const unsigned char*)d_bin, n, range, n_ranges, codes, stems, tot,
                                                         bleed_list);
}
int walk_run_records() { return KW_RUN; }
// runs [run_lo, run_hi) of the n records at d_bin (the records before run_lo * KW_RUN must be there too)
void launch_walk_runs(const void* d_bin, u64 n, u64 run_lo, u64 run_hi, u32* codes, u32* stems, CompressTotals* tot,
                      u32* park_list, u64* park_count, u64* bleed_list, cudaStream_t s)
{
    if (run_hi <= run_lo) return;
    u64 blocks = (run_hi - run_lo + KW_THREADS - 1) / KW_THREADS;
    const u64 resident = (u64)sm_count() * KW_MIN_BLOCKS;  // one wave of persistent blocks
    if (blocks > resident) blocks = resident;
    k_walk_runs<<<(unsigned)blocks, KW_THREADS, 0, s>>>((const unsigned char*)d_bin, n, run_lo, run_hi, codes, stems, tot,
                                                       park_list, park_count, bleed_list);
}
void launch_walk_items(const void* d_bin, u64 n, u32* codes, u32* stems, CompressTotals* tot, const u32* items, u64 n_items,
                       u32* park_list, u64* park_count, u64* bleed_list, cudaStream_t s)
{
    if (n_items == 0) return;
    const u64 blocks = (n_items + KW_THREADS - 1) / KW_THREADS;
    k_walk_items<<<(unsigned)blocks, KW_THREADS, 0, s>>>((const unsigned char*)d_bin, n, codes, stems, tot, items, n_items,
                                                        park_list, park_count, bleed_list);
}
u64 scan_tiles(u64 n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }
void launch_tile_aggregate(const u32* codes, u64 n, Agg* tile_agg, cudaStream_t s)
{
    if (n == 0) return;
    k_tile_aggregate<<<(unsigned)scan_tiles(n), SCAN_THREADS, 0, s>>>(codes, n, tile_agg);
}
u64 scan_blocks(u64 ntiles) { return (ntiles + AGGSCAN_THREADS - 1) / AGGSCAN_THREADS; }
void launch_scan_aggregates(Agg* tile_agg, u64 ntiles, Agg* block_tot, CompressTotals* tot, cudaStream_t s)
{
    const u64 nb = scan_blocks(ntiles);
    if (nb > 0) k_scan_aggregates_local<<<(unsigned)nb, AGGSCAN_THREADS, 0, s>>>(tile_agg, ntiles, block_tot);
    k_scan_aggregates_top<<<1, AGGSCAN_THREADS, 0, s>>>(block_tot, nb, tot);
    if (nb > 0) k_scan_aggregates_apply<<<(unsigned)nb, AGGSCAN_THREADS, 0, s>>>(tile_agg, ntiles, block_tot);
}
void launch_write_payload(const u32* codes, const u32* stems, u64 n, const Agg* tile_prefix, u32* payload,
                          u64* head_off, bool dense, cudaStream_t s)
{
    if (n == 0) return;
    if (dense)
        k_write_payload<false, true><<<(unsigned)scan_tiles(n), SCAN_THREADS, 0, s>>>(codes, stems, n, tile_prefix, payload,
                                                                                     head_off, nullptr, 0, 0);
    else
        k_write_payload<false, false><<<(unsigned)scan_tiles(n), SCAN_THREADS, 0, s>>>(codes, stems, n, tile_prefix, payload,
                                                                                      head_off, nullptr, 0, 0);
}
void launch_write_bleed(const u32* codes, u64 n, const Agg* tile_prefix, u32* payload, const u64* bleed_list, u64 bleed_count,
                        u64 rec_base, cudaStream_t s)
{
    if (n == 0 || bleed_count == 0) return;
    k_write_payload<true, false><<<(unsigned)scan_tiles(n), SCAN_THREADS, 0, s>>>(codes, nullptr, n, tile_prefix, payload, nullptr,
                                                                                   bleed_list, bleed_count, rec_base);
}
void launch_head_next(const u64* head_off, u64 heads, u32* next, CompressTotals* tot, cudaStream_t s)
{
    if (heads > 0) k_head_next<<<(unsigned)((heads + 255) / 256), 256, 0, s>>>(head_off, tot, next);
}
void launch_chunk_orbit(const u64* head_off, const u32* next, CompressTotals* tot, u64* seg_off, u64 max_chunks, u64 base,
                        u64 carry, bool speculate, cudaStream_t s)
{
    k_chunk_orbit<<<1, 32, 0, s>>>(head_off, next, tot, seg_off, max_chunks, base, carry, speculate ? 1 : 0);
}
void launch_orbit_table(const u64* head_off, const u32* next, const CompressTotals* tot, u64* table, u64 entries, cudaStream_t s)
{
    k_orbit_table<<<(unsigned)((entries + 127) / 128), 128, 0, s>>>(head_off, next, tot, table, entries);
}
void launch_orbit_resolve(const u64* tables, u64 entries, const u64* sizes, int world, int rank, u64* out, cudaStream_t s)
{
    k_orbit_resolve<<<1, 32, 0, s>>>(tables, entries, sizes, world, rank, out);
}
void launch_emit_chunks(const void* payload, const u64* seg_off, u64 chunks, void* out, u64 last_size, cudaStream_t s)
{
    const u64 n_seg = chunks + 1;  // 2^31 / 8 segments = 256 TiB of payload per launch
    k_emit_chunks<<<(unsigned)(n_seg * EMIT_BLOCKS_PER_SEG), EMIT_THREADS, 0, s>>>((const unsigned char*)payload, seg_off,
                                                                                    (unsigned char*)out, last_size, n_seg);
}
void launch_find_head(const u32* codes, u64 n, u64 start, u64* out, cudaStream_t s)
{
    k_find_head<<<1, 32, 0, s>>>(codes, n, start, out);
}

}  // namespace nnp
