// decompress.cu -- .binpack -> .bin kernels (decompressBin, compress_file.cpp:1376-1412, and the
// CompressedTrainingDataEntryReader it drives, :1128-1214).
//
// The reference walks a chunk strictly sequentially: chain k+1 starts where chain k's
// movetext ends, and that length is only known after decoding it (:815-818). Chunk-level
// parallelism (a few hundred chunks) cannot feed 148 SMs, so the chains are found
// speculatively and the reader's walk is verified instead of followed:
//
//   k_walk_chunks         one thread follows the 8-byte BINP headers (:500-521); large files are walked in
//                         parallel segments first (k_seg_find / k_seg_walk / k_seg_prefix / k_seg_finish)
//   k_chunk_heads_only    chunks of nothing but single positions (34-byte chains) need no discovery:
//                         k_emit_heads_only decodes a file made of them, k_emit_heads_chunks such chunks
//                         among ordinary ones (each is ONE entry of the candidate list, k_collapsed_tiles)
//   k_candidates_scan     every byte offset of every chunk is tested for "could be a stem the
//                         reference writer emits", in three stages of increasing cost, each run
//                         on the compacted survivors of the one before; result: a bitmap per tile
//   k_candidates_list     the flagged offsets in file order, with the position count of their header
//   k_mark_conflicts      two candidates closer than a stem cannot both be chain starts
//   k_emit_chains_verify  the optimistic strategy: one thread per candidate decodes its chain
//                         (doMove / nextMoveScore per ply, :669-813), writes the 40-byte records
//                         (SfenPacker::pack :266-312 through the spliced stream) at the index the
//                         prefix sum of the header counts gives it, and checks its link of the
//                         reader's walk; no violation anywhere = the output is the reader's
//   k_probe_chains, k_resolve_chunks, k_emit_chains
//                         the exhaustive strategy, taken when a link does not hold: decode every
//                         candidate to learn where it ends, follow offset 0 -> next -> next per
//                         chunk skipping false candidates, emit the real chains
//   k_slow_count, k_slow_emit
//                         sequential per-chunk walk for chunks even that cannot resolve
//                         (correctness net; exercised through the test hooks only)
//   k_emit_chains_text, k_slow_emit_text
//                         the same walks writing emitPlainEntry text (decompressPlain :1299-1335)
#include "common.cuh"
#include "kernels.h"
#include "chain.cuh"
#include "stream.cuh"
#include "text.cuh"
#include "verify.cuh"

namespace nnp {

// ------------------------------------------------------------------ small block scan (u32)

template <int THREADS>
__device__ __forceinline__ u32 block_exclusive_sum(u32 v, u32& total, u32* warp_tot /* [THREADS/32] shared */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
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

// single block: out[i] = sum of in[0..i) as u64, out[n] = total
constexpr int SUM_THREADS = 1024;
__global__ void __launch_bounds__(SUM_THREADS)
k_exclusive_sum(const u32* __restrict__ in, u64 n, u64* __restrict__ out)
{
    __shared__ u64 part[SUM_THREADS];
    const u64 per = (n + SUM_THREADS - 1) / SUM_THREADS;
    const u64 lo = (u64)threadIdx.x * per;
    u64 hi = lo + per;
    if (hi > n) hi = n;
    u64 s = 0;
    for (u64 i = lo; i < hi; ++i) s += in[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 run = 0;
        for (int i = 0; i < SUM_THREADS; ++i) {
            const u64 v = part[i];
            part[i] = run;
            run += v;
        }
        out[n] = run;
    }
    __syncthreads();
    u64 run = part[threadIdx.x];
    for (u64 i = lo; i < hi; ++i) {
        out[i] = run;
        run += in[i];
    }
}

// single block: out[i] = sum of in[0..i) for u64 input (per-chunk text sizes), out[n] = total
__global__ void __launch_bounds__(SUM_THREADS)
k_exclusive_sum64(const u64* __restrict__ in, u64 n, u64* __restrict__ out)
{
    __shared__ u64 part[SUM_THREADS];
    const u64 per = (n + SUM_THREADS - 1) / SUM_THREADS;
    const u64 lo = (u64)threadIdx.x * per;
    u64 hi = lo + per;
    if (hi > n) hi = n;
    u64 s = 0;
    for (u64 i = lo; i < hi; ++i) s += in[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 run = 0;
        for (int i = 0; i < SUM_THREADS; ++i) {
            const u64 v = part[i];
            part[i] = run;
            run += v;
        }
        out[n] = run;
    }
    __syncthreads();
    u64 run = part[threadIdx.x];
    for (u64 i = lo; i < hi; ++i) {
        const u64 v = in[i];
        out[i] = run;
        run += v;
    }
}

// ------------------------------------------------------------------ chunk table

// One warp: lane 0 follows the headers, the other lanes pull the lines of the next few headers into L2
// while it waits. A hop is one dependent DRAM access (~0.7 us), and the writer's flush rule
// (:1076-1080) makes a chunk 2^20 bytes plus a part of one chain, so the header j chunks ahead lies
// within a few hundred bytes behind j * (2^20 + 8): a wrong guess only costs the miss it would have
// been anyway. `world` / `rank`: the chunk range [chunks * rank / world, chunks * (rank + 1) / world)
// and its bytes are left in the info block (sharded decompression; world = 1 otherwise).
__global__ void __launch_bounds__(32) k_walk_chunks(const unsigned char* __restrict__ in, u64 n, ChunkTable tab, u64 max_chunks,
                                                     u32 world, u32 rank, const u64* __restrict__ seg_done)
{
    if (seg_done && *seg_done) return;  // the segmented walk below has filled the table
    const int lane = threadIdx.x;
    u64 pos = 0, k = 0, tiles = 0;
    int status = 0;
    // which header this lane guesses at: lane 0 the next one (exact), then 5, 9 and 13 lines for the
    // headers two, three and four chunks ahead
    const int ahead = lane == 0 ? 0 : lane < 6 ? 1 : lane < 15 ? 2 : lane < 28 ? 3 : -1;
    const int first = lane == 0 ? 0 : lane < 6 ? 1 : lane < 15 ? 6 : 15;
    while (pos < n) {  // hasNextChunk: peek() / eof (:468-472)
        u32 size = 0;
        int st = 0;
        if (lane == 0) {
            if (n - pos < 8 || in[pos] != 'B' || in[pos + 1] != 'I' || in[pos + 2] != 'N' || in[pos + 3] != 'P') {
                st = -1;  // NNP_ERR_BAD_MAGIC (:504-507)
            } else {
                size = (u32)in[pos + 4] | ((u32)in[pos + 5] << 8) | ((u32)in[pos + 6] << 16) | ((u32)in[pos + 7] << 24);
                if (size > MAX_CHUNK_SIZE) st = -2;                    // NNP_ERR_CHUNK_TOO_LARGE (:515-518)
                else if (n - pos - 8 < size || size < 34) st = -4;      // NNP_ERR_TRUNCATED
            }
        }
        size = __shfl_sync(0xffffffffu, size, 0);
        st = __shfl_sync(0xffffffffu, st, 0);
        if (st != 0) { status = st; break; }
        const u64 next = pos + 8 + (u64)size;
        if (ahead >= 0) {
            const u64 guess = next + (u64)ahead * (CHUNK_THRESHOLD + 8) + (u64)(lane - first) * 128;
            if (guess < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(in + guess));
        }
        if (lane == 0 && k < max_chunks) {
            tab.start[k] = pos + 8;
            tab.len[k] = size;
            tab.tile_base[k] = tiles;
            tiles += ((u64)size + CAND_TILE - 1) / CAND_TILE;
        }
        ++k;
        pos = next;
    }
    if (lane != 0) return;
    if (k <= max_chunks) tab.tile_base[k] = tiles;
    tab.info->chunks = k;
    tab.info->status = status;
    tab.info->tiles = tiles;
    // the rank's share: balanced contiguous chunk range (the same split as sharding.shard_bounds)
    const u64 w = world ? world : 1, r = rank;
    const u64 base = k / w, extra = k % w;
    const u64 lo = r * base + (r < extra ? r : extra);
    const u64 hi = lo + base + (r < extra ? 1 : 0);
    tab.info->range_lo = lo;
    tab.info->range_hi = hi;
    tab.info->byte_lo = tab.info->byte_hi = 0;
    if (k <= max_chunks && hi > lo) {
        tab.info->byte_lo = tab.start[lo] - 8;
        tab.info->byte_hi = tab.start[hi - 1] + tab.len[hi - 1];
    } else if (k <= max_chunks) {
        tab.info->byte_lo = tab.info->byte_hi = lo < k ? tab.start[lo] - 8 : pos;
    }
}

// ---- the header walk of a large file, in parallel
//
// The walk is a chain of dependent loads (~0.5 us per chunk for one warp): 1 ms for the 1.7 GB file eight
// ranks write together, 19 ms for 34 GB of single positions. For files of 64 MiB and more it is done in
// segments of 8 MiB: k_seg_find looks for the first header at or behind every segment's nominal start
// (a block scans for the magic with a plausible size field), k_seg_walk follows the headers of every
// segment in parallel (one warp each), and the segments are only believed if every walk ends exactly on the
// next segment's start -- by induction from offset 0 these are then the headers the sequential walk
// visits (a "BINP" inside a payload is stepped over by the walk of the segment in front of it and makes
// that check fail). Anything else (a broken header, a mismatch, a table that is too small) leaves the
// file to the sequential kernel, which also defines the error behaviour.
static u64 g_seg_bytes = 8ull << 20;  // test hook: nnp_debug_config("walk_seg_bytes", v); files of 8 segments and more
void set_walk_segment_bytes(u64 v) { g_seg_bytes = v ? v : (8ull << 20); }
constexpr int SEG_MAX = 2048;
constexpr int SEG_FIND_THREADS = 1024;
// scratch layout (u64): [0] done flag, [1] ok flag, then per segment: start[SEG_MAX + 1], count, tiles, base, tile_base
struct SegScratch {
    u64 done, ok;
    u64 start[SEG_MAX + 1];
    u64 count[SEG_MAX], tiles[SEG_MAX], base[SEG_MAX + 1], tile_base[SEG_MAX + 1];
};
__device__ __forceinline__ bool plausible_header(const unsigned char* __restrict__ in, u64 n, u64 p)
{
    if (n - p < 8) return false;
    if (in[p] != 'B' || in[p + 1] != 'I' || in[p + 2] != 'N' || in[p + 3] != 'P') return false;
    const u64 size = (u64)in[p + 4] | ((u64)in[p + 5] << 8) | ((u64)in[p + 6] << 16) | ((u64)in[p + 7] << 24);
    return size <= MAX_CHUNK_SIZE && size >= 34 && n - p - 8 >= size;
}
// SEG_FIND_PARTS blocks per segment, each looking at one window of SEG_FIND_WINDOW bytes behind the segment's
// nominal start: the lowest plausible header wins (atomicMin; the array is preset to ~0 = "none within
// 2 MiB", which sends the file to the sequential walk)
constexpr u32 SEG_FIND_PARTS = 16;
constexpr u64 SEG_FIND_WINDOW = 128ull << 10;
__global__ void __launch_bounds__(SEG_FIND_THREADS) k_seg_find(const unsigned char* __restrict__ in, u64 n, u32 segs, SegScratch* S)
{
    const u32 s = blockIdx.x / SEG_FIND_PARTS, part = blockIdx.x % SEG_FIND_PARTS;
    if (s == 0) {
        if (part == 0 && threadIdx.x == 0) {
            atomicMin(reinterpret_cast<unsigned long long*>(&S->start[0]), 0ull);
            atomicMin(reinterpret_cast<unsigned long long*>(&S->start[segs]), (unsigned long long)n);
            S->done = 0;
            S->ok = 1;
        }
        return;
    }
    const u64 o = (n / segs) * s;
    const u64 a0 = (o & ~15ull) + (u64)part * SEG_FIND_WINDOW;
    u64 best = ~0ull;
    for (u64 base = a0; base < a0 + SEG_FIND_WINDOW && base < n; base += (u64)SEG_FIND_THREADS * 16) {
        const u64 p0 = base + (u64)threadIdx.x * 16;  // 16 bytes per thread and step (+ the three behind them)
        if (p0 + 4 <= n) {
            const uint4 v = p0 + 16 <= n ? *reinterpret_cast<const uint4*>(in + p0) : load16_clipped(in + p0, in, in + n);
            const u32 nx = p0 + 20 <= n ? *reinterpret_cast<const u32*>(in + p0 + 16) : 0u;
            const u32 w[5] = {v.x, v.y, v.z, v.w, nx};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const u32 m = __funnelshift_r(w[j >> 2], w[(j >> 2) + 1], (j & 3) * 8);
                if (m == 0x504E4942u && p0 + j >= o && p0 + j < best && plausible_header(in, n, p0 + j)) best = p0 + j;
            }
        }
    }
    if (best != ~0ull) atomicMin(reinterpret_cast<unsigned long long*>(&S->start[s]), (unsigned long long)best);
}
// mode 0: count the chunks and candidate tiles of the segment and check that its walk ends on the next start;
// mode 1: write the segment's chunks into the table at the indices k_seg_prefix gave it
__global__ void __launch_bounds__(32) k_seg_walk(const unsigned char* __restrict__ in, u64 n, u32 segs, SegScratch* S, ChunkTable tab,
                                                 u64 max_chunks, int mode)
{
    const u32 s = blockIdx.x;
    if (threadIdx.x != 0) return;
    if (mode == 1 && !S->done) return;
    u64 pos = S->start[s];
    const u64 end = S->start[s + 1];
    u64 k = mode ? S->base[s] : 0, tiles = mode ? S->tile_base[s] : 0;
    bool ok = pos <= end && end != ~0ull;
    while (ok && pos < end) {
        if (!plausible_header(in, n, pos)) { ok = false; break; }
        const u32 size = (u32)in[pos + 4] | ((u32)in[pos + 5] << 8) | ((u32)in[pos + 6] << 16) | ((u32)in[pos + 7] << 24);
        if (mode && k < max_chunks) {
            tab.start[k] = pos + 8;
            tab.len[k] = size;
            tab.tile_base[k] = tiles;
        }
        tiles += ((u64)size + CAND_TILE - 1) / CAND_TILE;
        ++k;
        pos += 8 + (u64)size;
    }
    if (mode == 0) {
        if (!ok || pos != end) S->ok = 0;  // (every writer stores 0)
        S->count[s] = k;
        S->tiles[s] = tiles;
    }
}
// prefix sums over the segments, the totals, and (after mode 1) the rank's range; one thread
__global__ void k_seg_prefix(u32 segs, SegScratch* S, ChunkTable tab, u64 max_chunks)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    u64 k = 0, tiles = 0;
    for (u32 s = 0; s < segs; ++s) {
        S->base[s] = k;
        S->tile_base[s] = tiles;
        k += S->count[s];
        tiles += S->tiles[s];
    }
    S->base[segs] = k;
    S->tile_base[segs] = tiles;
    S->done = (S->ok && k <= max_chunks && k > 0) ? 1 : 0;
}
__global__ void k_seg_finish(u64 n, u32 segs, SegScratch* S, ChunkTable tab, u32 world, u32 rank)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || !S->done) return;
    const u64 k = S->base[segs];
    tab.tile_base[k] = S->tile_base[segs];
    tab.info->chunks = k;
    tab.info->status = 0;
    tab.info->tiles = S->tile_base[segs];
    const u64 w = world ? world : 1, r = rank;
    const u64 base = k / w, extra = k % w;
    const u64 lo = r * base + (r < extra ? r : extra);
    const u64 hi = lo + base + (r < extra ? 1 : 0);
    tab.info->range_lo = lo;
    tab.info->range_hi = hi;
    if (hi > lo) {
        tab.info->byte_lo = tab.start[lo] - 8;
        tab.info->byte_hi = tab.start[hi - 1] + tab.len[hi - 1];
    } else {
        tab.info->byte_lo = tab.info->byte_hi = lo < k ? tab.start[lo] - 8 : n;
    }
}

// ------------------------------------------------------------------ candidate discovery

constexpr int CAND_THREADS = 256;
static_assert(CAND_TILE == CAND_THREADS * 16, "stage 1 of k_candidates_scan tests 16 offsets per thread");

// Could the 34 bytes at s[0..34) be a stem + numPlies as written by packEntry (:997-1020)?
// Two groups of tests. (1) Properties every stem of the reference writer has, whatever the input:
// rule50 is a uint8_t stored big-endian in 16 bits (byte 30 == 0); unused nibbles stay zero
// (CompressedPosition() zero-initialises m_packedState, Position.h:1218-1222); nibble 13/14 only on
// the corner squares that carry the castling right, nibble 12 only on rank 4/5 matching the side to
// move (Position.h:1274-1333); a non-null move has from != to and promotion bits only with type
// Promotion (Chess.h:1071-1096). (2) Properties of stems made from legal chess positions: exactly
// one king per side, at most 32 pieces, no pawn on rank 1/8, the stored move starts on a piece of
// the side to move and does not land on an own piece. A real stem that fails group (2) (possible
// only for inputs outside that domain) merely sends the file through the exhaustive path.
// bytes 0..27 of a would-be stem as seven little-endian words, from aligned shared-memory loads
__device__ __forceinline__ void stem_words(const unsigned char* s, u32 (&v)[7])
{
    const u32* wp = reinterpret_cast<const u32*>(reinterpret_cast<uintptr_t>(s) & ~(uintptr_t)3);
    const int sh = (int)(reinterpret_cast<uintptr_t>(s) & 3) * 8;
    u32 prev = wp[0];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const u32 next = wp[j + 1];
        v[j] = __funnelshift_r(prev, next, sh);
        prev = next;
    }
}
__device__ __forceinline__ int stem_piece_count(const u32 (&v)[7])
{
    const u64 occ = ((u64)__byte_perm(v[0], 0, 0x0123) << 32) | __byte_perm(v[1], 0, 0x0123);  // big-endian (:1245-1257)
    return popc64(occ);
}

// stage 2a: 2..32 occupied squares and nothing but zeros behind their nibbles. The cheapest of the
// strong tests: in files of short chains it is what most zero-byte survivors fail.
__device__ __forceinline__ bool plausible_stem_count(const unsigned char* s)
{
    if (s[30] != 0) return false;
    u32 v[7];
    stem_words(s, v);
    const int n = stem_piece_count(v);
    if (n < 2 || n > 32) return false;
    // nibble k of the 16 nibble bytes is bits 4(k%8).. of word 2 + k/8; nibbles n.. must be zero
    u32 stray = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) stray |= v[2 + j] & ~stream_low_mask(4 * n, j);
    return stray == 0;
}

// stage 2b: among the first n nibbles there is exactly one white king (10) and one black king
// (11, or 15 = black to move), counted nibble-parallel on the four words
__device__ __forceinline__ bool plausible_stem_kings(const unsigned char* s)
{
    u32 v[7];
    stem_words(s, v);
    const int n = stem_piece_count(v);
    int wk = 0, bk = 0, k15 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const u32 x = v[2 + j];
        const u32 flags = 0x11111111u & stream_low_mask(4 * n, j);
        auto count_eq = [&](u32 pattern) {
            const u32 t = x ^ pattern;
            const u32 nz = (t | (t >> 1) | (t >> 2) | (t >> 3));
            return __popc(~nz & flags);
        };
        wk += count_eq(0xAAAAAAAAu);
        bk += count_eq(0xBBBBBBBBu);
        k15 += count_eq(0xFFFFFFFFu);
    }
    return wk == 1 && bk + k15 == 1;
}

// stage 3: the per-square tests, for offsets that passed plausible_stem_count and plausible_stem_kings
__device__ __forceinline__ bool plausible_stem_squares(const unsigned char* s)
{
    u64 occ = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) occ = (occ << 8) | s[i];
    int stm = WHITE;
    for (int i = 0; i < 16; ++i) stm |= ((s[8 + i] & 15) == 15) | ((s[8 + i] >> 4) == 15);
    const u32 cm = ((u32)s[24] << 8) | s[25];
    const int mtype = (int)(cm >> 14), from = (int)((cm >> 8) & 63), to = (int)((cm >> 2) & 63);
    if (cm != 0 && (from == to || ((cm & 3u) && mtype != MT_PROMOTION))) return false;
    int k = 0, n12 = 0, from_nib = -1, to_nib = -1;
    for (u64 b = occ; b; b &= b - 1, ++k) {
        const int sq = lsb64(b);
        const int nib = (s[8 + (k >> 1)] >> ((k & 1) * 4)) & 15;
        const int rank = sq >> 3;
        if (nib == 13 && sq != 0 && sq != 7) return false;
        if (nib == 14 && sq != 56 && sq != 63) return false;
        if (nib == 12) {
            if (++n12 > 1) return false;
            if (!((rank == 3 && stm == BLACK) || (rank == 4 && stm == WHITE))) return false;
        }
        if ((nib <= 1 || nib == 12) && (rank == 0 || rank == 7)) return false;
        if (sq == from) from_nib = nib;
        if (sq == to) to_nib = nib;
    }
    if (cm != 0) {
        // colour of a nibble: 12 is the pawn that just moved (the side NOT to move), 13 white, 14/15 black
        auto colour = [&](int nib) { return nib < 12 ? (nib & 1) : nib == 12 ? (stm ^ 1) : nib == 13 ? WHITE : BLACK; };
        if (from_nib < 0 || colour(from_nib) != stm) return false;
        if (to_nib >= 0 && colour(to_nib) == stm && mtype != MT_CASTLE) return false;
    }
    return true;
}

__device__ __forceinline__ u64 find_chunk(const u64* base, u64 chunks, u64 tile)
{
    u64 lo = 0, hi = chunks;  // last c with base[c] <= tile
    while (hi - lo > 1) {
        const u64 mid = (lo + hi) >> 1;
        if (base[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// Chunks that hold nothing but single positions (files of chain length 1: every chain is a stem with
// numPlies == 0, 34 bytes): the reader's walk (Reader::next :1154-1190) then visits the multiples of 34,
// whatever the bytes say, so the candidates are known without looking at every offset. One block per
// chunk tests "numPlies == 0 at every multiple of 34"; an ordinary chunk fails at one of its first
// stems and the block leaves at once. k_candidates_scan lists the multiples of 34 for a flagged chunk;
// the emit kernel verifies the walk as for any other candidate set.
constexpr int HEADS_ONLY_THREADS = 256;
__global__ void __launch_bounds__(HEADS_ONLY_THREADS)
k_chunk_heads_only(const unsigned char* __restrict__ in, ChunkTable tab, u32* __restrict__ chunk_flag,
                   u32* __restrict__ chunk_stems, u64* __restrict__ flagged)
{
    __shared__ int fail;
    const u64 c = blockIdx.x;
    const unsigned char* s = in + tab.start[c];
    const u32 stems = tab.len[c] / 34;
    if (threadIdx.x == 0) {
        int f = stems < 1;
        for (u32 k = 0; k < 4 && k < stems; ++k) f |= s[34 * k + 32] | s[34 * k + 33];
        fail = f;
    }
    __syncthreads();
    if (fail) {
        if (threadIdx.x == 0) chunk_flag[c] = 0;
        return;
    }
    int bad = 0;
    for (u32 k = threadIdx.x; k < stems; k += HEADS_ONLY_THREADS) bad |= s[34 * (u64)k + 32] | s[34 * (u64)k + 33];
    if (bad) fail = 1;  // benign race: every writer stores 1
    __syncthreads();
    if (threadIdx.x == 0) {
        chunk_flag[c] = fail ? 0u : 1u;
        if (!fail) {
            chunk_stems[c] = stems;  // the reader's record count for this chunk
            atomicAdd(flagged, 1ull);
        }
    }
}

// k_emit_heads_only: a file whose chunks ALL hold nothing but single positions needs no candidates at all --
// record r of chunk c is the stem at byte 34 * r of the chunk, whatever it contains (Reader::next :1154-1213
// reads 32 + 2 bytes, finds numPlies == 0 and is at the next stem). One thread per record; the chunk is found
// from the prefix sums of the chunks' stem counts (one search per block, then a step or two per thread).
#ifndef HEADS_EMIT_MIN_BLOCKS
#define HEADS_EMIT_MIN_BLOCKS 8  // 63 registers, no spills: 6.6 ms per 100 M start positions against 7.0 at 5 blocks (77 registers)
#endif
constexpr int HEADS_EMIT_THREADS = 128;
__global__ void __launch_bounds__(HEADS_EMIT_THREADS, HEADS_EMIT_MIN_BLOCKS)
k_emit_heads_only(const unsigned char* __restrict__ in, ChunkTable tab, u64 chunks, const u64* __restrict__ chunk_base,
                  unsigned char* __restrict__ out, u64 rec_limit)
{
    __shared__ u32 scratch[8 * HEADS_EMIT_THREADS];
    __shared__ StepTables T;
    __shared__ u64 first_chunk;
    step_tables_fill(T);
    const u64 r0 = (u64)blockIdx.x * HEADS_EMIT_THREADS;
    if (threadIdx.x == 0) {
        u64 lo = 0, hi = chunks;  // the last chunk whose first record is <= r0
        while (hi - lo > 1) {
            const u64 mid = (lo + hi) >> 1;
            if (chunk_base[mid] <= r0) lo = mid; else hi = mid;
        }
        first_chunk = lo;
    }
    __syncthreads();
    const u64 r = r0 + threadIdx.x;
    if (r >= chunk_base[chunks]) return;
    u64 c = first_chunk;
    while (chunk_base[c + 1] <= r) ++c;
    const unsigned char* s = in + tab.start[c] + 34 * (r - chunk_base[c]);
    u32 consumed = 0;
    emit_chain_bin(s, 0u, out, r, rec_limit, scratch + threadIdx.x, HEADS_EMIT_THREADS, consumed, &T, nullptr);
}

// pass 0: tests every offset of the tile, stores the tile's flag bitmap (CAND_TILE / 32 words) and
//         tile_count[tile].
// Stage 1 looks at one byte per offset -- the high byte of the stem's rule50 field, which packEntry
// always writes as zero -- 16 offsets per thread with byte-parallel compares, and queues the
// survivors (about 7 % of the offsets of real movetext); stage 2 runs the full test on the queue
// with dense warps.
__global__ void __launch_bounds__(CAND_THREADS)
k_candidates_scan(const unsigned char* __restrict__ in, u64 n_in, ChunkTable tab, u32* __restrict__ tile_count,
                  u32* __restrict__ tile_flags, u32 debug_reject_mod, const u32* __restrict__ chunk_flag,
                  const u64* __restrict__ scan_prefix)
{
    __shared__ __align__(16) unsigned char sm[CAND_TILE + 96];  // the tile from its 16-byte aligned base
    __shared__ u32 flags[CAND_TILE / 32];
    __shared__ unsigned short queue[CAND_TILE], queue2[CAND_TILE];
    __shared__ u32 warp_tot[CAND_THREADS / 32];
    __shared__ u32 nq, nq2;
    const int t = threadIdx.x, lane = t & 31;
    // `scan_prefix` (collapse mode): the grid only covers the tiles of the chunks that are not collapsed
    const u64 c = scan_prefix ? find_chunk(scan_prefix, tab.info->chunks, blockIdx.x) : find_chunk(tab.tile_base, tab.info->chunks, blockIdx.x);
    const u64 tile = scan_prefix ? tab.tile_base[c] + (blockIdx.x - scan_prefix[c]) : blockIdx.x;
    const u64 clen = tab.len[c];
    const u64 off0 = (tile - tab.tile_base[c]) * CAND_TILE;
    const unsigned char* src = in + tab.start[c] + off0;
    const u64 avail = clen - off0;  // > 0 by construction
    if (chunk_flag[c]) {
        // a chunk of single positions: the chains start at the multiples of 34 (at most one per bitmap word)
        u32 mine = 0;
        if (t < CAND_TILE / 32) {
            const u64 base = off0 + 32ull * t;
            const u64 m = (base + 33) / 34 * 34;
            u32 word = 0;
            if (m < base + 32 && m + 34 <= clen) {
                word = 1u << (u32)(m - base);
                if (debug_reject_mod && (u32)((m * 2654435761ull) >> 11) % debug_reject_mod == 0) word = 0;  // test hook
            }

            tile_flags[tile * (CAND_TILE / 32) + t] = word;
            mine = word != 0;
        }
        u32 total;
        block_exclusive_sum<CAND_THREADS>(mine, total, warp_tot);
        if (t == 0) tile_count[tile] = total;
        return;
    }
    const int nload = (int)(avail < (u64)(CAND_TILE + 34) ? avail : (u64)(CAND_TILE + 34));
    const unsigned char* base = reinterpret_cast<const unsigned char*>((uintptr_t)src & ~(uintptr_t)15);
    const int delta = (int)(src - base);
    const int nvec = (delta + nload + 15) >> 4;
    for (int i = t; i < nvec; i += CAND_THREADS)
        reinterpret_cast<uint4*>(sm)[i] = load16_clipped(base + 16 * i, in, in + n_in);
    if (t < CAND_TILE / 32) flags[t] = 0;
    if (t == 0) nq = nq2 = 0;
    __syncthreads();
    const unsigned char* s0 = sm + delta;  // s0[o] = byte at chunk offset off0 + o

    // stage 1: offsets 16t .. 16t+15, byte s0[o + 30]
    {
        const int A = delta + 30 + 16 * t;
        const u32* wp = reinterpret_cast<const u32*>(sm) + (A >> 2);
        const int sh = (A & 3) * 8;
        u32 x[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) x[i] = wp[i];
        u32 mask = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u32 y = __funnelshift_r(x[i], x[i + 1], sh);
            // bit 7 of every zero byte (a 0x01 byte above a zero byte is flagged too: stage 2 re-tests)
            const u32 z = (y - 0x01010101u) & ~y & 0x80808080u;
            mask |= ((((z >> 7) * 0x00204081u) >> 21) & 15u) << (4 * i);
        }
        // offsets with a whole stem + numPlies inside the chunk
        const long long last = (long long)avail - 34 - 16 * t;  // last valid local bit index
        if (last < 15) mask &= last < 0 ? 0u : ((2u << (int)last) - 1u);
        const u32 cnt = __popc(mask);
        u32 inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        u32 wbase = 0;
        if (lane == 31 && inc) wbase = atomicAdd(&nq, inc);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        u32 slot = wbase + inc - cnt;
        while (mask) {
            const int bit = __ffs((int)mask) - 1;
            mask &= mask - 1;
            queue[slot++] = (unsigned short)(16 * t + bit);
        }
    }
    __syncthreads();
    // stages 2a / 2b: the word-parallel tests on the survivors; stage 3: the per-square tests on what
    // is left (in files of single positions that is every 34th offset: it has to run dense, too).
    // The queues alternate: queue -> queue2 -> queue -> flags.
    const u32 n_queued = nq;
    for (u32 q = t; q < n_queued; q += CAND_THREADS) {
        const int o = queue[q];
        if (plausible_stem_count(s0 + o)) queue2[atomicAdd(&nq2, 1u)] = (unsigned short)o;
    }
    __syncthreads();
    const u32 n_counted = nq2;
    if (t == 0) nq = 0;
    __syncthreads();
    for (u32 q = t; q < n_counted; q += CAND_THREADS) {
        const int o = queue2[q];
        if (plausible_stem_kings(s0 + o)) queue[atomicAdd(&nq, 1u)] = (unsigned short)o;
    }
    __syncthreads();
    const u32 n_kings = nq;
    for (u32 q = t; q < n_kings; q += CAND_THREADS) {
        const int o = queue[q];
        bool ok = plausible_stem_squares(s0 + o);
        // test hook: drop a pseudo-random subset of candidates to exercise the fallbacks
        if (debug_reject_mod && (u32)(((off0 + (u64)o) * 2654435761ull) >> 11) % debug_reject_mod == 0) ok = false;
        if (ok) atomicOr(&flags[o >> 5], 1u << (o & 31));
    }
    __syncthreads();
    u32 mine = 0;
    if (t < CAND_TILE / 32) {
        const u32 word = flags[t];
        tile_flags[tile * (CAND_TILE / 32) + t] = word;
        mine = __popc(word);
    }
    u32 total;
    block_exclusive_sum<CAND_THREADS>(mine, total, warp_tot);
    if (t == 0) tile_count[tile] = total;
}

// pass 1: lists the flagged offsets at cand_*[tile_prefix[tile] ...] in ascending offset order,
//         with the position count 1 + numPlies read from the chain header (:1175)
constexpr int LIST_THREADS = CAND_TILE / 32;  // one thread per bitmap word
__global__ void __launch_bounds__(LIST_THREADS)
k_candidates_list(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ tile_flags,
                  const u64* __restrict__ tile_prefix, u32* __restrict__ cand_chunk, u32* __restrict__ cand_off,
                  u32* __restrict__ cand_cnt, const u32* __restrict__ collapsed)
{
    __shared__ u32 warp_tot[LIST_THREADS / 32];
    const u64 tile = blockIdx.x;
    if (tile_prefix[tile + 1] == tile_prefix[tile]) return;
    const u64 c = find_chunk(tab.tile_base, tab.info->chunks, tile);
    const u64 off0 = (tile - tab.tile_base[c]) * CAND_TILE;
    const unsigned char* src = in + tab.start[c] + off0;
    u32 word = tile_flags[tile * (CAND_TILE / 32) + threadIdx.x];
    u32 total;
    const u32 rank0 = block_exclusive_sum<LIST_THREADS>(__popc(word), total, warp_tot);
    u64 slot = tile_prefix[tile] + rank0;
    while (word) {
        const int b = __ffs((int)word) - 1;
        word &= word - 1;
        const u32 o = threadIdx.x * 32 + b;
        cand_chunk[slot] = (u32)c;
        cand_off[slot] = (u32)(off0 + o);
        cand_cnt[slot] = collapsed && collapsed[c] ? tab.len[c] / 34u : 1u + (((u32)src[o + 32] << 8) | (u32)src[o + 33]);
        ++slot;
    }
}

// ------------------------------------------------------------------ chain walking

// one thread per candidate: where does its chain end?
__device__ __forceinline__ u32 plain_len(const ChainCursor& c)
{
    CountSink s;
    put_plain_entry(s, c.pos, c.mv, c.score, c.ply, c.result);
    return s.n;
}

// TEXT: also sum the .plain size of the chain's records (decompressPlain needs output offsets)
constexpr int PROBE_THREADS = 128;
#ifndef PROBE_MIN_BLOCKS
#define PROBE_MIN_BLOCKS 8
#endif
template <bool TEXT>
__global__ void __launch_bounds__(PROBE_THREADS, PROBE_MIN_BLOCKS)
k_probe_chains(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ cand_chunk,
               const u32* __restrict__ cand_off, u64 ncand, u32* __restrict__ cand_next, u32* __restrict__ cand_cnt,
               u32* __restrict__ cand_tlen)
{
    const u64 i = (u64)blockIdx.x * PROBE_THREADS + threadIdx.x;
    if (i >= ncand) return;
    const u32 c = cand_chunk[i], off = cand_off[i];
    const u32 clen = tab.len[c];
    const unsigned char* s = in + tab.start[c] + off;
    ChainCursor cc;
    chain_open(s, cc);
    u32 next = off + 34;
    bool ok = true;
    u32 tlen = TEXT ? plain_len(cc) : 0u;
    if (cc.num_plies > 0) {
        // the stored first move must start on a piece of the side to move
        const int pc = cc.mv.from < 64 ? pos_piece_at(cc.pos, cc.mv.from) : NO_PIECE;
        if (pc == NO_PIECE || (pc & 1) != cc.pos.stm) ok = false;
        BitReader r;
        r.init(s + 34, (u64)(clen - off - 34));
        for (u32 k = 0; ok && k < cc.num_plies; ++k) {
            ok = chain_step(cc, r, true);
            if (TEXT && ok) tlen += plain_len(cc);
        }
        next = off + 34 + ((r.pos + 7) >> 3);  // numReadBytes (:815-818)
    }
    cand_next[i] = ok ? next : 0xFFFFFFFFu;
    cand_cnt[i] = 1u + cc.num_plies;
    if (TEXT) cand_tlen[i] = tlen;
}

// One warp per chunk: Reader::next / fetchNextChunkIfNeeded (:1154-1213) over the candidate list.
// The walk "offset 0 -> next -> next ..." is sequential in principle, but on real files every
// candidate is a real chain and candidate k+1 starts exactly where candidate k ends, so the warp
// verifies 32 links per step with one ballot and only falls out of lock-step at a false positive.
constexpr int RESOLVE_WARPS = 4;
__global__ void __launch_bounds__(RESOLVE_WARPS * 32)
k_resolve_chunks(ChunkTable tab, const u64* __restrict__ tile_prefix, const u32* __restrict__ cand_off,
                 const u32* __restrict__ cand_next, const u32* __restrict__ cand_cnt, u32* __restrict__ cand_base,
                 u32* __restrict__ chunk_count, u32* __restrict__ chunk_slow, const u32* __restrict__ cand_tlen,
                 u64* __restrict__ cand_tbase, u64* __restrict__ chunk_tbytes)
{
    const u64 c = (u64)blockIdx.x * RESOLVE_WARPS + (threadIdx.x >> 5);
    if (c >= tab.info->chunks) return;
    const int lane = threadIdx.x & 31;
    const u64 cb = tile_prefix[tab.tile_base[c]], ce = tile_prefix[tab.tile_base[c + 1]];
    const u32 clen = tab.len[c];
    u64 i = cb;
    u32 cur = 0, count = 0;
    u64 tbytes = 0;
    bool slow = false;
    while ((u64)cur + 34 <= clen) {
        if (i >= ce) { slow = true; break; }  // the expected chain start is not a candidate
        const u64 k = i + lane;
        const bool in = k < ce;
        const u32 o = in ? cand_off[k] : 0xFFFFFFFFu;
        const u32 nx = in ? cand_next[k] : 0xFFFFFFFFu;
        const u32 cn = in ? cand_cnt[k] : 0u;
        u32 expect = __shfl_up_sync(0xffffffffu, nx, 1);
        if (lane == 0) expect = cur;
        // a lane is "linked" when it starts where its predecessor ends, its own chain decodes, and
        // the predecessor did not already finish the chunk
        const bool linked = in && o == expect && nx != 0xFFFFFFFFu && (u64)expect + 34 <= clen;
        const u32 bad = ~__ballot_sync(0xffffffffu, linked);
        const int good = bad ? (__ffs((int)bad) - 1) : 32;  // lanes [0, good) are real chains
        // exclusive prefix of the counts over the good lanes
        u32 inc = lane < good ? cn : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 v = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += v;
        }
        if (lane < good) cand_base[k] = count + inc - cn;
        count += __shfl_sync(0xffffffffu, inc, 31);
        if (cand_tlen) {
            const u64 tl = lane < good ? (u64)cand_tlen[k] : 0ull;
            u64 tinc = tl;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u64 v = __shfl_up_sync(0xffffffffu, tinc, d);
                if (lane >= d) tinc += v;
            }
            if (lane < good) cand_tbase[k] = tbytes + tinc - tl;
            tbytes += __shfl_sync(0xffffffffu, tinc, 31);
        }
        if (good > 0) cur = __shfl_sync(0xffffffffu, nx, good - 1);
        i += good;
        if (good < 32 && (u64)cur + 34 <= clen) {
            // lane `good` broke the lock-step: candidate i is either a false positive that lies
            // inside the chain just walked (skip it) or the walk left the candidate set
            const u32 o_bad = __shfl_sync(0xffffffffu, o, good);
            const u32 nx_bad = __shfl_sync(0xffffffffu, nx, good);
            if (i >= ce) { slow = true; break; }
            if (o_bad < cur) {
                if (lane == 0) cand_base[i] = 0xFFFFFFFFu;
                ++i;
            } else {  // o_bad > cur, or the chain at cur does not decode
                (void)nx_bad;
                slow = true;
                break;
            }
        }
    }
    // whatever is left are false positives behind the last chain
    for (u64 k = i + lane; k < ce; k += 32) cand_base[k] = 0xFFFFFFFFu;
    if (slow) {
        __syncwarp();
        for (u64 k = cb + lane; k < ce; k += 32) cand_base[k] = 0xFFFFFFFFu;
        count = 0;  // filled in by k_slow_count
        tbytes = 0;
    }
    if (lane == 0) {
        chunk_count[c] = count;
        chunk_slow[c] = slow ? 1u : 0u;
        if (chunk_tbytes) chunk_tbytes[c] = tbytes;
    }
}

// ------------------------------------------------------------------ record emission

#ifndef EMIT_MIN_BLOCKS
#define EMIT_MIN_BLOCKS 5
#endif
constexpr int EMITC_THREADS = 128;
__global__ void __launch_bounds__(EMITC_THREADS, EMIT_MIN_BLOCKS)
k_emit_chains(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ cand_chunk,
              const u32* __restrict__ cand_off, const u32* __restrict__ cand_base, u64 ncand,
              const u64* __restrict__ chunk_base, unsigned char* __restrict__ out, DecompressTotals* tot,
              const u64* __restrict__ placed_rec, const u32* __restrict__ placed_next)
{
    __shared__ u32 scratch[8 * EMITC_THREADS];
    const u64 i = (u64)blockIdx.x * EMITC_THREADS + threadIdx.x;
    if (i >= ncand) return;
    const u32 b = cand_base[i];
    const u32 c = cand_chunk[i], off = cand_off[i];
    if (b == 0xFFFFFFFFu) {
        const u64 k = atomicAdd(&tot->false_candidates, 1ull);
        if (k < 8) tot->false_sample[k] = ((u64)c << 32) | off;
        return;
    }
    const u64 rec0 = chunk_base[c] + b;
    // after a failed optimistic walk: the chain was decoded then and already lies where it belongs
    if (placed_rec && placed_rec[i] == rec0 && placed_next[i] != 0xFFFFFFFFu) return;
    const unsigned char* s = in + tab.start[c] + off;
    u32 consumed = 0;
    const bool ok = emit_chain_bin(s, tab.len[c] - off - 34, out, rec0, ~0ull, scratch + threadIdx.x, EMITC_THREADS, consumed);
    if (!ok) atomicMin(&tot->error_chunk, (u64)c);
}

// Two candidates less than 34 bytes apart overlap, so at most one of them starts a chain. (This is
// how the rare false candidate arises in practice: 27 bytes before a real stem whose occupancy has
// empty middle ranks, the bytes "ff ff 00 00 00 00 ff" land on the score / ply / rule50 / numPlies
// fields.) The thread of the lower one decides: offset 0 always starts a chain; otherwise the
// chain of the candidate before the pair is walked to see which of the two it ends on. The loser
// gets cand_cnt = 0 (a chain holds at least one position), which takes it out of the prefix sum
// and of the walk below. Anything this heuristic gets wrong is caught by k_emit_chains_verify.
__global__ void __launch_bounds__(128)
k_mark_conflicts(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ cand_chunk,
                 const u32* __restrict__ cand_off, u32* __restrict__ cand_cnt, u64 ncand)
{
    const u64 i = (u64)blockIdx.x * 128 + threadIdx.x;
    if (i + 1 >= ncand) return;
    const u32 c = cand_chunk[i];
    if (cand_chunk[i + 1] != c) return;
    const u32 off = cand_off[i], off2 = cand_off[i + 1];
    if (off2 - off >= 34) return;
    if (i == 0 || cand_chunk[i - 1] != c) {
        if (off == 0) cand_cnt[i + 1] = 0;
        return;
    }
    const u32 poff = cand_off[i - 1];
    u32 consumed = 0;
    if (!walk_chain(in + tab.start[c] + poff, tab.len[c] - poff - 34, [](const ChainCursor&, u32) {}, consumed)) return;
    const u32 end = poff + consumed;
    if (end == off) cand_cnt[i + 1] = 0;
    else if (end == off2) cand_cnt[i] = 0;
}

// Optimistic single walk: assume every (unmarked) candidate is a real chain. Each thread emits its
// chain at the record index given by the prefix sum of the header counts and then verifies its link
// of the reader's walk (Reader::next / fetchNextChunkIfNeeded, :1154-1213): the chunk's first
// candidate sits at offset 0, every chain ends exactly where the next candidate starts, and
// after the last chain fewer than 34 bytes remain. If every link of every chunk holds, the
// candidates ARE the reader's chains and the output is final; otherwise *violations is raised
// and the caller reruns the file through the exhaustive path (probe / resolve / emit).
__global__ void __launch_bounds__(EMITC_THREADS, EMIT_MIN_BLOCKS)
k_emit_chains_verify(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ cand_chunk,
                     const u32* __restrict__ cand_off, const u32* __restrict__ cand_cnt, const u64* __restrict__ cand_rec,
                     u64 ncand, u64 cand_lo, u64 cand_hi, unsigned char* __restrict__ out, u64 rec_limit,
                     u32* __restrict__ cand_next, u64* __restrict__ violations, const u32* __restrict__ collapsed)
{
    __shared__ u32 scratch[8 * EMITC_THREADS];
    __shared__ StepTables T;
    // Record staging (chain.cuh: OutStage) is a build option, -DNNP_OUT_STAGE (160-byte vector drain) or
    // -DNNP_OUT_STAGE -DNNP_BULK_STORE (cp.async.bulk): both bring the kernel's DRAM traffic from 1.59x to
    // 1.11x / 1.17x of its algorithmic bytes and both make it 5-6 % slower, because the kernel is bound by
    // the integer pipe, not by memory (100 M positions: 4.80 ms plain, 5.05 ms staged, 5.09 ms bulk).
#ifdef NNP_OUT_STAGE
    __shared__ alignas(16) unsigned char stage[EMITC_THREADS * OUT_SLOT_BYTES];
    // (an output buffer that is only 8-byte aligned keeps the plain stores)
    unsigned char* const slot = (reinterpret_cast<uintptr_t>(out) & 15) ? nullptr : stage + threadIdx.x * OUT_SLOT_BYTES;
#else
    unsigned char* const slot = nullptr;
#endif
    step_tables_fill(T);
    const u64 i = cand_lo + (u64)blockIdx.x * EMITC_THREADS + threadIdx.x;
    if (i >= cand_hi) return;
    if (cand_cnt[i] == 0) {  // marked by k_mark_conflicts
        cand_next[i] = 0xFFFFFFFFu;
        return;
    }
    const u32 c = cand_chunk[i], off = cand_off[i];
    const u32 clen = tab.len[c];
    const u64 rec0 = cand_rec[i];
    const unsigned char* s = in + tab.start[c] + off;
    u32 consumed = 0;
    bool ok = true;
    if (collapsed && collapsed[c])
        consumed = clen / 34u * 34u;  // a chunk of single positions, listed as one entry: k_emit_heads_chunks writes it
    else
        ok = emit_chain_bin(s, clen - off - 34, out, rec0, rec_limit, scratch + threadIdx.x, EMITC_THREADS, consumed, &T, slot);
    // where the chain ends: should a link not hold anywhere, the reader's walk is resolved from these
    // (k_resolve_chunks) without decoding every candidate a second time
    cand_next[i] = ok ? off + consumed : 0xFFFFFFFFu;
    if (rec0 + cand_cnt[i] > rec_limit) ok = false;  // (a false candidate's count pushed records past the buffer)
    if (!reader_links_hold(cand_chunk, cand_off, cand_cnt, ncand, i, off + consumed, clen)) ok = false;
    if (!ok) atomicAdd(violations, 1ull);
}

// every chunk needs at least one candidate (its first chain); chunks without any are violations
__global__ void k_check_chunks(ChunkTable tab, const u64* __restrict__ tile_prefix, u64* __restrict__ violations)
{
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= tab.info->chunks) return;
    if (tile_prefix[tab.tile_base[c]] == tile_prefix[tab.tile_base[c + 1]]) atomicAdd(violations, 1ull);
}

// the same walk, writing emitPlainEntry text (decompressPlain :1299-1335)
__global__ void __launch_bounds__(EMITC_THREADS)
k_emit_chains_text(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ cand_chunk,
                   const u32* __restrict__ cand_off, const u32* __restrict__ cand_base, const u64* __restrict__ cand_tbase,
                   u64 ncand, const u64* __restrict__ chunk_tbase, unsigned char* __restrict__ out, DecompressTotals* tot)
{
    __shared__ uint4 windows[4 * EMITC_THREADS];
    const u64 i = (u64)blockIdx.x * EMITC_THREADS + threadIdx.x;
    if (i >= ncand) return;
    if (cand_base[i] == 0xFFFFFFFFu) return;
    const u32 c = cand_chunk[i], off = cand_off[i];
    const unsigned char* s = in + tab.start[c] + off;
    BufferedSink sink;
    sink.init(out + chunk_tbase[c] + cand_tbase[i], windows + threadIdx.x, EMITC_THREADS);
    u32 consumed = 0;
    const bool ok = walk_chain(s, tab.len[c] - off - 34,
                               [&](const ChainCursor& cc, u32) { put_plain_entry(sink, cc.pos, cc.mv, cc.score, cc.ply, cc.result); },
                               consumed);
    sink.finish();
    if (!ok) atomicMin(&tot->error_chunk, (u64)c);
}

// sequential fallback, one thread per flagged chunk
__global__ void k_slow_count(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ chunk_slow,
                             u32* __restrict__ chunk_count, u64* __restrict__ chunk_tbytes, DecompressTotals* tot)
{
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= tab.info->chunks || !chunk_slow[c]) return;
    const u32 clen = tab.len[c];
    const unsigned char* base = in + tab.start[c];
    u32 cur = 0, count = 0;
    u64 tbytes = 0;
    while ((u64)cur + 34 <= clen) {
        u32 consumed = 0;
        const bool ok = walk_chain(base + cur, clen - cur - 34,
                                   [&](const ChainCursor& cc, u32) { ++count; if (chunk_tbytes) tbytes += plain_len(cc); },
                                   consumed);
        if (!ok) { atomicMin(&tot->error_chunk, (u64)c); break; }
        cur += consumed;
    }
    chunk_count[c] = count;
    if (chunk_tbytes) chunk_tbytes[c] = tbytes;
    atomicAdd(&tot->slow_chunks, 1ull);
}
__global__ void k_slow_emit(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ chunk_slow,
                            const u64* __restrict__ chunk_base, unsigned char* __restrict__ out)
{
    __shared__ u32 scratch[8 * 32];  // launched with 32 threads per block
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= tab.info->chunks || !chunk_slow[c]) return;
    const u32 clen = tab.len[c];
    const unsigned char* base = in + tab.start[c];
    u32 cur = 0;
    u64 rec = chunk_base[c];
    const u64 rec_end = chunk_base[c + 1];
    while ((u64)cur + 34 <= clen) {
        u32 consumed = 0;
        const u32 plies = ((u32)base[cur + 32] << 8) | (u32)base[cur + 33];
        if (!emit_chain_bin(base + cur, clen - cur - 34, out, rec, rec_end, scratch + threadIdx.x, 32, consumed)) break;
        rec += 1 + plies;
        cur += consumed;
    }
}

__global__ void k_slow_emit_text(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ chunk_slow,
                                 const u64* __restrict__ chunk_tbase, unsigned char* __restrict__ out)
{
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= tab.info->chunks || !chunk_slow[c]) return;
    const u32 clen = tab.len[c];
    const unsigned char* base = in + tab.start[c];
    u32 cur = 0;
    WriteSink sink{out + chunk_tbase[c]};
    unsigned char* const end = out + chunk_tbase[c + 1];
    while ((u64)cur + 34 <= clen) {
        u32 consumed = 0;
        const bool ok = walk_chain(base + cur, clen - cur - 34,
                                   [&](const ChainCursor& cc, u32) {
                                       if (sink.p + plain_len(cc) <= end) put_plain_entry(sink, cc.pos, cc.mv, cc.score, cc.ply, cc.result);
                                   },
                                   consumed);
        if (!ok) break;
        cur += consumed;
    }
}

// ------------------------------------------------------------------ host launchers

void init_tables_decompress(cudaStream_t s) { k_step_tables_init<<<1, 256, 0, s>>>(); }


size_t walk_scratch_bytes() { return sizeof(SegScratch); }
// `scratch` (walk_scratch_bytes(), 16-byte aligned, may be null): enables the segmented walk for large files
void launch_walk_chunks(const void* d_in, u64 n, ChunkTable tab, u64 max_chunks, u32 world, u32 rank, void* scratch, cudaStream_t s)
{
    const unsigned char* in = (const unsigned char*)d_in;
    const u64* done = nullptr;
    if (scratch && n >= 8 * g_seg_bytes && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0) {
        SegScratch* S = reinterpret_cast<SegScratch*>(scratch);
        u64 segs64 = n / g_seg_bytes;
        const u32 segs = (u32)(segs64 > SEG_MAX ? SEG_MAX : segs64);
        cudaMemsetAsync(S->start, 0xFF, sizeof(u64) * (segs + 1), s);
        k_seg_find<<<segs * SEG_FIND_PARTS, SEG_FIND_THREADS, 0, s>>>(in, n, segs, S);
        k_seg_walk<<<segs, 32, 0, s>>>(in, n, segs, S, tab, max_chunks, 0);
        k_seg_prefix<<<1, 32, 0, s>>>(segs, S, tab, max_chunks);
        k_seg_walk<<<segs, 32, 0, s>>>(in, n, segs, S, tab, max_chunks, 1);
        k_seg_finish<<<1, 32, 0, s>>>(n, segs, S, tab, world, rank);
        done = &S->done;
    }
    k_walk_chunks<<<1, 32, 0, s>>>(in, n, tab, max_chunks, world, rank, done);
}
// k_collapsed_tiles: in collapse mode a chunk of single positions is ONE entry of the candidate list (offset 0, as
// many positions as the chunk has stems; k_emit_heads_chunks writes its records): its first tile lists bit 0, its
// other tiles nothing, and k_candidates_scan is not launched for any of them (scan_tiles[c] = 0).
__global__ void __launch_bounds__(CAND_TILE / 32)
k_collapsed_tiles(ChunkTable tab, const u32* __restrict__ chunk_flag, u32* __restrict__ tile_count,
                  u32* __restrict__ tile_flags, u32* __restrict__ scan_tiles)
{
    const u64 c = blockIdx.x;
    const u64 tb = tab.tile_base[c], ntiles = tab.tile_base[c + 1] - tb;
    if (!chunk_flag[c]) {
        if (threadIdx.x == 0) scan_tiles[c] = (u32)ntiles;
        return;
    }
    if (threadIdx.x == 0) scan_tiles[c] = 0;
    tile_flags[tb * (CAND_TILE / 32) + threadIdx.x] = threadIdx.x == 0 ? 1u : 0u;
    for (u64 j = threadIdx.x; j < ntiles; j += CAND_TILE / 32) tile_count[tb + j] = j == 0 ? 1u : 0u;
}

// k_emit_heads_chunks: the records of the chunks that the candidate list holds as one entry each (`collapsed`):
// record k of chunk c is the stem at byte 34 * k, its index the entry's record offset + k. blockIdx.x = chunk,
// blockIdx.y strides over the chunk's stems.
__global__ void __launch_bounds__(HEADS_EMIT_THREADS, HEADS_EMIT_MIN_BLOCKS)
k_emit_heads_chunks(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ collapsed,
                    const u64* __restrict__ tile_prefix, const u64* __restrict__ cand_rec, unsigned char* __restrict__ out,
                    u64 rec_limit)
{
    __shared__ u32 scratch[8 * HEADS_EMIT_THREADS];
    __shared__ StepTables T;
    const u64 c = blockIdx.x;
    if (!collapsed[c]) return;
    const u32 stems = tab.len[c] / 34u;
    if ((u64)blockIdx.y * HEADS_EMIT_THREADS >= stems) return;
    step_tables_fill(T);
    const u64 rec0 = cand_rec[tile_prefix[tab.tile_base[c]]];  // the chunk's entry is the first of its first tile
    const unsigned char* s = in + tab.start[c];
    for (u32 k = blockIdx.y * HEADS_EMIT_THREADS + threadIdx.x; k < stems; k += gridDim.y * HEADS_EMIT_THREADS) {
        u32 consumed = 0;
        emit_chain_bin(s + 34ull * k, 0u, out, rec0 + k, rec_limit, scratch + threadIdx.x, HEADS_EMIT_THREADS, consumed, &T, nullptr);
    }
}

void launch_chunk_heads_only(const void* d_in, ChunkTable tab, u64 chunks, u32* chunk_flag, u32* chunk_stems, u64* flagged,
                             cudaStream_t s)
{
    if (chunks == 0) return;
    k_chunk_heads_only<<<(unsigned)chunks, HEADS_ONLY_THREADS, 0, s>>>((const unsigned char*)d_in, tab, chunk_flag, chunk_stems,
                                                                       flagged);
}
void launch_emit_heads_only(const void* d_in, ChunkTable tab, u64 chunks, const u64* chunk_base, u64 positions, void* d_out,
                            u64 rec_limit, cudaStream_t s)
{
    if (positions == 0) return;
    k_emit_heads_only<<<(unsigned)((positions + HEADS_EMIT_THREADS - 1) / HEADS_EMIT_THREADS), HEADS_EMIT_THREADS, 0, s>>>(
        (const unsigned char*)d_in, tab, chunks, chunk_base, (unsigned char*)d_out, rec_limit);
}
// (chunk_flag: filled by launch_chunk_heads_only beforehand; scan_prefix / scan_total: collapse mode, see
// launch_collapsed_tiles)
void launch_candidates_scan(const void* d_in, u64 n_in, ChunkTable tab, u64 chunks, u64 tiles, u32* tile_count, u32* tile_flags,
                            u32 debug_reject_mod, const u32* chunk_flag, const u64* scan_prefix, u64 scan_total, cudaStream_t s)
{
    if (tiles == 0 || chunks == 0) return;
    const u64 blocks = scan_prefix ? scan_total : tiles;
    if (blocks == 0) return;
    k_candidates_scan<<<(unsigned)blocks, CAND_THREADS, 0, s>>>((const unsigned char*)d_in, n_in, tab, tile_count, tile_flags,
                                                              debug_reject_mod, chunk_flag, scan_prefix);
}
void launch_collapsed_tiles(ChunkTable tab, u64 chunks, const u32* chunk_flag, u32* tile_count, u32* tile_flags, u32* scan_tiles,
                            cudaStream_t s)
{
    if (chunks == 0) return;
    k_collapsed_tiles<<<(unsigned)chunks, CAND_TILE / 32, 0, s>>>(tab, chunk_flag, tile_count, tile_flags, scan_tiles);
}
void launch_emit_heads_chunks(const void* d_in, ChunkTable tab, u64 chunks, const u32* collapsed, const u64* tile_prefix,
                              const u64* cand_rec, void* d_out, u64 rec_limit, cudaStream_t s)
{
    if (chunks == 0) return;
    // 256 blocks of 128 stems cover the 30 841 stems of an ordinary chunk in one pass; larger chunks are strided
    k_emit_heads_chunks<<<dim3((unsigned)chunks, 256), HEADS_EMIT_THREADS, 0, s>>>((const unsigned char*)d_in, tab, collapsed,
                                                                                   tile_prefix, cand_rec, (unsigned char*)d_out,
                                                                                   rec_limit);
}
void launch_mark_conflicts(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off, u32* cand_cnt,
                           u64 ncand, cudaStream_t s)
{
    if (ncand < 2) return;
    k_mark_conflicts<<<(unsigned)((ncand + 127) / 128), 128, 0, s>>>((const unsigned char*)d_in, tab, cand_chunk, cand_off,
                                                                    cand_cnt, ncand);
}
void launch_candidates_list(const void* d_in, ChunkTable tab, u64 tiles, const u32* tile_flags, const u64* tile_prefix,
                            u32* cand_chunk, u32* cand_off, u32* cand_cnt, const u32* collapsed, cudaStream_t s)
{
    if (tiles == 0) return;
    k_candidates_list<<<(unsigned)tiles, LIST_THREADS, 0, s>>>((const unsigned char*)d_in, tab, tile_flags, tile_prefix,
                                                             cand_chunk, cand_off, cand_cnt, collapsed);
}
void launch_check_chunks(ChunkTable tab, u64 chunks, const u64* tile_prefix, u64* violations, cudaStream_t s)
{
    if (chunks == 0) return;
    k_check_chunks<<<(unsigned)((chunks + 127) / 128), 128, 0, s>>>(tab, tile_prefix, violations);
}
void launch_emit_chains_verify(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off,
                               const u32* cand_cnt, const u64* cand_rec, u64 ncand, u64 cand_lo, u64 cand_hi, void* out,
                               u64 rec_limit, u32* cand_next, u64* violations, const u32* collapsed, cudaStream_t s)
{
    if (cand_hi <= cand_lo) return;
    k_emit_chains_verify<<<(unsigned)((cand_hi - cand_lo + EMITC_THREADS - 1) / EMITC_THREADS), EMITC_THREADS, 0, s>>>(
        (const unsigned char*)d_in, tab, cand_chunk, cand_off, cand_cnt, cand_rec, ncand, cand_lo, cand_hi, (unsigned char*)out,
        rec_limit, cand_next, violations, collapsed);
}
void launch_exclusive_sum(const u32* in, u64 n, u64* out, cudaStream_t s)
{
    k_exclusive_sum<<<1, SUM_THREADS, 0, s>>>(in, n, out);
}
void launch_exclusive_sum64(const u64* in, u64 n, u64* out, cudaStream_t s)
{
    k_exclusive_sum64<<<1, SUM_THREADS, 0, s>>>(in, n, out);
}
void launch_probe_chains(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off, u64 ncand,
                         u32* cand_next, u32* cand_cnt, u32* cand_tlen, cudaStream_t s)
{
    if (ncand == 0) return;
    const unsigned blocks = (unsigned)((ncand + PROBE_THREADS - 1) / PROBE_THREADS);
    if (cand_tlen)
        k_probe_chains<true><<<blocks, PROBE_THREADS, 0, s>>>((const unsigned char*)d_in, tab, cand_chunk, cand_off, ncand,
                                                             cand_next, cand_cnt, cand_tlen);
    else
        k_probe_chains<false><<<blocks, PROBE_THREADS, 0, s>>>((const unsigned char*)d_in, tab, cand_chunk, cand_off, ncand,
                                                              cand_next, cand_cnt, cand_tlen);
}
void launch_resolve_chunks(ChunkTable tab, u64 chunks, const u64* tile_prefix, const u32* cand_off, const u32* cand_next,
                           const u32* cand_cnt, u32* cand_base, u32* chunk_count, u32* chunk_slow, const u32* cand_tlen,
                           u64* cand_tbase, u64* chunk_tbytes, cudaStream_t s)
{
    if (chunks == 0) return;
    k_resolve_chunks<<<(unsigned)((chunks + RESOLVE_WARPS - 1) / RESOLVE_WARPS), RESOLVE_WARPS * 32, 0, s>>>(
        tab, tile_prefix, cand_off, cand_next, cand_cnt, cand_base, chunk_count, chunk_slow, cand_tlen, cand_tbase, chunk_tbytes);
}
void launch_slow_count(const void* d_in, ChunkTable tab, u64 chunks, const u32* chunk_slow, u32* chunk_count,
                       u64* chunk_tbytes, DecompressTotals* tot, cudaStream_t s)
{
    if (chunks == 0) return;
    k_slow_count<<<(unsigned)((chunks + 31) / 32), 32, 0, s>>>((const unsigned char*)d_in, tab, chunk_slow, chunk_count,
                                                               chunk_tbytes, tot);
}
void launch_emit_chains_text(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off, const u32* cand_base,
                             const u64* cand_tbase, u64 ncand, const u64* chunk_tbase, void* out, DecompressTotals* tot,
                             cudaStream_t s)
{
    if (ncand == 0) return;
    k_emit_chains_text<<<(unsigned)((ncand + EMITC_THREADS - 1) / EMITC_THREADS), EMITC_THREADS, 0, s>>>(
        (const unsigned char*)d_in, tab, cand_chunk, cand_off, cand_base, cand_tbase, ncand, chunk_tbase, (unsigned char*)out, tot);
}
void launch_slow_emit_text(const void* d_in, ChunkTable tab, u64 chunks, const u32* chunk_slow, const u64* chunk_tbase,
                           void* out, cudaStream_t s)
{
    if (chunks == 0) return;
    k_slow_emit_text<<<(unsigned)((chunks + 31) / 32), 32, 0, s>>>((const unsigned char*)d_in, tab, chunk_slow, chunk_tbase,
                                                                   (unsigned char*)out);
}
void launch_emit_chains(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off, const u32* cand_base,
                        u64 ncand, const u64* chunk_base, void* out, DecompressTotals* tot, const u64* placed_rec,
                        const u32* placed_next, cudaStream_t s)
{
    if (ncand == 0) return;
    k_emit_chains<<<(unsigned)((ncand + EMITC_THREADS - 1) / EMITC_THREADS), EMITC_THREADS, 0, s>>>(
        (const unsigned char*)d_in, tab, cand_chunk, cand_off, cand_base, ncand, chunk_base, (unsigned char*)out, tot, placed_rec,
        placed_next);
}
void launch_slow_emit(const void* d_in, ChunkTable tab, u64 chunks, const u32* chunk_slow, const u64* chunk_base, void* out,
                      cudaStream_t s)
{
    if (chunks == 0) return;
    k_slow_emit<<<(unsigned)((chunks + 31) / 32), 32, 0, s>>>((const unsigned char*)d_in, tab, chunk_slow, chunk_base,
                                                              (unsigned char*)out);
}

}  // namespace nnp
