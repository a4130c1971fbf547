// halfkp.cu -- .binpack / .bin -> HalfKP feature rows (the decode fused into its
// consumer). The chain kernels are the ones of decompress.cu with the record writer replaced:
//
//   k_emit_chains_halfkp_verify   optimistic strategy (one walk, links verified, see decompress.cu)
//   k_emit_chains_halfkp          exhaustive strategy, after probe / resolve
//   k_slow_emit_halfkp            sequential per-chunk net
//   k_bin_halfkp                  one thread per 40-byte record (pos_from_packed_sfen :364-446), the row
//                                 listed from the record's piece tokens as they are decoded
//
// Rows are written at the position's index in reader order, i.e. row r describes the record r of the
// .bin file decompressBin (:1376-1412) writes for the same input.
#include "common.cuh"
#include "kernels.h"
#include "chain.cuh"
#include "halfkp.cuh"
#include "verify.cuh"

namespace nnp {

constexpr int HKP_THREADS = 128;
#ifndef HKP_MIN_BLOCKS
#define HKP_MIN_BLOCKS 5
#endif

// One chain per lane, the warp in lock-step over the plies (lanes whose chain is shorter idle, as they
// would in a divergent loop). The row of the chain head is built from its position; every ply then
// updates it in place (halfkp_apply_move) and after every step the warp writes the rows of all its
// lanes together (halfkp_store_warp). The walk is walk_chain's (chain.cuh). All 32 lanes must call
// this; `valid` says whether the lane has a chain. Rows rec0, rec0 + 1, ... are written.
// warp_x: the warp's [32][HALFKP_STAGE] staged rows, warp_map: its [64][32] square -> slot bytes.
__device__ __forceinline__ bool warp_emit_chains_halfkp(bool valid, const unsigned char* s, u32 bytes_after_stem,
                                                        const HalfKpOut& o, u64 rec0, int* warp_x, unsigned char* warp_map,
                                                        u32& consumed, const StepTables* T)
{
    const int lane = threadIdx.x & 31;
    int* x = warp_x + lane * HALFKP_STAGE;
    unsigned char* map = warp_map + lane;
    ChainCursor cc;
    BitReader r;
    HalfKpRow R;
    R.n = R.wbase = R.bbase = 0;
    u32 steps = 0;
    if (valid) {
        chain_open(s, cc);
        r.init(s + 34, (u64)bytes_after_stem);
        steps = cc.num_plies + 1;
    }
    bool ok = true;
    const u32 max_steps = __reduce_max_sync(0xffffffffu, steps);
    for (u32 k = 0; k < max_steps; ++k) {
        bool act = ok && k < steps;
        bool rebuild = k == 0;
        if (act && k > 0) {
            if (cc.mv.from > 63 || cc.mv.to > 63) {
                ok = act = false;
            } else {
                const int moved = pos_piece_at(cc.pos, cc.mv.from);
                rebuild = !halfkp_apply_move(cc.pos, cc.mv, moved, R, x, map, 32);
                if (!chain_step(cc, r, false, moved, T)) ok = act = false;
            }
        }
        if (act) {
            const int pieces = popc64(pos_all(cc.pos) & ~pos_type_bb(cc.pos, PT_KING));
            if (rebuild || pieces != R.n) halfkp_rebuild<true>(cc.pos, R, x, map, 32);  // also: more than 32 pieces
            else halfkp_bases(cc.pos, R);
            o.meta[rec0 + k] = halfkp_meta(cc.pos, cc.score, cc.ply, cc.result, R.n);
        }
        __syncwarp();
        halfkp_store_warp(__ballot_sync(0xffffffffu, act), rec0 + k, R, warp_x, o);
        __syncwarp();
    }
    if (valid && ok) consumed = 34 + ((r.pos + 7) >> 3);
    return ok;
}

__global__ void __launch_bounds__(HKP_THREADS, HKP_MIN_BLOCKS)
k_emit_chains_halfkp_verify(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ cand_chunk,
                            const u32* __restrict__ cand_off, const u32* __restrict__ cand_cnt,
                            const u64* __restrict__ cand_rec, u64 ncand, HalfKpOut out, u64* __restrict__ violations)
{
    __shared__ __align__(16) int stage[HALFKP_STAGE * HKP_THREADS];
    __shared__ unsigned char maps[64 * HKP_THREADS];
    __shared__ StepTables T;
    step_tables_fill(T);
    const u64 i = (u64)blockIdx.x * HKP_THREADS + threadIdx.x;
    const bool valid = i < ncand && cand_cnt[i] != 0;  // 0: marked by k_mark_conflicts
    u32 c = 0, off = 0, clen = 34;
    u64 rec0 = 0;
    if (valid) {
        c = cand_chunk[i];
        off = cand_off[i];
        clen = tab.len[c];
        rec0 = cand_rec[i];
    }
    u32 consumed = 0;
    bool ok = warp_emit_chains_halfkp(valid, in + (valid ? tab.start[c] + off : 0), clen - off - 34, out, rec0,
                                      stage + (threadIdx.x & ~31) * HALFKP_STAGE, maps + (threadIdx.x & ~31) * 64, consumed, &T);
    if (!valid) return;
    if (!reader_links_hold(cand_chunk, cand_off, cand_cnt, ncand, i, off + consumed, clen)) ok = false;
    if (!ok) atomicAdd(violations, 1ull);
}

__global__ void __launch_bounds__(HKP_THREADS, HKP_MIN_BLOCKS)
k_emit_chains_halfkp(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ cand_chunk,
                     const u32* __restrict__ cand_off, const u32* __restrict__ cand_base, u64 ncand,
                     const u64* __restrict__ chunk_base, HalfKpOut out, DecompressTotals* tot)
{
    __shared__ __align__(16) int stage[HALFKP_STAGE * HKP_THREADS];
    __shared__ unsigned char maps[64 * HKP_THREADS];
    __shared__ StepTables T;
    step_tables_fill(T);
    const u64 i = (u64)blockIdx.x * HKP_THREADS + threadIdx.x;
    const bool valid = i < ncand && cand_base[i] != 0xFFFFFFFFu;  // else: not a chain start
    u32 c = 0, off = 0, clen = 34;
    u64 rec0 = 0;
    if (valid) {
        c = cand_chunk[i];
        off = cand_off[i];
        clen = tab.len[c];
        rec0 = chunk_base[c] + cand_base[i];
    }
    u32 consumed = 0;
    const bool ok = warp_emit_chains_halfkp(valid, in + (valid ? tab.start[c] + off : 0), clen - off - 34, out, rec0,
                                            stage + (threadIdx.x & ~31) * HALFKP_STAGE, maps + (threadIdx.x & ~31) * 64, consumed,
                                            &T);
    if (valid && !ok) atomicMin(&tot->error_chunk, (u64)c);
}

// sequential fallback, one thread per flagged chunk; launched with 32 threads per block
__global__ void k_slow_emit_halfkp(const unsigned char* __restrict__ in, ChunkTable tab, const u32* __restrict__ chunk_slow,
                                   const u64* __restrict__ chunk_base, HalfKpOut out)
{
    __shared__ int stage[HALFKP_STAGE * 32];
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= tab.info->chunks || !chunk_slow[c]) return;
    const u32 clen = tab.len[c];
    const unsigned char* base = in + tab.start[c];
    int* mine = stage + threadIdx.x * HALFKP_STAGE;
    u32 cur = 0;
    u64 rec = chunk_base[c];
    const u64 rec_end = chunk_base[c + 1];
    while ((u64)cur + 34 <= clen) {
        u32 consumed = 0;
        const u32 plies = ((u32)base[cur + 32] << 8) | (u32)base[cur + 33];
        const bool ok = walk_chain(
            base + cur, clen - cur - 34,
            [&](const ChainCursor& cc, u32 k) {
                if (rec + k < rec_end) halfkp_emit_thread(cc.pos, cc.score, cc.ply, cc.result, rec + k, out, mine);
            },
            consumed);
        if (!ok) break;
        rec += 1 + plies;
        cur += consumed;
    }
}

__global__ void __launch_bounds__(HKP_THREADS)
k_bin_halfkp(const unsigned char* __restrict__ bin, u64 n, HalfKpOut out, CompressTotals* tot)
{
    __shared__ __align__(16) int stage[HALFKP_STAGE * HKP_THREADS];
    const u64 i = (u64)blockIdx.x * HKP_THREADS + threadIdx.x;
    bool wr = false;
    HalfKpRow R;
    R.n = R.wbase = R.bbase = 0;
    if (i < n) {
        const u32* w = reinterpret_cast<const u32*>(bin + i * 40);
        Pos p;
        int* x = stage + threadIdx.x * HALFKP_STAGE;
        int listed = 0;
        // the Huffman stream names every non-king piece in turn: the row is listed while it is decoded
        wr = sfen_decode([&](int j) { return w[j]; }, p, [&](int sq, u32 tok) {
            if (listed < HALFKP_ROW) x[listed] = 64 * (int)(((tok >> 1) & 7u) * 2u + ((tok >> 4) & 1u)) + sq;
            ++listed;
        });
        if (wr) {
            const u32 w8 = w[8], w9 = w[9];
            if (listed > HALFKP_ROW) {
                halfkp_rebuild<false>(p, R, x, nullptr, 0);  // more than 32 pieces (no writer emits that): the first 32 by (kind, square)
            } else {
                R.n = listed;
                halfkp_bases(p, R);
            }
            out.meta[i] = halfkp_meta(p, (int)(short)(w8 & 0xFFFF), (int)(w9 & 0xFFFF), (int)(signed char)((w9 >> 16) & 0xFF), R.n);
        } else {
            atomicMin(&tot->error_index, i);
        }
    }
    __syncwarp();
    halfkp_store_warp(__ballot_sync(0xffffffffu, wr), i, R, stage + (threadIdx.x & ~31) * HALFKP_STAGE, out);
}

// ------------------------------------------------------------------ host launchers

void init_tables_halfkp(cudaStream_t s) { k_step_tables_init<<<1, 256, 0, s>>>(); }


static HalfKpOut make_out(int* white, int* black, void* meta) { return HalfKpOut{white, black, reinterpret_cast<uint2*>(meta)}; }

void launch_emit_chains_halfkp_verify(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off,
                                      const u32* cand_cnt, const u64* cand_rec, u64 ncand, int* white, int* black, void* meta,
                                      u64* violations, cudaStream_t s)
{
    if (ncand == 0) return;
    k_emit_chains_halfkp_verify<<<(unsigned)((ncand + HKP_THREADS - 1) / HKP_THREADS), HKP_THREADS, 0, s>>>(
        (const unsigned char*)d_in, tab, cand_chunk, cand_off, cand_cnt, cand_rec, ncand, make_out(white, black, meta), violations);
}
void launch_emit_chains_halfkp(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off,
                               const u32* cand_base, u64 ncand, const u64* chunk_base, int* white, int* black, void* meta,
                               DecompressTotals* tot, cudaStream_t s)
{
    if (ncand == 0) return;
    k_emit_chains_halfkp<<<(unsigned)((ncand + HKP_THREADS - 1) / HKP_THREADS), HKP_THREADS, 0, s>>>(
        (const unsigned char*)d_in, tab, cand_chunk, cand_off, cand_base, ncand, chunk_base, make_out(white, black, meta), tot);
}
void launch_slow_emit_halfkp(const void* d_in, ChunkTable tab, u64 chunks, const u32* chunk_slow, const u64* chunk_base,
                             int* white, int* black, void* meta, cudaStream_t s)
{
    if (chunks == 0) return;
    k_slow_emit_halfkp<<<(unsigned)((chunks + 31) / 32), 32, 0, s>>>((const unsigned char*)d_in, tab, chunk_slow, chunk_base,
                                                                     make_out(white, black, meta));
}
void launch_bin_halfkp(const void* d_bin, u64 n, int* white, int* black, void* meta, CompressTotals* tot, cudaStream_t s)
{
    if (n == 0) return;
    k_bin_halfkp<<<(unsigned)((n + HKP_THREADS - 1) / HKP_THREADS), HKP_THREADS, 0, s>>>((const unsigned char*)d_bin, n,
                                                                                        make_out(white, black, meta), tot);
}

}  // namespace nnp
