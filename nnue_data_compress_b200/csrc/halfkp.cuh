// halfkp.cuh -- HalfKP feature rows written straight from a decoded position.
//
// The consumer of a .binpack is the NNUE trainer's data loader (the reference's README.md:3 points at
// it); what it feeds the network per position is, for each perspective, the list of active
// (own king square, piece square, piece kind) features. Emitting those rows from the chain walk
// skips the 40-byte .bin record and the Huffman decode the loader would do on it.
//
// The reference holds no feature code, so the index is the published HalfKP definition of the
// Stockfish NNUE trainers (nodchip's learner `kpp_board_index` / nnue-pytorch `halfkp_idx`):
//
//   orient(P, sq) = P == white ? sq : sq ^ 63                       (black sees the board rotated)
//   kind(P, pc)   = 2 * type(pc) + (colour(pc) != P)                (pawn .. queen = 0 .. 4; no kings)
//   index         = 1 + orient(P, sq) + 64 * kind + 641 * orient(P, king square of P)
//
// 641 = 10 * 64 + 1 planes per king square, 64 * 641 = 41024 features. Entry j of the white and
// of the black row describe the same piece; rows are padded with -1 to HALFKP_ROW entries. Order of
// the entries: a chain head's row lists the pieces by kind as white sees it (white pawn, black pawn,
// white knight, ... black queen) and by ascending square within a kind, so the white row ascends; a
// .bin record's row lists them as its Huffman stream does (rank 8 first, files a to h). Along a chain the row is UPDATED, not rebuilt (a move touches one
// or two entries; rebuilding costs more than decoding the move): a piece keeps its slot while it
// stands, and the slot of a captured piece is taken over by the row's last entry. The order is thus
// deterministic but depends on the chain's history; the consumer (a sparse feature transformer sums
// over the row) does not depend on it. A side without a king (malformed input only) uses king
// square 0 like the packing path (stream_from_pos); pieces beyond the 32nd (by kind, square) are dropped.
#pragma once
#include "chess.cuh"

namespace nnp {

constexpr int HALFKP_ROW = 32;
constexpr int HALFKP_PLANES = 641;
constexpr int HALFKP_STAGE = HALFKP_ROW + 4;  // words per staged row: 16-byte aligned rows, eight lanes of a row read 128 B

struct HalfKpOut {
    int* white;   // [positions][HALFKP_ROW]
    int* black;   // [positions][HALFKP_ROW]
    uint2* meta;  // [positions] nnp_halfkp_meta
};

// A staged row: x[j] = 64 * kind + square of entry j (kind = 2 * type + colour = the piece code) in
// shared memory, king-independent, so that one value serves both perspectives:
//   white index = wbase + x,  black index = bbase + (x ^ 127)     (127 flips the colour bit of the kind
//   and rotates the square), wbase = 1 + 641 * wk, bbase = 1 + 641 * (bk ^ 63).
struct HalfKpRow {
    int n;             // entries in use
    int wbase, bbase;
};

__device__ __forceinline__ void halfkp_bases(const Pos& p, HalfKpRow& R)
{
    const u64 kings = pos_type_bb(p, PT_KING);
    const u64 wkb = kings & p.occ[0], bkb = kings & p.occ[1];
    const int wk = wkb ? lsb64(wkb) : 0, bk = bkb ? lsb64(bkb) : 0;
    R.wbase = 1 + HALFKP_PLANES * wk;
    R.bbase = 1 + HALFKP_PLANES * (bk ^ 63);
}

// The row of `p` from scratch, in (kind, square) order. `map` (optional) = the lane's square -> slot
// bytes, `map_stride` apart, kept for halfkp_apply_move.
template <bool WITH_MAP>
__device__ __forceinline__ void halfkp_rebuild(const Pos& p, HalfKpRow& R, int* x, unsigned char* map, int map_stride)
{
    int n = 0;
#pragma unroll
    for (int t = PT_PAWN; t <= PT_QUEEN; ++t) {
        const u64 tm = ((t & 1) ? p.t0 : ~p.t0) & ((t & 2) ? p.t1 : ~p.t1) & ((t & 4) ? p.t2 : ~p.t2);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const u64 bb = tm & p.occ[c];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                u32 m = (u32)(bb >> (32 * h));
                while (m && n < HALFKP_ROW) {
                    const int s = __ffs((int)m) - 1 + 32 * h;
                    m &= m - 1;
                    x[n] = 64 * (2 * t + c) + s;
                    if (WITH_MAP) map[s * map_stride] = (unsigned char)n;
                    ++n;
                }
            }
        }
    }
    R.n = n;
    halfkp_bases(p, R);
}

// Applies move `m`, about to be made in position `p` (the position BEFORE the move; `moved` = the piece on
// m.from), to the staged row: the token changes of Board::doMove (Position.h:300-439) as restated in
// board_do_move. Returns false for moves outside its domain (kings captured or promoted, pieces
// appearing from nothing, castling onto occupied squares: nothing a writer emits from legal games);
// the caller then rebuilds the row from the position after the move. King squares are not part of
// the staged values: the caller refreshes the bases with halfkp_bases.
__device__ __forceinline__ bool halfkp_apply_move(const Pos& p, const Move& m, int moved, HalfKpRow& R, int* x,
                                                  unsigned char* map, int map_stride)
{
    const int from = m.from, to = m.to;
    if (moved == NO_PIECE || from == to) return false;
    const u64 all = pos_all(p);
    const u64 kings = pos_type_bb(p, PT_KING);
    auto occupied = [&](int sq) { return (int)((all >> sq) & 1); };
    auto has_king = [&](int sq) { return (int)((kings >> sq) & 1); };
    const bool king_moves = (moved >> 1) == PT_KING;
    int gone = -1;                   // square whose piece disappears from the row
    int src = -1, dst = 0, kind = 0;  // the piece on src stands on dst afterwards, as `kind`
    if (m.type == MT_CASTLE) {
        if (!king_moves || !occupied(to) || has_king(to)) return false;
        const int base = (moved & 1) ? 56 : 0;
        const bool is_short = (to & 7) == 7;
        const int rt = base + (is_short ? 5 : 3), kt = base + (is_short ? 6 : 2);
        if ((occupied(rt) && rt != from && rt != to) || (occupied(kt) && kt != from && kt != to)) return false;
        src = to;
        dst = rt;
        kind = pos_piece_at(p, to);
    } else {
        if (has_king(to)) return false;
        if (m.type == MT_ENPASSANT) {
            const int cap = (to & 7) | (from & 56);
            if (king_moves || occupied(to) || cap == from || cap == to || has_king(cap)) return false;
            if (occupied(cap)) gone = cap;
            src = from; dst = to; kind = moved;
        } else {
            if (occupied(to)) gone = to;
            if (m.type == MT_PROMOTION) {
                if (king_moves || m.promo == NO_PIECE) return false;
                src = from; dst = to; kind = m.promo;
            } else if (!king_moves) {
                src = from; dst = to; kind = moved;
            }
        }
    }
    int n = R.n;
    if (gone >= 0) {  // the last entry takes over the slot
        const int c = map[gone * map_stride], last = n - 1;
        const int xl = x[last];
        x[c] = xl;
        map[(xl & 63) * map_stride] = (unsigned char)c;
        n = last;
    }
    if (src >= 0) {
        const int sl = map[src * map_stride];
        x[sl] = 64 * kind + dst;
        map[dst * map_stride] = (unsigned char)sl;
    }
    R.n = n;
    return true;
}

// nnp_halfkp_meta: int16 score, uint16 ply, int8 result (narrowed as in
// trainingDataEntryToPackedSfenValue :570-585), uint8 stm, uint8 n_active, uint8 0
__device__ __forceinline__ uint2 halfkp_meta(const Pos& p, int score, int ply, int result, int n)
{
    return make_uint2(((u32)score & 0xFFFFu) | (((u32)ply & 0xFFFFu) << 16),
                      ((u32)result & 0xFFu) | ((u32)p.stm << 8) | ((u32)n << 16));
}

#ifndef NNP_HOST_SIM
// The warp writes the staged rows of its lanes in `mask` (lane r's rows go to record `rec` of lane r):
// eight lanes per row, so that one 16-byte store instruction covers four whole 128-byte rows instead
// of a sixteenth of 32 different ones. All 32 lanes must call this (between __syncwarp()s).
__device__ __forceinline__ void halfkp_store_warp(u32 mask, u64 rec, const HalfKpRow& R, const int* warp_x, const HalfKpOut& o)
{
    const int lane = threadIdx.x & 31;
    const int sub = lane >> 3, chunk = lane & 7, j = 4 * chunk;
    const u32 bases = (u32)R.wbase | ((u32)R.bbase << 16);  // both below 1 + 641 * 64
#pragma unroll 4
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + sub;
        const u64 rr = __shfl_sync(0xffffffffu, rec, r);
        const int left = __shfl_sync(0xffffffffu, R.n, r) - j;  // entries of the row at or behind this lane's first
        const u32 pk = __shfl_sync(0xffffffffu, bases, r);
        if (!((mask >> (4 * i)) & 15u)) continue;
        if ((mask >> r) & 1u) {
            const int4 x = *reinterpret_cast<const int4*>(warp_x + r * HALFKP_STAGE + j);
            const int wb = (int)(pk & 0xFFFFu), bb = (int)(pk >> 16);
            const int m0 = left > 0 ? 0 : -1, m1 = left > 1 ? 0 : -1, m2 = left > 2 ? 0 : -1, m3 = left > 3 ? 0 : -1;  // padding
            reinterpret_cast<int4*>(o.white + rr * HALFKP_ROW)[chunk] =
                make_int4((wb + x.x) | m0, (wb + x.y) | m1, (wb + x.z) | m2, (wb + x.w) | m3);
            reinterpret_cast<int4*>(o.black + rr * HALFKP_ROW)[chunk] =
                make_int4((bb + (x.x ^ 127)) | m0, (bb + (x.y ^ 127)) | m1, (bb + (x.z ^ 127)) | m2, (bb + (x.w ^ 127)) | m3);
        }
    }
}

#endif  // NNP_HOST_SIM

// one thread on its own (the sequential fallback): every row from scratch
__device__ __forceinline__ void halfkp_emit_thread(const Pos& p, int score, int ply, int result, u64 rec, const HalfKpOut& o, int* x)
{
    HalfKpRow R;
    halfkp_rebuild<false>(p, R, x, nullptr, 0);
    for (int j = 0; j < HALFKP_ROW; ++j) {
        o.white[rec * HALFKP_ROW + j] = j < R.n ? R.wbase + x[j] : -1;
        o.black[rec * HALFKP_ROW + j] = j < R.n ? R.bbase + (x[j] ^ 127) : -1;
    }
    o.meta[rec] = halfkp_meta(p, score, ply, result, R.n);
}

}  // namespace nnp
