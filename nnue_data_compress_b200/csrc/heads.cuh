// heads.cuh -- a .bin record that starts a chain needs no position either: its 32-byte stem (packEntry,
// compress_file.cpp:997-1020) is the record's PackedSfen transcoded. The stream names the kings in its
// header and then has one token per non-king square, rank 8 first (pos_from_packed_sfen :364-446); the stem
// wants the occupied squares a1 -> h8 with one nibble each (Position::compress, Position.h:1374-1406).
// The squares are visited in stream order, every rank's nibbles are gathered in file order and pushed
// in FRONT of the ranks already seen (rank 1 ends up first), and the three nibbles that depend on the
// stream's tail -- rooks that still carry a castling right, the pawn an en-passant capture would take,
// the black king when black is to move -- are patched afterwards. Same bytes as sfen_decode + stem_pack
// wherever it answers HEADS_OK; everything irregular (both kings on one square, more than 32 pieces) is
// left to that route (HEADS_OTHER), malformed streams are reported as the decoder reports them
// (HEADS_BAD). Files of single positions are nothing but chain heads: k_heads_transcode (compress.cu).
#pragma once
#include "chess.cuh"

namespace nnp {

enum : int { HEADS_OK = 0, HEADS_OTHER = 1, HEADS_BAD = 2 };

// nibble k of a 32-nibble string held as two 64-bit halves
__device__ __forceinline__ int nib_get(u64 lo, u64 hi, int k) { return (int)(((k & 16) ? hi : lo) >> ((k & 15) * 4)) & 15; }
__device__ __forceinline__ void nib_set(u64& lo, u64& hi, int k, int v)
{
    const u64 m = 15ull << ((k & 15) * 4), b = (u64)v << ((k & 15) * 4);
    if (k & 16) hi = (hi & ~m) | b; else lo = (lo & ~m) | b;
}

// the position a stem's occupancy + plain piece nibbles (0..11) describe; for the en-passant test only
__device__ __forceinline__ void pos_from_nibbles(u64 occ, u64 nlo, u64 nhi, int stm, Pos& p)
{
    pos_clear(p);
    p.stm = stm;
    u64 o0 = 0, o1 = 0, t0 = 0, t1 = 0, t2 = 0;
    int k = 0;
    for (u64 b = occ; b; b &= b - 1, ++k) {
        const u64 bit = b & (0ull - b);
        const int piece = nib_get(nlo, nhi, k), t = piece >> 1;
        if (piece & 1) o1 |= bit; else o0 |= bit;
        if (t & 1) t0 |= bit;
        if (t & 2) t1 |= bit;
        if (t & 4) t2 |= bit;
    }
    p.occ[0] = o0; p.occ[1] = o1; p.t0 = t0; p.t1 = t1; p.t2 = t2;
}
static __device__ __noinline__ bool heads_ep_possible(u64 occ, u64 nlo, u64 nhi, int stm, int ep)
{
    Pos p;
    pos_from_nibbles(occ, nlo, nhi, stm, p);
    return ep_possible(p, ep, stm);  // setEpSquare Position.h:868-872 (post-move test)
}

// `W(j)` returns 32-bit word j of the 40-byte record (j < 10); `out` receives the stem's eight words
template <typename WordFn>
__device__ __forceinline__ int record_to_stem(WordFn W, u32 (&out)[8])
{
    const u32 w0 = W(0), w1 = W(1);
    const int stm = (int)(w0 & 1u), wk = (int)((w0 >> 1) & 63u), bk = (int)((w0 >> 7) & 63u);
    if (wk == bk) return HEADS_OTHER;
    // the stream as a 64-bit window, low bit next, refilled a word at a time
    u64 win = (((u64)w1 << 32) | w0) >> 13;
    int avail = 51, nextw = 2;
    u64 occ = 0, nlo = 0, nhi = 0;
    int pieces = 0;
    u32 err = 0;
#pragma unroll 1
    for (int r = 7; r >= 0; --r) {
        u32 rank_nibs = 0, rank_occ = 0;
        int cnt = 0;
#pragma unroll
        for (int f = 0; f < 8; ++f) {
            const int sq = 8 * r + f;
            if (avail < 32) {  // room for a word: at least five bits are always there
                const u32 nw = nextw < 10 ? W(nextw) : 0u;
                ++nextw;
                win |= (u64)nw << avail;
                avail += 32;
            }
            int nib = -1;
            if (sq == wk) nib = (PT_KING << 1) | WHITE;            // 10
            else if (sq == bk) nib = (PT_KING << 1) | BLACK;      // 11; patched below when black is to move
            else {
                const u32 tok = (u32)win;  // '0', or 1 + type (3 bits, LSB first) + colour
                const bool piece = tok & 1u;
                if (piece) {
                    const u32 type = (tok >> 1) & 7u;
                    err |= type > (u32)PT_QUEEN ? 1u : 0u;
                    nib = (int)((type << 1) | ((tok >> 4) & 1u));
                }
                const int used = piece ? 5 : 1;
                win >>= used;
                avail -= used;
            }
            if (nib >= 0) {
                rank_nibs |= (u32)nib << (4 * cnt);
                ++cnt;
                rank_occ |= 1u << f;
            }
        }
        // this rank lies in front of the ranks above it in the stem's a1 -> h8 order
        const int s = 4 * cnt;
        if (s) {
            nhi = s == 64 ? nlo : (nhi << s) | (nlo >> (64 - s));
            nlo = s == 64 ? 0ull : nlo << s;
            nlo |= rank_nibs;
        }
        occ |= (u64)rank_occ << (8 * r);
        pieces += cnt;
    }
    if (pieces > 32) return HEADS_OTHER;
    // tail: castling(4) ep(1[+6]) rule50(6) fullmove(8)
    if (avail < 32) {
        const u32 nw = nextw < 10 ? W(nextw) : 0u;
        ++nextw;
        win |= (u64)nw << avail;
        avail += 32;
    }
    u32 tail = (u32)win;
    int used = 4;
    const int cr = (int)(tail & 15u);
    tail >>= 4;
    int ep = SQ_NONE;
    if (tail & 1u) {
        ep = (int)((tail >> 1) & 63u);
        tail >>= 7;
        used += 7;
    } else {
        tail >>= 1;
        used += 1;
    }
    const u32 rule50 = tail & 63u;
    used += 14;
    const int cursor = 32 * nextw - avail + used;
    if (cursor > 256 || err) return HEADS_BAD;  // "Improperly encoded bin sfen" (:407-408, :441-442), type codes 5..7
    if (ep != SQ_NONE && !heads_ep_possible(occ, nlo, nhi, stm, ep)) ep = SQ_NONE;
    // the nibbles that depend on the tail (stem_nibble)
    if (ep != SQ_NONE) {
        const int sq = (ep & 7) + (stm == BLACK ? 24 : 32);  // the pawn on the ep file, rank 4 / rank 5
        if ((occ >> sq) & 1) {
            const int k = popc64(occ & before64(sq));
            if ((nib_get(nlo, nhi, k) >> 1) == PT_PAWN) nib_set(nlo, nhi, k, 12);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int right = c == 0 ? CR_WQ : c == 1 ? CR_WK : c == 2 ? CR_BQ : CR_BK;
        const int sq = c == 0 ? 0 : c == 1 ? 7 : c == 2 ? 56 : 63;
        const int rook = (PT_ROOK << 1) | (c >> 1);
        if ((cr & right) && ((occ >> sq) & 1)) {
            const int k = popc64(occ & before64(sq));
            if (nib_get(nlo, nhi, k) == rook) nib_set(nlo, nhi, k, 13 + (c >> 1));
        }
    }
    if (stm == BLACK) nib_set(nlo, nhi, popc64(occ & before64(bk)), 15);  // the black king says who is to move
    const u64 be = bswap64(occ);
    out[0] = (u32)be;
    out[1] = (u32)(be >> 32);
    out[2] = (u32)nlo; out[3] = (u32)(nlo >> 32); out[4] = (u32)nhi; out[5] = (u32)(nhi >> 32);
    // move, score, ply | result, rule50 (packEntry :1005-1019): the record's own words 8 and 9
    const u32 w8 = W(8), w9 = W(9);
    const Move mv = sfmove_to_move(w8 >> 16);
    const int score = (int)(short)(w8 & 0xFFFF), ply = (int)(w9 & 0xFFFF), result = (int)(signed char)((w9 >> 16) & 0xFF);
    u32 cm = 0;  // CompressedMove(Move) Chess.h:1071-1096
    if (mv.from != mv.to) {
        cm = ((u32)mv.type << 14) | ((u32)mv.from << 8) | ((u32)mv.to << 2);
        if (mv.type == MT_PROMOTION) cm |= (u32)((mv.promo >> 1) - PT_KNIGHT);
        cm &= 0xFFFF;
    }
    const u32 sc = zz_enc(score);
    const u32 pr = ((u32)ply | (zz_enc(result) << 14)) & 0xFFFF;
    out[6] = (cm >> 8) | ((cm & 0xFF) << 8) | ((sc >> 8) << 16) | ((sc & 0xFF) << 24);
    out[7] = (pr >> 8) | ((pr & 0xFF) << 8) | ((rule50 & 0xFF) << 24);
    return HEADS_OK;
}

}  // namespace nnp
