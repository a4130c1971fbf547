// heads.cuh -- a .bin record that starts a chain needs no position either: its 32-byte stem (packEntry,
// compress_file.cpp:997-1020) is the record's PackedSfen transcoded. The stream names the kings in its
// header and then has one token per non-king square, rank 8 first (pos_from_packed_sfen :364-446); the stem
// wants the occupied squares a1 -> h8 with one nibble each (Position::compress, Position.h:1374-1406).
// The tokens are read as one flat sequence (record_to_stem below), the kings are inserted, the ranks are put
// into the stem's order (rank 1 first), and the three nibbles that depend on the stream's tail -- rooks that
// still carry a castling right, the pawn an en-passant capture would take, the black king when black is to
// move -- are patched afterwards. Same bytes as sfen_decode + stem_pack
// wherever it answers HEADS_OK; everything irregular (both kings on one square, more than 32 pieces) is
// left to that route (HEADS_OTHER), malformed streams are reported as the decoder reports them
// (HEADS_BAD). Files of single positions are nothing but chain heads: k_heads_transcode, k_heads_direct (compress.cu).
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

// inserts a set bit at position k of x (the bits at and above k move up by one)
__device__ __forceinline__ u64 insert_one(u64 x, int k)
{
    const u64 below = (1ull << k) - 1;
    return (x & below) | (1ull << k) | ((x & ~below) << 1);
}
// inserts nibble v at nibble index k of the 32-nibble string hi:lo (the nibbles at and above k move up by one)
__device__ __forceinline__ void insert_nibble(u64& lo, u64& hi, int k, int v)
{
    if (k < 16) {
        const u64 below = (1ull << (4 * k)) - 1;  // k < 16: the shift is below 64
        hi = (hi << 4) | (lo >> 60);
        lo = (lo & below) | ((u64)v << (4 * k)) | ((lo & ~below) << 4);
    } else {
        const u64 below = (1ull << (4 * (k - 16))) - 1;
        hi = (hi & below) | ((u64)v << (4 * (k - 16))) | ((hi & ~below) << 4);
    }
}

// `W(j)` returns 32-bit word j of the 40-byte record (j < 10); `out` receives the stem's eight words.
//
// The 62 tokens of the non-king squares are read as one flat sequence, six per refill of a 32-bit window
// (a token is '0' or 1 + 4 bits: six of them fit), with the same handful of instructions for every token and no
// branch that depends on the position: occupancy bit j = "token j is a piece", the nibbles are gathered in
// stream order. Afterwards the kings are inserted (a bit into the occupancy, a nibble into the string, each at
// the place the header names) and the ranks are put in the stem's order: the stream has rank 8 first, the stem
// rank 1 -- the occupancy in stream order IS the stem's big-endian occupancy field, and the nibble string is
// re-cut rank by rank (a rank's nibble count is the popcount of its occupancy byte).
template <typename WordFn>
__device__ __forceinline__ int record_to_stem(WordFn W, u32 (&out)[8])
{
    const u32 w0 = W(0), w1 = W(1);
    const int stm = (int)(w0 & 1u), wk = (int)((w0 >> 1) & 63u), bk = (int)((w0 >> 7) & 63u);
    if (wk == bk) return HEADS_OTHER;
    u64 win = (((u64)w1 << 32) | w0) >> 13;  // the stream as a 64-bit window, low bit next
    int avail = 51, nextw = 2, cursor = 13;
    u32 o_lo = 0, o_hi = 0;              // bit j: token j is a piece
    u32 s0 = 0, s1 = 0, s2 = 0, s3 = 0;  // their nibbles, in stream order
    int bits = 0;                        // 4 * pieces so far
#pragma unroll
    for (int g = 0; g < 11; ++g) {
        if (avail < 32) {  // a group needs at most 30 bits
            const u32 nw = nextw < 10 ? W(nextw) : 0u;
            ++nextw;
            win |= (u64)nw << avail;
            avail += 32;
        }
        // (the place of the next nibble is kept as a power of two and nibbles are multiplied into place: the
        // kernel is bound by the integer ALU pipe, multiply-adds issue on the other one)
        u32 c = (u32)win, acc = 0, pw = 1;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int j = 6 * g + k;
            if (j < 62) {
                const u32 piece = c & 1u;  // '0', or 1 + type (3 bits, LSB first) + colour
                const u32 nib = (c & 14u) | ((c >> 4) & 1u);  // type * 2 + colour
                acc += nib * piece * pw;
                pw *= 15u * piece + 1u;
                if (j < 32) o_lo += piece * (1u << j); else o_hi += piece * (1u << (j - 32));
                c >>= 1u + 4u * piece;
            }
        }
        const int a = 31 - __clz((int)pw);                // 4 * pieces of the group
        const int used = (g < 10 ? 6 : 2) + a;            // one bit per token, four more per piece
        // the group's nibbles (at most 24 bits) go behind the ones gathered so far
        const int wi = bits >> 5, sh = bits & 31;
        const u32 lo = acc << sh, hi = __funnelshift_l(acc, 0u, sh);
        s0 |= wi == 0 ? lo : 0u;
        s1 |= wi == 1 ? lo : wi == 0 ? hi : 0u;
        s2 |= wi == 2 ? lo : wi == 1 ? hi : 0u;
        s3 |= wi == 3 ? lo : wi == 2 ? hi : 0u;
        bits += a;
        win >>= used;
        avail -= used;
        cursor += used;
    }
    if (bits > 4 * 30) return HEADS_OTHER;  // more than 32 pieces with the kings: the stem has no room (stem_pack drops them)
    // type codes 5..7 (nibbles 10..15) never terminate the reference's table search (:336-352)
    u32 err = 0;
    err |= s0 & 0x88888888u & ((s0 << 1) | (s0 << 2));
    err |= s1 & 0x88888888u & ((s1 << 1) | (s1 << 2));
    err |= s2 & 0x88888888u & ((s2 << 1) | (s2 << 2));
    err |= s3 & 0x88888888u & ((s3 << 1) | (s3 << 2));
    // tail: castling(4) ep(1[+6]) rule50(6) fullmove(8)
    if (avail < 32) {
        const u32 nw = nextw < 10 ? W(nextw) : 0u;
        ++nextw;
        win |= (u64)nw << avail;
        avail += 32;
    }
    u32 tail = (u32)win;
    const int cr = (int)(tail & 15u);
    tail >>= 4;
    cursor += 4;
    int ep = SQ_NONE;
    if (tail & 1u) {
        ep = (int)((tail >> 1) & 63u);
        tail >>= 7;
        cursor += 7;
    } else {
        tail >>= 1;
        cursor += 1;
    }
    const u32 rule50 = tail & 63u;
    cursor += 14;
    if (cursor > 256 || err) return HEADS_BAD;  // "Improperly encoded bin sfen" (:407-408, :441-442)
    // the kings: stream square = square ^ 56; the one that comes first in the stream goes in first
    const int ka = wk ^ 56, kb = bk ^ 56;
    const int k1 = ka < kb ? ka : kb, k2 = ka < kb ? kb : ka;
    u64 so = insert_one(insert_one(((u64)o_hi << 32) | o_lo, k1), k2);  // occupancy, stream order
    u64 slo = ((u64)s1 << 32) | s0, shi = ((u64)s3 << 32) | s2;
    insert_nibble(slo, shi, popc64(so & ((1ull << k1) - 1)), k1 == ka ? 10 : 11);
    insert_nibble(slo, shi, popc64(so & ((1ull << k2) - 1)), k2 == ka ? 10 : 11);
    // ranks: byte r of `so` is rank 8 - r; its nibbles leave the front of the stream-ordered string and are
    // pushed in front of the board-ordered one
    s0 = (u32)slo; s1 = (u32)(slo >> 32); s2 = (u32)shi; s3 = (u32)(shi >> 32);
    u32 b0 = 0, b1 = 0, b2 = 0, b3 = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int n = 4 * __popc((u32)(so >> (8 * r)) & 0xFFu);  // 0..32 bits
        const u32 seg = s0 & __funnelshift_lc(0xFFFFFFFFu, 0u, n);
        b3 = __funnelshift_lc(b2, b3, n);
        b2 = __funnelshift_lc(b1, b2, n);
        b1 = __funnelshift_lc(b0, b1, n);
        b0 = __funnelshift_lc(0u, b0, n) | seg;
        s0 = __funnelshift_rc(s0, s1, n);
        s1 = __funnelshift_rc(s1, s2, n);
        s2 = __funnelshift_rc(s2, s3, n);
        s3 = __funnelshift_rc(s3, 0u, n);
    }
    const u64 occ = bswap64(so);
    u64 nlo = ((u64)b1 << 32) | b0, nhi = ((u64)b3 << 32) | b2;
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
