// stream.cuh -- the PackedSfen Huffman stream (SfenPacker::pack, compress_file.cpp:266-312) kept
// up to date along a chain instead of being rebuilt for every position.
//
// Layout of the 256-bit little-endian stream: stm(1) wK(6) bK(6), then one token per non-king
// square in stream order (square ^ 56: rank 8 first, file a first): '0' for an empty square,
// 1 + type(3) + colour(1) for a piece; then castling(4) ep(1[+6]) rule50(6) fullmove(8).
// A move changes two to four tokens, so the board part of the next record's stream is the current
// one with those tokens replaced and everything above them shifted by the width difference:
// a few dozen word operations instead of a loop over all pieces. The tail is re-appended for every
// record. Moves outside the domain of the splice (no or several kings of a colour, king captures,
// degenerate castling / en-passant geometry: nothing the reference writer emits from legal games)
// make stream_apply_move return false and the caller rebuilds the stream with stream_from_pos.
#pragma once
#include "chess.cuh"

namespace nnp {

// bits of 32-bit word k that lie below stream position p
__device__ __forceinline__ u32 stream_low_mask(int p, int k)
{
    const int t = max(p - 32 * k, 0);
    return __funnelshift_lc(0xffffffffu, 0u, (u32)t);  // clamps the shift to 32
}

// Replaces the `wo` bits at stream position X by the `wn` low bits of `val` (wo, wn <= 5); all bits
// above move by wn - wo. Header + board tokens of a position with at most 32 non-king pieces end
// below bit 13 + 62 + 4 * 32 = 203, so only words 0..6 are edited: word 7 (bits 224..255) never
// holds a token, it only ever receives tail bits (stream_with_tail) or, in the compressor, the
// junk behind the compared range.
constexpr int STREAM_BOARD_WORDS = 7;
__device__ __forceinline__ void stream_edit(u32 (&W)[8], int X, int wo, int wn, u32 val)
{
    const u32 up = (u32)(wn - wo + 8);  // 3 .. 13: shift up, applied to the stream moved down one byte
    const int wi = X >> 5, sb = X & 31;
    const u32 field = (1u << wn) - 1u;  // the inserted bits: cleared in the shifted stream, then set
    const u32 clo = field << sb, chi = __funnelshift_l(field, 0u, sb);
    const u32 vlo = val << sb, vhi = __funnelshift_l(val, 0u, sb);
    u32 N[STREAM_BOARD_WORDS];
    u32 below = W[0] << 24;
#pragma unroll
    for (int k = 0; k < STREAM_BOARD_WORDS; ++k) {
        const u32 next = W[k + 1];
        const u32 b8 = __byte_perm(W[k], next, 0x4321);   // stream bits 32k+8 .. 32k+39
        const u32 sh = __funnelshift_l(below, b8, up);     // bit i of the result = old bit i - (wn - wo)
        below = b8;
        const u32 m = stream_low_mask(X, k);
        u32 v = (W[k] & m) | (sh & ~m);
        if (k == wi) v = (v & ~clo) | vlo;
        if (k == wi + 1) v = (v & ~chi) | vhi;
        N[k] = v;
    }
#pragma unroll
    for (int k = 0; k < STREAM_BOARD_WORDS; ++k) W[k] = N[k];
}

__device__ __forceinline__ u32 stream_token(int piece) { return 1u | ((u32)(piece >> 1) << 1) | ((u32)(piece & 1) << 4); }

// end of the board part for a position with one king per side
__device__ __forceinline__ int stream_board_end(const Pos& p) { return 13 + 62 + 4 * (popc64(pos_all(p)) - 2); }

// Header + board tokens of `p` built from scratch into W; returns the bit length. `col` is a
// per-thread scratch column of 8 words with the given stride (shared memory: the word index of a
// token is data dependent).
__device__ __forceinline__ int stream_from_pos(const Pos& p, u32* col, int stride, u32 (&W)[8])
{
    const u64 all = pos_all(p);
    const u64 kings = pos_type_bb(p, PT_KING);
    const u64 wkb = kings & p.occ[0], bkb = kings & p.occ[1];
    const int wk = wkb ? lsb64(wkb) : 0, bk = bkb ? lsb64(bkb) : 0;  // kingSquare Position.h:742-745
#pragma unroll
    for (int k = 1; k < 8; ++k) col[k * stride] = 0;
    col[0] = (u32)p.stm | ((u32)wk << 1) | ((u32)bk << 7);
    const u64 s_k = bswap64(kings);
    u64 s_np = bswap64(all) & ~s_k;
    int np = 0;
    while (s_np) {
        const int s = lsb64(s_np);
        s_np &= s_np - 1;
        const int pos = 13 + s - popc64(s_k & before64(s)) + 4 * np;
        const u32 tok = stream_token(pos_piece_at(p, s ^ 56));
        const int wi = pos >> 5, sb = pos & 31;
        if (wi < 8) col[wi * stride] |= tok << sb;
        if (sb > 27 && wi < 7) col[(wi + 1) * stride] |= tok >> (32 - sb);
        ++np;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) W[k] = col[k * stride];
    return 13 + (64 - popc64(kings)) + 4 * np;
}

// Applies move `m`, made in position `p` (the position BEFORE the move), to the stream of `p`.
// The token changes mirror Board::doMove (Position.h:300-439) as restated in board_do_move.
// General form: up to four token edits applied one after the other (castling, en passant, and the
// reference for the fused form below in the CPU suite).
__device__ __forceinline__ bool stream_apply_move_generic(u32 (&W)[8], const Pos& p, const Move& m, int moved = -1)
{
    const u64 all = pos_all(p);
    const u64 kings = pos_type_bb(p, PT_KING);
    const u64 wkb = kings & p.occ[0], bkb = kings & p.occ[1];
    if (popc64(wkb) != 1 || popc64(bkb) != 1 || popc64(all) > 34) return false;
    const int from = m.from, to = m.to;
    if (from > 63 || to > 63 || from == to) return false;
    const int pc = moved >= 0 ? moved : pos_piece_at(p, from);  // the caller may have looked it up already
    if (pc == NO_PIECE) return false;
    const bool king_moves = (pc >> 1) == PT_KING;
    auto occupied = [&](int sq) { return (int)((all >> sq) & 1); };
    auto has_king = [&](int sq) { return (int)((kings >> sq) & 1); };

    // up to four token edits, packed as stream square | old width << 6 | new width << 9 | new bits << 12
    auto edit = [](int sq, int wo, int wn, u32 v) { return (u32)(sq ^ 56) | ((u32)wo << 6) | ((u32)wn << 9) | (v << 12); };
    u32 e0, e1, e2 = 0, e3 = 0;
    int n = 2;
    int new_king_sq = -1;
    if (m.type == MT_CASTLE) {
        const int rook = pos_piece_at(p, to);
        if (!king_moves || rook == NO_PIECE || (rook >> 1) == PT_KING) return false;
        const int base = (pc & 1) ? 56 : 0;
        const bool is_short = (to & 7) == 7;
        const int rt = base + (is_short ? 5 : 3), kt = base + (is_short ? 6 : 2);
        if (rt == from || rt == to || kt == from || kt == to || has_king(rt) || has_king(kt)) return false;
        e0 = edit(from, 0, 1, 0);
        e1 = edit(to, 5, 1, 0);
        e2 = edit(rt, occupied(rt) ? 5 : 1, 5, stream_token(rook));
        e3 = edit(kt, occupied(kt) ? 5 : 1, 0, 0);
        n = 4;
        new_king_sq = kt;
    } else {
        if (has_king(to)) return false;
        const int wto = occupied(to) ? 5 : 1;
        if (king_moves) {
            if (m.type != MT_NORMAL) return false;
            e0 = edit(from, 0, 1, 0);
            e1 = edit(to, wto, 0, 0);
            new_king_sq = to;
        } else {
            const int placed = m.type == MT_PROMOTION ? m.promo : pc;
            if (placed == NO_PIECE) return false;
            e0 = edit(from, 5, 1, 0);
            e1 = edit(to, wto, 5, stream_token(placed));
            if (m.type == MT_ENPASSANT) {
                const int cap = (to & 7) | (from & 56);
                if ((pc >> 1) != PT_PAWN || cap == from || cap == to || has_king(cap)) return false;
                if (occupied(cap)) {
                    e2 = edit(cap, 5, 1, 0);
                    n = 3;
                }
            }
        }
    }
    const u64 s_all = bswap64(all);
    const int ks1 = lsb64(wkb) ^ 56, ks2 = lsb64(bkb) ^ 56;
    // one loop body for all edits (the kernels are instruction-cache sensitive); edits applied
    // earlier at lower stream squares have moved the later ones
    u32 done_lo = 0;  // packed (square, delta + 8) of the edits already applied, 10 bits each
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
        const u32 e = j == 0 ? e0 : j == 1 ? e1 : j == 2 ? e2 : e3;
        const int sq = (int)(e & 63u), wo = (int)((e >> 6) & 7u), wn = (int)((e >> 9) & 7u);
        // position of stream square sq in the stream of `p`: 13 + non-king squares before it
        // + 4 bits for every non-king piece before it
        const int ab = popc64(s_all & before64(sq));
        const int kb = (sq > ks1) + (sq > ks2);
        int x = 13 + sq + 4 * ab - 5 * kb;
        for (u32 d = done_lo; d; d >>= 10)
            if ((int)(d & 63u) < sq) x += (int)((d >> 6) & 15u) - 8;
        stream_edit(W, x, wo, wn, e >> 12);
        done_lo = (done_lo << 10) | (u32)sq | ((u32)(wn - wo + 8) << 6);
    }
    u32 w0 = W[0] ^ 1u;  // side to move
    if (new_king_sq >= 0) {
        const int f = (pc & 1) ? 7 : 1;
        w0 = (w0 & ~(63u << f)) | ((u32)new_king_sq << f);
    }
    W[0] = w0;
    return true;
}

// the eight-word mask "bits below stream position pos"
__device__ __forceinline__ void stream_low_masks(int pos, u32 (&M)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) M[k] = stream_low_mask(pos, k);
}

// Normal moves and promotions (everything but castling and en passant: 99.7 % of the plies of a game)
// change exactly two tokens, and both edits are applied in ONE pass over the seven board words:
//
//   new = old[0, X1) ++ v1 ++ old[X1 + wo1, X2) ++ v2 ++ old[X2 + wo2, ...)
//
// with X1 < X2 the stream positions of the two squares in the OLD stream, wo the old and wn the new
// token widths (a non-king piece leaves '0' behind: 5 -> 1; it arrives as a 5-bit token on an empty
// square or on a captured piece: 1 | 5 -> 5; a king occupies no stream bits, so it leaves a '0':
// 0 -> 1, and its destination token disappears: 1 | 5 -> 0). The bits between the edits move by
// d1 = wn1 - wo1, the bits above the second edit by d1 + d2, which is 0 for a quiet move and -4 for a
// capture. Each output word is therefore a masked merge of the old word, the old stream shifted by d1
// and the old stream shifted by d1 + d2; the at most one 5-bit token is OR-ed in last, and the '0'
// bits are simply left uncovered by the three masks.
__device__ __forceinline__ bool stream_apply_move(u32 (&W)[8], const Pos& p, const Move& m, int moved = -1,
                                                  const StepTables* T = nullptr)
{
#ifdef NNP_SPLICE_GENERIC
    return stream_apply_move_generic(W, p, m, moved);
#endif
    if (m.type == MT_CASTLE || m.type == MT_ENPASSANT) return stream_apply_move_generic(W, p, m, moved);
    const u64 all = pos_all(p);
    const u64 kings = pos_type_bb(p, PT_KING);
    const u64 wkb = kings & p.occ[0], bkb = kings & p.occ[1];
    if (popc64(wkb) != 1 || popc64(bkb) != 1 || popc64(all) > 34) return false;
    const int from = m.from, to = m.to;
    if (from > 63 || to > 63 || from == to) return false;
    const int pc = moved >= 0 ? moved : pos_piece_at(p, from);
    if (pc == NO_PIECE) return false;
    if ((kings >> to) & 1) return false;
    const bool king_moves = (pc >> 1) == PT_KING;
    const int placed = m.type == MT_PROMOTION ? m.promo : pc;
    if (king_moves ? m.type != MT_NORMAL : placed == NO_PIECE) return false;
    const int wto = (int)((all >> to) & 1) * 4 + 1;  // old width of the destination token: 1 or 5

    // stream squares and their positions in the old stream (see stream_apply_move_generic)
    const u64 s_all = bswap64(all);
    const int ks1 = lsb64(wkb) ^ 56, ks2 = lsb64(bkb) ^ 56;
    const int sf = from ^ 56, st = to ^ 56;
#ifdef NNP_NO_SMALL_LUT
    const u64 bf = before64(sf), bt = before64(st);
#else
    const u64 bf = T ? T->before[sf] : before64(sf), bt = T ? T->before[st] : before64(st);
#endif
    const int xf = 13 + sf + 4 * popc64(s_all & bf) - 5 * ((sf > ks1) + (sf > ks2));
    const int xt = 13 + st + 4 * popc64(s_all & bt) - 5 * ((st > ks1) + (st > ks2));
    const int wof = king_moves ? 0 : 5, wnt = king_moves ? 0 : 5;  // the from-square always becomes '0' (width 1)
    const bool from_first = sf < st;
    const int X1 = from_first ? xf : xt, X2 = from_first ? xt : xf;
    const int wo1 = from_first ? wof : wto, wn1 = from_first ? 1 : wnt;
    const int wn2 = from_first ? wnt : 1;
    const int d1 = wn1 - wo1;                       // -5 .. +4
    const int dd = 1 - wof + wnt - wto;             // d1 + d2: 0 or -4
    const int X2n = X2 + d1;                        // second edit in the new stream
    u32 M1[8], M1e[8], M2[8], M2e[8];
    stream_low_masks(X1, M1);
    stream_low_masks(X1 + wn1, M1e);
    stream_low_masks(X2n, M2);
    stream_low_masks(X2n + wn2, M2e);
    // the arriving piece's token (none for a king) at its position in the new stream
    const int Xtok = from_first ? X2n : X1;
    const u32 tok = king_moves ? 0u : stream_token(placed);
    const int wi = Xtok >> 5, sb = Xtok & 31;
    const u32 vlo = tok << sb, vhi = __funnelshift_l(tok, 0u, sb);
    const u32 upA = (u32)(d1 + 8), upB = (u32)(dd + 8);
    u32 N[STREAM_BOARD_WORDS];
    u32 below = W[0] << 24;
#pragma unroll
    for (int k = 0; k < STREAM_BOARD_WORDS; ++k) {
        const u32 b8 = __byte_perm(W[k], W[k + 1], 0x4321);  // stream bits 32k+8 .. 32k+39
        const u32 A = __funnelshift_l(below, b8, upA);        // old stream moved by d1
        const u32 B = __funnelshift_l(below, b8, upB);        // old stream moved by d1 + d2
        below = b8;
        u32 v = (W[k] & M1[k]) | (A & M2[k] & ~M1e[k]) | (B & ~M2e[k]);
        if (k == wi) v |= vlo;
        if (k == wi + 1) v |= vhi;
        N[k] = v;
    }
#pragma unroll
    for (int k = 0; k < STREAM_BOARD_WORDS; ++k) W[k] = N[k];
    u32 w0 = W[0] ^ 1u;  // side to move
    if (king_moves) {
        const int f = (pc & 1) ? 7 : 1;
        w0 = (w0 & ~(63u << f)) | ((u32)to << f);
    }
    W[0] = w0;
    return true;
}

// W (header + board, `end` bits) followed by castling / ep / rule50 / fullmove of `p` (:290-311)
__device__ __forceinline__ void stream_with_tail(const u32 (&W)[8], int end, const Pos& p, u32 (&out)[8])
{
    u32 T = (u32)p.cr & 15u;
    int n = 5;
    if (p.ep != SQ_NONE) {
        T |= 16u | ((u32)(p.ep & 63) << 5);
        n = 11;
    }
    T |= ((u32)p.rule50 & 63u) << n;
    T |= (u32)(((p.ply + 1) >> 1) & 0xFF) << (n + 6);  // halfMove() Position.h:933-936, 8 bits
    const int wi = end >> 5, sb = end & 31;
    const u32 lo = T << sb, hi = __funnelshift_l(T, 0u, sb);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        u32 v = W[k];
        if (k == wi) v |= lo;
        if (k == wi + 1) v |= hi;
        out[k] = v;
    }
}

}  // namespace nnp
