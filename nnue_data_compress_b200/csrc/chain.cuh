// chain.cuh -- walking one chain of a .binpack: the device-side form of
// CompressedTrainingDataEntryReader::next (compress_file.cpp:1154-1190) and
// PackedMoveScoreListReader::nextEntry (:669-678), shared by the decompression kernels.
#pragma once
#include "chess.cuh"
#include "stream.cuh"

namespace nnp {

struct ChainCursor {
    Pos pos;
    Move mv;
    int moved;  // the piece on mv.from in pos when the ply decoder looked it up, else -1
    int score, ply, result;
    int last_score;
    u32 num_plies;
};

__device__ __forceinline__ void chain_open(const unsigned char* s, ChainCursor& c)
{
    stem_unpack([&](int i) { return (u32)s[i]; }, c.pos, c.mv, c.score, c.ply, c.result);
    c.num_plies = ((u32)s[32] << 8) | (u32)s[33];
    c.moved = -1;
    c.last_score = (int)(short)(-c.score);  // PackedMoveScoreListReader ctor (:618)
}

// PackedMoveScoreListReader::nextEntry (:669-678)
__device__ __forceinline__ bool chain_step(ChainCursor& c, BitReader& r, bool strict, int moved = -1,
                                           const StepTables* T = nullptr)
{
    if (strict && (c.mv.from > 63 || c.mv.to > 63)) return false;  // null move followed by plies
    pos_do_move(c.pos, c.mv, moved, T);
    Move m;
    int sc;
    int looked_up = -1;
    if (!decode_ply(r, c.pos, c.last_score, m, sc, strict, T, &looked_up)) return false;
    c.mv = m;
    c.moved = looked_up;
    c.score = sc;
    c.ply = (c.ply + 1) & 0xFFFF;
    c.result = (int)(short)(-c.result);
    return true;
}

// walks one chain from its stem, calling emit(cursor, k) for k = 0..numPlies and before_move(cursor)
// just before the cursor's move is made; returns false when the movetext runs off the chunk
template <typename MoveFn, typename EmitFn>
__device__ __forceinline__ bool walk_chain(const unsigned char* s, u32 bytes_after_stem, MoveFn before_move, EmitFn emit,
                                           u32& consumed, const StepTables* T = nullptr)
{
    ChainCursor cc;
    chain_open(s, cc);
    emit(cc, 0u);
    BitReader r;
    r.init(s + 34, (u64)bytes_after_stem);
    for (u32 k = 0; k < cc.num_plies; ++k) {
        if (cc.mv.from > 63 || cc.mv.to > 63) return false;
        const int moved = before_move(cc);  // the piece on cc.mv.from if the hook looked it up, else -1
        if (!chain_step(cc, r, false, moved, T)) return false;
        emit(cc, k + 1);
    }
    consumed = 34 + ((r.pos + 7) >> 3);
    return true;
}
template <typename EmitFn>
__device__ __forceinline__ bool walk_chain(const unsigned char* s, u32 bytes_after_stem, EmitFn emit, u32& consumed)
{
    return walk_chain(s, bytes_after_stem, [](const ChainCursor&) { return -1; }, emit, consumed);
}

// Staging of a chain's 40-byte records (k_emit_chains_verify). A record is 40 bytes at a 40-byte stride:
// written as five 8-byte stores when it is made, it leaves a partial 32-byte sector at its end that
// the next record of the chain completes a whole step later, by which time L2 has usually written
// it back: the DRAM interface then sees 1.3x the bytes and reads back what it merges. Four records
// starting at a record index that is a multiple of four are 160 bytes = five whole sectors, 16-byte
// aligned at both ends: a lane gathers them in its shared-memory slot and writes them with ten 16-byte
// stores (or, with NNP_BULK_STORE, hands them to the copy engine as one cp.async.bulk, bulk.cuh --
// measured slower here: the copy is a uniform-datapath instruction, so a warp issues its 32 lanes'
// copies one after the other). Records in front of the first multiple of four and behind
// the last whole group of a chain go out with plain stores.
constexpr int OUT_GROUP = 4;
constexpr int OUT_SLOT_BYTES = OUT_GROUP * 40 + 16;  // + 16: the slots of a quarter-warp start in different banks
struct OutStage {
    unsigned char* slot;  // 16-byte aligned, OUT_GROUP * 40 bytes; nullptr = plain stores only
    u64 first;            // record index of slot record 0
    int staged;
};
__device__ __forceinline__ void store_record(unsigned char* out, u64 rec, const u32 (&w)[8], u32 w8, u32 w9)
{
    uint2* d = reinterpret_cast<uint2*>(out + rec * 40);
    d[0] = make_uint2(w[0], w[1]);
    d[1] = make_uint2(w[2], w[3]);
    d[2] = make_uint2(w[4], w[5]);
    d[3] = make_uint2(w[6], w[7]);
    d[4] = make_uint2(w8, w9);
}
// what is left in the slot at the end of a chain (fewer than OUT_GROUP records)
__device__ __forceinline__ void out_stage_finish(OutStage& st, unsigned char* out)
{
    for (int m = 0; m < st.staged; ++m) {
        const uint2* sp = reinterpret_cast<const uint2*>(st.slot + m * 40);
        uint2* d = reinterpret_cast<uint2*>(out + (st.first + m) * 40);
#pragma unroll
        for (int i = 0; i < 5; ++i) d[i] = sp[i];
    }
    st.staged = 0;
#if defined(NNP_BULK_STORE) && defined(__CUDACC__)
    if (st.slot) bulk_store_wait_read();  // the slot must outlive the copy engine's reads of it
#endif
}
__device__ __forceinline__ void out_stage_put(OutStage& st, unsigned char* out, u64 rec, const u32 (&w)[8], u32 w8, u32 w9)
{
    if (!st.slot || (st.staged == 0 && (rec & (OUT_GROUP - 1)))) {  // no staging, or in front of a group boundary
        store_record(out, rec, w, w8, w9);
        return;
    }
    if (st.staged == 0) {
#if defined(NNP_BULK_STORE) && defined(__CUDACC__)
        bulk_store_wait_read();  // the copy engine is done with the slot's previous content
#endif
        st.first = rec;
    }
    uint2* d = reinterpret_cast<uint2*>(st.slot + st.staged * 40);
    d[0] = make_uint2(w[0], w[1]);
    d[1] = make_uint2(w[2], w[3]);
    d[2] = make_uint2(w[4], w[5]);
    d[3] = make_uint2(w[6], w[7]);
    d[4] = make_uint2(w8, w9);
    if (++st.staged == OUT_GROUP) {
#if defined(NNP_BULK_STORE) && defined(__CUDACC__)
        bulk_store(out + st.first * 40, st.slot, OUT_GROUP * 40);
#else
        const uint4* sp = reinterpret_cast<const uint4*>(st.slot);
        uint4* g = reinterpret_cast<uint4*>(out + st.first * 40);  // 160 * (first / 4): 32-byte aligned
#pragma unroll
        for (int i = 0; i < OUT_GROUP * 40 / 16; ++i) g[i] = sp[i];
#endif
        st.staged = 0;
    }
}

// A chain of a single position (numPlies == 0: all of a file of chain length 1) needs no position at all:
// its 40-byte record is the stem transcoded. The stem lists the occupied squares a1 -> h8 with one nibble
// each (CompressedPosition, Position.h:1374-1406), the PackedSfen stream wants a token per non-king
// square rank 8 first (SfenPacker::pack :266-312): the pieces are visited in stream order, each one's
// nibble looked up by its rank among the occupied squares, and its token dropped at 13 + stream square
// - kings so far + 4 * pieces so far. Same bytes as stem_unpack + stream_from_pos + stream_with_tail
// (one white and one black king, at most 32 pieces; anything else returns false and takes that route),
// at a fifth of the instructions. `col` is the thread's 8-word scratch column in shared memory.
__device__ __forceinline__ bool stem_to_record(const unsigned char* s, u32* col, int stride, u32 (&w)[8], u32& w8, u32& w9)
{
    // the 32 stem bytes as eight little-endian words, from aligned loads
    u32 v[8];
    {
        const u32* wp = reinterpret_cast<const u32*>(reinterpret_cast<uintptr_t>(s) & ~(uintptr_t)3);
        const int sh = (int)(reinterpret_cast<uintptr_t>(s) & 3) * 8;
        u32 prev = wp[0];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const u32 next = (sh || j < 7) ? wp[j + 1] : 0u;  // (aligned stems end with wp[7])
            v[j] = __funnelshift_r(prev, next, sh);
            prev = next;
        }
    }
    const u64 occ = ((u64)__byte_perm(v[0], 0, 0x0123) << 32) | __byte_perm(v[1], 0, 0x0123);  // big-endian (:1245-1257)
    if (popc64(occ) > 32) return false;
    const int np = popc64(occ);
    // the nibbles, np of them; what lies behind them is not looked at (stem_unpack visits occupied squares only)
    u32 nb4[4];  // the nibble string, eight to a word
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int keep = min(max(4 * np - 32 * j, 0), 32);
        nb4[j] = v[2 + j] & __funnelshift_lc(0xFFFFFFFFu, 0u, keep);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) col[k * stride] = 0;
    // The nibbles above 9 carry state (Position.h:1374-1456): 10 / 11 the kings, 15 the black king with black to
    // move, 12 the pawn that just made a double push, 13 / 14 a rook that may still castle. They are found in
    // all 32 nibbles at once (bit-sliced compares) and rewritten to the plain piece they stand for, so that the
    // loop over the squares below has no cases.
    u32 k10[4], kbk[4], e12[4], e13[4], e14[4];
    u32 any15 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const u32 M = 0x11111111u, x = nb4[j];
        const u32 b0 = x & M, b1 = (x >> 1) & M, b2 = (x >> 2) & M, b3 = (x >> 3) & M;
        k10[j] = b3 & ~b2 & b1 & ~b0;
        kbk[j] = b3 & b1 & b0;  // 1011 and 1111
        e12[j] = b3 & b2 & ~b1 & ~b0;
        e13[j] = b3 & b2 & ~b1 & b0;
        e14[j] = b3 & b2 & b1 & ~b0;
        const u32 e15 = kbk[j] & b2;
        any15 |= e15;
        nb4[j] = x ^ (e13[j] * 0xBu) ^ (e14[j] * 0x9u) ^ (e15 * 0x4u) ^ (e12[j] * 0xCu);  // -> 6, 7, 11, 0
    }
    const int nwk = __popc(k10[0]) + __popc(k10[1]) + __popc(k10[2]) + __popc(k10[3]);
    const int nbk = __popc(kbk[0]) + __popc(kbk[1]) + __popc(kbk[2]) + __popc(kbk[3]);
    const int n12 = __popc(e12[0]) + __popc(e12[1]) + __popc(e12[2]) + __popc(e12[3]);
    // (several ep nibbles -- corrupted input only -- are visited in a different order by stem_unpack, where the
    // last one counts: the other route)
    if (nwk != 1 || nbk != 1 || n12 > 1) return false;
    // nibble index of the one marked nibble: words below the one that holds it count 8 (a zero word minus one is
    // all ones), the word itself counts the nibbles below the mark, words above it are switched off
    auto only = [](const u32 (&m)[4]) {
        const u32 L = 0x11111111u;
        const int c0 = __popc((m[0] - 1u) & L), c1 = __popc((m[1] - 1u) & L), c2 = __popc((m[2] - 1u) & L),
                  c3 = __popc((m[3] - 1u) & L);
        const bool s0 = m[0] != 0u, s1 = s0 || m[1] != 0u, s2 = s1 || m[2] != 0u;
        return c0 + (s0 ? 0 : c1) + (s1 ? 0 : c2) + (s2 ? 0 : c3);
    };
    auto marked = [](const u32 (&m)[4], int k) {  // is nibble k marked?
        const u64 lo = ((u64)m[1] << 32) | m[0], hi = ((u64)m[3] << 32) | m[2];
        return (u32)(((k & 16) ? hi : lo) >> (4 * (k & 15))) & 1u;
    };
    const int stm = any15 ? BLACK : WHITE;
    const int wk = nth_set_bit(occ, (u32)only(k10)), bk = nth_set_bit(occ, (u32)only(kbk));
    // castling rights: a 13 on a1 is the queen side, anywhere else the king side; 14 likewise with a8
    int cr = 0;
    {
        const u32 a1 = (u32)(occ & 1ull) & e13[0];  // nibble 0 is a1's when a1 is occupied
        if (a1) cr |= CR_WQ;
        if ((e13[0] ^ a1) | e13[1] | e13[2] | e13[3]) cr |= CR_WK;
        const int q = popc64(occ & before64(56));
        const bool a8 = ((occ >> 56) & 1ull) && marked(e14, q);
        if (a8) cr |= CR_BQ;
        const int n14 = __popc(e14[0]) + __popc(e14[1]) + __popc(e14[2]) + __popc(e14[3]);
        if (n14 > (a8 ? 1 : 0)) cr |= CR_BK;
    }
    int ep = SQ_NONE;
    if (n12) {  // the pawn that just made a double push (Position.h:1440-1456): white on rank 4, else black
        const int k = only(e12), sq = nth_set_bit(occ, (u32)k);
        if ((sq >> 3) == 3) {
            ep = (sq - 8) & 0xFF;
        } else {
            ep = (sq + 8) & 0xFF;
            const u32 bit = 1u << (4 * (k & 7));  // a black pawn is piece 1
            if ((k >> 3) == 0) nb4[0] |= bit; else if ((k >> 3) == 1) nb4[1] |= bit; else if ((k >> 3) == 2) nb4[2] |= bit; else nb4[3] |= bit;
        }
    }
    // Square by square in stream order (rank 8 first): one '0' for an empty square, 1 + type + colour for a piece,
    // nothing for a king. The rank's nibbles are consecutive in the stem; every half rank gathers its tokens in
    // a 32-bit register (at most 4 x 5 bits) and the rank is dropped into the stream column in one piece.
    int cursor = 13;  // stream position of the next token
    // the string is consumed from its top: moved up so that the last piece's nibble is the topmost one, the
    // rank's nibbles are the top 4 * (pieces of the rank) bits, and the string moves up by as much afterwards
    u32 t0 = nb4[0], t1 = nb4[1], t2 = nb4[2], t3 = nb4[3];
    {
        const int up = 4 * (32 - np);  // 0 .. 120 (a stem has at least the two kings)
        if (up & 64) { t3 = t1; t2 = t0; t1 = 0u; t0 = 0u; }
        if (up & 32) { t3 = t2; t2 = t1; t1 = t0; t0 = 0u; }
        t3 = __funnelshift_l(t2, t3, up);
        t2 = __funnelshift_l(t1, t2, up);
        t1 = __funnelshift_l(t0, t1, up);
        t0 <<= (up & 31);
    }
#pragma unroll
    for (int r = 7; r >= 0; --r) {
        const u32 rb = (u32)(occ >> (8 * r)) & 0xFFu;
        const int take = 4 * __popc(rb);                 // 0 .. 32 bits
        u32 nw = __funnelshift_rc(t3, 0u, 32 - take);    // the rank's nibbles, file a lowest
        t3 = __funnelshift_lc(t2, t3, take);
        t2 = __funnelshift_lc(t1, t2, take);
        t1 = __funnelshift_lc(t0, t1, take);
        t0 = __funnelshift_lc(0u, t0, take);
        u32 hb[2];
        int hn[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // (the place of the next token is kept as a power of two and tokens are multiplied into place: the
            // kernels that call this are bound by the integer ALU pipe, multiply-adds issue on the other one)
            u32 b = 0, pw = 1;
#pragma unroll
            for (int f = 4 * h; f < 4 * h + 4; ++f) {
                const u32 p = (rb >> f) & 1u;
                const u32 nib = nw & 15u;
                const u32 king = nib >= 10u ? p : 0u;
                const u32 emit = p ^ king;
                const u32 tok = 1u | (nib & 14u) | ((nib & 1u) << 4);  // stream_token
                b += tok * emit * pw;
                pw *= 30u * emit + 2u - king;  // x 32 behind a token, x 2 behind a '0', x 1 for a king
                nw >>= 4 * p;
            }
            hb[h] = b;
            hn[h] = 31 - __clz((int)pw);
        }
        const u64 bits = (u64)hb[0] | ((u64)hb[1] << hn[0]);
        {
            const int ci = cursor >> 5, sb = cursor & 31;
            const u32 b0 = (u32)bits, b1 = (u32)(bits >> 32);
            col[ci * stride] |= b0 << sb;
            const u32 mid = __funnelshift_l(b0, b1, sb);  // bits 32 - sb .. 63 - sb
            if (ci < 7) col[(ci + 1) * stride] |= mid;
            if (ci < 6 && sb) col[(ci + 2) * stride] |= b1 >> (32 - sb);
        }
        cursor += hn[0] + hn[1];
    }
    // header, then castling / ep / rule50 / full move behind the board (stream_with_tail)
    u32 T = (u32)cr & 15u;
    int n = 5;
    if (ep != SQ_NONE) {
        T |= 16u | ((u32)(ep & 63) << 5);
        n = 11;
    }
    const u32 pr = __byte_perm(v[7], 0, 0x4401) & 0xFFFFu;  // bytes 28, 29 big-endian
    const int ply = (int)(pr & 0x3FFF);
    const u32 rule50 = v[7] >> 24;                       // byte 31
    T |= (rule50 & 63u) << n;
    T |= (u32)(((ply + 1) >> 1) & 0xFF) << (n + 6);
    const int end = cursor;  // 13 + 62 + 4 * pieces
    {
        const int wi = end >> 5, sb = end & 31;
        col[wi * stride] |= T << sb;
        if (wi < 7) col[(wi + 1) * stride] |= __funnelshift_l(T, 0u, sb);
    }
    col[0] |= (u32)stm | ((u32)wk << 1) | ((u32)bk << 7);
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = col[k * stride];
    // score, move, ply, result (unpackEntry :1022-1043, trainingDataEntryToPackedSfenValue :570-585)
    const u32 cm = __byte_perm(v[6], 0, 0x4401) & 0xFFFFu;   // bytes 24, 25 big-endian
    const u32 sc = __byte_perm(v[6], 0, 0x4423) & 0xFFFFu;   // bytes 26, 27 big-endian
    Move mv;
    if (cm == 0) {
        mv.from = mv.to = SQ_NONE;  // Move::null()
        mv.type = MT_NORMAL;
        mv.promo = NO_PIECE;
    } else {
        mv.type = (int)(cm >> 14);
        mv.from = (int)((cm >> 8) & 63);
        mv.to = (int)((cm >> 2) & 63);
        mv.promo = NO_PIECE;
        if (mv.type == MT_PROMOTION) mv.promo = ((PT_KNIGHT + (int)(cm & 3)) << 1) | ((mv.to >> 3) == 0 ? BLACK : WHITE);
    }
    const int score = zz_dec(sc), result = zz_dec(pr >> 14);
    w8 = ((u32)score & 0xFFFFu) | (move_to_sfmove(mv) << 16);
    w9 = ((u32)ply & 0xFFFFu) | (((u32)result & 0xFFu) << 16) | 0xFF000000u;
    return true;
}

// walk_chain writing 40-byte .bin records rec0, rec0 + 1, ... (below rec_limit). The Huffman stream
// of the position is carried along the chain (stream.cuh): built once for the chain head, then
// spliced per move. `col` is the thread's 8-word scratch column in shared memory; `slot` (optional)
// its record staging slot for bulk stores.
__device__ __forceinline__ bool emit_chain_bin(const unsigned char* s, u32 bytes_after_stem, unsigned char* out, u64 rec0,
                                               u64 rec_limit, u32* col, int stride, u32& consumed,
                                               const StepTables* T = nullptr, unsigned char* slot = nullptr)
{
    if ((((u32)s[32] << 8) | (u32)s[33]) == 0u) {  // a single position: transcode the stem
        u32 w[8], w8, w9;
        if (stem_to_record(s, col, stride, w, w8, w9)) {
            if (rec0 < rec_limit) store_record(out, rec0, w, w8, w9);
            consumed = 34;
            return true;
        }
    }
    u32 W[8];
    bool spliced = false;
    OutStage st{slot, 0, 0};
    const bool ok = walk_chain(
        s, bytes_after_stem,
        [&](const ChainCursor& cc) {
            const int moved = cc.moved >= 0 ? cc.moved : pos_piece_at(cc.pos, cc.mv.from);
            spliced = stream_apply_move(W, cc.pos, cc.mv, moved, T);
            return moved;
        },
        [&](const ChainCursor& cc, u32 k) {
            const int end = (k == 0 || !spliced) ? stream_from_pos(cc.pos, col, stride, W) : stream_board_end(cc.pos);
            if (rec0 + k >= rec_limit) return;
            u32 w[8];
            stream_with_tail(W, end, cc.pos, w);
            // trainingDataEntryToPackedSfenValue (:570-585): score, move, gamePly, result, padding 0xFF
            const u32 w8 = ((u32)cc.score & 0xFFFFu) | (move_to_sfmove(cc.mv) << 16);
            const u32 w9 = ((u32)cc.ply & 0xFFFFu) | (((u32)cc.result & 0xFFu) << 16) | 0xFF000000u;
            out_stage_put(st, out, rec0 + k, w, w8, w9);
        },
        consumed, T);
    if (slot) out_stage_finish(st, out);
    return ok;
}

}  // namespace nnp
