// walk.cuh -- the chain-walking form of the .bin compressor's per-record step.
//
// CompressedTrainingDataEntryWriter::addTrainingDataEntry (compress_file.cpp:1061-1092) asks, for
// every record, whether it continues its predecessor: isContinuation (:587-593) compares the
// predecessor's position after its move with the record's decoded position. Decoding every
// PackedSfen from scratch (pos_from_packed_sfen :364-446, a loop over up to 30 Huffman tokens) is
// what the reference spends 90 % of its compression time on. Along a chain it is not needed: the
// expected stream of the next record is the current one with two to four tokens replaced
// (stream.cuh), so "decode and compare positions" becomes "splice and compare bits":
//
//   * equal header + board bits  <=>  equal side to move, king squares and board, because the
//     token stream of a board is unique (kings are named in the header, every other square is
//     '0' or 1+type+colour, and type codes above Queen are rejected by the decoder);
//   * castling rights and the (post-move nullified, Position.h:868-872) ep square are read from the
//     record's tail right behind the board bits and compared as values;
//   * rule50 and the full-move counter are ignored, as Position::operator== does (Position.h:977-984).
//
// Only anchors (the record before a run, and chain heads) are decoded from scratch.
#pragma once
#include "link.cuh"
#include "stream.cuh"

namespace nnp {

__device__ __forceinline__ RecordFields record_fields(u32 w8, u32 w9)
{
    RecordFields f;
    f.score = (int)(short)(w8 & 0xFFFF);
    f.mv = sfmove_to_move(w8 >> 16);
    f.ply = (int)(w9 & 0xFFFF);
    f.result = (int)(signed char)((w9 >> 16) & 0xFF);
    return f;
}

// the two field tests of isContinuation (:589-590), on the raw words 9 of both records
__device__ __forceinline__ bool fields_link(u32 prev_w9, u32 cur_w9)
{
    const int pr = (int)(signed char)((prev_w9 >> 16) & 0xFF), cr = (int)(signed char)((cur_w9 >> 16) & 0xFF);
    const int pp = (int)(prev_w9 & 0xFFFF), cp = (int)(cur_w9 & 0xFFFF);
    return pr == -cr && pp + 1 == cp;
}

// From-scratch decode of record `rec` (pos_from_packed_sfen :364-446). Not inlined: the kernels that
// walk chains keep one copy of the Huffman loop, off their hot path -- and the compact token-driven loop, not
// the unrolled flat form (sfen_decode_flat: 900 straight-line instructions): k_walk_chains already waits for
// instructions (no_instruction 0.47 stalls per issue), with the flat form its anchors cost 12 % more at 8 plies
// per chain (7.47 against 6.5 ms per 100 M) although the flat form executes half the instructions.
static __device__ __noinline__ bool decode_record(const unsigned char* __restrict__ bin, u64 rec, Pos& P)
{
    const u32* w = reinterpret_cast<const u32*>(bin + rec * 40);
    pos_clear(P);
    return sfen_decode([&](int j) { return w[j]; }, P, [](int, u32) {});
}

static __device__ __noinline__ void store_stem_cold(const Pos& P, u32 w8, u32 w9, u32* stem_out)
{
    store_stem(P, record_fields(w8, w9), stem_out);
}

// The reference's own route for one record, taken when the splice cannot decide: decode the record
// and compare positions (isContinuation :587-593). On entry Q is the predecessor's position after
// its move; on return it is the record's own position. Emits the stem when the record starts a chain.
static __device__ __noinline__ bool walk_slow(const unsigned char* __restrict__ bin, u64 rec, Pos& Q, u32 c8, u32 c9,
                                              u32* __restrict__ stems, bool& ok)
{
    Pos C;
    ok = decode_record(bin, rec, C);
    const bool cont = ok && pos_equal(Q, C);
    if (!cont) store_stem(C, record_fields(c8, c9), stems + rec * 8);
    Q = C;
    return cont;
}

// The state a thread carries along a chain: the previous record's raw stream, score / move / ply words
// and decoded position, and the next record's words already on their way from memory.
struct WalkState {
    u32 Wp[8], p8, p9;
    Pos P;
    bool valid;
    Move pmv;               // the move of the record in front, decoded once
    int pmoved;             // the piece on pmv.from in P if the encoder looked it up, else -1
    uint2 n0, n1, n2, n3, n4;  // record `rec` (software pipeline: loaded while the record before it is processed)
    u64 rec;                // the next record to look at
};

// Opens a walk at record `a`: decodes it from scratch (the anchor; when it is a chain head also emits its
// code 0 and its stem) and starts the loads of record a + 1 (records at or behind `e` are never touched).
// on_error(rec) reports "Improperly encoded bin sfen" (:407-408, :441-442).
template <typename ErrFn>
__device__ __forceinline__ void walk_open(WalkState& S, const unsigned char* __restrict__ bin, u64 a, bool a_is_head, u64 e,
                                          u32* __restrict__ codes, u32* __restrict__ stems, ErrFn on_error)
{
    {
        const uint2* src = reinterpret_cast<const uint2*>(bin + a * 40);
        const uint2 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4];
        S.Wp[0] = v0.x; S.Wp[1] = v0.y; S.Wp[2] = v1.x; S.Wp[3] = v1.y; S.Wp[4] = v2.x; S.Wp[5] = v2.y; S.Wp[6] = v3.x; S.Wp[7] = v3.y;
        S.p8 = v4.x; S.p9 = v4.y;
    }
    // The out-of-line helpers take a position by reference; they get a copy of their own so that the
    // walk's position never has its address taken and stays in registers through the loop.
    {
        Pos anchor;
        S.valid = decode_record(bin, a, anchor);
        if (a_is_head) {
            codes[a] = 0u;
            store_stem_cold(anchor, S.p8, S.p9, stems + a * 8);
        }
        S.P = anchor;
    }
    if (!S.valid) on_error(a);
    S.pmv = sfmove_to_move(S.p8 >> 16);
    S.pmoved = -1;
    S.rec = a + 1;
    S.n0 = S.n1 = S.n2 = S.n3 = S.n4 = make_uint2(0u, 0u);
    if (S.rec < e) {
        const uint2* src = reinterpret_cast<const uint2*>(bin + S.rec * 40);
        S.n0 = src[0]; S.n1 = src[1]; S.n2 = src[2]; S.n3 = src[3]; S.n4 = src[4];
    }
}

// One record (S.rec < e): false when its ply / result fields rule out a continuation -- it is a chain head
// whatever its position and is left untouched -- otherwise the record gets its code (and its stem when its
// position turns out not to continue the chain) and S moves on to the next one.
template <typename ErrFn>
__device__ __forceinline__ bool walk_step(WalkState& S, const unsigned char* __restrict__ bin, u64 e, u32* __restrict__ codes,
                                          u32* __restrict__ stems, ErrFn on_error, const StepTables* T, const BleedLog* B)
{
    const u64 rec = S.rec;
    u32 Wc[8], c8, c9;
    Wc[0] = S.n0.x; Wc[1] = S.n0.y; Wc[2] = S.n1.x; Wc[3] = S.n1.y; Wc[4] = S.n2.x; Wc[5] = S.n2.y; Wc[6] = S.n3.x; Wc[7] = S.n3.y;
    c8 = S.n4.x; c9 = S.n4.y;
    if (!S.valid || !fields_link(S.p9, c9)) return false;  // rec starts a chain
    if (rec + 1 < e) {
        const uint2* src = reinterpret_cast<const uint2*>(bin + (rec + 1) * 40);
        S.n0 = src[0]; S.n1 = src[1]; S.n2 = src[2]; S.n3 = src[3]; S.n4 = src[4];
    }
    const Move pm = S.pmv;
    const Move cm = sfmove_to_move(c8 >> 16);
    const int moved = S.pmoved >= 0 ? S.pmoved : pm.from < 64 ? pos_piece_at(S.P, pm.from) : NO_PIECE;
    const bool spliced = stream_apply_move(S.Wp, S.P, pm, moved, T);  // Wp becomes the expected stream
    pos_do_move(S.P, pm, moved, T);                                   // Position::afterMove
    bool cont = false;
    if (spliced) {
        const int end = stream_board_end(S.P);
        u32 diff = 0;
#pragma unroll
        for (int k = 0; k < STREAM_BOARD_WORDS; ++k) diff |= (S.Wp[k] ^ Wc[k]) & stream_low_mask(end, k);
        if (diff == 0) {
            // same side to move, kings and board; castling(4) and ep(1[+6]) follow the board bits
            const u32* cw = reinterpret_cast<const u32*>(bin + rec * 40);
            const int wi = end >> 5;  // end <= 203: wi + 1 <= 7
            const u32 t = __funnelshift_r(cw[wi], cw[wi + 1], end & 31);
            const int cr = (int)(t & 15u);
            int ep = SQ_NONE;
            if (t & 16u) {
                const int sq = (int)((t >> 5) & 63u);
                if (ep_possible(S.P, sq, S.P.stm)) ep = sq;  // setEpSquare Position.h:868-872
            }
            cont = cr == S.P.cr && ep == S.P.ep;
        }
    }
    if (!cont) {
        bool ok;
        Pos Q = S.P;
        cont = walk_slow(bin, rec, Q, c8, c9, stems, ok);
        S.P = Q;
        if (!ok) on_error(rec);
        S.valid = ok;
    }
    u32 code = 0u;
    int cmoved = -1;
    if (cont) {
        int nbits;
        u32 bleed = 0;
        const u32 bits = encode_ply(S.P, cm, (int)(short)(c8 & 0xFFFF), (int)(short)(-(int)(short)(S.p8 & 0xFFFF)), nbits, T,
                                    &bleed, &cmoved);
        code = bits | (1u << (31 - nbits));
        if (bleed && B) bleed_report(*B, rec, bleed);
    }
    codes[rec] = code;
#pragma unroll
    for (int k = 0; k < 8; ++k) S.Wp[k] = Wc[k];
    S.p8 = c8;
    S.p9 = c9;
    S.pmv = cm;
    S.pmoved = cmoved;
    S.rec = rec + 1;
    return true;
}

// One work item: open at record `a`, then produce the codes of records a+1 .. e-1 by walking. The walk
// stops at the first chain head. at_head(rec, a) decides what happens to it: true = the walk goes on from it
// (it becomes the next anchor), false = the walk ends (the caller parked it for a later round, or it
// belongs to another thread). at_end(rec) is called when the walk reaches `e` without having met a head
// there.
template <typename ErrFn, typename HeadFn, typename EndFn>
__device__ __forceinline__ void walk_item(const unsigned char* __restrict__ bin, u64 a, bool a_is_head, u64 e,
                                          u32* __restrict__ codes, u32* __restrict__ stems, ErrFn on_error, HeadFn at_head,
                                          EndFn at_end, const StepTables* T = nullptr, const BleedLog* B = nullptr)
{
    for (;;) {
        WalkState S;
        walk_open(S, bin, a, a_is_head, e, codes, stems, on_error);
        while (S.rec < e && walk_step(S, bin, e, codes, stems, on_error, T, B)) {}
        if (S.rec >= e) {
            at_end(S.rec);
            return;
        }
        if (!at_head(S.rec, a)) return;
        a = S.rec;
        a_is_head = true;
    }
}

}  // namespace nnp
