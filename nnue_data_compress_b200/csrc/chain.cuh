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
    int score, ply, result;
    int last_score;
    u32 num_plies;
};

__device__ __forceinline__ void chain_open(const unsigned char* s, ChainCursor& c)
{
    stem_unpack([&](int i) { return (u32)s[i]; }, c.pos, c.mv, c.score, c.ply, c.result);
    c.num_plies = ((u32)s[32] << 8) | (u32)s[33];
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
    if (!decode_ply(r, c.pos, c.last_score, m, sc, strict, T)) return false;
    c.mv = m;
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

// walk_chain writing 40-byte .bin records rec0, rec0 + 1, ... (below rec_limit). The Huffman stream
// of the position is carried along the chain (stream.cuh): built once for the chain head, then
// spliced per move. `col` is the thread's 8-word scratch column in shared memory.
__device__ __forceinline__ bool emit_chain_bin(const unsigned char* s, u32 bytes_after_stem, unsigned char* out, u64 rec0,
                                               u64 rec_limit, u32* col, int stride, u32& consumed,
                                               const StepTables* T = nullptr)
{
    u32 W[8];
    bool spliced = false;
    return walk_chain(
        s, bytes_after_stem,
        [&](const ChainCursor& cc) {
            const int moved = pos_piece_at(cc.pos, cc.mv.from);
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
            uint2* d = reinterpret_cast<uint2*>(out + (rec0 + k) * 40);
            d[0] = make_uint2(w[0], w[1]);
            d[1] = make_uint2(w[2], w[3]);
            d[2] = make_uint2(w[4], w[5]);
            d[3] = make_uint2(w[6], w[7]);
            d[4] = make_uint2(w8, w9);
        },
        consumed, T);
}

}  // namespace nnp
