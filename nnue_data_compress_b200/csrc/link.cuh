// link.cuh -- the per-record half of CompressedTrainingDataEntryWriter::addTrainingDataEntry
// (compress_file.cpp:1061-1092), shared by the .bin and the .plain compressors: decide whether a
// record continues its predecessor (isContinuation :587-593) and produce either its movetext bit
// string (addMoveScore :877-989) or its 32-byte stem (packEntry :997-1020).
#pragma once
#include "chess.cuh"

namespace nnp {

// plies whose ids do not fit their fields (see encode_ply): record index | packed ids << 32, appended in
// no particular order; *count may run past the capacity (the host then refuses the input)
constexpr u64 BLEED_CAP = 1u << 20;
struct BleedLog {
    u64* list;
    u64* count;
};
__device__ __forceinline__ void bleed_report(const BleedLog& B, u64 rec, u32 packed)
{
#ifdef NNP_HOST_SIM
    const u64 i = (*B.count)++;
#else
    const u64 i = atomicAdd(reinterpret_cast<unsigned long long*>(B.count), 1ull);
#endif
    if (i < BLEED_CAP) B.list[i] = (rec & 0xFFFFFFFFull) | ((u64)packed << 32);
}

struct RecordFields {
    Move mv;
    int score;   // int16
    int ply;     // uint16
    int result;  // int16 (int8 widened on the .bin path)
};

// Returns the record's code word: 0 for a chain head, otherwise the ply's bits left-aligned and
// terminated by a single 1 bit (never zero), so that the bit count is 32 - ffs(code).
__device__ __forceinline__ u32 link_code(bool has_prev, const Pos& prev, const RecordFields& pf, const Pos& cur,
                                         const RecordFields& cf, u32* bleed = nullptr)
{
    if (bleed) *bleed = 0u;
    bool cont = false;
    if (has_prev && pf.result == -cf.result && pf.ply + 1 == cf.ply) {  // short-circuit order of :589-592
        Pos a = prev;
        pos_do_move(a, pf.mv);  // Position::afterMove
        cont = pos_equal(a, cur);
    }
    if (!cont) return 0u;
    int nbits;
    const int last_score = (int)(short)(-pf.score);  // m_lastScore (:838, :986)
    const u32 bits = encode_ply(cur, cf.mv, cf.score, last_score, nbits, nullptr, bleed);
    return bits | (1u << (31 - nbits));
}

// packEntry (:997-1020) of a chain head into 8 words at `stem_out` (32-byte aligned)
__device__ __forceinline__ void store_stem(const Pos& cur, const RecordFields& cf, u32* stem_out)
{
    u32 s[8];
    stem_pack(cur, cf.mv, cf.score, cf.ply, cf.result, s);
    uint4* d = reinterpret_cast<uint4*>(stem_out);
    d[0] = make_uint4(s[0], s[1], s[2], s[3]);
    d[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

__device__ __forceinline__ u32 link_and_encode(bool has_prev, const Pos& prev, const RecordFields& pf, const Pos& cur,
                                               const RecordFields& cf, u32* stem_out, u32* bleed = nullptr)
{
    const u32 code = link_code(has_prev, prev, pf, cur, cf, bleed);
    if (code == 0u) store_stem(cur, cf, stem_out);
    return code;
}

}  // namespace nnp
