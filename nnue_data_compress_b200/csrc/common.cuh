// common.cuh -- shared declarations of the libnnuepack kernels and their host drivers.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "chess.cuh"

namespace nnp {

constexpr u32 CHUNK_THRESHOLD = 1u << 20;          // suggestedChunkSize compress_file.cpp:20
constexpr u32 MAX_CHUNK_SIZE = 100u * (1u << 20);  // maxChunkSize compress_file.cpp:22
constexpr u64 NO_ERROR_IDX = ~0ull;
constexpr u64 NO_CARRY = ~0ull;      // chunk orbit: no chunk was opened before this shard
constexpr u64 NATURAL_SIZE = ~0ull;  // chunk emission: the last chunk ends with the payload

// Summary of a run of records for the segmented payload scan ("payload scan").
// A run is described by what happens before its first chain head (bits/plies that still
// belong to the chain that was open when the run began), the bytes that are fully
// determined inside it, and the chain left open at its end.
struct Agg {
    u64 bytes;       // from the first head's stem through the last head's stem + numPlies field
    u32 pre_bits;    // movetext bits before the first head
    u32 pre_plies;
    u32 post_bits;   // movetext bits after the last head
    u32 post_plies;
    u32 heads;
    u32 pad;
};

__host__ __device__ __forceinline__ u64 ceil8(u64 bits) { return (bits + 7) >> 3; }

__host__ __device__ __forceinline__ Agg agg_combine(const Agg& a, const Agg& b)
{
    Agg r;
    r.pad = 0;
    if (b.heads == 0) {
        if (a.heads == 0) {
            r.bytes = 0;
            r.pre_bits = a.pre_bits + b.pre_bits;
            r.pre_plies = a.pre_plies + b.pre_plies;
            r.post_bits = 0;
            r.post_plies = 0;
            r.heads = 0;
        } else {
            r = a;
            r.post_bits = a.post_bits + b.pre_bits;
            r.post_plies = a.post_plies + b.pre_plies;
        }
    } else {
        if (a.heads == 0) {
            r = b;
            r.pre_bits = a.pre_bits + b.pre_bits;
            r.pre_plies = a.pre_plies + b.pre_plies;
        } else {
            r.bytes = a.bytes + ceil8((u64)a.post_bits + b.pre_bits) + b.bytes;
            r.pre_bits = a.pre_bits;
            r.pre_plies = a.pre_plies;
            r.post_bits = b.post_bits;
            r.post_plies = b.post_plies;
            r.heads = a.heads + b.heads;
        }
    }
    return r;
}

// totals written by the tile-aggregate scan and read back by the host driver
struct CompressTotals {
    u64 payload_bytes;  // sum over chains of 34 + ceil(bits/8)
    u64 heads;          // number of chains
    u64 chunks;         // written by the chunk-orbit kernel
    u64 error_index;    // first record whose sfen is malformed, or NO_ERROR_IDX
    u64 parked[2];      // chain heads parked by the current / next round of the chain walk
    u64 bleeds;         // plies whose ids do not fit their fields (see encode_ply); listed in BleedLog::list
};


struct DecompressTotals {
    u64 positions;
    u64 error_chunk;   // first chunk with a decode error (truncated movetext), or NO_ERROR_IDX
    u64 slow_chunks;   // chunks that needed the sequential fallback
    u64 violations;    // links of the optimistic walk that did not hold
    u64 false_candidates;      // candidates the exhaustive walk found not to be chain starts
    u64 false_sample[8];       // the first few of them: chunk << 32 | offset (diagnostics)
    u64 heads_only_chunks;     // chunks that hold nothing but single positions (k_chunk_heads_only)
};

// one parsed .plain record (TrainingDataEntry, compress_file.cpp:548-555), 64 bytes
struct Entry {
    u64 occ0, occ1, t0, t1, t2;
    u32 meta;       // stm | ep << 1 | castling << 8 | rule50 << 12
    u32 pos_ply;    // Position::m_ply
    u32 mv;         // from | to << 6 | type << 12 | promoted piece << 14
    u32 score_ply;  // score (int16) | ply << 16
    u32 result;     // int16
    u32 pad;
};

struct PlainTotals {
    u64 error_pos;  // byte offset of the first construct the parser rejects, or NO_ERROR_IDX
    u64 committed;  // scratch for the flush-boundary orbit
    u64 inherit;    // some record does not define all five keys itself: the inheritance passes are needed
};

#ifdef __CUDACC__
// 16 bytes at p, zero-filled outside [lo, hi)
__device__ __forceinline__ uint4 load16_clipped(const unsigned char* p, const unsigned char* lo, const unsigned char* hi)
{
    if (p >= lo && p + 16 <= hi) return *reinterpret_cast<const uint4*>(p);  // p is 16-byte aligned by construction
    u32 w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; ++i)
        if (p + i >= lo && p + i < hi) w[i >> 2] |= (u32)p[i] << ((i & 3) * 8);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

#endif

}  // namespace nnp
