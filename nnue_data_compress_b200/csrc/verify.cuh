// verify.cuh -- the link test of the optimistic decode strategy (see k_emit_chains_verify).
#pragma once
#include "common.cuh"

namespace nnp {

// Candidate i of chunk cand_chunk[i] (payload length clen) starts at cand_off[i] and its chain ends
// at `end`. Its link of the reader's walk (Reader::next / fetchNextChunkIfNeeded, :1154-1213) holds
// when the chunk's first candidate sits at offset 0, the chain ends exactly where the next candidate
// starts, and behind the chunk's last chain fewer than 34 bytes remain. Candidates marked by
// k_mark_conflicts (cand_cnt == 0) are not part of the walk.
__device__ __forceinline__ bool reader_links_hold(const u32* __restrict__ cand_chunk, const u32* __restrict__ cand_off,
                                                  const u32* __restrict__ cand_cnt, u64 ncand, u64 i, u32 end, u32 clen)
{
    const u32 c = cand_chunk[i];
    u64 prev = i, next = i + 1;  // neighbours in the chunk, skipping marked candidates
    while (prev > 0 && cand_chunk[prev - 1] == c && cand_cnt[prev - 1] == 0) --prev;
    while (next < ncand && cand_chunk[next] == c && cand_cnt[next] == 0) ++next;
    const bool first = prev == 0 || cand_chunk[prev - 1] != c;
    const bool last = next == ncand || cand_chunk[next] != c;
    if (first && cand_off[i] != 0) return false;
    if (last) return (u64)end + 34 > clen;
    return cand_off[next] == end;
}

}  // namespace nnp
