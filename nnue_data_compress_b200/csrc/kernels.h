// kernels.h -- host-side launchers of the libnnuepack kernels (one per kernel, no logic).
#pragma once
#include "common.cuh"

namespace nnp {

// ---- lookup tables (chess.cuh: one copy per translation unit), filled once per device at nnp_init
void init_tables_compress(cudaStream_t s);
void init_tables_decompress(cudaStream_t s);
void init_tables_halfkp(cudaStream_t s);

// ---- compress (.bin -> .binpack), compress.cu
void launch_heads_transcode(const void* d_bin, u64 n, u32* codes, u32* stems, CompressTotals* tot, u64* bleed_list,
                            cudaStream_t s);
u64 heads_direct_bytes(u64 n);  // size of the .binpack of n single-position chains
void launch_heads_direct(const void* d_bin, u64 n, void* d_out, u32* fallback, cudaStream_t s);
void launch_decode_link_encode(const void* d_bin, u64 n, u32* codes, u32* stems, CompressTotals* tot, u64* bleed_list,
                               cudaStream_t s);
u64 walk_runs(u64 n);  // number of runs = upper bound of the parked heads of any round
void launch_sample_heads(const void* d_bin, u64 n, u64 stride, u64 samples, u64* heads, cudaStream_t s);
int walk_run_records();
void launch_walk_runs(const void* d_bin, u64 n, u64 run_lo, u64 run_hi, u32* codes, u32* stems, CompressTotals* tot,
                      u32* park_list, u64* park_count, u64* bleed_list, cudaStream_t s);
void launch_walk_items(const void* d_bin, u64 n, u32* codes, u32* stems, CompressTotals* tot, const u32* items, u64 n_items,
                       u32* park_list, u64* park_count, u64* bleed_list, cudaStream_t s);
void launch_walk_chains(const void* d_bin, u64 n, u32* codes, u32* stems, CompressTotals* tot, u64* bleed_list, cudaStream_t s);
constexpr u64 BLEED_LIST_CAP = 1u << 20;  // = BLEED_CAP (link.cuh)
void launch_write_bleed(const u32* codes, u64 n, const Agg* tile_prefix, u32* payload, const u64* bleed_list, u64 bleed_count,
                        u64 rec_base, cudaStream_t s);
u64 scan_tiles(u64 n);
void launch_tile_aggregate(const u32* codes, u64 n, Agg* tile_agg, cudaStream_t s);
u64 scan_blocks(u64 ntiles);  // entries of the block_tot scratch
void launch_scan_aggregates(Agg* tile_agg, u64 ntiles, Agg* block_tot, CompressTotals* tot, cudaStream_t s);
void launch_write_payload(const u32* codes, const u32* stems, u64 n, const Agg* tile_prefix, u32* payload,
                          u64* head_off, bool dense, cudaStream_t s);
void launch_head_next(const u64* head_off, u64 heads, u32* next, CompressTotals* tot, cudaStream_t s);
void launch_chunk_orbit(const u64* head_off, const u32* next, CompressTotals* tot, u64* seg_off, u64 max_chunks, u64 base,
                        u64 carry, bool speculate, cudaStream_t s);
void launch_orbit_table(const u64* head_off, const u32* next, const CompressTotals* tot, u64* table, u64 entries, cudaStream_t s);
void launch_orbit_resolve(const u64* tables, u64 entries, const u64* sizes, int world, int rank, u64* out, cudaStream_t s);
void launch_emit_chunks(const void* payload, const u64* seg_off, u64 chunks, void* out, u64 last_size, cudaStream_t s);
void launch_find_head(const u32* codes, u64 n, u64 start, u64* out, cudaStream_t s);

// ---- decompress (.binpack -> .bin), decompress.cu
constexpr int CAND_TILE = 4096;  // byte offsets tested per block of the candidate kernels

struct ChunkInfo {
    u64 chunks;  // chunks found (may exceed the table capacity: then re-walk with a larger table)
    u64 tiles;   // candidate tiles over all stored chunks
    int status;  // 0 or the nnp_status of the header that stopped the walk
    int pad;
    u64 range_lo, range_hi;  // the chunk range of (world, rank) and its bytes in the file, headers included
    u64 byte_lo, byte_hi;
};
struct ChunkTable {
    u64* start;      // file offset of the chunk payload
    u32* len;        // payload bytes
    u64* tile_base;  // [chunks + 1] exclusive prefix of candidate tiles
    ChunkInfo* info;
};

size_t walk_scratch_bytes();
void set_walk_segment_bytes(u64 v);  // test hook (0 = default)
void launch_walk_chunks(const void* d_in, u64 n, ChunkTable tab, u64 max_chunks, u32 world, u32 rank, void* scratch, cudaStream_t s);
void launch_chunk_heads_only(const void* d_in, ChunkTable tab, u64 chunks, u32* chunk_flag, u32* chunk_stems, u64* flagged,
                             cudaStream_t s);
void launch_emit_heads_only(const void* d_in, ChunkTable tab, u64 chunks, const u64* chunk_base, u64 positions, void* d_out,
                            u64 rec_limit, cudaStream_t s);
void launch_candidates_scan(const void* d_in, u64 n_in, ChunkTable tab, u64 chunks, u64 tiles, u32* tile_count, u32* tile_flags,
                            u32 debug_reject_mod, const u32* chunk_flag, const u64* scan_prefix, u64 scan_total, cudaStream_t s);
// collapse mode: the tiles of the chunks of single positions get their (one) entry here, scan_tiles[c] = tiles of
// chunk c that k_candidates_scan still has to look at (its grid is their sum, `scan_prefix` their exclusive sum)
void launch_collapsed_tiles(ChunkTable tab, u64 chunks, const u32* chunk_flag, u32* tile_count, u32* tile_flags, u32* scan_tiles,
                            cudaStream_t s);
// `collapsed` (nullable, per chunk): the chunk is one entry of the candidate list (see k_candidates_scan)
void launch_emit_heads_chunks(const void* d_in, ChunkTable tab, u64 chunks, const u32* collapsed, const u64* tile_prefix,
                              const u64* cand_rec, void* d_out, u64 rec_limit, cudaStream_t s);
void launch_mark_conflicts(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off, u32* cand_cnt,
                           u64 ncand, cudaStream_t s);
void launch_candidates_list(const void* d_in, ChunkTable tab, u64 tiles, const u32* tile_flags, const u64* tile_prefix,
                            u32* cand_chunk, u32* cand_off, u32* cand_cnt, const u32* collapsed, cudaStream_t s);
void launch_check_chunks(ChunkTable tab, u64 chunks, const u64* tile_prefix, u64* violations, cudaStream_t s);
// candidates [cand_lo, cand_hi)
void launch_emit_chains_verify(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off,
                               const u32* cand_cnt, const u64* cand_rec, u64 ncand, u64 cand_lo, u64 cand_hi, void* out,
                               u64 rec_limit, u32* cand_next, u64* violations, const u32* collapsed, cudaStream_t s);
void launch_exclusive_sum(const u32* in, u64 n, u64* out, cudaStream_t s);
void launch_exclusive_sum64(const u64* in, u64 n, u64* out, cudaStream_t s);
void launch_probe_chains(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off, u64 ncand,
                         u32* cand_next, u32* cand_cnt, u32* cand_tlen, cudaStream_t s);
void launch_resolve_chunks(ChunkTable tab, u64 chunks, const u64* tile_prefix, const u32* cand_off, const u32* cand_next,
                           const u32* cand_cnt, u32* cand_base, u32* chunk_count, u32* chunk_slow, const u32* cand_tlen,
                           u64* cand_tbase, u64* chunk_tbytes, cudaStream_t s);
void launch_slow_count(const void* d_in, ChunkTable tab, u64 chunks, const u32* chunk_slow, u32* chunk_count,
                       u64* chunk_tbytes, DecompressTotals* tot, cudaStream_t s);
void launch_emit_chains_text(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off, const u32* cand_base,
                             const u64* cand_tbase, u64 ncand, const u64* chunk_tbase, void* out, DecompressTotals* tot,
                             cudaStream_t s);
void launch_slow_emit_text(const void* d_in, ChunkTable tab, u64 chunks, const u32* chunk_slow, const u64* chunk_tbase,
                           void* out, cudaStream_t s);
void launch_emit_chains(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off, const u32* cand_base,
                        u64 ncand, const u64* chunk_base, void* out, DecompressTotals* tot, const u64* placed_rec,
                        const u32* placed_next, cudaStream_t s);
void launch_slow_emit(const void* d_in, ChunkTable tab, u64 chunks, const u32* chunk_slow, const u64* chunk_base, void* out,
                      cudaStream_t s);

// ---- HalfKP feature rows, halfkp.cu (white / black: [positions][32] int, meta: [positions] nnp_halfkp_meta)
void launch_emit_chains_halfkp_verify(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off,
                                      const u32* cand_cnt, const u64* cand_rec, u64 ncand, int* white, int* black, void* meta,
                                      u64* violations, cudaStream_t s);
void launch_emit_chains_halfkp(const void* d_in, ChunkTable tab, const u32* cand_chunk, const u32* cand_off,
                               const u32* cand_base, u64 ncand, const u64* chunk_base, int* white, int* black, void* meta,
                               DecompressTotals* tot, cudaStream_t s);
void launch_slow_emit_halfkp(const void* d_in, ChunkTable tab, u64 chunks, const u32* chunk_slow, const u64* chunk_base,
                             int* white, int* black, void* meta, cudaStream_t s);
void launch_bin_halfkp(const void* d_bin, u64 n, int* white, int* black, void* meta, CompressTotals* tot, cudaStream_t s);

// ---- .plain text, plain.cu
u64 large_sum_tiles(u64 n);
void launch_exclusive_sum_large(const u32* in, u64 n, u64* out, u32* tile_sum, u64* tile_prefix, cudaStream_t s);
u64 find_tiles(u64 n);
void launch_find_records(bool write, const void* text, u64 n, u32* tile_count, const u64* tile_prefix, u64* rec_pos,
                         PlainTotals* tot, cudaStream_t s);
void launch_parse_records(const void* text, u64 n, const u64* rec_pos, u64 nrec, Entry* entries, PlainTotals* tot,
                          cudaStream_t s);
u64 defs_tiles(u64 nrec);
void launch_parse_inherited(const void* text, u64 n, const u64* rec_pos, u64 nrec, u64* defs, u64* tile_max, Entry* entries,
                            PlainTotals* tot, cudaStream_t s);
void launch_entries_link_encode(const Entry* entries, u64 n, u32* codes, u32* stems, CompressTotals* tot, u64* bleed_list,
                                cudaStream_t s);
void launch_entries_to_bin(const Entry* entries, u64 n, void* out, cudaStream_t s);
void launch_bin_text(bool write, const void* bin, u64 n, u32* lens, const u64* offs, void* out, CompressTotals* tot,
                     cudaStream_t s);
void launch_text_flush_orbit(const void* text, u64 limit, int drop_last, u64* committed, cudaStream_t s);

// ---- synthetic input, generate.cu
void launch_play_games(bool write, u64 n_games, u32 max_plies, u64 seed, u32* game_len, const u64* game_base, void* out,
                       u64 n_positions, cudaStream_t s);

}  // namespace nnp
