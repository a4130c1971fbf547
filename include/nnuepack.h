/*
 * nnuepack.h -- C ABI of libnnuepack.so, the B200-native (sm_100a) replacement for the
 * conversion hot path of Sopel97/nnue_data_compress.
 *
 * The reference has no library API: its six file drivers are file-local functions of
 * src/compress_file.cpp, reached only from convert() (:1593-1621). Each entry point below
 * is the buffer-level equivalent of one of those drivers -- same formats, same bytes --
 * and is what a maintainer binds in place of the driver body.
 * All citations are relative to /root/reference.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every buffer; nothing is allocated
 *     on the caller's behalf except through nnp_host_alloc().
 *   - "host" entry points take host pointers (pageable or pinned; pinned memory from
 *     nnp_host_alloc() gives full PCIe speed) and do H2D, kernels, D2H internally.
 *   - "_dev" entry points take device pointers of the current device and never touch
 *     the host data path; they are what the throughput numbers are measured on.
 *   - return value: NNP_OK (0) or a negative nnp_status. On NNP_ERR_BAD_SFEN,
 *     NNP_ERR_BAD_MAGIC and NNP_ERR_CHUNK_TOO_LARGE the output holds exactly the bytes the
 *     reference tool would have written before it printed the matching message and
 *     exited (compress_file.cpp:407-408, :441-442, :504-518, :1094-1106, :1704-1708).
 *   - capacity query: pass out == NULL; *out_bytes receives a safe upper bound.
 *   - there is NO CPU fallback: every conversion runs CUDA kernels; without a usable
 *     device nnp_init() fails with NNP_ERR_NO_DEVICE and every other call with
 *     NNP_ERR_NOT_INITIALISED.
 *   - one context per device. A host thread works on the device it bound itself to with
 *     nnp_init() / nnp_bind_device(); a thread that never did uses the first device the process
 *     initialised. Calls on one device are serialised, calls on different devices run
 *     concurrently: one process per GPU (torchrun) and one host thread per GPU (the *_multi
 *     drivers, the CLI with NNP_DEVICES) are both supported.
 */
#ifndef NNUEPACK_H
#define NNUEPACK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum nnp_status {
    NNP_OK = 0,
    NNP_ERR_BAD_MAGIC = -1,       /* "Invalid binpack file or chunk."                      :504-507 */
    NNP_ERR_CHUNK_TOO_LARGE = -2, /* "Chunks size larger than supported. Malformed file?"  :515-518 */
    NNP_ERR_BAD_SFEN = -3,        /* "Improperly encoded bin sfen"                         :407-408, :441-442 */
    NNP_ERR_TRUNCATED = -4,       /* chunk/movetext runs past the end of the input */
    NNP_ERR_NOMEM = -5,           /* device or host allocation failed */
    NNP_ERR_BAD_ARG = -6,
    NNP_ERR_BAD_TEXT = -7,        /* .plain input on which the reference's std::stoi would throw */
    NNP_ERR_CAPACITY = -8,        /* output buffer too small; *out_bytes = bytes required */
    NNP_ERR_NO_DEVICE = -9,       /* no CUDA device / extension cannot run: there is no CPU path */
    NNP_ERR_NOT_INITIALISED = -10,
    NNP_ERR_CUDA = -11,           /* a CUDA runtime call failed; see nnp_last_cuda_error() */
    NNP_ERR_WINDOW = -12          /* sharded compression: the overlap window does not reach the next chain head */
} nnp_status;

/* ---- lifecycle ------------------------------------------------------------------ */

/* Binds the process to CUDA device `device` (0-based; for one-process-per-GPU launches pass
 * LOCAL_RANK), creates the library's streams and uploads the attack tables (the GPU
 * counterpart of the reference's static initialisation, src/chess/Bitboard.cpp:460-464). */
int nnp_init(int device);
/* Initialises devices 0 .. n_devices - 1 (n_devices <= 0: every visible device) and binds the calling
 * thread to device 0. Returns the number of devices (> 0) or a negative nnp_status. This is the
 * "all GPUs of the box" form SURVEY.md 8b asks for; convert() of a multi-GPU CLI calls it once. */
int nnp_init_all(int n_devices);
/* Binds the calling host thread to an initialised device: its further calls run there. */
int nnp_bind_device(int device);
/* number of initialised devices, and the CUDA ordinal of the index-th of them (-1 past the end) */
int nnp_device_count(void);
int nnp_device_at(int index);
/* Releases every device of the process. */
void nnp_shutdown(void);
const char* nnp_strerror(int status);
/* Runs all further work on the caller's CUDA stream (a cudaStream_t of the bound device, e.g.
 * torch's current stream) so that the caller's events bracket it; NULL restores the library's own
 * stream. Calls still return only after their result size is known. The library's own stream is
 * non-blocking: device buffers written by work on another stream must be complete (or that stream
 * must be the one given here) before a *_dev entry point reads them. To name the legacy default
 * stream pass cudaStreamLegacy ((cudaStream_t)0x1), since NULL selects the library's stream. */
int nnp_set_stream(void* cuda_stream);
const char* nnp_last_cuda_error(void);
/* number of kernel launches issued by this library since nnp_init (for bench accounting) */
uint64_t nnp_kernel_launches(void);

/* Pinned host memory for full-speed H2D/D2H. */
void* nnp_host_alloc(size_t bytes);
void nnp_host_free(void* p);

/* ---- the six drivers, host buffers ------------------------------------------------ */

/* compressBin        compress_file.cpp:1338-1374  (.bin -> .binpack) */
int nnp_bin_to_binpack(const void* bin, size_t bin_bytes, void* out, size_t out_cap, size_t* out_bytes);
/* decompressBin      compress_file.cpp:1376-1412  (.binpack -> .bin) */
int nnp_binpack_to_bin(const void* binpack, size_t binpack_bytes, void* out, size_t out_cap, size_t* out_bytes);
/* compressPlain      compress_file.cpp:1246-1297  (.plain -> .binpack) */
int nnp_plain_to_binpack(const void* plain, size_t plain_bytes, void* out, size_t out_cap, size_t* out_bytes);
/* decompressPlain    compress_file.cpp:1299-1335  (.binpack -> .plain) */
int nnp_binpack_to_plain(const void* binpack, size_t binpack_bytes, void* out, size_t out_cap, size_t* out_bytes);
/* convertBinToPlain  compress_file.cpp:1414-1465  (.bin -> .plain) */
int nnp_bin_to_plain(const void* bin, size_t bin_bytes, void* out, size_t out_cap, size_t* out_bytes);
/* convertPlainToBin  compress_file.cpp:1467-1533  (.plain -> .bin) */
int nnp_plain_to_bin(const void* plain, size_t plain_bytes, void* out, size_t out_cap, size_t* out_bytes);

/* ---- the same drivers on device-resident buffers ------------------------------------ */
/* Input and output live in HBM of the bound device. The call returns after the result size
 * is known (it synchronises the library's stream once at the end). `d_out` must hold at
 * least the capacity returned by the out == NULL query. */
int nnp_bin_to_binpack_dev(const void* d_bin, size_t bin_bytes, void* d_out, size_t out_cap, size_t* out_bytes);
int nnp_binpack_to_bin_dev(const void* d_binpack, size_t binpack_bytes, void* d_out, size_t out_cap, size_t* out_bytes);
int nnp_plain_to_binpack_dev(const void* d_plain, size_t plain_bytes, void* d_out, size_t out_cap, size_t* out_bytes);
int nnp_binpack_to_plain_dev(const void* d_binpack, size_t binpack_bytes, void* d_out, size_t out_cap, size_t* out_bytes);
int nnp_bin_to_plain_dev(const void* d_bin, size_t bin_bytes, void* d_out, size_t out_cap, size_t* out_bytes);
int nnp_plain_to_bin_dev(const void* d_plain, size_t plain_bytes, void* d_out, size_t out_cap, size_t* out_bytes);

/* ---- .bin -> .binpack over several GPUs, byte-identical to one reference run ------------------
 *
 * compressBin is sequential in two places: chains (a record joins the chain of its predecessor,
 * compress_file.cpp:587-593, :1061-1092) and chunks (a chunk is flushed when a chain head arrives and
 * at least 1 MiB has gathered since the last flush, :1076-1080). Both are kept exact when the records
 * are split over ranks (one process per GPU):
 *
 *   1. every rank calls nnp_shard_compress_begin_dev() on the records it owns plus one halo record
 *      before them and an overlap window behind them. A chain belongs to the rank its head lies in.
 *   2. the ranks all-gather `payload_bytes` (8 bytes each): the exclusive prefix sum is each rank's
 *      `payload_base`, the offset of its chains in the header-less payload of the whole file.
 *   3. in rank order, nnp_shard_compress_orbit(payload_base, carry_in, ...) replays the flush rule
 *      over the rank's chain heads; `carry_out` (the global offset of the last chunk start so far)
 *      and the running number of chunks are handed to the next rank (16 bytes, NO payload).
 *   4. the ranks all-gather `first_start`; nnp_shard_compress_emit_dev(next_start, ...) writes the
 *      rank's slice of the .binpack: its payload with a BINP header in front of every chunk start.
 *      `next_start` = the first chunk start of any later rank, or the total payload size: it closes
 *      the rank's last chunk. The slice belongs at file offset payload_base + 8 * chunks_before.
 *
 * nnue_data_compress_b200/sharding.py drives these calls over torch.distributed (NCCL / gloo).
 * A malformed record (NNP_ERR_BAD_SFEN) is reported with its index and ends the sharded run: the
 * partial-output rule of the single-GPU entry point is not replayed across ranks.
 * The sequence begin .. emit keeps its state in the device's context: another compressor call on the same
 * device in between (nnp_bin_to_binpack*, nnp_plain_to_binpack*) reuses that memory and ends the sequence --
 * its remaining calls answer NNP_ERR_BAD_ARG until the next begin. */
typedef struct nnp_shard_info {
    uint64_t first_owned_record; /* index (in the buffer) of the first chain head >= own_lo */
    uint64_t end_owned_record;   /* index of the first chain head >= own_hi (records before it are owned) */
    uint64_t payload_bytes;      /* bytes of the owned chains: sum of 34 + ceil(movetext bits / 8) */
    uint64_t chains;
    uint64_t first_bad_record;   /* NNP_ERR_BAD_SFEN: index of the first malformed record, else ~0 */
} nnp_shard_info;
#define NNP_NO_CARRY (~(uint64_t)0)

/* d_bin: records [g0, g1) of the file, device memory; own_lo / own_hi: the rank's nominal range as
 * indices into that buffer (own_lo = 1 when a halo record is present, 0 for the first rank);
 * reaches_eof: the buffer ends with the file, so a chain that crosses own_hi may run to its end
 * (otherwise NNP_ERR_WINDOW asks for a larger overlap window). */
int nnp_shard_compress_begin_dev(const void* d_bin, size_t n_records, size_t own_lo, size_t own_hi, int reaches_eof,
                                 nnp_shard_info* info);
/* carry_in: NNP_NO_CARRY for the first rank that has any payload. first_start = NNP_NO_CARRY when
 * no chunk starts inside this rank's payload. */
int nnp_shard_compress_orbit(uint64_t payload_base, uint64_t carry_in, uint64_t* n_chunk_starts, uint64_t* first_start,
                             uint64_t* carry_out);
/* d_out == NULL: *out_bytes = size of the slice (payload_bytes + 8 * n_chunk_starts). */
int nnp_shard_compress_emit_dev(uint64_t next_start, void* d_out, size_t out_cap, size_t* out_bytes);

/* Step 3 without the rank-order wait. The carry into a rank lies before its payload, so a chunk can
 * only be entered at one of the chain heads of the rank's first MiB (at most 2^20 / 34 of them, the
 * size of a chain being at least 34 bytes). nnp_shard_compress_table_dev() follows the flush rule from
 * every such head to the end of the rank's payload and writes NNP_ORBIT_TABLE_ENTRIES entries of
 * three uint64 {offset of the entry head in the rank's payload (~0 = unused), offset of the last chunk
 * start, number of chunk starts} into d_table (device memory). The ranks all-gather their tables
 * (0.7 MB each); nnp_shard_compress_resolve_dev() then replays the whole chain locally by table lookup
 * and returns what nnp_shard_compress_orbit() and _emit_dev() need for this rank, plus the number of
 * chunks of the whole file. */
#define NNP_ORBIT_TABLE_ENTRIES 30848
int nnp_shard_compress_table_dev(void* d_table);
int nnp_shard_compress_resolve_dev(const void* d_tables, const uint64_t* payload_bytes, int world, int rank, uint64_t* carry_in,
                                   uint64_t* chunks_before, uint64_t* next_start, uint64_t* total_chunks);

/* ---- .binpack -> .bin over several GPUs: ONE file, chunk ranges per rank (BASELINE configs[2]) ------
 *
 * decompressBin (compress_file.cpp:1376-1412) reads chunk after chunk (hasNextChunk / readNextChunk,
 * :468-480) and every chunk starts with a stem (:1128-1214), so chunks decode independently: rank r of
 * `world` takes the contiguous chunk range [chunks * r / world, chunks * (r + 1) / world), decodes it
 * into its own buffer, the ranks all-gather their position counts (8 bytes each, the only exchange)
 * and rank r's records belong at byte 40 * (positions of the ranks before it) of the .bin file.
 * nnp_binpack_chunk_range() is the header walk alone on host memory (the rank then only needs
 * bytes [byte_lo, byte_hi) of the file on its device and calls nnp_binpack_to_bin_dev on them);
 * nnp_shard_decompress_dev() does walk + decode on a file that is resident on the device.
 * Status: a broken header (NNP_ERR_BAD_MAGIC / _CHUNK_TOO_LARGE / _TRUNCATED) is reported by every
 * rank after the chunks in front of it have been decoded; the reference's 1 MiB flush rule for the
 * records before an error is a property of the whole run and is not replayed per rank. */
typedef struct nnp_chunk_range {
    uint64_t chunks_total;       /* well-formed chunks of the whole file */
    uint64_t chunk_lo, chunk_hi; /* the rank's chunks [lo, hi) */
    uint64_t byte_lo, byte_hi;   /* their bytes in the file, chunk headers included */
    uint64_t positions;          /* nnp_shard_decompress_dev: positions the rank decoded */
} nnp_chunk_range;
int nnp_binpack_chunk_range(const void* binpack, size_t binpack_bytes, int world, int rank, nnp_chunk_range* range);
int nnp_binpack_chunk_range_dev(const void* d_binpack, size_t binpack_bytes, int world, int rank, nnp_chunk_range* range);
/* d_out == NULL: count only (*out_bytes = 40 * positions of the rank's range). */
int nnp_shard_decompress_dev(const void* d_binpack, size_t binpack_bytes, int world, int rank, void* d_out, size_t out_cap,
                             size_t* out_bytes, nnp_chunk_range* range);

/* ---- whole files of any size, on every device of the process ------------------
 * File-to-file forms of the two headline drivers: the input is processed in slabs of about
 * `slab_bytes` (0 = default: 256 MiB of records / 32 MiB of chunks) that are dealt to all initialised
 * devices (nnp_init_all); per device a loader, a compute and a drainer thread work on double buffers,
 * so that reading + H2D, the kernels and D2H + writing of consecutive slabs overlap and neither host nor
 * device memory has to hold the file. .bin -> .binpack runs the heavy part of every slab (the sharded
 * compressor's begin) concurrently and replays the chunk-flush rule slab after slab in file order in
 * host memory: it writes exactly the file one reference run writes, whatever the slab size and the
 * number of devices; .binpack -> .bin decodes groups of whole chunks. `append` != 0 appends to the
 * output file (the tool's -a). *positions (may be NULL) receives the number of positions converted.
 * The reference errors return their status after leaving in the file what the reference tool would
 * have left.
 * The *_multi forms are the same pipelines between host buffers (the whole-buffer host drivers above
 * use one device): out == NULL is the capacity query (for .binpack -> .bin a count pass). */
int nnp_bin_to_binpack_file(const char* in_path, const char* out_path, int append, size_t slab_bytes, uint64_t* positions);
int nnp_binpack_to_bin_file(const char* in_path, const char* out_path, int append, size_t slab_bytes, uint64_t* positions);
int nnp_bin_to_binpack_multi(const void* bin, size_t bin_bytes, void* out, size_t out_cap, size_t* out_bytes);
int nnp_binpack_to_bin_multi(const void* binpack, size_t binpack_bytes, void* out, size_t out_cap, size_t* out_bytes);

/* ---- decode fused into its consumer: HalfKP feature rows ------------ */

/* What the NNUE trainer's data loader (README.md:3 of the reference points at it) computes from every
 * decoded position, produced on the device straight from the chain walk, without the .bin record in
 * between. The reference itself has no feature code: the index is the published HalfKP definition
 * (nodchip learner / nnue-pytorch halfkp_idx), restated in csrc/halfkp.cuh and oracle/oracle.c:
 *   index = 1 + orient(P, sq) + 64 * (2 * type + (colour != P)) + 641 * orient(P, king square of P),
 *   orient(white, sq) = sq, orient(black, sq) = sq ^ 63, type pawn..queen = 0..4, kings excluded.
 * Row r belongs to record r of the .bin file nnp_binpack_to_bin writes for the same input (for the
 * .bin entry point: to record r of the input). d_white / d_black: [positions][NNP_HALFKP_ROW] int32,
 * the non-king pieces, the same piece at the same slot of both rows, padded with -1 at the end. A
 * chain head's row is ordered by (2 * type + colour, square), i.e. the white row ascends; a .bin
 * record's row follows its Huffman stream (rank 8 first, files a to h); along a chain the row is updated in place (a piece keeps its slot while it stands,
 * the row's last entry takes over the slot of a captured piece), so the order is deterministic but
 * follows the chain's history -- a sparse feature transformer sums over the row and does not care; d_meta: [positions] nnp_halfkp_meta. All three 16-byte aligned. d_white == NULL: count
 * only (*positions is set). NNP_ERR_CAPACITY when cap_positions is too small (*positions = needed).
 * Malformed input: same status codes as nnp_binpack_to_bin_dev / nnp_bin_to_plain_dev; *positions
 * then is the number of leading rows that are valid. */
#define NNP_HALFKP_ROW 32
#define NNP_HALFKP_FEATURES 41024 /* 64 * 641 */
typedef struct nnp_halfkp_meta {
    int16_t score;    /* as stored: from the side to move's point of view */
    uint16_t ply;
    int8_t result;    /* as stored in the .bin record (int8 narrowing, compress_file.cpp:570-585) */
    uint8_t stm;      /* 0 white, 1 black */
    uint8_t n_active; /* entries of each row that are not padding */
    uint8_t reserved; /* 0 */
} nnp_halfkp_meta;
int nnp_binpack_to_halfkp_dev(const void* d_binpack, size_t binpack_bytes, int32_t* d_white, int32_t* d_black,
                              nnp_halfkp_meta* d_meta, size_t cap_positions, size_t* positions);
int nnp_bin_to_halfkp_dev(const void* d_bin, size_t bin_bytes, int32_t* d_white, int32_t* d_black, nnp_halfkp_meta* d_meta,
                          size_t cap_positions, size_t* positions);

/* ---- helpers around the path --------------------------------------------------------- */

/* Number of positions a binpack holds = sum over chains of (1 + numPlies)
 * (CompressedTrainingDataEntryReader, compress_file.cpp:1154-1190); reads chunk headers and
 * chain headers only after the chain starts are known, so it costs one decode pass. */
int nnp_binpack_count_dev(const void* d_binpack, size_t binpack_bytes, uint64_t* n_positions);

/* Synthetic input: random legal-move games from the start position written
 * as 40-byte .bin records, at most `max_plies` positions per game, generated on the device.
 * Deterministic in (n_positions, max_plies, seed). Not a reference driver: bench/test input. */
int nnp_generate_bin_dev(void* d_out, size_t n_positions, uint32_t max_plies, uint64_t seed);

/* Diagnostics of the .binpack decoder: out14[0] = decodes finished by the
 * optimistic single walk, [1] = decodes that fell back to the exhaustive walk, [2] = chain-start
 * candidates of the last decode, [3] = positions their headers claim, [4] = walk violations of the
 * last optimistic attempt (~0 when it was not attempted), [5] = candidates the last exhaustive walk
 * found not to be chain starts, [6..13] = the first of those as chunk << 32 | offset. */
int nnp_decode_stats(uint64_t* out14);

/* Test hooks (the environment variables NNP_DEBUG_EXHAUSTIVE / NNP_DEBUG_REJECT_MOD set the same
 * switches at nnp_init; NNP_DEBUG_SINGLES=0 sets "k1_direct" = "dec_direct" = 2): "exhaustive" != 0 skips the optimistic decode strategy; "reject_mod" = m
 * drops the chain-start candidates whose hashed offset is 0 mod m (0 = off), which forces the
 * fallback strategies; "k1_per_record" / "k1_walk" / "k1_runs" / "k1_heads" != 0 pin the compressor's first
 * kernel to its record-parallel, chain-owning, run-based or chain-head-transcoding form (by default a sample of
 * the chain-head density picks); "k1_direct" = 1 always tries / = 2 never tries the one-kernel route for .bin files of
 * single positions, "dec_direct" = 2 never takes the candidate-free routes for .binpack chunks of single positions;
 * "walk_seg_bytes" = v sets the segment size of the parallel chunk-header walk (0 = default 8 MiB; files
 * shorter than eight segments are walked sequentially).
 * Results never depend on these switches. */
int nnp_debug_config(const char* key, uint64_t value);

/* Timing of the last *_dev call, measured with CUDA events on the library's stream:
 * total milliseconds, and the slice spent in the dominant kernel of that direction. */
int nnp_last_timing(float* total_ms, float* dominant_kernel_ms);
/* name of the kernel that slice belongs to (which form of K1 / which decode strategy ran) */
const char* nnp_last_dominant_kernel(void);
/* positions converted by the last driver call of the calling thread's device (what the reference's
 * "Processed N bytes and M positions." lines count, compress_file.cpp:1369-1372, :1395-1410) */
uint64_t nnp_last_positions(void);

#ifdef __cplusplus
}
#endif
#endif /* NNUEPACK_H */
