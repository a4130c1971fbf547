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
 *   - one call at a time per process (the library owns one context per process, bound
 *     to one GPU: the multi-GPU model is one process per GPU).
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
    NNP_ERR_CUDA = -11            /* a CUDA runtime call failed; see nnp_last_cuda_error() */
} nnp_status;

/* ---- lifecycle ------------------------------------------------------------------ */

/* Binds the process to CUDA device `device` (0-based; for one-process-per-GPU launches pass
 * LOCAL_RANK), creates the library's streams and uploads the attack tables (the GPU
 * counterpart of the reference's static initialisation, src/chess/Bitboard.cpp:460-464). */
int nnp_init(int device);
void nnp_shutdown(void);
const char* nnp_strerror(int status);
/* Runs all further work on the caller's CUDA stream (a cudaStream_t of the bound device, e.g.
 * torch's current stream) so that the caller's events bracket it; NULL restores the library's own
 * stream. Calls still return only after their result size is known. */
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
 * switches at nnp_init): "exhaustive" != 0 skips the optimistic decode strategy; "reject_mod" = m
 * drops the chain-start candidates whose hashed offset is 0 mod m (0 = off), which forces the
 * fallback strategies. Results never depend on these switches. */
int nnp_debug_config(const char* key, uint64_t value);

/* Timing of the last *_dev call, measured with CUDA events on the library's stream:
 * total milliseconds, and the slice spent in the dominant kernel of that direction. */
int nnp_last_timing(float* total_ms, float* dominant_kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* NNUEPACK_H */
