/*
 * oracle.h -- TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.
 *
 * Plain-C, single-threaded restatement of the conversion hot path of
 * Sopel97/nnue_data_compress (src/compress_file.cpp and the parts of src/chess it
 * reaches). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libnnuepack.so) never
 * links, loads or calls it.
 *
 * Parity status: PINNED. The reference ships no tests or golden vectors (SURVEY.md
 * section 4), so the oracle is pinned against outputs of the reference itself built
 * in this container (oracle/Makefile -> oracle/_ref/nnue_data_compression): the
 * committed fixtures under tests/golden/ were produced by that binary
 * (tests/golden/make_golden.py) and tests/test_oracle_vs_reference.py re-runs the
 * comparison live whenever oracle/_ref is present.
 *
 * Parity domain: byte-exact for every input on which the reference's Board keeps its
 * mailbox and its bitboards consistent, i.e. every stored move starts on a square
 * holding a piece (legality is NOT required). Malformed inputs that make the
 * reference run undefined behaviour (lookups with no king on the board, Huffman
 * codes that never terminate, movelists above its 10 KiB buffer slack) are reported
 * as errors here instead.
 */
#ifndef NNP_ORACLE_H
#define NNP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_OK = 0,
    ORC_ERR_BAD_MAGIC = -1,      /* compress_file.cpp:504-507 "Invalid binpack file or chunk." */
    ORC_ERR_CHUNK_TOO_LARGE = -2,/* compress_file.cpp:515-518 */
    ORC_ERR_BAD_SFEN = -3,       /* compress_file.cpp:407-408, :441-442 "Improperly encoded bin sfen" */
    ORC_ERR_TRUNCATED = -4,      /* chunk shorter than its header says / movetext runs off the chunk */
    ORC_ERR_NOMEM = -5,
    ORC_ERR_BAD_MODE = -6,
    ORC_ERR_BAD_TEXT = -7        /* .plain input the reference would crash on (stoi on non-number) */
};

/* conversion directions == the reference's six file drivers (compress_file.cpp:1246-1533) */
enum {
    ORC_BIN_TO_BINPACK = 0,   /* compressBin        :1338 */
    ORC_BINPACK_TO_BIN = 1,   /* decompressBin      :1376 */
    ORC_PLAIN_TO_BINPACK = 2, /* compressPlain      :1246 */
    ORC_BINPACK_TO_PLAIN = 3, /* decompressPlain    :1299 */
    ORC_BIN_TO_PLAIN = 4,     /* convertBinToPlain  :1414 */
    ORC_PLAIN_TO_BIN = 5      /* convertPlainToBin  :1467 */
};

/* Converts a whole in-memory file. *out is malloc'ed (free with orc_free). Returns ORC_OK
 * or a negative error; on error *out holds whatever the reference would already have
 * flushed (the partial chunk, compress_file.cpp:1094-1106) and *out_len its size. */
int orc_convert(int mode, const uint8_t* in, size_t in_len, uint8_t** out, size_t* out_len);
void orc_free(void* p);
const char* orc_strerror(int code);

/* Fine-grained entry points for unit tests. */
/* 32-byte sfen -> FEN string of the decoded Position (incl. post-move ep nullification). */
int orc_sfen_to_fen(const uint8_t sfen[32], char* fen_out, size_t cap);
/* FEN -> 32-byte sfen (Position::fromFen then SfenPacker::pack). */
int orc_fen_to_sfen(const char* fen, uint8_t sfen_out[32]);
/* isContinuation(rec_a, rec_b) on two 40-byte records; returns 0/1 or a negative error. */
int orc_is_continuation(const uint8_t a[40], const uint8_t b[40]);
/* number of positions a binpack buffer holds (sum of 1 + numPlies over chains) */
int64_t orc_binpack_count(const uint8_t* in, size_t in_len);


/* HalfKP feature rows of n 40-byte .bin records (checker of nnp_bin_to_halfkp_dev and, composed with
 * ORC_BINPACK_TO_BIN, of nnp_binpack_to_halfkp_dev). NOT a restatement of reference code: the
 * reference has no feature extraction. It restates the published HalfKP index of the Stockfish NNUE
 * trainers (nodchip learner kpp_board_index / nnue-pytorch halfkp_idx):
 *   1 + orient(P, sq) + 64 * (2 * type + (colour != P)) + 641 * orient(P, king of P), orient(black) = sq ^ 63.
 * PARITY UNPINNED for the index formula itself (nothing under /root/reference computes it); the
 * positions it is applied to come from sfen_unpack, which is pinned. white / black: [n][32] int32,
 * ordered by (2 * type + colour, square) = ascending white index, -1 padded; meta: [n][8] bytes = int16 score, uint16 ply, int8 result,
 * uint8 stm, uint8 n_active, 0. Returns ORC_OK or ORC_ERR_BAD_SFEN (then *bad_index = the record). */
int orc_bin_to_halfkp(const uint8_t* bin, size_t n, int32_t* white, int32_t* black, uint8_t* meta, size_t* bad_index);

#ifdef __cplusplus
}
#endif
#endif
