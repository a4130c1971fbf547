// chess.cuh -- device-side position model for the conversion kernels (sm_100a).
//
// A position is five 64-bit planes held in registers (white/black occupancy + three
// bit planes of the piece type) plus a few scalars: everything the path reads fits in
// 12 registers, comparisons are five XORs, and there is no per-square mailbox to index
// dynamically. Square numbering follows the reference (a1 = 0 ... h8 = 63,
// src/chess/Chess.h:593-755); piece codes are type<<1|colour with type order
// Pawn,Knight,Bishop,Rook,Queen,King and 12 = none (src/chess/Chess.h:74-84, :142-205).
//
// Citations are relative to /root/reference.
#pragma once
#include <cstdint>
#ifdef NNP_HOST_SIM
#include "host_sim.h"  // tests/host_sim: CPU stand-ins for the intrinsics, test harness only
#else
#include <cuda_runtime.h>
#include "bulk.cuh"
#endif

namespace nnp {

typedef unsigned long long u64;
typedef unsigned int u32;

enum : int { PT_PAWN = 0, PT_KNIGHT, PT_BISHOP, PT_ROOK, PT_QUEEN, PT_KING, PT_NONE };
enum : int { WHITE = 0, BLACK = 1 };
enum : int { NO_PIECE = 12, SQ_NONE = 64 };
enum : int { MT_NORMAL = 0, MT_PROMOTION = 1, MT_CASTLE = 2, MT_ENPASSANT = 3 };  // Chess.h:905-911
enum : int { CR_WK = 1, CR_WQ = 2, CR_BK = 4, CR_BQ = 8, CR_ALL = 15 };           // Chess.h:1195-1205

constexpr u64 FILE_A = 0x0101010101010101ull;
constexpr u64 FILE_H = 0x8080808080808080ull;
constexpr u64 RANK_1 = 0x00000000000000FFull;
constexpr u64 DIAG_A1H8 = 0x8040201008040201ull;
constexpr u64 DIAG_H1A8 = 0x0102040810204080ull;

__device__ __forceinline__ u64 bit64(int sq) { return 1ull << sq; }
__device__ __forceinline__ int popc64(u64 b) { return __popcll(b); }
__device__ __forceinline__ int lsb64(u64 b) { return __ffsll((long long)b) - 1; }
__device__ __forceinline__ u64 brev64(u64 b) { return __brevll(b); }
__device__ __forceinline__ u64 bswap64(u64 b)
{
    u32 lo = (u32)b, hi = (u32)(b >> 32);
    return ((u64)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}
// bb::before (Bitboard.h:730-733)
__device__ __forceinline__ u64 before64(int sq) { return sq >= 64 ? ~0ull : (1ull << sq) - 1ull; }

// usedBitsSafe (compress_file.cpp:600-604): ceil(log2(n)), 0 for n <= 1
__device__ __forceinline__ int used_bits(u32 n) { return n <= 1 ? 0 : 32 - __clz(n - 1); }

// index of the n-th set bit (nthSetBitIndex, util/ArithmeticUtility.h:186-209), branch-light
__device__ __forceinline__ int nth_set_bit(u64 v, u32 n)
{
    u32 w = (u32)v;
    int base = 0;
    u32 c = __popc(w);
    if (n >= c) { w = (u32)(v >> 32); n -= c; base = 32; }
    c = __popc(w & 0xFFFFu);
    if (n >= c) { w >>= 16; n -= c; base += 16; }
    c = __popc(w & 0xFFu);
    if (n >= c) { w >>= 8; n -= c; base += 8; }
    c = __popc(w & 0xFu);
    if (n >= c) { w >>= 4; n -= c; base += 4; }
    c = __popc(w & 0x3u);
    if (n >= c) { w >>= 2; n -= c; base += 2; }
    if (n >= (w & 1u)) base += 1;
    return base & 63;
}

// zigzag (signedToUnsigned / unsignedToSigned, compress_file.cpp:524-546)
__device__ __forceinline__ u32 zz_enc(int v16)
{
    u32 r = (u32)v16 & 0xFFFFu;
    if (r & 0x8000u) r ^= 0x7FFFu;
    return ((r << 1) | (r >> 15)) & 0xFFFFu;
}
__device__ __forceinline__ int zz_dec(u32 r)
{
    r &= 0xFFFFu;
    r = ((r << 15) | (r >> 1)) & 0xFFFFu;
    if (r & 0x8000u) r ^= 0x7FFFu;
    return (int)(short)r;
}

// ---------------------------------------------------------------- attack sets
// Table-free, exact equivalents of bb::pseudoAttacks / bb::attacks (Bitboard.h:837-880);
// the reference's 800 KiB fancy-magic tables (Bitboard.cpp:401-464) are replaced by the
// o^(o-2r) line trick with a full 64-bit reversal, which needs no memory at all.

__device__ __forceinline__ u64 knight_attacks(int sq)
{
    u64 b = bit64(sq);
    u64 l1 = (b >> 1) & ~FILE_H, l2 = (b >> 2) & 0x3f3f3f3f3f3f3f3full;
    u64 r1 = (b << 1) & ~FILE_A, r2 = (b << 2) & 0xfcfcfcfcfcfcfcfcull;
    u64 h1 = l1 | r1, h2 = l2 | r2;
    return (h1 << 16) | (h1 >> 16) | (h2 << 8) | (h2 >> 8);
}
__device__ __forceinline__ u64 king_attacks(int sq)
{
    u64 b = bit64(sq);
    u64 h = ((b >> 1) & ~FILE_H) | ((b << 1) & ~FILE_A);
    u64 m = h | b;
    return h | (m << 8) | (m >> 8);
}
// bb::pawnAttacks (Bitboard.cpp:500-510)
__device__ __forceinline__ u64 pawn_attacks(u64 pawns, int color)
{
    u64 e = pawns & ~FILE_H, w = pawns & ~FILE_A;
    return color == WHITE ? ((e << 9) | (w << 7)) : ((e >> 7) | (w >> 9));
}
__device__ __forceinline__ u64 line_attacks(u64 occ, u64 mask, int sq)
{
    u64 o = occ & mask;
    u64 fwd = o - 2 * bit64(sq);
    u64 rev = brev64(brev64(o) - 2 * bit64(63 - sq));
    return (fwd ^ rev) & mask;
}
__device__ __forceinline__ u64 diag_mask(int sq)
{
    int d = (sq & 7) - (sq >> 3);
    return d >= 0 ? (DIAG_A1H8 >> (8 * d)) : (DIAG_A1H8 << (8 * -d));
}
__device__ __forceinline__ u64 anti_mask(int sq)
{
    int d = (sq & 7) + (sq >> 3) - 7;
    return d >= 0 ? (DIAG_H1A8 << (8 * d)) : (DIAG_H1A8 >> (8 * -d));
}
__device__ __forceinline__ u64 bishop_attacks(int sq, u64 occ)
{
    return line_attacks(occ, diag_mask(sq), sq) | line_attacks(occ, anti_mask(sq), sq);
}
__device__ __forceinline__ u64 rook_attacks(int sq, u64 occ)
{
    return line_attacks(occ, FILE_A << (sq & 7), sq) | line_attacks(occ, RANK_1 << (sq & 56), sq);
}
// bb::attacks(pt, sq, occ) (Bitboard.h:864-880); Pawn / None give the empty set
__device__ __forceinline__ u64 piece_attacks(int pt, int sq, u64 occ)
{
    u64 a = 0;
    if (pt == PT_KNIGHT) a = knight_attacks(sq);
    else if (pt == PT_KING) a = king_attacks(sq);
    else {
        if (pt == PT_BISHOP || pt == PT_QUEEN) a = bishop_attacks(sq, occ);
        if (pt == PT_ROOK || pt == PT_QUEEN) a |= rook_attacks(sq, occ);
    }
    return a;
}

// Lookup tables in shared memory (filled by the block from the functions above and below). The chain
// kernels are bound by the integer ALU pipe, so everything that is a pure function of a square or of a
// bit position is looked up (LSU pipe) instead of computed: knight / king step attacks, pawn capture
// sets, the two diagonal masks, 1 << sq and (1 << sq) - 1 and the castling rights a move from / to a
// square preserves. (A 7 KB table of the splice's "bits below stream position p" masks was measured
// too: it costs the L1 cache more than it saves)
constexpr int STEP_TABLE_ROWS = 65;
struct StepTables {
    u64 knight[64], king[64];
    u64 pawn[2][64];     // pawn_attacks(bit64(sq), colour)
    u64 diag[64], anti[64];
    u64 bit[64], before[65];
    unsigned char keep_cr[64];  // preserved_cr(sq)
    alignas(16) u32 pad[2];
};
__device__ __forceinline__ u64 diag_mask(int sq);
__device__ __forceinline__ u64 anti_mask(int sq);
__device__ __forceinline__ int preserved_cr(int sq);
__device__ __forceinline__ void step_tables_entry(StepTables& T, int i)
{
    if (i < 64) {
        T.knight[i] = knight_attacks(i);
        T.king[i] = king_attacks(i);
        T.pawn[0][i] = pawn_attacks(bit64(i), 0);
        T.pawn[1][i] = pawn_attacks(bit64(i), 1);
        T.diag[i] = diag_mask(i);
        T.anti[i] = anti_mask(i);
        T.bit[i] = bit64(i);
        T.keep_cr[i] = (unsigned char)preserved_cr(i);
    }
    if (i < 65) T.before[i] = before64(i);
}
#ifdef __CUDACC__
// One copy per translation unit and device, computed once (k_step_tables_init, at nnp_init) and then
// copied into every block's shared memory with 16-byte loads: computing the tables per block would
// cost more than one percent of a block's work.
static __device__ StepTables g_step_tables;
static __global__ void k_step_tables_init()
{
    for (int i = threadIdx.x; i < STEP_TABLE_ROWS; i += blockDim.x) step_tables_entry(g_step_tables, i);
}
static_assert(sizeof(StepTables) % 16 == 0, "moved as one bulk copy");
// whole block; call before any early return. One thread asks the copy engine for the 4 KB (cp.async.bulk,
// bulk.cuh); everybody waits on the mbarrier the copy completes on.
__device__ __forceinline__ void step_tables_fill(StepTables& T)
{
    __shared__ alignas(8) unsigned long long bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (unsigned)sizeof(StepTables));
        bulk_load(&T, &g_step_tables, (unsigned)sizeof(StepTables), &bar);
    }
    mbar_wait(&bar, 0);
}
#endif
__device__ __forceinline__ u64 bishop_attacks(int sq, u64 occ, const StepTables* T)
{
#ifdef NNP_NO_SMALL_LUT
    T = nullptr;
#endif
    if (!T) return bishop_attacks(sq, occ);
    return line_attacks(occ, T->diag[sq], sq) | line_attacks(occ, T->anti[sq], sq);
}
__device__ __forceinline__ u64 piece_attacks(int pt, int sq, u64 occ, const StepTables* T)
{
    if (!T) return piece_attacks(pt, sq, occ);
    u64 a = 0;
    if (pt == PT_KNIGHT || pt == PT_KING) a = pt == PT_KNIGHT ? T->knight[sq] : T->king[sq];
    else {
        if (pt == PT_BISHOP || pt == PT_QUEEN) a = bishop_attacks(sq, occ, T);
        if (pt == PT_ROOK || pt == PT_QUEEN) a |= rook_attacks(sq, occ);
    }
    return a;
}

// ---------------------------------------------------------------- position

struct Pos {
    u64 occ[2];  // by colour
    u64 t0, t1, t2;  // piece-type bit planes, zero on empty squares
    int stm;     // 0 white, 1 black
    int ep;      // SQ_NONE if none
    int cr;      // castling rights mask
    int rule50;  // std::uint8_t in the reference (Position.h:1014)
    int ply;     // std::uint16_t m_ply (Position.h:1015)
};

struct Move {
    int from, to, type, promo;  // promo = piece code or NO_PIECE
};
// (the piece standing on `from`, which decode_ply / encode_ply look up anyway, travels next to the move
// as a separate `moved` value: -1 = not looked up)

__device__ __forceinline__ void pos_clear(Pos& p)  // Position() Position.h:828-836
{
    p.occ[0] = p.occ[1] = p.t0 = p.t1 = p.t2 = 0;
    p.stm = WHITE;
    p.ep = SQ_NONE;
    p.cr = CR_ALL;
    p.rule50 = 0;
    p.ply = 0;
}
__device__ __forceinline__ u64 pos_all(const Pos& p) { return p.occ[0] | p.occ[1]; }
// colour-indexed occupancy as a select (keeps the planes in registers)
__device__ __forceinline__ u64 pos_occ(const Pos& p, int c) { return c ? p.occ[1] : p.occ[0]; }
__device__ __forceinline__ int pos_piece_at(const Pos& p, int sq)
{
    u64 all = pos_all(p);
    if (!((all >> sq) & 1)) return NO_PIECE;
    int t = (int)((p.t0 >> sq) & 1) | ((int)((p.t1 >> sq) & 1) << 1) | ((int)((p.t2 >> sq) & 1) << 2);
    return (t << 1) | (int)((p.occ[1] >> sq) & 1);
}
__device__ __forceinline__ void pos_remove(Pos& p, int sq)
{
    u64 m = ~bit64(sq);
    p.occ[0] &= m; p.occ[1] &= m; p.t0 &= m; p.t1 &= m; p.t2 &= m;
}
__device__ __forceinline__ void pos_put(Pos& p, int sq, int piece)  // Board::place Position.h:260-275
{
    pos_remove(p, sq);
    if (piece == NO_PIECE) return;
    u64 b = bit64(sq);
    int t = piece >> 1;
    if (piece & 1) p.occ[1] |= b; else p.occ[0] |= b;
    if (t & 1) p.t0 |= b;
    if (t & 2) p.t1 |= b;
    if (t & 4) p.t2 |= b;
}
__device__ __forceinline__ u64 pos_type_bb(const Pos& p, int t)
{
    u64 a = (t & 1) ? p.t0 : ~p.t0;
    u64 b = (t & 2) ? p.t1 : ~p.t1;
    u64 c = (t & 4) ? p.t2 : ~p.t2;
    return a & b & c & pos_all(p);
}
__device__ __forceinline__ bool pos_equal(const Pos& a, const Pos& b)  // Position.h:977-984
{
    u64 d = (a.occ[0] ^ b.occ[0]) | (a.occ[1] ^ b.occ[1]) | (a.t0 ^ b.t0) | (a.t1 ^ b.t1) | (a.t2 ^ b.t2);
    return d == 0 && a.stm == b.stm && a.ep == b.ep && a.cr == b.cr;
}

// bb::isAttackedBySlider (Bitboard.cpp:536-554)
__device__ __forceinline__ bool attacked_by_slider(int sq, u64 bq, u64 rq, u64 occ)
{
    if (bishop_attacks(sq, occ) & bq) return true;
    return (rook_attacks(sq, occ) & rq) != 0;
}

// Position::isEpPossible + isEpPossibleColdPath (Position.cpp:824-883), evaluated on the
// board currently held in `p`: post-move when called from setEpSquare / FEN parsing,
// PRE-move when called from pos_do_move (Position.cpp:647-657) -- both variants are needed
// for byte parity.
static __device__ __noinline__ bool ep_possible_cold(const Pos& p, int ep, u64 attackers, int side)
{
    u64 all = pos_all(p);
    u64 kings = pos_type_bb(p, PT_KING) & pos_occ(p, side);
    if (!kings) return true;  // outside the parity domain (reference: undefined)
    int ksq = lsb64(kings);
    u64 theirs = pos_occ(p, side ^ 1);
    u64 bq = (pos_type_bb(p, PT_BISHOP) | pos_type_bb(p, PT_QUEEN)) & theirs;
    u64 rq = (pos_type_bb(p, PT_ROOK) | pos_type_bb(p, PT_QUEEN)) & theirs;
    u64 queen_lines = bishop_attacks(ksq, 0) | rook_attacks(ksq, 0);
    if (((bq | rq) & queen_lines) == 0) return true;
    while (attackers) {
        int sq = lsb64(attackers);
        attackers &= attackers - 1;
        int captured = (ep & 7) | (sq & 56);
        u64 occ = ((all ^ bit64(sq)) | bit64(ep)) ^ bit64(captured);
        if (!attacked_by_slider(ksq, bq, rq, occ)) return true;
    }
    return false;
}
__device__ __forceinline__ bool ep_possible(const Pos& p, int ep, int side)
{
    u64 attackers = pawn_attacks(bit64(ep), side ^ 1) & pos_type_bb(p, PT_PAWN) & pos_occ(p, side);
    if (!attackers) return false;
    const Pos copy = p;  // the out-of-line test gets its own copy: the caller's position stays in registers
    return ep_possible_cold(copy, ep, attackers, side);
}

// Board::isSquareAttacked as used by Position::trySet (Position.cpp:505) and the game generator
__device__ __forceinline__ bool square_attacked(const Pos& p, int sq, int by, u64 occ)
{
    const u64 them = pos_occ(p, by);
    if (pawn_attacks(bit64(sq), by ^ 1) & pos_type_bb(p, PT_PAWN) & them) return true;
    if (knight_attacks(sq) & pos_type_bb(p, PT_KNIGHT) & them) return true;
    if (king_attacks(sq) & pos_type_bb(p, PT_KING) & them) return true;
    const u64 bq = (pos_type_bb(p, PT_BISHOP) | pos_type_bb(p, PT_QUEEN)) & them;
    const u64 rq = (pos_type_bb(p, PT_ROOK) | pos_type_bb(p, PT_QUEEN)) & them;
    if (bq && (bishop_attacks(sq, occ) & bq)) return true;
    if (rq && (rook_attacks(sq, occ) & rq)) return true;
    return false;
}

// detail::lookup::preservedCastlingRights (Position.cpp:605-624)
__device__ __forceinline__ int preserved_cr(int sq)
{
    int m = CR_ALL;
    if (sq == 4) m = CR_ALL & ~(CR_WK | CR_WQ);
    if (sq == 60) m = CR_ALL & ~(CR_BK | CR_BQ);
    if (sq == 7) m = CR_ALL & ~CR_WK;
    if (sq == 0) m = CR_ALL & ~CR_WQ;
    if (sq == 63) m = CR_ALL & ~CR_BK;
    if (sq == 56) m = CR_ALL & ~CR_BQ;
    return m;
}

// Board::doMove for MoveType::Normal / Promotion (Position.h:300-350): `placed` lands on `to`, whatever
// stood there disappears, `from` is vacated -- place(piece, to) then place(none, from), so a move
// with from == to leaves the square empty. One masked update per plane instead of three generic
// square operations.
__device__ __forceinline__ void pos_move_piece(Pos& p, int from, int to, int placed)
{
    const u64 bf = bit64(from), bt = bit64(to);
    const u64 keep = ~(bf | bt);
    const u64 set = (from == to || placed == NO_PIECE) ? 0ull : bt;
    const int t = placed >> 1;
    p.occ[0] = (p.occ[0] & keep) | ((placed & 1) ? 0ull : set);
    p.occ[1] = (p.occ[1] & keep) | ((placed & 1) ? set : 0ull);
    p.t0 = (p.t0 & keep) | ((t & 1) ? set : 0ull);
    p.t1 = (p.t1 & keep) | ((t & 2) ? set : 0ull);
    p.t2 = (p.t2 & keep) | ((t & 4) ? set : 0ull);
}

// Board::doMove + doMoveColdPath (Position.h:300-439), mailbox semantics on the planes
// `moved` = pos_piece_at(p, m.from), which the callers have at hand
__device__ __forceinline__ void board_do_move(Pos& p, const Move& m, int moved)
{
    if (m.type == MT_NORMAL) {
        pos_move_piece(p, m.from, m.to, moved);
    } else if (m.type == MT_PROMOTION) {
        pos_move_piece(p, m.from, m.to, m.promo);
    } else if (m.type == MT_ENPASSANT) {
        int pc = moved;
        pos_put(p, m.to, pc);
        pos_remove(p, m.from);
        pos_remove(p, (m.to & 7) | (m.from & 56));
    } else {
        int rook = pos_piece_at(p, m.to), king = moved;
        int base = (king & 1) ? 56 : 0;  // king.color(); Piece::none() counts as white
        bool is_short = (m.to & 7) == 7;  // CastlingTraits::moveCastlingType CastlingTraits.h:37-40
        pos_remove(p, m.to);
        pos_remove(p, m.from);
        pos_put(p, base + (is_short ? 5 : 3), rook);  // rookDestination f/d CastlingTraits.h:11
        pos_put(p, base + (is_short ? 6 : 2), king);  // kingDestination g/c CastlingTraits.h:12
    }
}

// Position::doMove (Position.cpp:626-662)
// `moved` (optional) = pos_piece_at(p, m.from) when the caller has already looked it up
__device__ __forceinline__ void pos_do_move(Pos& p, const Move& m, int moved = -1, const StepTables* T = nullptr)
{
    if (moved < 0) moved = pos_piece_at(p, m.from);
    int moved_type = moved >> 1;
    p.ply = (p.ply + 1) & 0xFFFF;
    p.rule50 = (p.rule50 + 1) & 0xFF;
    if (m.type != MT_CASTLE && (moved_type == PT_PAWN || ((pos_all(p) >> m.to) & 1))) p.rule50 = 0;
#ifdef NNP_NO_SMALL_LUT
    T = nullptr;
#endif
    if (T && (m.from | m.to) < 64) p.cr &= (int)T->keep_cr[m.from] & (int)T->keep_cr[m.to];
    else p.cr &= preserved_cr(m.from) & preserved_cr(m.to);
    p.ep = SQ_NONE;
    if (moved_type == PT_PAWN && ((m.to ^ m.from) == 16)) {
        int cand = (m.to + m.from) >> 1;
        if (ep_possible(p, cand, p.stm ^ 1)) p.ep = cand;  // on the PRE-move board
    }
    board_do_move(p, m, moved);
    p.stm ^= 1;
}

// ---------------------------------------------------------------- .bin record fields

// StockfishMove::toMove (compress_file.cpp:63-83)
__device__ __forceinline__ Move sfmove_to_move(u32 raw)
{
    Move m;
    m.to = raw & 63;
    m.from = (raw >> 6) & 63;
    int flag = (raw >> 14) & 3;
    m.type = flag == 1 ? MT_PROMOTION : flag == 2 ? MT_ENPASSANT : flag == 3 ? MT_CASTLE : MT_NORMAL;
    m.promo = NO_PIECE;
    if (m.type == MT_PROMOTION) m.promo = ((PT_KNIGHT + ((raw >> 12) & 3)) << 1) | ((m.to >> 3) == 7 ? WHITE : BLACK);
    return m;
}
// StockfishMove::fromMove (compress_file.cpp:35-61), same shift/or sequence on 16 bits
__device__ __forceinline__ u32 move_to_sfmove(const Move& m)
{
    u32 flag = m.type == MT_PROMOTION ? 1 : m.type == MT_ENPASSANT ? 2 : m.type == MT_CASTLE ? 3 : 0;
    u32 promo = m.type == MT_PROMOTION ? (u32)((m.promo >> 1) - PT_KNIGHT) : 0;
    u32 raw = flag;
    raw = (raw << 2) & 0xFFFF; raw |= promo;
    raw = (raw << 6) & 0xFFFF; raw |= (u32)m.from;
    raw = (raw << 6) & 0xFFFF; raw |= (u32)m.to;
    return raw & 0xFFFF;
}

// ---------------------------------------------------------------- sfen (Huffman) codec

// pos_from_packed_sfen (compress_file.cpp:364-446). `W(j)` returns 32-bit word j of the
// 40-byte record (j < 10; the stream is one little-endian 256-bit integer consumed from
// bit 0, BitStream :126-185; bits past 256 come from the score/move fields exactly as in
// the reference's struct, and any decode that gets there fails the cursor check anyway).
// Returns false for "Improperly encoded bin sfen" (:407-408, :441-442) and for the
// 3-bit type codes 5..7 on which the reference's table search never terminates (:336-352).
// One half of the board (32 stream squares). `wlo:whi` is a 64-bit window of the stream with
// `avail` valid bits, refilled one 32-bit word at a time; runs of empty squares are consumed with one
// count-trailing-zeros, every piece sets one bit in each of the planes it belongs to. The planes of
// this half are 32-bit registers: stream square s (rank 8 first) is square s ^ 56, i.e. bit
// (s ^ 24) & 31 of the other 32-bit half.
// on_piece(square, token) sees every piece token in stream order (the HalfKP row of a .bin record is
// listed straight from it; the other callers pass nothing).
template <typename WordFn, typename PieceFn>
__device__ __forceinline__ void sfen_decode_half(WordFn W, u32& wlo, u32& whi, int& avail, int& nextw, int& idx,
                                                 const int idx_end, const int ka, const int kb, u32& err, u32& o0,
                                                 u32& o1, u32& q0, u32& q1, u32& q2, PieceFn on_piece)
{
    while (idx < idx_end) {
        if (avail < 32) {  // insert the next word above the valid bits
            const u32 nw = nextw < 10 ? W(nextw) : 0u;
            ++nextw;
            wlo |= nw << avail;
            whi |= __funnelshift_l(nw, 0u, avail);
            avail += 32;
        }
        int z = __clz(__brev(wlo));  // empty squares ahead (32 when the low word is all zero)
        z = min(z, idx_end - idx);
        idx += z;
        wlo = __funnelshift_rc(wlo, whi, z);
        whi = __funnelshift_rc(whi, 0u, z);
        avail -= z;
        if (idx >= idx_end) break;
        if (avail < 5) continue;  // the 5-bit token is not in the window yet
        const u32 tok = wlo;      // 1, type (3 bits, LSB first), colour
        err |= ((tok >> 1) & 7u) > (u32)PT_QUEEN ? 1u : 0u;
        int s = idx;              // token ordinal -> stream square: skip the king squares
        s += (s >= ka);
        s += (s >= kb);
        const u32 b = 1u << ((s ^ 24) & 31);
        if (tok & 16u) o1 |= b; else o0 |= b;
        if (tok & 2u) q0 |= b;
        if (tok & 4u) q1 |= b;
        if (tok & 8u) q2 |= b;
        on_piece(s ^ 56, tok);
        idx += 1;
        wlo = __funnelshift_rc(wlo, whi, 5);
        whi >>= 5;
        avail -= 5;
    }
}

template <typename WordFn, typename PieceFn>
__device__ __forceinline__ bool sfen_decode(WordFn W, Pos& p, PieceFn on_piece)
{
    const u32 w0 = W(0), w1 = W(1);
    p.stm = w0 & 1;
    const int wk = (w0 >> 1) & 63, bk = (w0 >> 7) & 63;
    // token ordinal -> stream square (rank 8 first, file a first) skips the king squares
    int ka = wk ^ 56, kb = bk ^ 56;
    if (ka > kb) { int t = ka; ka = kb; kb = t; }
    int ntok = 62;
    if (ka == kb) { kb = 64; ntok = 63; }
    // number of token squares among stream squares 0..31 (ranks 8..5)
    const int half_tok = 32 - (ka < 32) - (kb < 32);
    u32 wlo = __funnelshift_r(w0, w1, 13), whi = w1 >> 13;
    int avail = 51, nextw = 2, idx = 0;
    u32 err = 0;
    u32 o0h = 0, o1h = 0, q0h = 0, q1h = 0, q2h = 0;  // squares 32..63 (ranks 5..8)
    u32 o0l = 0, o1l = 0, q0l = 0, q1l = 0, q2l = 0;  // squares 0..31  (ranks 1..4)
    sfen_decode_half(W, wlo, whi, avail, nextw, idx, half_tok, ka, kb, err, o0h, o1h, q0h, q1h, q2h, on_piece);
    sfen_decode_half(W, wlo, whi, avail, nextw, idx, ntok, ka, kb, err, o0l, o1l, q0l, q1l, q2l, on_piece);
    p.occ[0] = ((u64)o0h << 32) | o0l;
    p.occ[1] = ((u64)o1h << 32) | o1l;
    p.t0 = ((u64)q0h << 32) | q0l;
    p.t1 = ((u64)q1h << 32) | q1l;
    p.t2 = ((u64)q2h << 32) | q2l;
    p.ep = SQ_NONE;
    // kings: white first, black second (a black king on the same square replaces it, :376-377)
    pos_put(p, wk, (PT_KING << 1) | WHITE);
    pos_put(p, bk, (PT_KING << 1) | BLACK);
    // tail: castling(4) ep(1[+6]) rule50(6) fullmove(8) = at most 25 bits
    if (avail < 32) {
        const u32 nw = nextw < 10 ? W(nextw) : 0u;
        ++nextw;
        wlo |= nw << avail;
        whi |= __funnelshift_l(nw, 0u, avail);
        avail += 32;
    }
    u32 tail = wlo;
    int used = 0;
    p.cr = (int)(tail & 15u);  // WK,WQ,BK,BQ in stream order == CastlingRights bit order (:414-427)
    tail >>= 4; used += 4;
    if (tail & 1u) {
        const int ep = (int)((tail >> 1) & 63u);
        tail >>= 7; used += 7;
        p.ep = ep_possible(p, ep, p.stm) ? ep : SQ_NONE;  // setEpSquare Position.h:868-872
    } else {
        tail >>= 1; used += 1;
    }
    p.rule50 = (int)(tail & 63u);
    const int hm = (int)((tail >> 6) & 255u);
    used += 14;
    p.ply = (2 * hm - 1 + (p.stm == BLACK)) & 0xFFFF;  // setHalfMove Position.h:938-941
    const int cursor = 32 * nextw - avail + used;
    // the reference checks the cursor after every piece and at the end (:407-408, :441-442); it only
    // grows, so the final value decides. Type codes 5..7 never terminate its table search (:336-352).
    return cursor <= 256 && err == 0;
}

// x with a zero inserted at bit k (the bits at and above k move up by one)
__device__ __forceinline__ u64 insert_zero_bit(u64 x, int k)
{
    const u64 below = (1ull << k) - 1;
    return (x & below) | ((x & ~below) << 1);
}

// The same decode without a piece callback, as a flat loop: the 62 tokens of the non-king squares are read four
// per step with the same instructions whatever the position (no count-trailing-zeros runs, no loop whose length
// is the piece count of the busiest lane of the warp): a token's five bits -- piece, type (3), colour -- are
// parked a byte apart in one register, and one multiply per bit plane gathers the four tokens' bits into a
// nibble of that plane ((x & 0x01010101) * 0x01020408 >> 24). The planes are built over the token squares in
// stream order; the kings' squares are then opened up (a zero bit inserted at each) and the byte-swapped planes
// are the board's (stream square s is square s ^ 56). Half the instructions of the token-driven loop above.
template <typename WordFn>
__device__ __forceinline__ bool sfen_decode_flat(WordFn W, Pos& p)
{
    const u32 w0 = W(0), w1 = W(1);
    p.stm = w0 & 1;
    const int wk = (w0 >> 1) & 63, bk = (w0 >> 7) & 63;  // wk != bk (the caller's business)
    u64 win = (((u64)w1 << 32) | w0) >> 13;
    int avail = 51, nextw = 2, cursor = 13;
    u32 lo[5] = {0, 0, 0, 0, 0}, hi[5] = {0, 0, 0, 0, 0};  // planes: piece, type bit 0 / 1 / 2, colour; bit j = token j
#pragma unroll
    for (int g = 0; g < 16; ++g) {
        if (avail < 32) {  // four tokens need at most 20 bits
            const u32 nw = nextw < 10 ? W(nextw) : 0u;
            ++nextw;
            win |= (u64)nw << avail;
            avail += 32;
        }
        u32 c = (u32)win, acc = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (4 * g + k < 62) {
                const u32 piece = c & 1u;          // '0', or 1 + type (3 bits, LSB first) + colour
                acc += ((c & 31u) * piece) << (8 * k);
                c >>= 1u + 4u * piece;
            }
        }
        const int used = (g < 15 ? 4 : 2) + 4 * __popc(acc & 0x01010101u);
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            const u32 nib = (((acc >> b) & 0x01010101u) * 0x01020408u) >> 24;  // token k of the group -> bit k
            if (g < 8) lo[b] |= nib << (4 * g); else hi[b] |= nib << (4 * g - 32);
        }
        win >>= used;
        avail -= used;
        cursor += used;
    }
    const u32 err = (lo[3] & (lo[2] | lo[1])) | (hi[3] & (hi[2] | hi[1]));  // type codes 5..7
    // the kings' squares: stream square = square ^ 56, the one that comes first in the stream is opened first
    const int ka = wk ^ 56, kb = bk ^ 56;
    const int k1 = ka < kb ? ka : kb, k2 = ka < kb ? kb : ka;
    u64 pl[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) pl[b] = insert_zero_bit(insert_zero_bit(((u64)hi[b] << 32) | lo[b], k1), k2);
    const u64 kings = (1ull << ka) | (1ull << kb);
    pl[0] |= kings;
    pl[1] |= kings;  // PT_KING = 5: type bits 0 and 2
    pl[3] |= kings;
    pl[4] |= 1ull << kb;
    p.occ[1] = bswap64(pl[4]);
    p.occ[0] = bswap64(pl[0] ^ pl[4]);
    p.t0 = bswap64(pl[1]);
    p.t1 = bswap64(pl[2]);
    p.t2 = bswap64(pl[3]);
    p.ep = SQ_NONE;
    // tail: castling(4) ep(1[+6]) rule50(6) fullmove(8) = at most 25 bits
    if (avail < 32) {
        const u32 nw = nextw < 10 ? W(nextw) : 0u;
        ++nextw;
        win |= (u64)nw << avail;
        avail += 32;
    }
    u32 tail = (u32)win;
    p.cr = (int)(tail & 15u);
    tail >>= 4; cursor += 4;
    if (tail & 1u) {
        const int ep = (int)((tail >> 1) & 63u);
        tail >>= 7; cursor += 7;
        p.ep = ep_possible(p, ep, p.stm) ? ep : SQ_NONE;  // setEpSquare Position.h:868-872
    } else {
        tail >>= 1; cursor += 1;
    }
    p.rule50 = (int)(tail & 63u);
    const int hm = (int)((tail >> 6) & 255u);
    cursor += 14;
    p.ply = (2 * hm - 1 + (p.stm == BLACK)) & 0xFFFF;  // setHalfMove Position.h:938-941
    return cursor <= 256 && err == 0;
}

template <typename WordFn>
__device__ __forceinline__ bool sfen_decode(WordFn W, Pos& p)
{
    const u32 w0 = W(0);
    if (((w0 >> 1) & 63u) != ((w0 >> 7) & 63u)) return sfen_decode_flat(W, p);
    return sfen_decode(W, p, [](int, u32) {});  // both kings on one square: 63 tokens, the black king replaces the white one
}

// SfenPacker::pack (compress_file.cpp:266-312) into eight 32-bit words out[0..7].
// Tokens are emitted piece by piece in stream order; `out` may live in registers because
// the word index only grows (a current-word accumulator is flushed when it is passed).
__device__ __forceinline__ void sfen_encode(const Pos& p, u32* out /* [8], any address space */)
{
    u64 all = pos_all(p);
    u64 kings = pos_type_bb(p, PT_KING);
    u64 wkb = kings & p.occ[0], bkb = kings & p.occ[1];
    int wk = wkb ? lsb64(wkb) : 0, bk = bkb ? lsb64(bkb) : 0;  // kingSquare Position.h:742-745
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = 0;
    u32 cw = (u32)p.stm | ((u32)wk << 1) | ((u32)bk << 7);
    int wi = 0;  // index of the word held in cw
    auto put = [&](u32 v, int pos, int n) {
        // append n (<= 8) bits of v at absolute bit position pos (positions never decrease)
        int w = pos >> 5, sh = pos & 31;
        while (wi < w) { if (wi < 8) out[wi] = cw; cw = 0; ++wi; }
        cw |= v << sh;
        if (sh + n > 32) { if (wi < 8) out[wi] = cw; cw = v >> (32 - sh); ++wi; }
    };
    // stream order: s = sq ^ 56; every king square (any colour, any count) is skipped (:286-287)
    u64 s_all = bswap64(all), s_kings = bswap64(kings);
    u64 s_pieces = s_all & ~s_kings;
    int npieces = 0;
    while (s_pieces) {
        int s = lsb64(s_pieces);
        s_pieces &= s_pieces - 1;
        int kings_before = popc64(s_kings & before64(s));
        int pos = 13 + (s - kings_before) + 4 * npieces;
        int pc = pos_piece_at(p, s ^ 56);
        put(1u | ((u32)(pc >> 1) << 1) | ((u32)(pc & 1) << 4), pos, 5);
        ++npieces;
    }
    int cursor = 13 + (64 - popc64(kings)) + 4 * npieces;
    put((u32)p.cr & 15u, cursor, 4);
    cursor += 4;
    if (p.ep == SQ_NONE) {
        put(0u, cursor, 1);
        cursor += 1;
    } else {
        put(1u | ((u32)(p.ep & 63) << 1), cursor, 7);
        cursor += 7;
    }
    put((u32)p.rule50 & 63u, cursor, 6);
    cursor += 6;
    put((u32)(((p.ply + 1) >> 1) & 0xFF), cursor, 8);  // halfMove() Position.h:933-936, 8 bits
    while (wi < 8) { out[wi] = cw; cw = 0; ++wi; }
}

// ---------------------------------------------------------------- stem (32-byte chain head)

// Position::compress (Position.h:1374-1406) nibble for one occupied square
__device__ __forceinline__ int stem_nibble(const Pos& p, int sq, int pc)
{
    int t = pc >> 1, c = pc & 1;
    if (t == PT_PAWN) {
        if (p.ep != SQ_NONE && (sq & 7) == (p.ep & 7) &&
            (((sq >> 3) == 3 && p.stm == BLACK) || ((sq >> 3) == 4 && p.stm == WHITE)))
            return 12;
        return pc;
    }
    if (t == PT_ROOK) {
        if (c == WHITE && ((sq == 0 && (p.cr & CR_WQ)) || (sq == 7 && (p.cr & CR_WK)))) return 13;
        if (c == BLACK && ((sq == 56 && (p.cr & CR_BQ)) || (sq == 63 && (p.cr & CR_BK)))) return 14;
        return pc;
    }
    if (t == PT_KING) return c == WHITE ? 10 : (p.stm == WHITE ? 11 : 15);
    return pc;
}

// packEntry (compress_file.cpp:997-1020) into eight words that are the stem's bytes 0..31
// in memory order (little-endian words).
__device__ __forceinline__ void stem_pack(const Pos& p, const Move& mv, int score, int ply, int result, u32* out /* [8] */)
{
    u64 all = pos_all(p);
    u64 be = bswap64(all);  // big-endian occupancy, Position.h:1245-1257
    out[0] = (u32)be;
    out[1] = (u32)(be >> 32);
    u32 nib[4] = {0, 0, 0, 0};
    int k = 0;
    u64 b = all;
    while (b) {
        int sq = lsb64(b);
        b &= b - 1;
        u32 n = (u32)stem_nibble(p, sq, pos_piece_at(p, sq));
        u32 v = n << ((k & 7) * 4);
        int w = k >> 3;
        // a 33rd piece has no nibble (the reference writes past m_packedState[16], Position.h:1374-1406): dropped
        if (w == 0) nib[0] |= v; else if (w == 1) nib[1] |= v; else if (w == 2) nib[2] |= v; else if (w == 3) nib[3] |= v;
        ++k;
    }
    out[2] = nib[0]; out[3] = nib[1]; out[4] = nib[2]; out[5] = nib[3];
    u32 cm = 0;  // CompressedMove(Move) Chess.h:1071-1096
    if (mv.from != mv.to) {
        cm = ((u32)mv.type << 14) | ((u32)mv.from << 8) | ((u32)mv.to << 2);
        if (mv.type == MT_PROMOTION) cm |= (u32)((mv.promo >> 1) - PT_KNIGHT);
        cm &= 0xFFFF;
    }
    u32 sc = zz_enc(score);
    u32 pr = ((u32)ply | (zz_enc(result) << 14)) & 0xFFFF;
    u32 r50 = (u32)p.rule50 & 0xFF;
    // bytes 24..27: move hi, move lo, score hi, score lo; 28..31: pr hi, pr lo, 0, rule50
    out[6] = (cm >> 8) | ((cm & 0xFF) << 8) | ((sc >> 8) << 16) | ((sc & 0xFF) << 24);
    out[7] = (pr >> 8) | ((pr & 0xFF) << 8) | (0u << 16) | (r50 << 24);
}

// unpackEntry (compress_file.cpp:1022-1043) + CompressedPosition::decompress
// (Position.h:1408-1505) + CompressedMove::decompress (Chess.h:1142-1172).
// `B(i)` returns byte i of the stem.
template <typename ByteFn>
__device__ __forceinline__ void stem_unpack(ByteFn B, Pos& p, Move& mv, int& score, int& ply, int& result)
{
    pos_clear(p);
    p.cr = 0;
    u64 occ = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) occ = (occ << 8) | (u64)B(i);
    // the occupied squares are distinct, so the planes are built by OR-ing one bit per piece
    // (Board::place on an empty board) instead of generic square updates
    u64 o0 = 0, o1 = 0, t0 = 0, t1 = 0, t2 = 0;
    int k = 0;
    u64 b = occ;
    while (b) {
        const int sq = lsb64(b);
        const u64 bit = b & (0ull - b);
        b &= b - 1;
        const int nib = (B(8 + (k >> 1)) >> ((k & 1) * 4)) & 15;
        ++k;
        int piece = nib;  // 0..11: Piece ordinal (type << 1 | colour)
        if (nib == 12) {  // pawn that just made a double push: rank 4 = white, else black (Position.h:1440-1456)
            if ((sq >> 3) == 3) { piece = (PT_PAWN << 1) | WHITE; p.ep = (sq - 8) & 0xFF; }
            else { piece = (PT_PAWN << 1) | BLACK; p.ep = (sq + 8) & 0xFF; }
        } else if (nib == 13) {
            piece = (PT_ROOK << 1) | WHITE;
            p.cr |= (sq == 0) ? CR_WQ : CR_WK;
        } else if (nib == 14) {
            piece = (PT_ROOK << 1) | BLACK;
            p.cr |= (sq == 56) ? CR_BQ : CR_BK;
        } else if (nib == 15) {
            piece = (PT_KING << 1) | BLACK;
            p.stm = BLACK;
        }
        const int t = piece >> 1;
        if (piece & 1) o1 |= bit; else o0 |= bit;
        if (t & 1) t0 |= bit;
        if (t & 2) t1 |= bit;
        if (t & 4) t2 |= bit;
    }
    p.occ[0] = o0; p.occ[1] = o1; p.t0 = t0; p.t1 = t1; p.t2 = t2;
    u32 cm = ((u32)B(24) << 8) | (u32)B(25);
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
    score = zz_dec(((u32)B(26) << 8) | (u32)B(27));
    u32 pr = ((u32)B(28) << 8) | (u32)B(29);
    ply = (int)(pr & 0x3FFF);
    p.ply = ply;
    result = zz_dec(pr >> 14);
    p.rule50 = (int)B(31);  // setRule50Counter(uint8_t) of a 16-bit big-endian value
}

// ---------------------------------------------------------------- movetext codec

// destination set of a pawn (addMoveScore :890-919 / nextMoveScore :701-730)
__device__ __forceinline__ u64 pawn_destinations(const Pos& p, int from, u64 ours, u64 theirs, const StepTables* T = nullptr)
{
#ifdef NNP_NO_SMALL_LUT
    T = nullptr;
#endif
    u64 occ = ours | theirs;
    u64 targets = theirs;
    if (p.ep != SQ_NONE) targets |= T ? T->bit[p.ep & 63] : bit64(p.ep & 63);
    u64 dest = (T ? T->pawn[p.stm & 1][from] : pawn_attacks(bit64(from), p.stm)) & targets;
    int s1 = p.stm == WHITE ? from + 8 : from - 8;
    if (s1 >= 0 && s1 < 64 && !((occ >> s1) & 1)) {
        dest |= bit64(s1);
        int s2 = p.stm == WHITE ? s1 + 8 : s1 - 8;
        int start_rank = p.stm == WHITE ? 1 : 6;
        if ((from >> 3) == start_rank && s2 >= 0 && s2 < 64 && !((occ >> s2) & 1)) dest |= bit64(s2);
    }
    return dest;
}

// PackedMoveScoreList::addMoveScore (compress_file.cpp:877-989) for one continuation ply.
// Returns the ply's bit string left-aligned in 32 bits; nbits <= 31 (6 + 5 + 20).
// Field values are masked to their widths here. The reference does not mask (addBitsLE8 :840-862 takes
// a std::uint8_t and ORs it in): an id that does not fit its field -- possible only for a stored move
// that is not pseudo-legal in its position -- sets bits of the byte the field starts in, above the
// field. Where that lands depends on the byte alignment of the chain, which only the payload writer
// knows, so such a ply is reported through `bleed` (packed raw ids and widths, 0 = none) and the
// writer ORs the stray bits in (k_write_payload<true>).
__device__ __forceinline__ u32 encode_ply(const Pos& p, const Move& mv, int score, int last_score, int& nbits,
                                          const StepTables* T = nullptr, u32* bleed = nullptr, int* moved_out = nullptr)
{
    int stm = p.stm;
    u64 ours = pos_occ(p, stm), theirs = pos_occ(p, stm ^ 1);
    u64 occ = ours | theirs;
    u32 piece_id = (u32)popc64(ours & before64(mv.from));
    int pc = pos_piece_at(p, mv.from);
    if (moved_out) *moved_out = pc;
    int pt = pc >> 1;
    // every mover type yields a destination set; counting it and ranking the destination happen
    // once behind the branches (a warp usually holds movers of all types)
    u64 dest;
    u32 extra_moves = 0;
    bool promotes = false;
    if (pt == PT_PAWN) {
        dest = pawn_destinations(p, mv.from, ours, theirs, T);
        promotes = (mv.from >> 3) == (stm == WHITE ? 6 : 1);
    } else if (pt == PT_KING) {
        const int our_mask = stm == WHITE ? (CR_WK | CR_WQ) : (CR_BK | CR_BQ);
        dest = (T ? T->king[mv.from] : king_attacks(mv.from)) & ~ours;
        extra_moves = (u32)__popc((u32)(p.cr & our_mask));
    } else {
        dest = piece_attacks(pt, mv.from, occ, T) & ~ours;
    }
    const u32 dest_n = (u32)popc64(dest);
    u32 num_moves = dest_n + extra_moves;
    u32 move_id = (u32)popc64(dest & before64(mv.to));
    if (promotes) {
        move_id = move_id * 4 + (u32)((mv.promo >> 1) - PT_KNIGHT);
        num_moves *= 4;
    }
    if (pt == PT_KING && mv.type == MT_CASTLE) {
        const int long_right = stm == WHITE ? CR_WQ : CR_BQ;
        move_id = dest_n - 1;
        if (p.cr & long_right) move_id += 1;
        if ((mv.to & 7) == 7) move_id += 1;
    }
    int w1 = used_bits((u32)popc64(ours)), w2 = used_bits(num_moves);
    if (bleed) {
        const u32 rp = piece_id & 0xFFu, rm = move_id & 0xFFu;  // the reference's std::uint8_t argument
        *bleed = ((rp >> w1) | (rm >> w2)) ? (rp | ((u32)w1 << 8) | (rm << 12) | ((u32)w2 << 20) | (1u << 31)) : 0u;
    }
    u64 acc = 0;  // bits accumulate at the low end, MSB-first order == append order
    int n = 0;
    acc = (acc << w1) | (piece_id & ((1u << w1) - 1u)); n += w1;
    acc = (acc << w2) | (move_id & ((1u << w2) - 1u)); n += w2;
    // addBitsVle16 (:864-874): 4-bit groups, low group first, continuation flag on top. All four possible
    // blocks are laid out first group leftmost and the unused ones shifted out, without a loop.
    const u32 v = zz_enc((int)(short)(score - last_score));
    const int nb = 1 + (v > 0xFu) + (v > 0xFFu) + (v > 0xFFFu);
    u32 blocks = ((v & 15u) << 15) | (((v >> 4) & 15u) << 10) | (((v >> 8) & 15u) << 5) | (v >> 12);
    blocks |= nb > 1 ? (16u << 15) : 0u;
    blocks |= nb > 2 ? (16u << 10) : 0u;
    blocks |= nb > 3 ? (16u << 5) : 0u;
    acc = (acc << (5 * nb)) | (blocks >> (5 * (4 - nb)));
    n += 5 * nb;
    nbits = n;
    return (u32)(acc << (32 - n));
}

// MSB-first bit reader over a byte span (PackedMoveScoreListReader::extractBitsLE8 :623-648): a
// 32-bit window refilled 16 bits at a time with aligned halfword loads, the next halfword already
// loaded one refill ahead, so that a field costs a few register operations and the load latency of
// the movetext is off the critical path (the byte-per-field form this replaces accounted for a third
// of the chain decoder's stall samples). `pos` counts the bits consumed (numReadBytes :815-818).
struct BitReader {
    const unsigned char* p;
    u32 nbits;    // bits available
    u32 pos;
    bool overrun;
    u32 win;      // the next `have` bits of the stream in its top bits, zeros below
    u32 have;
    u32 next16;   // the two bytes behind the window as loaded (little-endian halfword): swapped when they enter it,
                  // so that nothing waits for the load before the next refill
    u32 fetched;  // bytes behind the span's start that are in win or next16
    u32 nbytes;
    __device__ __forceinline__ u32 load16(u32 idx) const  // bytes idx (low) and idx + 1 (high), zeros past the span
    {
        if (idx + 1 < nbytes) return *reinterpret_cast<const unsigned short*>(p + idx);  // p + idx is even (init)
        return idx < nbytes ? (u32)p[idx] : 0u;
    }
    __device__ __forceinline__ void init(const unsigned char* p0, u64 bytes)
    {
        p = p0;
        const u64 bits = bytes * 8;
        nbits = bits > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (u32)bits;
        nbytes = nbits >> 3;
        pos = 0;
        overrun = false;
        win = 0;
        have = 0;
        fetched = 0;
        if ((reinterpret_cast<uintptr_t>(p0) & 1) && nbytes > 0) {  // an odd start: one byte, then aligned halfwords
            win = (u32)p0[0] << 24;
            have = 8;
            fetched = 1;
        }
        next16 = load16(fetched);
        fetched += 2;
    }
    // The decoder's reads: ensure16() makes at least 16 bits visible (one refill at most), take(n) removes
    // n <= 16 visible bits (n = 0 gives 0) without any test. Bits past the span read as zeros and are
    // noticed by the caller afterwards (past_end()), which is what get() reports field by field.
    __device__ __forceinline__ void ensure16()
    {
        if (have < 16) {
            const u32 be = ((next16 & 0xFFu) << 8) | (next16 >> 8);
            win |= be << (16 - have);
            have += 16;
            next16 = load16(fetched);
            fetched += 2;
        }
    }
    __device__ __forceinline__ u32 take(int n)
    {
        const u32 v = __funnelshift_l(win, 0u, (u32)n);  // win >> (32 - n), 0 for n = 0
        win <<= n;
        have -= n;
        pos += n;
        return v;
    }
    __device__ __forceinline__ bool past_end()
    {
        if (pos > nbits) overrun = true;
        return overrun;
    }
    __device__ __forceinline__ u32 get(int n)  // n <= 8
    {
        if (n == 0) return 0;
        if (pos + (u32)n > nbits) { overrun = true; pos += n; return 0; }
        if (have < 16) {  // room for 16 more bits (and n <= 8 <= have afterwards)
            const u32 be = ((next16 & 0xFFu) << 8) | (next16 >> 8);  // stream order: byte idx first
            win |= be << (16 - have);
            have += 16;
            next16 = load16(fetched);
            fetched += 2;
        }
        const u32 v = win >> (32 - n);
        win <<= n;
        have -= n;
        pos += n;
        return v;
    }
};

// PackedMoveScoreListReader::nextMoveScore (compress_file.cpp:685-813).
// `strict` rejects ids the reference encoder can never produce (used by the speculative
// chain discovery to kill false candidates early); returns false on such an id.
__device__ __forceinline__ bool decode_ply(BitReader& r, const Pos& p, int& last_score, Move& mv, int& score, bool strict,
                                           const StepTables* T = nullptr, int* moved_out = nullptr)
{
    int stm = p.stm;
    u64 ours = pos_occ(p, stm), theirs = pos_occ(p, stm ^ 1);
    u64 occ = ours | theirs;
    u32 n_ours = (u32)popc64(ours);
    r.ensure16();  // piece id (<= 6 bits) and move id (<= 8 bits) come out of one refill
    u32 piece_id = r.take(used_bits(n_ours));
    if (strict && piece_id >= n_ours) return false;
    // an id beyond the count (corrupted movetext; the reference indexes its lookup table out of range,
    // ArithmeticUtility.h:186-209) selects square 0, as the oracle does
    int from = piece_id < n_ours ? nth_set_bit(ours, piece_id) : 0;
    const int pc_from = pos_piece_at(p, from);
    int pt = pc_from >> 1;
    if (moved_out) *moved_out = pc_from;
    mv.from = from;
    mv.type = MT_NORMAL;
    mv.promo = NO_PIECE;
    mv.to = 0;
    // Every mover type yields a destination set and a move count; the id is then read and the
    // destination selected at ONE place, so that a warp whose lanes move different piece types
    // goes through the bit reader and the n-th-bit selection once, not once per type.
    u64 dest;
    u32 n, att_n = 0;
    bool promotes = false;
    if (pt == PT_PAWN) {
        dest = pawn_destinations(p, from, ours, theirs, T);
        n = (u32)popc64(dest);
        promotes = (from >> 3) == (stm == WHITE ? 6 : 1);
        if (promotes) n *= 4;
    } else if (pt == PT_KING) {
        const int our_mask = stm == WHITE ? (CR_WK | CR_WQ) : (CR_BK | CR_BQ);
        dest = (T ? T->king[from] : king_attacks(from)) & ~ours;
        att_n = (u32)popc64(dest);
        n = att_n + (u32)__popc((u32)(p.cr & our_mask));
    } else {
        dest = piece_attacks(pt, from, occ, T) & ~ours;
        n = (u32)popc64(dest);
    }
    const u32 id = r.take(min(used_bits(n), 8));
    if (strict && id >= n) return false;
    if (pt == PT_KING && id >= att_n) {
        const int long_right = stm == WHITE ? CR_WQ : CR_BQ;
        const bool is_long = (id - att_n == 0) && (p.cr & long_right);
        mv.from = stm == WHITE ? 4 : 60;  // Move::castle Chess.h:1029-1040
        mv.to = (stm == WHITE ? 0 : 56) + (is_long ? 0 : 7);
        mv.type = MT_CASTLE;
        if (moved_out && mv.from != from) *moved_out = -1;  // (a corrupted stream: the king is not on its home square)
    } else {
        mv.to = id < n ? nth_set_bit(dest, promotes ? id >> 2 : id) : 0;
        if (promotes) {
            mv.promo = ((PT_KNIGHT + (int)(id & 3)) << 1) | stm;
            mv.type = MT_PROMOTION;
        } else if (pt == PT_PAWN && mv.to == p.ep) {
            mv.type = MT_ENPASSANT;
        }
    }
    // extractVle16 (:650-667): 5-bit blocks, 4 payload bits (low group first) under a continuation flag.
    // Two blocks are looked at per refill, without a loop: nearly every score delta fits them.
    u32 v = 0;
    int off = 0;
    for (;;) {
        r.ensure16();
        const u32 two = r.win >> 22;                 // the next two blocks
        const u32 b0 = two >> 5, b1 = two & 31u;
        const bool more0 = (b0 >> 4) != 0;
        v |= (b0 & 15u) << off;
        if (more0) v |= (b1 & 15u) << (off + 4);
        r.take(more0 ? 10 : 5);
        if (!more0 || !(b1 >> 4) || r.pos > r.nbits) break;
        off += 8;
        if (off > 24) return false;  // a ninth block: the reference shifts a 16-bit value by 32 (undefined); no encoder emits it
    }
    if (r.past_end()) return false;
    score = (int)(short)(last_score + zz_dec(v & 0xFFFFu));
    last_score = (int)(short)(-score);
    return !r.overrun;
}

}  // namespace nnp
