/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH. See oracle.h.
 *
 * Sequential restatement of the reference algorithm, one function per reference
 * function, each citing the file:line (relative to /root/reference) it follows.
 * The position model is a 64-square mailbox; bitboards are derived from it on demand.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ---------------------------------------------------------------- basic types */

/* Piece ids: src/chess/Chess.h:142-205, id = type<<1 | colour; type order Pawn..King,None */
enum { PT_PAWN = 0, PT_KNIGHT, PT_BISHOP, PT_ROOK, PT_QUEEN, PT_KING, PT_NONE };
enum { WHITE = 0, BLACK = 1 };
#define PIECE(t, c) ((uint8_t)(((t) << 1) | (c)))
#define NO_PIECE PIECE(PT_NONE, WHITE) /* 12 */
#define P_TYPE(p) ((p) >> 1)
#define P_COLOR(p) ((p)&1)
#define SQ_NONE 64
/* MoveType order: src/chess/Chess.h:905-911 */
enum { MT_NORMAL = 0, MT_PROMOTION = 1, MT_CASTLE = 2, MT_ENPASSANT = 3 };
/* CastlingRights bits: src/chess/Chess.h:1195-1205 */
enum { CR_WK = 1, CR_WQ = 2, CR_BK = 4, CR_BQ = 8, CR_ALL = 15 };

typedef uint64_t bb_t;
#define BB(sq) ((bb_t)1 << (sq))

typedef struct {
    uint8_t sq[64];
    uint8_t stm;
    uint8_t ep; /* SQ_NONE if none */
    uint8_t cr;
    uint8_t rule50; /* std::uint8_t, Position.h:1014 */
    uint16_t ply;   /* std::uint16_t m_ply, Position.h:1015 */
} pos_t;

typedef struct {
    uint8_t from, to; /* may be SQ_NONE for the null move */
    uint8_t type;
    uint8_t promo; /* piece id or NO_PIECE */
} move_t;

typedef struct {
    pos_t pos;
    move_t move;
    int16_t score;
    uint16_t ply;
    int16_t result;
} entry_t; /* TrainingDataEntry compress_file.cpp:548-555 */

/* ---------------------------------------------------------------- growable buffer */

typedef struct {
    uint8_t* p;
    size_t n, cap;
    int oom;
} buf_t;

static void buf_reserve(buf_t* b, size_t extra)
{
    if (b->oom) return;
    if (b->n + extra <= b->cap) return;
    size_t nc = b->cap ? b->cap * 2 : 4096;
    while (nc < b->n + extra) nc *= 2;
    uint8_t* np = (uint8_t*)realloc(b->p, nc);
    if (!np) { b->oom = 1; return; }
    b->p = np;
    b->cap = nc;
}
static void buf_put(buf_t* b, const void* d, size_t n)
{
    buf_reserve(b, n);
    if (b->oom) return;
    memcpy(b->p + b->n, d, n);
    b->n += n;
}
static void buf_putc(buf_t* b, char c) { buf_put(b, &c, 1); }
static void buf_puts(buf_t* b, const char* s) { buf_put(b, s, strlen(s)); }

/* ---------------------------------------------------------------- bit helpers */

static int popcnt(bb_t b) { return __builtin_popcountll(b); }
static int lsb(bb_t b) { return __builtin_ctzll(b); }

/* bb::before, Bitboard.h:730-733 */
static bb_t bb_before(int sq) { return sq >= 64 ? ~(bb_t)0 : (BB(sq) - 1); }

/* usedBitsSafe compress_file.cpp:600-604 + util::usedBits ArithmeticUtility.h:211-217 */
static int used_bits_safe(unsigned v)
{
    if (v == 0) return 0;
    v -= 1;
    if (v == 0) return 0;
    return (31 - __builtin_clz(v)) + 1;
}

/* nthSetBitIndex ArithmeticUtility.h:186-209 (any exact select is equivalent) */
static int nth_set_bit(bb_t v, unsigned n)
{
    for (unsigned i = 0; i < n; ++i) v &= v - 1;
    return v ? lsb(v) : 0;
}

#define FILE_A 0x0101010101010101ull
#define FILE_H 0x8080808080808080ull

/* bb::pawnAttacks Bitboard.cpp:500-510 */
static bb_t pawn_attacks(bb_t pawns, int color)
{
    if (color == WHITE) return ((pawns & ~FILE_H) << 9) | ((pawns & ~FILE_A) << 7);
    return ((pawns & ~FILE_H) >> 7) | ((pawns & ~FILE_A) >> 9);
}

static bb_t step_attacks(int sq, const int (*offs)[2])
{
    bb_t b = 0;
    int f = sq & 7, r = sq >> 3;
    for (int i = 0; i < 8; ++i) {
        int nf = f + offs[i][0], nr = r + offs[i][1];
        if (nf >= 0 && nf < 8 && nr >= 0 && nr < 8) b |= BB(nr * 8 + nf);
    }
    return b;
}
/* Bitboard.cpp:17-18 offsets, :62-83 / :161-183 tables */
static bb_t knight_attacks(int sq)
{
    static const int o[8][2] = {{-1, -2}, {-1, 2}, {1, -2}, {1, 2}, {-2, -1}, {-2, 1}, {2, -1}, {2, 1}};
    return step_attacks(sq, o);
}
static bb_t king_attacks(int sq)
{
    static const int o[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
    return step_attacks(sq, o);
}
/* slider attacks: result-equivalent to the fancy-magic lookup (Bitboard.h:689-708, Bitboard.cpp:401-464) */
static bb_t ray_attacks(int sq, bb_t occ, const int (*dirs)[2], int nd)
{
    bb_t b = 0;
    for (int d = 0; d < nd; ++d) {
        int f = (sq & 7) + dirs[d][0], r = (sq >> 3) + dirs[d][1];
        while (f >= 0 && f < 8 && r >= 0 && r < 8) {
            int s = r * 8 + f;
            b |= BB(s);
            if (occ & BB(s)) break;
            f += dirs[d][0];
            r += dirs[d][1];
        }
    }
    return b;
}
static bb_t bishop_attacks(int sq, bb_t occ)
{
    static const int d[4][2] = {{1, 1}, {1, -1}, {-1, -1}, {-1, 1}};
    return ray_attacks(sq, occ, d, 4);
}
static bb_t rook_attacks(int sq, bb_t occ)
{
    static const int d[4][2] = {{0, 1}, {1, 0}, {0, -1}, {-1, 0}};
    return ray_attacks(sq, occ, d, 4);
}
/* bb::attacks(pt, sq, occ) Bitboard.h:864-880; Pawn/None -> empty (pseudoAttacks table rows are zero) */
static bb_t piece_attacks(int pt, int sq, bb_t occ)
{
    switch (pt) {
    case PT_KNIGHT: return knight_attacks(sq);
    case PT_BISHOP: return bishop_attacks(sq, occ);
    case PT_ROOK: return rook_attacks(sq, occ);
    case PT_QUEEN: return bishop_attacks(sq, occ) | rook_attacks(sq, occ);
    case PT_KING: return king_attacks(sq);
    default: return 0;
    }
}

/* ---------------------------------------------------------------- position */

static void pos_init(pos_t* p) /* Position() Position.h:828-836 */
{
    memset(p->sq, NO_PIECE, 64);
    p->stm = WHITE;
    p->ep = SQ_NONE;
    p->cr = CR_ALL;
    p->rule50 = 0;
    p->ply = 0;
}
static bb_t pos_piece_bb(const pos_t* p, uint8_t piece)
{
    bb_t b = 0;
    for (int s = 0; s < 64; ++s)
        if (p->sq[s] == piece) b |= BB(s);
    return b;
}
static bb_t pos_color_bb(const pos_t* p, int color)
{
    bb_t b = 0;
    for (int s = 0; s < 64; ++s)
        if (p->sq[s] != NO_PIECE && P_COLOR(p->sq[s]) == color) b |= BB(s);
    return b;
}
/* Board::kingSquare Position.h:742-745; -1 when there is no such king (reference: undefined) */
static int pos_king_sq(const pos_t* p, int color)
{
    bb_t b = pos_piece_bb(p, PIECE(PT_KING, color));
    return b ? lsb(b) : -1;
}

/* bb::isAttackedBySlider Bitboard.cpp:536-554 */
static int attacked_by_slider(int sq, bb_t bishops, bb_t rooks, bb_t queens, bb_t occ)
{
    if (bishop_attacks(sq, occ) & (bishops | queens)) return 1;
    return (rook_attacks(sq, occ) & (rooks | queens)) != 0;
}

/* Position::isEpPossible + isEpPossibleColdPath Position.cpp:824-883.
 * Evaluated on whatever board `p` currently holds: the caller decides whether that is the
 * post-move board (setEpSquare / FEN) or the PRE-move board (inside doMove, :647-657). */
static int is_ep_possible(const pos_t* p, int ep, int side)
{
    bb_t attackers = pawn_attacks(BB(ep), !side) & pos_piece_bb(p, PIECE(PT_PAWN, side));
    if (!attackers) return 0;
    bb_t all = pos_color_bb(p, WHITE) | pos_color_bb(p, BLACK);
    int ksq = pos_king_sq(p, side);
    bb_t bishops = pos_piece_bb(p, PIECE(PT_BISHOP, !side));
    bb_t rooks = pos_piece_bb(p, PIECE(PT_ROOK, !side));
    bb_t queens = pos_piece_bb(p, PIECE(PT_QUEEN, !side));
    if (ksq < 0) return 1; /* outside the parity domain */
    while (attackers) {
        int sq = lsb(attackers);
        attackers &= attackers - 1;
        bb_t queen_lines = bishop_attacks(ksq, 0) | rook_attacks(ksq, 0);
        if (((bishops | rooks | queens) & queen_lines) == 0) return 1;
        int captured = (ep & 7) | (sq & 56);
        bb_t occ = ((all ^ BB(sq)) | BB(ep)) ^ BB(captured);
        if (!attacked_by_slider(ksq, bishops, rooks, queens, occ)) return 1;
    }
    return 0;
}

/* Position::setEpSquare Position.h:868-872 + nullifyEpSquareIfNotPossible Position.cpp:885-891 */
static void pos_set_ep(pos_t* p, int ep)
{
    p->ep = (uint8_t)ep;
    if (p->ep != SQ_NONE && !is_ep_possible(p, p->ep, p->stm)) p->ep = SQ_NONE;
}

/* Board::doMove + doMoveColdPath Position.h:300-439 (mailbox part) */
static void board_do_move(pos_t* p, move_t m)
{
    if (m.type == MT_NORMAL) {
        uint8_t piece = p->sq[m.from];
        p->sq[m.to] = piece;
        p->sq[m.from] = NO_PIECE;
    } else if (m.type == MT_PROMOTION) {
        p->sq[m.to] = m.promo;
        p->sq[m.from] = NO_PIECE;
    } else if (m.type == MT_ENPASSANT) {
        uint8_t moved = p->sq[m.from];
        int cap = (m.to & 7) | (m.from & 56);
        p->sq[m.to] = moved;
        p->sq[m.from] = NO_PIECE;
        p->sq[cap] = NO_PIECE;
    } else {
        int rook_from = m.to, king_from = m.from;
        uint8_t rook = p->sq[rook_from], king = p->sq[king_from];
        int color = P_COLOR(king);
        int is_short = (m.to & 7) == 7; /* CastlingTraits::moveCastlingType CastlingTraits.h:37-40 */
        int rook_to = (color == WHITE ? 0 : 56) + (is_short ? 5 : 3); /* f/d CastlingTraits.h:11 */
        int king_to = (color == WHITE ? 0 : 56) + (is_short ? 6 : 2); /* g/c CastlingTraits.h:12 */
        p->sq[rook_from] = NO_PIECE;
        p->sq[king_from] = NO_PIECE;
        p->sq[rook_to] = rook;
        p->sq[king_to] = king;
    }
}

/* detail::lookup::preservedCastlingRights Position.cpp:605-624 */
static uint8_t preserved_cr(int sq)
{
    switch (sq) {
    case 4: return (uint8_t)(CR_ALL & ~(CR_WK | CR_WQ));
    case 60: return (uint8_t)(CR_ALL & ~(CR_BK | CR_BQ));
    case 7: return (uint8_t)(CR_ALL & ~CR_WK);
    case 0: return (uint8_t)(CR_ALL & ~CR_WQ);
    case 63: return (uint8_t)(CR_ALL & ~CR_BK);
    case 56: return (uint8_t)(CR_ALL & ~CR_BQ);
    default: return CR_ALL;
    }
}

/* Position::doMove Position.cpp:626-662 */
static void pos_do_move(pos_t* p, move_t m)
{
    int moved_type = P_TYPE(p->sq[m.from]);
    p->ply = (uint16_t)(p->ply + 1);
    p->rule50 = (uint8_t)(p->rule50 + 1);
    if (m.type != MT_CASTLE && (moved_type == PT_PAWN || p->sq[m.to] != NO_PIECE)) p->rule50 = 0;
    p->cr &= preserved_cr(m.from);
    p->cr &= preserved_cr(m.to);
    p->ep = SQ_NONE;
    if (moved_type == PT_PAWN && ((m.to ^ m.from) == 16)) {
        int cand = (m.to + m.from) >> 1;
        /* evaluated BEFORE the board is updated -- quirk Q1 */
        if (is_ep_possible(p, cand, !p->stm)) p->ep = (uint8_t)cand;
    }
    board_do_move(p, m);
    p->stm = !p->stm;
}

/* Position::operator== Position.h:977-984 + Board::operator== :243-258 */
static int pos_equal(const pos_t* a, const pos_t* b)
{
    return a->stm == b->stm && a->ep == b->ep && a->cr == b->cr && memcmp(a->sq, b->sq, 64) == 0;
}

/* ---------------------------------------------------------------- .bin record codec */

/* nodchip::BitStream compress_file.cpp:126-185: LSB-first over bytes */
typedef struct {
    uint8_t* data;
    int cursor;
} bitstream_t;
static int bs_read1(bitstream_t* s)
{
    /* the struct is 40 bytes, reads past bit 319 would leave the record */
    int b = s->cursor < 320 ? (s->data[s->cursor / 8] >> (s->cursor & 7)) & 1 : 0;
    ++s->cursor;
    return b;
}
static int bs_read(bitstream_t* s, int n)
{
    int r = 0;
    for (int i = 0; i < n; ++i) r |= bs_read1(s) ? (1 << i) : 0;
    return r;
}
static void bs_write1(bitstream_t* s, int b)
{
    if (b && s->cursor < 256) s->data[s->cursor / 8] |= (uint8_t)(1 << (s->cursor & 7));
    ++s->cursor;
}
static void bs_write(bitstream_t* s, int d, int n)
{
    for (int i = 0; i < n; ++i) bs_write1(s, d & (1 << i));
}

/* pos_from_packed_sfen compress_file.cpp:364-446 (+ read_board_piece_from_stream :336-360).
 * `rec` points at the 40-byte record so that overruns read the same bytes the reference reads. */
static int sfen_unpack(const uint8_t* rec, pos_t* pos)
{
    bitstream_t s = {(uint8_t*)rec, 0};
    pos_init(pos);
    pos->stm = (uint8_t)bs_read1(&s);
    int wk = bs_read(&s, 6);
    pos->sq[wk] = PIECE(PT_KING, WHITE);
    int bk = bs_read(&s, 6);
    pos->sq[bk] = PIECE(PT_KING, BLACK);
    for (int r = 7; r >= 0; --r) {
        for (int f = 0; f < 8; ++f) {
            int sq = r * 8 + f;
            if (P_TYPE(pos->sq[sq]) == PT_KING) continue;
            /* Huffman: '0' = empty; '1' + 3 bits (LSB-first) = piece type 0..4; + colour bit */
            if (!bs_read1(&s)) continue;
            int t = bs_read(&s, 3);
            if (t > PT_QUEEN) return ORC_ERR_BAD_SFEN; /* reference: never terminates */
            int c = bs_read1(&s);
            pos->sq[sq] = PIECE(t, c);
            if (s.cursor > 256) return ORC_ERR_BAD_SFEN; /* :407-408 */
        }
    }
    uint8_t cr = 0;
    if (bs_read1(&s)) cr |= CR_WK;
    if (bs_read1(&s)) cr |= CR_WQ;
    if (bs_read1(&s)) cr |= CR_BK;
    if (bs_read1(&s)) cr |= CR_BQ;
    pos->cr = cr;
    if (bs_read1(&s)) pos_set_ep(pos, bs_read(&s, 6)); /* :430-433 post-move nullification */
    pos->rule50 = (uint8_t)bs_read(&s, 6);
    int hm = bs_read(&s, 8);
    pos->ply = (uint16_t)(2 * hm - 1 + (pos->stm == BLACK)); /* setHalfMove Position.h:938-941 */
    if (s.cursor > 256) return ORC_ERR_BAD_SFEN; /* :441-442 */
    return ORC_OK;
}

/* SfenPacker::pack compress_file.cpp:266-312 */
static void sfen_pack(const pos_t* pos, uint8_t out[32])
{
    memset(out, 0, 32);
    bitstream_t s = {out, 0};
    bs_write1(&s, pos->stm);
    int wk = pos_king_sq(pos, WHITE), bk = pos_king_sq(pos, BLACK);
    bs_write(&s, wk < 0 ? 0 : wk, 6);
    bs_write(&s, bk < 0 ? 0 : bk, 6);
    for (int r = 7; r >= 0; --r) {
        for (int f = 0; f < 8; ++f) {
            uint8_t pc = pos->sq[r * 8 + f];
            if (P_TYPE(pc) == PT_KING) continue;
            if (pc == NO_PIECE) {
                bs_write1(&s, 0);
            } else {
                bs_write(&s, 1 | (P_TYPE(pc) << 1), 4); /* huffman_table :236-245 */
                bs_write1(&s, P_COLOR(pc));
            }
        }
    }
    bs_write1(&s, (pos->cr & CR_WK) != 0);
    bs_write1(&s, (pos->cr & CR_WQ) != 0);
    bs_write1(&s, (pos->cr & CR_BK) != 0);
    bs_write1(&s, (pos->cr & CR_BQ) != 0);
    if (pos->ep == SQ_NONE) {
        bs_write1(&s, 0);
    } else {
        bs_write1(&s, 1);
        bs_write(&s, pos->ep, 6);
    }
    bs_write(&s, pos->rule50, 6);
    bs_write(&s, (uint16_t)((pos->ply + 1) / 2), 8); /* halfMove() Position.h:933-936 */
}

/* StockfishMove::toMove compress_file.cpp:63-83 */
static move_t sfmove_to_move(uint16_t raw)
{
    move_t m;
    m.to = raw & 63;
    m.from = (raw >> 6) & 63;
    int promo_idx = (raw >> 12) & 3;
    int flag = (raw >> 14) & 3;
    m.type = flag == 1 ? MT_PROMOTION : flag == 2 ? MT_ENPASSANT : flag == 3 ? MT_CASTLE : MT_NORMAL;
    m.promo = NO_PIECE;
    if (m.type == MT_PROMOTION) m.promo = PIECE(PT_KNIGHT + promo_idx, (m.to >> 3) == 7 ? WHITE : BLACK);
    return m;
}
/* StockfishMove::fromMove compress_file.cpp:35-61 (same |= / <<= sequence on 16 bits) */
static uint16_t move_to_sfmove(move_t m)
{
    unsigned flag = m.type == MT_PROMOTION ? 1 : m.type == MT_ENPASSANT ? 2 : m.type == MT_CASTLE ? 3 : 0;
    unsigned promo = m.type == MT_PROMOTION ? (unsigned)(P_TYPE(m.promo) - PT_KNIGHT) : 0;
    uint16_t raw = 0;
    raw |= (uint16_t)flag;
    raw <<= 2;
    raw |= (uint16_t)promo;
    raw <<= 6;
    raw |= (uint16_t)m.from;
    raw <<= 6;
    raw |= (uint16_t)m.to;
    return raw;
}

static uint16_t rd16le(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

/* packedSfenValueToTrainingDataEntry compress_file.cpp:557-568; layout :95-122 */
static int record_to_entry(const uint8_t rec[40], entry_t* e)
{
    int rc = sfen_unpack(rec, &e->pos);
    if (rc != ORC_OK) return rc;
    e->score = (int16_t)rd16le(rec + 32);
    e->move = sfmove_to_move(rd16le(rec + 34));
    e->ply = rd16le(rec + 36);
    e->result = (int16_t)(int8_t)rec[38];
    return ORC_OK;
}
/* trainingDataEntryToPackedSfenValue compress_file.cpp:570-585 */
static void entry_to_record(const entry_t* e, uint8_t rec[40])
{
    sfen_pack(&e->pos, rec);
    uint16_t sc = (uint16_t)e->score, mv = move_to_sfmove(e->move);
    rec[32] = (uint8_t)sc;
    rec[33] = (uint8_t)(sc >> 8);
    rec[34] = (uint8_t)mv;
    rec[35] = (uint8_t)(mv >> 8);
    rec[36] = (uint8_t)e->ply;
    rec[37] = (uint8_t)(e->ply >> 8);
    rec[38] = (uint8_t)(int8_t)e->result;
    rec[39] = 0xff;
}

/* ---------------------------------------------------------------- stem + movetext codec */

/* signedToUnsigned / unsignedToSigned compress_file.cpp:524-546 */
static uint16_t s2u(int16_t a)
{
    uint16_t r = (uint16_t)a;
    if (r & 0x8000) r ^= 0x7FFF;
    return (uint16_t)((r << 1) | (r >> 15));
}
static int16_t u2s(uint16_t r)
{
    r = (uint16_t)((r << 15) | (r >> 1));
    if (r & 0x8000) r ^= 0x7FFF;
    return (int16_t)r;
}

/* isContinuation compress_file.cpp:587-593 */
static int is_continuation(const entry_t* lhs, const entry_t* rhs)
{
    if ((int)lhs->result != -(int)rhs->result) return 0;
    if ((int)lhs->ply + 1 != (int)rhs->ply) return 0;
    pos_t after = lhs->pos; /* Position::afterMove Position.cpp:813-822 */
    pos_do_move(&after, lhs->move);
    return pos_equal(&after, &rhs->pos);
}

/* Position::compress Position.h:1374-1406 with detail::compress* :1269-1353 */
static uint8_t compress_piece(const pos_t* p, int sq, uint8_t pc)
{
    switch (P_TYPE(pc)) {
    case PT_PAWN:
        if (p->ep == SQ_NONE) return pc;
        if ((sq & 7) == (p->ep & 7) &&
            ((((sq >> 3) == 3) & (p->stm == BLACK)) | (((sq >> 3) == 4) & (p->stm == WHITE))))
            return 12;
        return pc;
    case PT_ROOK:
        if (P_COLOR(pc) == WHITE && ((sq == 0 && (p->cr & CR_WQ)) || (sq == 7 && (p->cr & CR_WK)))) return 13;
        if (P_COLOR(pc) == BLACK && ((sq == 56 && (p->cr & CR_BQ)) || (sq == 63 && (p->cr & CR_BK)))) return 14;
        return pc;
    case PT_KING:
        if (P_COLOR(pc) == WHITE) return 10;
        return p->stm == WHITE ? 11 : 15;
    default: return pc;
    }
}

/* packEntry compress_file.cpp:997-1020; CompressedPosition::writeToBigEndian Position.h:1245-1257;
 * CompressedMove(Move) Chess.h:1071-1096 */
static void pack_stem(const entry_t* e, uint8_t out[32])
{
    memset(out, 0, 32);
    bb_t occ = pos_color_bb(&e->pos, WHITE) | pos_color_bb(&e->pos, BLACK);
    for (int i = 0; i < 8; ++i) out[i] = (uint8_t)(occ >> (56 - 8 * i));
    int k = 0;
    for (bb_t b = occ; b; b &= b - 1, ++k) {
        int sq = lsb(b);
        uint8_t nib = compress_piece(&e->pos, sq, e->pos.sq[sq]);
        out[8 + k / 2] |= (uint8_t)(nib << ((k & 1) * 4));
    }
    uint16_t cm = 0;
    if (e->move.from != e->move.to) {
        cm = (uint16_t)((e->move.type << 14) | (e->move.from << 8) | (e->move.to << 2));
        if (e->move.type == MT_PROMOTION) cm |= (uint16_t)(P_TYPE(e->move.promo) - PT_KNIGHT);
    }
    out[24] = (uint8_t)(cm >> 8);
    out[25] = (uint8_t)cm;
    uint16_t sc = s2u(e->score);
    out[26] = (uint8_t)(sc >> 8);
    out[27] = (uint8_t)sc;
    uint16_t pr = (uint16_t)(e->ply | (s2u(e->result) << 14));
    out[28] = (uint8_t)(pr >> 8);
    out[29] = (uint8_t)pr;
    out[30] = (uint8_t)(e->pos.rule50 >> 8); /* always 0: rule50Counter() is a uint8_t */
    out[31] = e->pos.rule50;
}

/* unpackEntry compress_file.cpp:1022-1043; CompressedPosition::decompress Position.h:1408-1505;
 * CompressedMove::decompress Chess.h:1142-1172 */
static void unpack_stem(const uint8_t in[32], entry_t* e)
{
    pos_t* p = &e->pos;
    pos_init(p);
    p->cr = 0;
    bb_t occ = 0;
    for (int i = 0; i < 8; ++i) occ = (occ << 8) | in[i];
    int k = 0;
    for (bb_t b = occ; b; b &= b - 1, ++k) {
        int sq = lsb(b);
        uint8_t nib = (in[8 + k / 2] >> ((k & 1) * 4)) & 15;
        switch (nib) {
        case 12:
            if ((sq >> 3) == 3) {
                p->sq[sq] = PIECE(PT_PAWN, WHITE);
                p->ep = (uint8_t)(sq - 8);
            } else {
                p->sq[sq] = PIECE(PT_PAWN, BLACK);
                p->ep = (uint8_t)(sq + 8);
            }
            break;
        case 13:
            p->sq[sq] = PIECE(PT_ROOK, WHITE);
            p->cr |= (sq == 0) ? CR_WQ : CR_WK;
            break;
        case 14:
            p->sq[sq] = PIECE(PT_ROOK, BLACK);
            p->cr |= (sq == 56) ? CR_BQ : CR_BK;
            break;
        case 15:
            p->sq[sq] = PIECE(PT_KING, BLACK);
            p->stm = BLACK;
            break;
        default: p->sq[sq] = nib; break;
        }
    }
    uint16_t cm = (uint16_t)((in[24] << 8) | in[25]);
    if (cm == 0) {
        e->move.from = e->move.to = SQ_NONE; /* Move::null() */
        e->move.type = MT_NORMAL;
        e->move.promo = NO_PIECE;
    } else {
        e->move.type = (uint8_t)(cm >> 14);
        e->move.from = (cm >> 8) & 63;
        e->move.to = (cm >> 2) & 63;
        e->move.promo = NO_PIECE;
        if (e->move.type == MT_PROMOTION)
            e->move.promo = PIECE(PT_KNIGHT + (cm & 3), (e->move.to >> 3) == 0 ? BLACK : WHITE);
    }
    e->score = u2s((uint16_t)((in[26] << 8) | in[27]));
    uint16_t pr = (uint16_t)((in[28] << 8) | in[29]);
    e->ply = pr & 0x3FFF;
    p->ply = e->ply;
    e->result = u2s(pr >> 14);
    p->rule50 = (uint8_t)((in[30] << 8) | in[31]);
}

/* destination set / index arithmetic shared by addMoveScore (:877-989) and nextMoveScore (:685-813) */
static bb_t pawn_destinations(const pos_t* pos, int from, bb_t ours, bb_t theirs)
{
    int stm = pos->stm;
    bb_t occ = ours | theirs;
    bb_t targets = theirs;
    if (pos->ep != SQ_NONE) targets |= BB(pos->ep);
    bb_t dest = pawn_attacks(BB(from), stm) & targets;
    int fwd = stm == WHITE ? 8 : -8;
    int s1 = from + fwd;
    if (s1 >= 0 && s1 < 64 && !(occ & BB(s1))) {
        dest |= BB(s1);
        int s2 = s1 + fwd;
        int start_rank = stm == WHITE ? 1 : 6;
        if ((from >> 3) == start_rank && s2 >= 0 && s2 < 64 && !(occ & BB(s2))) dest |= BB(s2);
    }
    return dest;
}

/* PackedMoveScoreList compress_file.cpp:827-994 */
typedef struct {
    buf_t text;
    unsigned bits_left;
    int16_t last_score;
    uint16_t num_plies;
} movelist_t;

static void ml_clear(movelist_t* ml, const entry_t* e) /* :832-838 */
{
    ml->num_plies = 0;
    ml->text.n = 0;
    ml->bits_left = 0;
    ml->last_score = (int16_t)(-e->score);
}
static void ml_add_bits(movelist_t* ml, uint8_t bits, unsigned count) /* addBitsLE8 :840-862 */
{
    if (count == 0) return;
    if (ml->bits_left == 0) {
        buf_putc(&ml->text, (char)(uint8_t)(bits << (8 - count)));
        ml->bits_left = 8;
    } else if (count <= ml->bits_left) {
        ml->text.p[ml->text.n - 1] |= (uint8_t)(bits << (ml->bits_left - count));
    } else {
        unsigned spill = count - ml->bits_left;
        ml->text.p[ml->text.n - 1] |= (uint8_t)(bits >> spill);
        buf_putc(&ml->text, (char)(uint8_t)(bits << (8 - spill)));
        ml->bits_left += 8;
    }
    ml->bits_left -= count;
}
static void ml_add_vle16(movelist_t* ml, uint16_t v) /* addBitsVle16 :864-874, block size 4 (:606) */
{
    for (;;) {
        uint8_t block = (uint8_t)((v & 15) | ((v > 15) << 4));
        ml_add_bits(ml, block, 5);
        v >>= 4;
        if (v == 0) break;
    }
}
static void ml_add_move_score(movelist_t* ml, const pos_t* pos, move_t mv, int16_t score) /* :877-989 */
{
    int stm = pos->stm;
    bb_t ours = pos_color_bb(pos, stm), theirs = pos_color_bb(pos, !stm);
    bb_t occ = ours | theirs;
    unsigned piece_id = (unsigned)popcnt(ours & bb_before(mv.from));
    unsigned num_moves = 0;
    int move_id = 0;
    int pt = P_TYPE(pos->sq[mv.from]);
    if (pt == PT_PAWN) {
        bb_t dest = pawn_destinations(pos, mv.from, ours, theirs);
        move_id = popcnt(dest & bb_before(mv.to));
        num_moves = (unsigned)popcnt(dest);
        int second_to_last = stm == WHITE ? 6 : 1;
        if ((mv.from >> 3) == second_to_last) {
            int promo_idx = P_TYPE(mv.promo) - PT_KNIGHT;
            move_id = move_id * 4 + promo_idx;
            num_moves *= 4;
        }
    } else if (pt == PT_KING) {
        uint8_t our_mask = stm == WHITE ? (CR_WK | CR_WQ) : (CR_BK | CR_BQ);
        bb_t att = king_attacks(mv.from) & ~ours;
        int att_n = popcnt(att);
        int n_cr = __builtin_popcount(pos->cr & our_mask);
        num_moves = (unsigned)(att_n + n_cr);
        if (mv.type == MT_CASTLE) {
            uint8_t long_right = stm == WHITE ? CR_WQ : CR_BQ;
            move_id = att_n - 1;
            if (pos->cr & long_right) move_id += 1;
            if ((mv.to & 7) == 7) move_id += 1;
        } else {
            move_id = popcnt(att & bb_before(mv.to));
        }
    } else {
        bb_t att = piece_attacks(pt, mv.from, occ) & ~ours;
        move_id = popcnt(att & bb_before(mv.to));
        num_moves = (unsigned)popcnt(att);
    }
    ml_add_bits(ml, (uint8_t)piece_id, (unsigned)used_bits_safe((unsigned)popcnt(ours)));
    ml_add_bits(ml, (uint8_t)move_id, (unsigned)used_bits_safe(num_moves));
    uint16_t delta = s2u((int16_t)(score - ml->last_score));
    ml_add_vle16(ml, delta);
    ml->last_score = (int16_t)(-score);
    ml->num_plies = (uint16_t)(ml->num_plies + 1);
}

/* PackedMoveScoreListReader compress_file.cpp:608-825: MSB-first bit reader over a bounded span */
typedef struct {
    const uint8_t* p;
    size_t nbytes; /* bytes available from p to the end of the chunk */
    size_t bitpos;
    int overrun;
} mreader_t;
static unsigned mr_bits(mreader_t* r, unsigned count) /* extractBitsLE8 :623-648 */
{
    unsigned v = 0;
    for (unsigned i = 0; i < count; ++i) {
        size_t byte = r->bitpos >> 3;
        unsigned bit = 0;
        if (byte < r->nbytes) bit = (r->p[byte] >> (7 - (r->bitpos & 7))) & 1;
        else r->overrun = 1;
        v = (v << 1) | bit;
        ++r->bitpos;
    }
    return v;
}
static uint16_t mr_vle16(mreader_t* r) /* extractVle16 :650-667 */
{
    uint16_t v = 0;
    unsigned offset = 0;
    for (;;) {
        unsigned block = mr_bits(r, 5);
        v |= (uint16_t)((block & 15) << offset);
        if (!(block >> 4)) break;
        offset += 4;
        if (r->overrun) break;
    }
    return v;
}
/* nextMoveScore :685-813 */
static void mr_next_move_score(mreader_t* r, const pos_t* pos, int16_t* last_score, move_t* mv, int16_t* score)
{
    int stm = pos->stm;
    bb_t ours = pos_color_bb(pos, stm), theirs = pos_color_bb(pos, !stm);
    bb_t occ = ours | theirs;
    unsigned piece_id = mr_bits(r, (unsigned)used_bits_safe((unsigned)popcnt(ours)));
    int from = nth_set_bit(ours, piece_id);
    int pt = P_TYPE(pos->sq[from]);
    mv->from = (uint8_t)from;
    mv->type = MT_NORMAL;
    mv->promo = NO_PIECE;
    if (pt == PT_PAWN) {
        bb_t dest = pawn_destinations(pos, from, ours, theirs);
        unsigned n = (unsigned)popcnt(dest);
        int promotion_rank = stm == WHITE ? 6 : 1;
        if ((from >> 3) == promotion_rank) {
            unsigned id = mr_bits(r, (unsigned)used_bits_safe(n * 4));
            mv->promo = PIECE(PT_KNIGHT + (id % 4), stm);
            mv->to = (uint8_t)nth_set_bit(dest, id / 4);
            mv->type = MT_PROMOTION;
        } else {
            unsigned id = mr_bits(r, (unsigned)used_bits_safe(n));
            mv->to = (uint8_t)nth_set_bit(dest, id);
            if (mv->to == pos->ep) mv->type = MT_ENPASSANT;
        }
    } else if (pt == PT_KING) {
        uint8_t our_mask = stm == WHITE ? (CR_WK | CR_WQ) : (CR_BK | CR_BQ);
        bb_t att = king_attacks(from) & ~ours;
        unsigned att_n = (unsigned)popcnt(att);
        unsigned n_cr = (unsigned)__builtin_popcount(pos->cr & our_mask);
        unsigned id = mr_bits(r, (unsigned)used_bits_safe(att_n + n_cr));
        if (id >= att_n) {
            unsigned idx = id - att_n;
            uint8_t long_right = stm == WHITE ? CR_WQ : CR_BQ;
            int is_long = (idx == 0) && (pos->cr & long_right);
            /* Move::castle Chess.h:1029-1040: fixed e1/e8 -> h/a file rook square */
            mv->from = (uint8_t)(stm == WHITE ? 4 : 60);
            mv->to = (uint8_t)((stm == WHITE ? 0 : 56) + (is_long ? 0 : 7));
            mv->type = MT_CASTLE;
        } else {
            mv->to = (uint8_t)nth_set_bit(att, id);
        }
    } else {
        bb_t att = piece_attacks(pt, from, occ) & ~ours;
        unsigned id = mr_bits(r, (unsigned)used_bits_safe((unsigned)popcnt(att)));
        mv->to = (uint8_t)nth_set_bit(att, id);
    }
    *score = (int16_t)(*last_score + u2s(mr_vle16(r)));
    *last_score = (int16_t)(-*score);
}

/* ---------------------------------------------------------------- writer (chunks) */

#define CHUNK_THRESHOLD (1024u * 1024u) /* suggestedChunkSize compress_file.cpp:20 */
#define MAX_CHUNK (100u * 1024u * 1024u) /* maxChunkSize :22 */

typedef struct {
    buf_t* out;
    buf_t chunk; /* m_packedEntries / m_packedSize */
    entry_t last;
    movelist_t ml;
    int is_first;
} writer_t;

static void writer_init(writer_t* w, buf_t* out) /* :1049-1059 */
{
    memset(w, 0, sizeof(*w));
    w->out = out;
    pos_init(&w->last.pos);
    w->last.move.from = w->last.move.to = 0;
    w->last.move.type = MT_NORMAL;
    w->last.move.promo = NO_PIECE;
    w->last.ply = 0xFFFF;
    w->last.result = 0x7FFF;
    w->is_first = 1;
}
static void writer_flush_chunk(writer_t* w) /* CompressedTrainingDataFile::append :462-466, :486-498 */
{
    uint8_t h[8] = {'B', 'I', 'N', 'P', 0, 0, 0, 0};
    uint32_t n = (uint32_t)w->chunk.n;
    h[4] = (uint8_t)n;
    h[5] = (uint8_t)(n >> 8);
    h[6] = (uint8_t)(n >> 16);
    h[7] = (uint8_t)(n >> 24);
    buf_put(w->out, h, 8);
    buf_put(w->out, w->chunk.p, w->chunk.n);
    w->chunk.n = 0;
}
static void writer_write_movelist(writer_t* w) /* :1116-1125 */
{
    buf_putc(&w->chunk, (char)(w->ml.num_plies >> 8));
    buf_putc(&w->chunk, (char)(w->ml.num_plies));
    if (w->ml.num_plies > 0) buf_put(&w->chunk, w->ml.text.p, w->ml.text.n);
}
static void writer_add(writer_t* w, const entry_t* e) /* addTrainingDataEntry :1061-1092 */
{
    if (is_continuation(&w->last, e)) {
        ml_add_move_score(&w->ml, &e->pos, e->move, e->score);
    } else {
        if (!w->is_first) writer_write_movelist(w);
        if (w->chunk.n >= CHUNK_THRESHOLD) writer_flush_chunk(w);
        uint8_t stem[32];
        pack_stem(e, stem);
        buf_put(&w->chunk, stem, 32);
        ml_clear(&w->ml, e);
        w->is_first = 0;
    }
    w->last = *e;
}
static void writer_finish(writer_t* w) /* ~CompressedTrainingDataEntryWriter :1094-1106 */
{
    if (w->chunk.n > 0) {
        if (!w->is_first) writer_write_movelist(w);
        writer_flush_chunk(w);
    }
    free(w->chunk.p);
    free(w->ml.text.p);
}

/* ---------------------------------------------------------------- reader (chunks) */

typedef struct {
    const uint8_t* in;
    size_t in_len, file_pos;
    const uint8_t* chunk;
    size_t chunk_len, offset;
    int is_end;
    int err;
    /* chain state (PackedMoveScoreListReader) */
    int in_chain;
    entry_t entry;
    uint16_t num_plies, read_plies;
    int16_t last_score;
    mreader_t mr;
} reader_t;

/* hasNextChunk / readNextChunk / readChunkHeader compress_file.cpp:468-521 */
static int reader_next_chunk(reader_t* r)
{
    if (r->file_pos >= r->in_len) return 0;
    if (r->in_len - r->file_pos < 8 || memcmp(r->in + r->file_pos, "BINP", 4) != 0) {
        r->err = ORC_ERR_BAD_MAGIC;
        return 0;
    }
    const uint8_t* h = r->in + r->file_pos;
    uint32_t size = (uint32_t)h[4] | ((uint32_t)h[5] << 8) | ((uint32_t)h[6] << 16) | ((uint32_t)h[7] << 24);
    if (size > MAX_CHUNK) {
        r->err = ORC_ERR_CHUNK_TOO_LARGE;
        return 0;
    }
    if (r->in_len - r->file_pos - 8 < size) {
        r->err = ORC_ERR_TRUNCATED;
        return 0;
    }
    r->chunk = h + 8;
    r->chunk_len = size;
    r->offset = 0;
    r->file_pos += 8 + (size_t)size;
    return 1;
}
static void reader_init(reader_t* r, const uint8_t* in, size_t n) /* :1132-1147 */
{
    memset(r, 0, sizeof(*r));
    r->in = in;
    r->in_len = n;
    if (!reader_next_chunk(r)) r->is_end = 1;
}
static void reader_fetch_if_needed(reader_t* r) /* fetchNextChunkIfNeeded :1199-1213 */
{
    if (r->offset + 34 > r->chunk_len) {
        if (!reader_next_chunk(r)) r->is_end = 1;
    }
}
/* CompressedTrainingDataEntryReader::next :1154-1190 + PackedMoveScoreListReader::nextEntry :669-678 */
static int reader_next(reader_t* r, entry_t* out)
{
    if (r->in_chain) {
        pos_do_move(&r->entry.pos, r->entry.move);
        move_t mv;
        int16_t sc;
        mr_next_move_score(&r->mr, &r->entry.pos, &r->last_score, &mv, &sc);
        if (r->mr.overrun) return ORC_ERR_TRUNCATED;
        r->entry.move = mv;
        r->entry.score = sc;
        r->entry.ply = (uint16_t)(r->entry.ply + 1);
        r->entry.result = (int16_t)(-r->entry.result);
        r->read_plies = (uint16_t)(r->read_plies + 1);
        *out = r->entry;
        if (!(r->read_plies < r->num_plies)) {
            r->offset += (r->mr.bitpos + 7) / 8; /* numReadBytes :815-818 */
            r->in_chain = 0;
            reader_fetch_if_needed(r);
        }
        return r->err;
    }
    if (r->offset + 34 > r->chunk_len) return ORC_ERR_TRUNCATED;
    unpack_stem(r->chunk + r->offset, &r->entry);
    r->offset += 32;
    r->num_plies = (uint16_t)((r->chunk[r->offset] << 8) | r->chunk[r->offset + 1]);
    r->offset += 2;
    *out = r->entry;
    if (r->num_plies > 0) {
        r->in_chain = 1;
        r->read_plies = 0;
        r->last_score = (int16_t)(-r->entry.score);
        r->mr.p = r->chunk + r->offset;
        r->mr.nbytes = r->chunk_len - r->offset;
        r->mr.bitpos = 0;
        r->mr.overrun = 0;
    } else {
        reader_fetch_if_needed(r);
    }
    return r->err;
}

/* ---------------------------------------------------------------- text: FEN / UCI / .plain */

static const char fen_piece_char[13] = {'P', 'p', 'N', 'n', 'B', 'b', 'R', 'r', 'Q', 'q', 'K', 'k', 'X'};

static void put_int(buf_t* b, long v)
{
    char t[32];
    snprintf(t, sizeof t, "%ld", v);
    buf_puts(b, t);
}
static void put_square(buf_t* b, int sq) /* appendSquareToString ParserBits.h:140-144 */
{
    buf_putc(b, (char)('a' + (sq & 7)));
    buf_putc(b, (char)('1' + (sq >> 3)));
}
/* Position::fen Position.cpp:583-603 + Board::fen :345-395 */
static void put_fen(buf_t* b, const pos_t* p)
{
    for (int r = 7; r >= 0; --r) {
        int empty = 0;
        for (int f = 0; f < 8; ++f) {
            uint8_t pc = p->sq[r * 8 + f];
            if (pc == NO_PIECE) {
                ++empty;
            } else {
                if (empty) buf_putc(b, (char)('0' + empty));
                empty = 0;
                buf_putc(b, fen_piece_char[pc]);
            }
        }
        if (empty) buf_putc(b, (char)('0' + empty));
        if (r > 0) buf_putc(b, '/');
    }
    buf_putc(b, ' ');
    buf_putc(b, p->stm == WHITE ? 'w' : 'b');
    buf_putc(b, ' ');
    if (p->cr == 0) {
        buf_putc(b, '-');
    } else {
        if (p->cr & CR_WK) buf_putc(b, 'K');
        if (p->cr & CR_WQ) buf_putc(b, 'Q');
        if (p->cr & CR_BK) buf_putc(b, 'k');
        if (p->cr & CR_BQ) buf_putc(b, 'q');
    }
    buf_putc(b, ' ');
    if (p->ep == SQ_NONE) buf_putc(b, '-');
    else put_square(b, p->ep);
    buf_putc(b, ' ');
    put_int(b, p->rule50);
    buf_putc(b, ' ');
    put_int(b, (uint16_t)((p->ply + 1) / 2));
}
/* uci::moveToUci Uci.cpp:14-39 */
static void put_uci(buf_t* b, const pos_t* p, move_t m)
{
    put_square(b, m.from);
    if (m.type == MT_CASTLE) {
        int is_short = (m.to & 7) == 7;
        put_square(b, (p->stm == WHITE ? 0 : 56) + (is_short ? 6 : 2));
    } else {
        put_square(b, m.to);
        if (m.type == MT_PROMOTION) buf_putc(b, "pnbrqk "[P_TYPE(m.promo)]);
    }
}
/* emitPlainEntry compress_file.cpp:1216-1237 */
static void emit_plain(buf_t* b, const entry_t* e)
{
    buf_puts(b, "fen ");
    put_fen(b, &e->pos);
    buf_puts(b, "\nmove ");
    put_uci(b, &e->pos, e->move);
    buf_puts(b, "\nscore ");
    put_int(b, e->score);
    buf_puts(b, "\nply ");
    put_int(b, e->ply);
    buf_puts(b, "\nresult ");
    put_int(b, e->result);
    buf_puts(b, "\ne\n");
}

/* Board::isSquareAttacked as used by Position::trySet Position.cpp:505 */
static int is_square_attacked(const pos_t* p, int sq, int by)
{
    bb_t occ = pos_color_bb(p, WHITE) | pos_color_bb(p, BLACK);
    bb_t bq = pos_piece_bb(p, PIECE(PT_BISHOP, by)) | pos_piece_bb(p, PIECE(PT_QUEEN, by));
    bb_t rq = pos_piece_bb(p, PIECE(PT_ROOK, by)) | pos_piece_bb(p, PIECE(PT_QUEEN, by));
    if (bishop_attacks(sq, occ) & bq) return 1;
    if (rook_attacks(sq, occ) & rq) return 1;
    if (king_attacks(sq) & pos_piece_bb(p, PIECE(PT_KING, by))) return 1;
    if (knight_attacks(sq) & pos_piece_bb(p, PIECE(PT_KNIGHT, by))) return 1;
    /* pawns of colour `by` that attack sq */
    if (pawn_attacks(BB(sq), !by) & pos_piece_bb(p, PIECE(PT_PAWN, by))) return 1;
    return 0;
}

static int stoi_prefix(const char* s, const char* end, long* out)
{
    /* std::stoi: skips leading whitespace, optional sign, then digits; fails without digits */
    while (s < end && (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r' || *s == '\v' || *s == '\f')) ++s;
    int neg = 0;
    if (s < end && (*s == '+' || *s == '-')) neg = (*s++ == '-');
    if (s >= end || *s < '0' || *s > '9') return 0;
    long v = 0;
    while (s < end && *s >= '0' && *s <= '9') {
        v = v * 10 + (*s++ - '0');
        if (v > 4000000000L) v = 4000000000L;
    }
    *out = neg ? -v : v;
    return 1;
}

/* Position::fromFen Position.cpp:563-568 -> trySet :478-561 (failure leaves the partial state) */
static int pos_from_fen(pos_t* p, const char* s, const char* end)
{
    pos_init(p);
    const char* part[6];
    const char* part_end[6];
    const char* cur = s;
    for (int i = 0; i < 6; ++i) { /* nextPart :481-495 */
        const char* e = cur;
        while (e < end && *e != ' ') ++e;
        part[i] = cur;
        part_end[i] = e;
        cur = e < end ? e + 1 : end;
    }
    { /* Board::trySet Position.h:45-138 */
        int f = 0, r = 7, last_skip = 0;
        for (const char* c = part[0]; c < part_end[0]; ++c) {
            uint8_t piece = NO_PIECE;
            switch (*c) {
            case 'r': piece = PIECE(PT_ROOK, BLACK); break;
            case 'n': piece = PIECE(PT_KNIGHT, BLACK); break;
            case 'b': piece = PIECE(PT_BISHOP, BLACK); break;
            case 'q': piece = PIECE(PT_QUEEN, BLACK); break;
            case 'k': piece = PIECE(PT_KING, BLACK); break;
            case 'p': piece = PIECE(PT_PAWN, BLACK); break;
            case 'R': piece = PIECE(PT_ROOK, WHITE); break;
            case 'N': piece = PIECE(PT_KNIGHT, WHITE); break;
            case 'B': piece = PIECE(PT_BISHOP, WHITE); break;
            case 'Q': piece = PIECE(PT_QUEEN, WHITE); break;
            case 'K': piece = PIECE(PT_KING, WHITE); break;
            case 'P': piece = PIECE(PT_PAWN, WHITE); break;
            case '1': case '2': case '3': case '4': case '5': case '6': case '7': case '8':
                if (last_skip) return 0;
                last_skip = 1;
                f += *c - '0';
                if (f > 8) return 0;
                break;
            case '/':
                last_skip = 0;
                if (f != 8) return 0;
                f = 0;
                --r;
                break;
            default: return 0;
            }
            if (piece != NO_PIECE) {
                last_skip = 0;
                if (f < 0 || f > 7 || r < 0 || r > 7) return 0; /* sq.isOk() */
                p->sq[r * 8 + f] = piece;
                ++f;
            }
        }
        if (f != 8) return 0;
        if (r != 0) return 0;
        /* isValid Position.h:35-41 */
        if (popcnt(pos_piece_bb(p, PIECE(PT_KING, WHITE))) != 1) return 0;
        if (popcnt(pos_piece_bb(p, PIECE(PT_KING, BLACK))) != 1) return 0;
        if ((pos_piece_bb(p, PIECE(PT_PAWN, WHITE)) | pos_piece_bb(p, PIECE(PT_PAWN, BLACK))) & 0xFF000000000000FFull) return 0;
    }
    if (part_end[1] - part[1] == 1 && *part[1] == 'w') p->stm = WHITE;
    else if (part_end[1] - part[1] == 1 && *part[1] == 'b') p->stm = BLACK;
    else return 0;
    if (is_square_attacked(p, pos_king_sq(p, !p->stm), p->stm)) return 0; /* :505 */
    { /* tryParseCastlingRights ParserBits.h:64-97 */
        uint8_t rights = 0;
        if (!(part_end[2] - part[2] == 1 && *part[2] == '-')) {
            for (const char* c = part[2]; c < part_end[2]; ++c) {
                uint8_t add = 0;
                switch (*c) {
                case 'K': add = CR_WK; break;
                case 'Q': add = CR_WQ; break;
                case 'k': add = CR_BK; break;
                case 'q': add = CR_BQ; break;
                }
                if ((rights & add) == add) return 0;
                rights |= add;
            }
        }
        p->cr = rights;
    }
    { /* tryParseEpSquare ParserBits.h:58-62 */
        size_t n = (size_t)(part_end[3] - part[3]);
        if (n == 1 && *part[3] == '-') p->ep = SQ_NONE;
        else if (n == 2 && part[3][0] >= 'a' && part[3][0] <= 'h' && part[3][1] >= '1' && part[3][1] <= '8')
            p->ep = (uint8_t)((part[3][0] - 'a') + 8 * (part[3][1] - '1'));
        else return 0;
    }
    {
        long v = 0;
        if (part_end[4] > part[4]) {
            /* std::stoi(rule50.data()) parses to the end of the *whole* string */
            if (!stoi_prefix(part[4], end, &v)) return -1;
            p->rule50 = (uint8_t)v;
        } else {
            p->rule50 = 0;
        }
        if (part_end[5] > part[5]) {
            if (!stoi_prefix(part[5], end, &v)) return -1;
            p->ply = (uint16_t)(v * 2 - (p->stm == WHITE));
        } else {
            p->ply = 0;
        }
    }
    if (p->ep != SQ_NONE && !is_ep_possible(p, p->ep, p->stm)) p->ep = SQ_NONE; /* :558 */
    return 1;
}

/* uci::uciToMove Uci.cpp:41-75 */
static move_t uci_to_move(const pos_t* p, const char* s, size_t n)
{
    move_t m;
    char c[5] = {0, 0, 0, 0, 0};
    for (size_t i = 0; i < n && i < 5; ++i) c[i] = s[i];
    int from = ((c[0] - 'a') + 8 * (c[1] - '1')) & 0xFF;
    int to = ((c[2] - 'a') + 8 * (c[3] - '1')) & 0xFF;
    m.from = (uint8_t)from;
    m.to = (uint8_t)to;
    m.type = MT_NORMAL;
    m.promo = NO_PIECE;
    if (n == 5) {
        static const char tbl[] = "PpNnBbRrQqKk ";
        const char* it = strchr(tbl, c[4]);
        int pt = it && c[4] ? (int)((it - tbl) / 2) : PT_NONE;
        m.type = MT_PROMOTION;
        m.promo = PIECE(pt, p->stm);
        return m;
    }
    if (from < 64 && to < 64 && P_TYPE(p->sq[from]) == PT_KING && abs((from & 7) - (to & 7)) > 1) {
        int is_short = (to & 7) == 6;
        m.from = (uint8_t)(p->stm == WHITE ? 4 : 60);
        m.to = (uint8_t)((p->stm == WHITE ? 0 : 56) + (is_short ? 7 : 0));
        m.type = MT_CASTLE;
    } else if (p->ep == to) {
        m.type = MT_ENPASSANT; /* quirk Q3: no check that the mover is a pawn */
    }
    return m;
}

/* The `>> key`, `>> std::ws`, getline tokeniser of compressPlain / convertPlainToBin
 * (compress_file.cpp:1264-1296, :1488-1524). Calls `sink` for every "e" record. */
typedef int (*entry_sink)(void* ctx, const entry_t* e);
static int is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }
static int parse_plain(const uint8_t* in, size_t n, entry_sink sink, void* ctx)
{
    const char* s = (const char*)in;
    const char* end = s + n;
    entry_t e;
    memset(&e, 0, sizeof e);
    pos_init(&e.pos); /* TrainingDataEntry e; -> Position() */
    e.move.promo = NO_PIECE;
    const char* move = "";
    size_t move_len = 0;
    for (;;) {
        while (s < end && is_ws(*s)) ++s;
        if (s >= end) break;
        const char* k = s;
        while (s < end && !is_ws(*s)) ++s;
        size_t klen = (size_t)(s - k);
        if (klen == 1 && k[0] == 'e') {
            if (move_len < 4) return ORC_ERR_BAD_TEXT;
            e.move = uci_to_move(&e.pos, move, move_len);
            if (e.move.from >= 64 || e.move.to >= 64) return ORC_ERR_BAD_TEXT;
            int rc = sink(ctx, &e);
            if (rc != ORC_OK) return rc;
            continue;
        }
        while (s < end && is_ws(*s)) ++s;
        const char* v = s;
        while (s < end && *s != '\n') ++s;
        const char* vend = s;
        if (s < end) ++s;
        long num = 0;
        if (klen == 3 && memcmp(k, "fen", 3) == 0) {
            if (pos_from_fen(&e.pos, v, vend) < 0) return ORC_ERR_BAD_TEXT;
        } else if (klen == 4 && memcmp(k, "move", 4) == 0) {
            move = v;
            move_len = (size_t)(vend - v);
        } else if (klen == 5 && memcmp(k, "score", 5) == 0) {
            if (!stoi_prefix(v, vend, &num)) return ORC_ERR_BAD_TEXT;
            e.score = (int16_t)num;
        } else if (klen == 3 && memcmp(k, "ply", 3) == 0) {
            if (!stoi_prefix(v, vend, &num)) return ORC_ERR_BAD_TEXT;
            e.ply = (uint16_t)num;
        } else if (klen == 6 && memcmp(k, "result", 6) == 0) {
            if (!stoi_prefix(v, vend, &num)) return ORC_ERR_BAD_TEXT;
            e.result = (int16_t)num;
        }
    }
    return ORC_OK;
}

/* ---------------------------------------------------------------- the six drivers */

static int sink_writer(void* ctx, const entry_t* e)
{
    writer_add((writer_t*)ctx, e);
    return ORC_OK;
}

/* decompress-side output buffering: the reference only hands `buffer` to the file once it
 * exceeds 1 MiB (compress_file.cpp:1318-1325, :1395-1402, :1448-1455, :1504-1511), so an
 * exception loses the unflushed tail. `committed` tracks what reached the file. */
typedef struct {
    buf_t* out;
    size_t committed;
    size_t pending_start;
    int is_bin;
} emitter_t;
static void emitter_add(emitter_t* em, const entry_t* e)
{
    if (em->is_bin) {
        uint8_t rec[40];
        entry_to_record(e, rec);
        buf_put(em->out, rec, 40);
    } else {
        emit_plain(em->out, e);
    }
    if (em->out->n - em->pending_start > 1024u * 1024u) {
        em->committed = em->out->n;
        em->pending_start = em->out->n;
    }
}
static int sink_emitter(void* ctx, const entry_t* e)
{
    emitter_add((emitter_t*)ctx, e);
    return ORC_OK;
}

int orc_convert(int mode, const uint8_t* in, size_t in_len, uint8_t** out, size_t* out_len)
{
    buf_t ob = {0, 0, 0, 0};
    int rc = ORC_OK;
    if (mode == ORC_BIN_TO_BINPACK || mode == ORC_PLAIN_TO_BINPACK) {
        writer_t w;
        writer_init(&w, &ob);
        if (mode == ORC_BIN_TO_BINPACK) {
            /* compressBin :1338-1374: a short trailing record is dropped (:1360) */
            for (size_t off = 0; off + 40 <= in_len; off += 40) {
                entry_t e;
                rc = record_to_entry(in + off, &e);
                if (rc != ORC_OK) break;
                writer_add(&w, &e);
            }
        } else {
            rc = parse_plain(in, in_len, sink_writer, &w);
        }
        writer_finish(&w); /* the destructor runs during unwinding too */
    } else if (mode == ORC_BINPACK_TO_BIN || mode == ORC_BINPACK_TO_PLAIN) {
        reader_t r;
        reader_init(&r, in, in_len);
        emitter_t em = {&ob, 0, 0, mode == ORC_BINPACK_TO_BIN};
        rc = r.err;
        while (rc == ORC_OK && !r.is_end) {
            entry_t e;
            rc = reader_next(&r, &e);
            if (rc != ORC_OK) break; /* the exception leaves next() before the entry is emitted */
            emitter_add(&em, &e);
        }
        if (rc != ORC_OK) ob.n = em.committed;
    } else if (mode == ORC_BIN_TO_PLAIN) {
        emitter_t em = {&ob, 0, 0, 0};
        for (size_t off = 0; off + 40 <= in_len; off += 40) {
            entry_t e;
            rc = record_to_entry(in + off, &e);
            if (rc != ORC_OK) break;
            emitter_add(&em, &e);
        }
        if (rc != ORC_OK) ob.n = em.committed;
    } else if (mode == ORC_PLAIN_TO_BIN) {
        emitter_t em = {&ob, 0, 0, 1};
        rc = parse_plain(in, in_len, sink_emitter, &em);
        if (rc != ORC_OK) ob.n = em.committed;
    } else {
        rc = ORC_ERR_BAD_MODE;
    }
    if (ob.oom) rc = ORC_ERR_NOMEM;
    if (!ob.p) ob.p = (uint8_t*)malloc(1);
    *out = ob.p;
    *out_len = ob.n;
    return rc;
}

void orc_free(void* p) { free(p); }

const char* orc_strerror(int code)
{
    switch (code) {
    case ORC_OK: return "ok";
    case ORC_ERR_BAD_MAGIC: return "Invalid binpack file or chunk.";
    case ORC_ERR_CHUNK_TOO_LARGE: return "Chunks size larger than supported. Malformed file?";
    case ORC_ERR_BAD_SFEN: return "Improperly encoded bin sfen";
    case ORC_ERR_TRUNCATED: return "truncated binpack chunk";
    case ORC_ERR_NOMEM: return "out of memory";
    case ORC_ERR_BAD_MODE: return "bad conversion mode";
    case ORC_ERR_BAD_TEXT: return "malformed .plain input";
    default: return "unknown error";
    }
}

int orc_sfen_to_fen(const uint8_t sfen[32], char* fen_out, size_t cap)
{
    uint8_t rec[40];
    memcpy(rec, sfen, 32);
    memset(rec + 32, 0, 8);
    pos_t p;
    int rc = sfen_unpack(rec, &p);
    if (rc != ORC_OK) return rc;
    buf_t b = {0, 0, 0, 0};
    put_fen(&b, &p);
    buf_putc(&b, 0);
    if (b.oom || b.n > cap) {
        free(b.p);
        return ORC_ERR_NOMEM;
    }
    memcpy(fen_out, b.p, b.n);
    free(b.p);
    return ORC_OK;
}

int orc_fen_to_sfen(const char* fen, uint8_t sfen_out[32])
{
    pos_t p;
    if (pos_from_fen(&p, fen, fen + strlen(fen)) < 0) return ORC_ERR_BAD_TEXT;
    sfen_pack(&p, sfen_out);
    return ORC_OK;
}

int orc_is_continuation(const uint8_t a[40], const uint8_t b[40])
{
    entry_t ea, eb;
    int rc = record_to_entry(a, &ea);
    if (rc != ORC_OK) return rc;
    rc = record_to_entry(b, &eb);
    if (rc != ORC_OK) return rc;
    return is_continuation(&ea, &eb);
}

int64_t orc_binpack_count(const uint8_t* in, size_t in_len)
{
    reader_t r;
    reader_init(&r, in, in_len);
    if (r.err) return r.err;
    int64_t n = 0;
    while (!r.is_end) {
        entry_t e;
        int rc = reader_next(&r, &e);
        if (rc != ORC_OK) return rc;
        ++n;
    }
    return n;
}

/* HalfKP rows: see oracle.h (published trainer formula, not reference code) */
int orc_bin_to_halfkp(const uint8_t* bin, size_t n, int32_t* white, int32_t* black, uint8_t* meta, size_t* bad_index)
{
    for (size_t r = 0; r < n; ++r) {
        const uint8_t* rec = bin + r * 40;
        pos_t pos;
        int rc = sfen_unpack(rec, &pos);
        if (rc != ORC_OK) {
            if (bad_index) *bad_index = r;
            return rc;
        }
        int ksq[2] = {0, 0}, have[2] = {0, 0};
        for (int sq = 0; sq < 64; ++sq) {
            uint8_t pc = pos.sq[sq];
            if (pc != NO_PIECE && P_TYPE(pc) == PT_KING && !have[P_COLOR(pc)]) {
                have[P_COLOR(pc)] = 1;
                ksq[P_COLOR(pc)] = sq;
            }
        }
        int32_t* w = white + r * 32;
        int32_t* b = black + r * 32;
        int cnt = 0;
        /* row order: by kind as white sees it (2 * type + colour), then by square */
        for (int kind = 0; kind < 10; ++kind) {
            for (int sq = 0; sq < 64 && cnt < 32; ++sq) {
                uint8_t pc = pos.sq[sq];
                if (pc == NO_PIECE || P_TYPE(pc) == PT_KING) continue;
                int t = P_TYPE(pc), c = P_COLOR(pc);
                if (2 * t + c != kind) continue;
                w[cnt] = 1 + sq + 64 * (2 * t + (c != WHITE)) + 641 * ksq[WHITE];
                b[cnt] = 1 + (sq ^ 63) + 64 * (2 * t + (c != BLACK)) + 641 * (ksq[BLACK] ^ 63);
                ++cnt;
            }
        }
        for (int k = cnt; k < 32; ++k) w[k] = b[k] = -1;
        uint8_t* m = meta + r * 8;
        m[0] = rec[32]; m[1] = rec[33];           /* score */
        m[2] = rec[36]; m[3] = rec[37];           /* gamePly */
        m[4] = rec[38];                           /* game_result */
        m[5] = pos.stm;
        m[6] = (uint8_t)cnt;
        m[7] = 0;
    }
    return ORC_OK;
}
