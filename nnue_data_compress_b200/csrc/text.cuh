// text.cuh -- device-side .plain text: FEN / UCI emit (emitPlainEntry, compress_file.cpp:1216-1237)
// and parse (compressPlain's tokeniser :1264-1296, Position::trySet Position.cpp:478-561,
// uci::uciToMove Uci.cpp:41-75). Emitters are templated on a sink so that the size pass and the
// write pass run the same code.
#pragma once
#include "chess.cuh"

namespace nnp {

struct CountSink {
    u32 n = 0;
    __device__ __forceinline__ void put(char) { ++n; }
};
struct WriteSink {
    unsigned char* p;
    __device__ __forceinline__ void put(char c) { *p++ = (unsigned char)c; }
};

// Text sink for the emitting kernels. A thread's text is a run of ~100 bytes per record at an
// arbitrary byte offset; storing it byte by byte costs one 32-byte sector operation per character
// and warp lane (measured: 1 047 M sector stores for 1 047 MB of text). Characters are therefore
// gathered in a 64-byte window in shared memory that is congruent to the destination address
// modulo 16 and leave the SM as 16-byte stores; only the ragged first and last groups of a thread's
// text go out bytewise.
struct BufferedSink {
    unsigned char* dst;  // global address of window byte 0 (16-byte aligned)
    uint4* win;          // group g of the window is win[g * stride] (shared memory, one column per thread)
    int stride;
    int n;               // next byte of the window
    int lo;              // first byte of the window that belongs to this thread (first window only)
    u32 word;            // the 4-byte group being filled
    __device__ __forceinline__ void init(unsigned char* out, uint4* column, int column_stride)
    {
        lo = n = (int)(reinterpret_cast<uintptr_t>(out) & 15);
        dst = out - lo;
        win = column;
        stride = column_stride;
        word = 0;
    }
    __device__ __forceinline__ void store_word(int i, u32 w) { reinterpret_cast<u32*>(win + (i >> 2) * stride)[i & 3] = w; }
    __device__ __forceinline__ void flush()
    {
        if (n & 3) store_word(n >> 2, word);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int a = max(lo, 16 * g), b = min(n, 16 * g + 16);
            if (a >= b) continue;
            if (b - a == 16) {
                *reinterpret_cast<uint4*>(dst + 16 * g) = win[g * stride];
            } else {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(win + g * stride);
                for (int i = a; i < b; ++i) dst[i] = src[i - 16 * g];
            }
        }
        dst += 64;
        n = 0;
        lo = 0;
        word = 0;
    }
    __device__ __forceinline__ void put(char c)
    {
        word |= (u32)(unsigned char)c << (8 * (n & 3));
        if ((n & 3) == 3) {
            store_word(n >> 2, word);
            word = 0;
        }
        if (++n == 64) flush();
    }
    __device__ __forceinline__ void finish() { if (n > lo) flush(); }
};

template <typename S>
__device__ __forceinline__ void put_str(S& s, const char* lit)
{
    for (int i = 0; lit[i]; ++i) s.put(lit[i]);
}
template <typename S>
__device__ __forceinline__ void put_uint(S& s, u32 v)  // std::to_string
{
    char buf[10];
    int n = 0;
    do {
        buf[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) s.put(buf[--n]);
}
template <typename S>
__device__ __forceinline__ void put_int(S& s, int v)
{
    if (v < 0) {
        s.put('-');
        put_uint(s, (u32)(-(long long)v));
    } else {
        put_uint(s, (u32)v);
    }
}
template <typename S>
__device__ __forceinline__ void put_square(S& s, int sq)  // appendSquareToString ParserBits.h:140-144
{
    s.put((char)('a' + (sq & 7)));
    s.put((char)('1' + (sq >> 3)));
}

// Position::fen (Position.cpp:583-603) + Board::fen (:345-395)
template <typename S>
__device__ __forceinline__ void put_fen(S& s, const Pos& p)
{
    const u64 all = pos_all(p);
    for (int r = 7; r >= 0; --r) {
        int empty = 0;
        for (int f = 0; f < 8; ++f) {
            const int sq = r * 8 + f;
            if (!((all >> sq) & 1)) {
                ++empty;
            } else {
                if (empty) s.put((char)('0' + empty));
                empty = 0;
                const int t = (int)((p.t0 >> sq) & 1) | ((int)((p.t1 >> sq) & 1) << 1) | ((int)((p.t2 >> sq) & 1) << 2);
                const char up = "PNBRQK??"[t];
                s.put(((p.occ[1] >> sq) & 1) ? (char)(up + 32) : up);
            }
        }
        if (empty) s.put((char)('0' + empty));
        if (r > 0) s.put('/');
    }
    s.put(' ');
    s.put(p.stm == WHITE ? 'w' : 'b');
    s.put(' ');
    if (p.cr == 0) {
        s.put('-');
    } else {
        if (p.cr & CR_WK) s.put('K');
        if (p.cr & CR_WQ) s.put('Q');
        if (p.cr & CR_BK) s.put('k');
        if (p.cr & CR_BQ) s.put('q');
    }
    s.put(' ');
    if (p.ep == SQ_NONE) s.put('-');
    else put_square(s, p.ep);
    s.put(' ');
    put_uint(s, (u32)p.rule50 & 0xFF);
    s.put(' ');
    put_uint(s, (u32)(((p.ply & 0xFFFF) + 1) >> 1));  // halfMove() Position.h:933-936
}

// uci::moveToUci (Uci.cpp:14-39)
template <typename S>
__device__ __forceinline__ void put_uci(S& s, const Pos& p, const Move& m)
{
    put_square(s, m.from);
    if (m.type == MT_CASTLE) {
        const bool is_short = (m.to & 7) == 7;
        put_square(s, (p.stm == WHITE ? 0 : 56) + (is_short ? 6 : 2));
    } else {
        put_square(s, m.to);
        if (m.type == MT_PROMOTION) s.put("pnbrqk "[(m.promo >> 1) > 6 ? 6 : (m.promo >> 1)]);
    }
}

// emitPlainEntry (compress_file.cpp:1216-1237)
template <typename S>
__device__ __forceinline__ void put_plain_entry(S& s, const Pos& p, const Move& m, int score, int ply, int result)
{
    put_str(s, "fen ");
    put_fen(s, p);
    put_str(s, "\nmove ");
    put_uci(s, p, m);
    put_str(s, "\nscore ");
    put_int(s, score);
    put_str(s, "\nply ");
    put_uint(s, (u32)ply & 0xFFFF);
    put_str(s, "\nresult ");
    put_int(s, result);
    put_str(s, "\ne\n");
}

// ---------------------------------------------------------------- parsing

__device__ __forceinline__ bool is_ws(unsigned char c)
{
    return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f';
}

// std::stoi on [s, end): leading whitespace, optional sign, digits; false without digits
__device__ __forceinline__ bool parse_int(const unsigned char* s, const unsigned char* end, long long& out)
{
    while (s < end && is_ws(*s)) ++s;
    bool neg = false;
    if (s < end && (*s == '+' || *s == '-')) neg = (*s++ == '-');
    if (s >= end || *s < '0' || *s > '9') return false;
    long long v = 0;
    while (s < end && *s >= '0' && *s <= '9') {
        v = v * 10 + (*s++ - '0');
        if (v > 4000000000LL) v = 4000000000LL;
    }
    out = neg ? -v : v;
    return true;
}

// Position::fromFen (Position.cpp:563-568) -> trySet (:478-561). A FEN the reference rejects
// leaves the partially set position behind exactly as trySet's early returns do. Returns false
// only where the reference would crash (std::stoi on a non-number).
__device__ __forceinline__ bool parse_fen(const unsigned char* s, const unsigned char* end, Pos& p)
{
    pos_clear(p);
    const unsigned char* part[6];
    const unsigned char* part_end[6];
    const unsigned char* cur = s;
#pragma unroll
    for (int i = 0; i < 6; ++i) {  // nextPart (:481-495): split on single spaces
        const unsigned char* e = cur;
        while (e < end && *e != ' ') ++e;
        part[i] = cur;
        part_end[i] = e;
        cur = e < end ? e + 1 : end;
    }
    {  // Board::trySet (Position.h:45-138)
        int f = 0, r = 7;
        bool last_skip = false;
        for (const unsigned char* c = part[0]; c < part_end[0]; ++c) {
            int piece = NO_PIECE;
            switch (*c) {
            case 'r': piece = (PT_ROOK << 1) | BLACK; break;
            case 'n': piece = (PT_KNIGHT << 1) | BLACK; break;
            case 'b': piece = (PT_BISHOP << 1) | BLACK; break;
            case 'q': piece = (PT_QUEEN << 1) | BLACK; break;
            case 'k': piece = (PT_KING << 1) | BLACK; break;
            case 'p': piece = (PT_PAWN << 1) | BLACK; break;
            case 'R': piece = (PT_ROOK << 1) | WHITE; break;
            case 'N': piece = (PT_KNIGHT << 1) | WHITE; break;
            case 'B': piece = (PT_BISHOP << 1) | WHITE; break;
            case 'Q': piece = (PT_QUEEN << 1) | WHITE; break;
            case 'K': piece = (PT_KING << 1) | WHITE; break;
            case 'P': piece = (PT_PAWN << 1) | WHITE; break;
            case '1': case '2': case '3': case '4': case '5': case '6': case '7': case '8':
                if (last_skip) return true;
                last_skip = true;
                f += *c - '0';
                if (f > 8) return true;
                break;
            case '/':
                last_skip = false;
                if (f != 8) return true;
                f = 0;
                --r;
                break;
            default: return true;
            }
            if (piece != NO_PIECE) {
                last_skip = false;
                if (f < 0 || f > 7 || r < 0 || r > 7) return true;
                pos_put(p, r * 8 + f, piece);
                ++f;
            }
        }
        if (f != 8 || r != 0) return true;
        const u64 kings = pos_type_bb(p, PT_KING);  // isValid (Position.h:35-41)
        if (popc64(kings & p.occ[0]) != 1 || popc64(kings & p.occ[1]) != 1) return true;
        if (pos_type_bb(p, PT_PAWN) & 0xFF000000000000FFull) return true;
    }
    if (part_end[1] - part[1] == 1 && *part[1] == 'w') p.stm = WHITE;
    else if (part_end[1] - part[1] == 1 && *part[1] == 'b') p.stm = BLACK;
    else return true;
    {
        const u64 k = pos_type_bb(p, PT_KING) & pos_occ(p, p.stm ^ 1);
        if (square_attacked(p, lsb64(k), p.stm, pos_all(p))) return true;  // :505
    }
    {  // tryParseCastlingRights (ParserBits.h:64-97)
        int rights = 0;
        if (!(part_end[2] - part[2] == 1 && *part[2] == '-')) {
            for (const unsigned char* c = part[2]; c < part_end[2]; ++c) {
                int add = 0;
                if (*c == 'K') add = CR_WK;
                else if (*c == 'Q') add = CR_WQ;
                else if (*c == 'k') add = CR_BK;
                else if (*c == 'q') add = CR_BQ;
                if ((rights & add) == add) return true;  // duplicate or unknown character
                rights |= add;
            }
        }
        p.cr = rights;
    }
    {  // tryParseEpSquare (ParserBits.h:58-62)
        const long n = (long)(part_end[3] - part[3]);
        if (n == 1 && *part[3] == '-') p.ep = SQ_NONE;
        else if (n == 2 && part[3][0] >= 'a' && part[3][0] <= 'h' && part[3][1] >= '1' && part[3][1] <= '8')
            p.ep = (part[3][0] - 'a') + 8 * (part[3][1] - '1');
        else return true;
    }
    long long v = 0;
    if (part_end[4] > part[4]) {
        if (!parse_int(part[4], end, v)) return false;  // std::stoi(rule50.data())
        p.rule50 = (int)(v & 0xFF);
    } else {
        p.rule50 = 0;
    }
    if (part_end[5] > part[5]) {
        if (!parse_int(part[5], end, v)) return false;
        p.ply = (int)((v * 2 - (p.stm == WHITE)) & 0xFFFF);
    } else {
        p.ply = 0;
    }
    if (p.ep != SQ_NONE && !ep_possible(p, p.ep, p.stm)) p.ep = SQ_NONE;  // :558
    return true;
}

// uci::uciToMove (Uci.cpp:41-75). Returns false where the reference would dereference an empty
// optional or index outside the board.
__device__ __forceinline__ bool parse_uci(const Pos& p, const unsigned char* s, int n, Move& m)
{
    if (n < 4) return false;
    const int from = (s[0] - 'a') + 8 * (s[1] - '1');
    const int to = (s[2] - 'a') + 8 * (s[3] - '1');
    if (from < 0 || from > 63 || to < 0 || to > 63) return false;
    m.from = from;
    m.to = to;
    m.type = MT_NORMAL;
    m.promo = NO_PIECE;
    if (n == 5) {
        int pt;
        switch (s[4]) {
        case 'n': case 'N': pt = PT_KNIGHT; break;
        case 'b': case 'B': pt = PT_BISHOP; break;
        case 'r': case 'R': pt = PT_ROOK; break;
        case 'q': case 'Q': pt = PT_QUEEN; break;
        default: return false;
        }
        m.type = MT_PROMOTION;
        m.promo = (pt << 1) | p.stm;
        return true;
    }
    const int d = (from & 7) - (to & 7);
    if ((pos_piece_at(p, from) >> 1) == PT_KING && (d > 1 || d < -1)) {
        const bool is_short = (to & 7) == 6;
        m.from = p.stm == WHITE ? 4 : 60;  // Move::castle Chess.h:1029-1040
        m.to = (p.stm == WHITE ? 0 : 56) + (is_short ? 7 : 0);
        m.type = MT_CASTLE;
    } else if (p.ep == to) {
        m.type = MT_ENPASSANT;  // quirk Q3: any mover landing on the ep square
    }
    return true;
}

}  // namespace nnp
