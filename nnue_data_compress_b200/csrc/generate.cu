// generate.cu -- device-side synthetic input: random legal-move games written as .bin records.
//
// This is the workload generator of SURVEY.md 8(d) (the recipe the reference-linked
// oracle/gen_ref.cpp follows with the reference's own movegen), re-done on the GPU so that
// 100M-position inputs exist in HBM without a host round trip: one thread plays one game.
// It is bench/test input, not a reference driver; the records it emits go through the same
// sfen_encode / pos_do_move as the decompressor, i.e. they look exactly like reference-made
// data (ep squares set by Position::doMove's own rule, rule50 and fullmove from the game).
#include "../../include/nnuepack.h"
#include "common.cuh"
#include "kernels.h"

namespace nnp {

struct Rng {
    u64 s;
    __device__ __forceinline__ u64 next()
    {  // splitmix64
        u64 z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};

__device__ __forceinline__ u64 destinations(const Pos& p, int from, int pt, u64 ours, u64 theirs)
{
    if (pt == PT_PAWN) return pawn_destinations(p, from, ours, theirs);
    return piece_attacks(pt, from, ours | theirs) & ~ours;
}

// castling by the rules of chess: right present, rook at home, path empty, king not in check
// and not passing through or landing on an attacked square
__device__ __forceinline__ int castle_moves(const Pos& p, Move* out /* [2] */)
{
    const int stm = p.stm;
    const int base = stm == WHITE ? 0 : 56;
    const u64 occ = pos_all(p);
    int n = 0;
    const int kpc = (PT_KING << 1) | stm, rpc = (PT_ROOK << 1) | stm;
    if (pos_piece_at(p, base + 4) != kpc) return 0;
    const int short_right = stm == WHITE ? CR_WK : CR_BK, long_right = stm == WHITE ? CR_WQ : CR_BQ;
    if ((p.cr & (short_right | long_right)) == 0) return 0;
    if (square_attacked(p, base + 4, stm ^ 1, occ)) return 0;
    if ((p.cr & short_right) && pos_piece_at(p, base + 7) == rpc && !(occ & (0x60ull << base)) &&
        !square_attacked(p, base + 5, stm ^ 1, occ) && !square_attacked(p, base + 6, stm ^ 1, occ)) {
        out[n].from = base + 4; out[n].to = base + 7; out[n].type = MT_CASTLE; out[n].promo = NO_PIECE;
        ++n;
    }
    if ((p.cr & long_right) && pos_piece_at(p, base + 0) == rpc && !(occ & (0x0Eull << base)) &&
        !square_attacked(p, base + 3, stm ^ 1, occ) && !square_attacked(p, base + 2, stm ^ 1, occ)) {
        out[n].from = base + 4; out[n].to = base + 0; out[n].type = MT_CASTLE; out[n].promo = NO_PIECE;
        ++n;
    }
    return n;
}

__device__ __forceinline__ bool leaves_king_safe(const Pos& p, const Move& m)
{
    Pos q = p;
    board_do_move(q, m, pos_piece_at(q, m.from));
    const u64 kings = pos_type_bb(q, PT_KING) & pos_occ(q, p.stm);
    if (!kings) return false;
    return !square_attacked(q, lsb64(kings), p.stm ^ 1, pos_all(q));
}

// index-th pseudo-legal non-castling move in (piece, destination, promotion) order; returns
// the number of such moves when index is out of range (index = ~0u counts them)
__device__ __forceinline__ u32 nth_pseudo_move(const Pos& p, u32 index, Move& out)
{
    const int stm = p.stm;
    const u64 ours = pos_occ(p, stm), theirs = pos_occ(p, stm ^ 1);
    const int promo_rank = stm == WHITE ? 6 : 1;
    u32 seen = 0;
    u64 b = ours;
    while (b) {
        const int from = lsb64(b);
        b &= b - 1;
        const int pt = pos_piece_at(p, from) >> 1;
        const u64 dest = destinations(p, from, pt, ours, theirs);
        const bool promo = pt == PT_PAWN && (from >> 3) == promo_rank;
        const u32 cnt = (u32)popc64(dest) * (promo ? 4u : 1u);
        if (index - seen < cnt) {
            const u32 k = index - seen;
            out.from = from;
            out.to = nth_set_bit(dest, promo ? (k >> 2) : k);
            out.type = MT_NORMAL;
            out.promo = NO_PIECE;
            if (promo) {
                out.type = MT_PROMOTION;
                out.promo = ((PT_KNIGHT + (int)(k & 3)) << 1) | stm;
            } else if (pt == PT_PAWN && out.to == p.ep) {
                out.type = MT_ENPASSANT;
            }
            return seen + cnt;
        }
        seen += cnt;
    }
    return seen;
}

// uniformly random legal move; false when there is none (mate / stalemate)
__device__ __forceinline__ bool random_legal_move(const Pos& p, Rng& rng, Move& out)
{
    Move castles[2];
    const u32 nc = (u32)castle_moves(p, castles);
    Move tmp;
    const u32 np = nth_pseudo_move(p, 0xFFFFFFFFu, tmp);
    const u32 total = np + nc;
    if (total == 0) return false;
    // rejection sampling over pseudo-legal moves is uniform over the legal ones
    for (int attempt = 0; attempt < 48; ++attempt) {
        const u32 r = (u32)(rng.next() % total);
        if (r >= np) { out = castles[r - np]; return true; }
        nth_pseudo_move(p, r, out);
        if (leaves_king_safe(p, out)) return true;
    }
    // rarely reached (almost every pseudo-legal move illegal): enumerate
    u32 legal = nc;
    for (u32 i = 0; i < np; ++i) {
        nth_pseudo_move(p, i, tmp);
        legal += leaves_king_safe(p, tmp) ? 1u : 0u;
    }
    if (legal == 0) return false;
    u32 r = (u32)(rng.next() % legal);
    if (r < nc) { out = castles[r]; return true; }
    r -= nc;
    for (u32 i = 0; i < np; ++i) {
        nth_pseudo_move(p, i, tmp);
        if (leaves_king_safe(p, tmp)) {
            if (r == 0) { out = tmp; return true; }
            --r;
        }
    }
    return false;
}

__device__ __forceinline__ void start_position(Pos& p)
{
    pos_clear(p);
    p.occ[0] = 0x000000000000FFFFull;
    p.occ[1] = 0xFFFF000000000000ull;
    // types: rank 1/8 = R N B Q K B N R (3 1 2 4 5 2 1 3), pawns = 0
    const u64 back = 0xFF000000000000FFull;
    const u64 t0 = 0xD3ull, t1 = 0xA5ull, t2 = 0x18ull;  // bit planes of 3,1,2,4,5,2,1,3 over files a..h
    p.t0 = (t0 | (t0 << 56)) & back;
    p.t1 = (t1 | (t1 << 56)) & back;
    p.t2 = (t2 | (t2 << 56)) & back;
    p.cr = CR_ALL;
}

// WRITE == false: game_len[g] = number of positions game g yields (<= max_plies)
// WRITE == true : writes them at record game_base[g] .. , clipped to n_positions
template <bool WRITE>
__global__ void __launch_bounds__(128)
k_play_games(u64 n_games, u32 max_plies, u64 seed, u32* __restrict__ game_len, const u64* __restrict__ game_base,
             unsigned char* __restrict__ out, u64 n_positions)
{
    const u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_games) return;
    Rng rng;
    rng.s = seed * 0xD1342543DE82EF95ull + g * 0x2545F4914F6CDD1Dull + 0x1234567ull;
    (void)rng.next();
    Pos p;
    start_position(p);
    int score = (int)(rng.next() % 200) - 100;
    int result = (int)(rng.next() % 3) - 1;
    int ply = 0;
    u64 rec = WRITE ? game_base[g] : 0;
    u32 len = 0;
    for (u32 i = 0; i < max_plies; ++i) {
        if (WRITE && rec >= n_positions) break;
        if (p.rule50 >= 63) break;
        Move m;
        if (!random_legal_move(p, rng, m)) break;
        if (WRITE) {
            u32 w[10];
            sfen_encode(p, w);
            w[8] = ((u32)score & 0xFFFFu) | (move_to_sfmove(m) << 16);
            w[9] = ((u32)ply & 0xFFFFu) | (((u32)result & 0xFFu) << 16) | 0xFF000000u;
            uint2* d = reinterpret_cast<uint2*>(out + rec * 40);
#pragma unroll
            for (int k = 0; k < 5; ++k) d[k] = make_uint2(w[2 * k], w[2 * k + 1]);
            ++rec;
        }
        ++len;
        pos_do_move(p, m);
        ply += 1;
        result = -result;
        int s = -score + (int)(rng.next() % 61) - 30;
        s = s > 3000 ? 3000 : s;
        s = s < -3000 ? -3000 : s;
        score = s;
    }
    if (!WRITE) game_len[g] = len;
}

void launch_play_games(bool write, u64 n_games, u32 max_plies, u64 seed, u32* game_len, const u64* game_base, void* out,
                       u64 n_positions, cudaStream_t s)
{
    if (n_games == 0) return;
    const unsigned blocks = (unsigned)((n_games + 127) / 128);
    if (write)
        k_play_games<true><<<blocks, 128, 0, s>>>(n_games, max_plies, seed, game_len, game_base, (unsigned char*)out, n_positions);
    else
        k_play_games<false><<<blocks, 128, 0, s>>>(n_games, max_plies, seed, game_len, game_base, (unsigned char*)out, n_positions);
}

}  // namespace nnp
