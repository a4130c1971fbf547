// sim.cpp -- TEST INFRASTRUCTURE ONLY (see host_sim.h). Runs device functions of chess.cuh /
// stream.cuh on the CPU over a .bin record array so that the CPU suite can check them against the
// oracle and the golden vectors.
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "host_sim.h"
#include "../../nnue_data_compress_b200/csrc/chess.cuh"
#include "../../nnue_data_compress_b200/csrc/stream.cuh"
#include "../../nnue_data_compress_b200/csrc/walk.cuh"
#include "../../nnue_data_compress_b200/csrc/chain.cuh"
#include "../../nnue_data_compress_b200/csrc/heads.cuh"
#include "../../nnue_data_compress_b200/csrc/halfkp.cuh"
#include <algorithm>
#include <vector>

using namespace nnp;

namespace {
struct Rec {
    u32 w[10];
};
Rec load(const unsigned char* bin, size_t i)
{
    Rec r;
    std::memcpy(r.w, bin + 40 * i, 40);
    return r;
}
// the lookup tables the kernels keep in shared memory, filled with the same per-entry function
const StepTables* tables()
{
    static StepTables* T = nullptr;
    if (!T) {
        T = new StepTables();
        for (int i = 0; i < STEP_TABLE_ROWS; ++i) step_tables_entry(*T, i);
    }
    return T;
}
// every form of the splice on one move: the fused two-edit pass with and without the tables and the
// general edit-by-edit form must accept the same moves and give the same stream
bool splice_all_forms(const u32 (&W)[8], const Pos& p, const Move& m, u32 (&out)[8], bool& agree)
{
    u32 a[8], b[8], c[8];
    std::memcpy(a, W, 32);
    std::memcpy(b, W, 32);
    std::memcpy(c, W, 32);
    const bool oa = stream_apply_move(a, p, m, -1, tables());
    const bool ob = stream_apply_move(b, p, m, -1, nullptr);
    const bool oc = stream_apply_move_generic(c, p, m);
    agree = oa == ob && ob == oc;
    if (oa && agree) agree = std::memcmp(a, b, 28) == 0 && std::memcmp(a, c, 28) == 0;  // the seven board words
    std::memcpy(out, a, 32);
    return oa;
}
}  // namespace

extern "C" {

// For every record: decode, re-encode from scratch with sfen_encode and with stream_from_pos +
// stream_with_tail; count disagreements between the two encoders (-> *enc_mismatch).
// For every consecutive pair (i, i+1): apply record i's move to record i's position with
// stream_apply_move (+ pos_do_move) and compare with sfen_encode of the resulting position
// (-> *splice_mismatch); *spliced counts the moves inside the splice domain.
int sim_stream_check(const unsigned char* bin, size_t n, uint64_t* enc_mismatch, uint64_t* splice_mismatch,
                     uint64_t* spliced, uint64_t* first_bad)
{
    *enc_mismatch = *splice_mismatch = *spliced = 0;
    *first_bad = ~0ull;
    u32 col[8];
    for (size_t i = 0; i < n; ++i) {
        const Rec r = load(bin, i);
        Pos p;
        pos_clear(p);
        if (!sfen_decode([&](int j) { return r.w[j]; }, p)) continue;
        u32 a[8], W[8], b[8];
        sfen_encode(p, a);
        const int end = stream_from_pos(p, col, 1, W);
        stream_with_tail(W, end, p, b);
        if (std::memcmp(a, b, 32) != 0) {
            ++*enc_mismatch;
            if (*first_bad == ~0ull) *first_bad = i;
        }
        const Move m = sfmove_to_move(r.w[8] >> 16);
        u32 W2[8];
        bool agree;
        const bool ok = splice_all_forms(W, p, m, W2, agree);
        if (!agree) {
            ++*splice_mismatch;
            if (*first_bad == ~0ull) *first_bad = i;
        }
        if (!ok) continue;
        ++*spliced;
        Pos q = p, q2 = p;
        pos_do_move(q, m);
        pos_do_move(q2, m, -1, tables());
        if (!pos_equal(q, q2) || q.rule50 != q2.rule50 || q.ply != q2.ply) ++*splice_mismatch;
        u32 c[8], d[8];
        sfen_encode(q, c);
        stream_with_tail(W2, stream_board_end(q), q, d);
        if (std::memcmp(c, d, 32) != 0) {
            ++*splice_mismatch;
            if (*first_bad == ~0ull) *first_bad = i;
        }
    }
    return 0;
}

// The same splice check with pseudo-random (mostly illegal) moves of all four types on every
// position: whenever stream_apply_move accepts a move, the result must equal the from-scratch
// encoding of pos_do_move's result. Returns the number of accepted moves; *mismatch counts errors.
uint64_t sim_stream_fuzz(const unsigned char* bin, size_t n, int moves_per_pos, uint64_t seed, uint64_t* mismatch)
{
    *mismatch = 0;
    uint64_t accepted = 0, x = seed * 2862933555777941757ull + 3037000493ull;
    u32 col[8];
    for (size_t i = 0; i < n; ++i) {
        const Rec r = load(bin, i);
        Pos p;
        pos_clear(p);
        if (!sfen_decode([&](int j) { return r.w[j]; }, p)) continue;
        u32 W[8];
        stream_from_pos(p, col, 1, W);
        for (int k = 0; k < moves_per_pos; ++k) {
            x = x * 6364136223846793005ull + 1442695040888963407ull;
            const u32 v = (u32)(x >> 33);
            Move m;
            m.from = v & 63;
            m.to = (v >> 6) & 63;
            m.type = (v >> 12) & 3;
            m.promo = NO_PIECE;
            // bias towards moves that start on a piece and towards castling-like geometry
            if ((v >> 20) & 1) {
                u64 all = pos_all(p);
                m.from = nth_set_bit(all, (v >> 21) % (u32)popc64(all));
            }
            if (m.type == MT_CASTLE && ((v >> 14) & 1)) {
                m.from = ((v >> 15) & 1) ? 4 : 60;
                m.to = (m.from & 56) + (((v >> 16) & 1) ? 7 : 0);
            }
            if (m.type == MT_PROMOTION) m.promo = ((PT_KNIGHT + (int)((v >> 17) & 3)) << 1) | (int)((v >> 19) & 1);
            u32 W2[8];
            bool agree;
            const bool ok = splice_all_forms(W, p, m, W2, agree);
            if (!agree) ++*mismatch;
            if (!ok) continue;
            ++accepted;
            Pos q = p, q2 = p;
            pos_do_move(q, m);
            pos_do_move(q2, m, -1, tables());
            if (!pos_equal(q, q2) || q.rule50 != q2.rule50) ++*mismatch;
            u32 c[8], d[8];
            sfen_encode(q, c);
            stream_with_tail(W2, stream_board_end(q), q, d);
            if (std::memcmp(c, d, 32) != 0) ++*mismatch;
        }
    }
    return accepted;
}

// The chain-walking compressor step (walk.cuh) against the record-parallel one (link.cuh) on a
// whole .bin: runs of `run` records, parked heads worked off in rounds exactly as
// k_walk_link_encode does. Returns the number of records whose code or stem differs; *parked and
// *errors report how many heads were queued and how many decode errors were seen.
uint64_t sim_walk_check(const unsigned char* bin, size_t n, int run, uint64_t* parked, uint64_t* first_error)
{
    std::vector<u32> codes_a(n, 0xDEADBEEFu), codes_b(n, 0xDEADBEEFu);
    std::vector<u32> stems_a(n * 8 + 8, 0u), stems_b(n * 8 + 8, 0u);
    uint64_t err_a = ~0ull, err_b = ~0ull;
    *parked = 0;
    // (a) record-parallel reference form
    for (size_t i = 0; i < n; ++i) {
        const Rec r = load(bin, i);
        Pos cur, prev;
        pos_clear(cur);
        pos_clear(prev);
        const bool ok = sfen_decode([&](int j) { return r.w[j]; }, cur);
        if (!ok && i < err_a) err_a = i;
        bool prev_ok = false;
        RecordFields pf = record_fields(r.w[8], r.w[9]);
        if (i > 0) {
            const Rec q = load(bin, i - 1);
            prev_ok = sfen_decode([&](int j) { return q.w[j]; }, prev);
            pf = record_fields(q.w[8], q.w[9]);
        }
        const RecordFields cf = record_fields(r.w[8], r.w[9]);
        const u32 code = link_code(i > 0 && ok && prev_ok, prev, pf, cur, cf);
        codes_a[i] = code;
        if (code == 0u) store_stem(cur, cf, &stems_a[i * 8]);
    }
    // (b) chain walk
    std::vector<uint64_t> queue, next;
    auto on_error = [&](u64 rec) { if (rec < err_b) err_b = rec; };
    if (run < 0) {
        // the chain-owning form (k_walk_chains): a range's thread walks the chains whose heads lie in it
        const size_t range = (size_t)-run;
        for (size_t r0 = 0; r0 < n; r0 += range) {
            const size_t r1 = r0 + range < n ? r0 + range : n;
            size_t h = r0;
            if (r0 > 0)
                while (h < r1 && fields_link(load(bin, h - 1).w[9], load(bin, h).w[9])) ++h;
            if (h >= r1) continue;
            walk_item(bin, h, true, n, codes_b.data(), stems_b.data(), on_error, [&](u64 rec, u64) { return rec < r1; },
                      [](u64) {}, (r0 / range) % 2 ? tables() : nullptr);
        }
        run = 1;  // nothing queued
    } else
    for (size_t r0 = 0; r0 < n; r0 += (size_t)run) {
        const size_t e = r0 + run < n ? r0 + run : n;
        bool head = r0 == 0;
        if (!head) head = !fields_link(load(bin, r0 - 1).w[9], load(bin, r0).w[9]);
        walk_item(bin, head ? r0 : r0 - 1, head, e, codes_b.data(), stems_b.data(), on_error,
                  [&](u64 rec, u64 a) { if (rec == a + 1) return true; queue.push_back(rec); return false; }, [](u64) {},
                  (r0 / run) % 2 ? tables() : nullptr);
    }
    while (!queue.empty()) {
        *parked += queue.size();
        next.clear();
        for (uint64_t rec : queue) {
            size_t e = (rec / run + 1) * run;
            if (e > n) e = n;
            walk_item(bin, rec, true, e, codes_b.data(), stems_b.data(), on_error,
                      [&](u64 r, u64 a) { if (r == a + 1) return true; next.push_back(r); return false; }, [](u64) {},
                      (rec / run) % 2 ? nullptr : tables());
        }
        queue.swap(next);
    }
    *first_error = err_b;
    uint64_t bad = err_a != err_b ? 1 : 0;
    const size_t lim = err_a == ~0ull ? n : (size_t)err_a;  // records behind the first error are never used
    for (size_t i = 0; i < lim; ++i) {
        if (codes_a[i] != codes_b[i]) ++bad;
        else if (codes_a[i] == 0u && std::memcmp(&stems_a[i * 8], &stems_b[i * 8], 32) != 0) ++bad;
    }
    return bad;
}

// .binpack -> .bin with the device-side chain walker (chain.cuh: BitReader, decode_ply, pos_do_move,
// spliced stream): chunks and chains are followed sequentially as the reference reader does.
// Returns the number of records written, or -1 on a malformed chunk / chain.
long long sim_decode_binpack(const unsigned char* in, size_t n, unsigned char* out, size_t out_cap_records)
{
    size_t pos = 0;
    unsigned long long rec = 0;
    u32 col[8];
    while (pos < n) {
        if (n - pos < 8 || std::memcmp(in + pos, "BINP", 4) != 0) return -1;
        u32 size;
        std::memcpy(&size, in + pos + 4, 4);
        if (n - pos - 8 < size) return -1;
        const unsigned char* chunk = in + pos + 8;
        u32 cur = 0;
        while ((unsigned long long)cur + 34 <= size) {
            const u32 plies = ((u32)chunk[cur + 32] << 8) | chunk[cur + 33];
            u32 consumed = 0;
            if (!emit_chain_bin(chunk + cur, size - cur - 34, out, rec, out_cap_records, col, 1, consumed,
                                (rec & 1) ? tables() : nullptr))
                return -1;
            rec += 1ull + plies;
            cur += consumed;
        }
        pos += 8 + (size_t)size;
    }
    return (long long)rec;
}

namespace {
// the staged row kept up to date by halfkp_apply_move must hold the same (kind, square) values as a
// row rebuilt from the position, the map must point every listed piece at its slot
bool row_matches(const Pos& q, const HalfKpRow& R, const int* x, const unsigned char* map)
{
    HalfKpRow S;
    int y[HALFKP_STAGE];
    halfkp_rebuild<false>(q, S, y, nullptr, 0);
    if (S.n != R.n) return false;
    std::vector<int> a(x, x + R.n), b(y, y + S.n);
    std::sort(a.begin(), a.end());
    if (a != b) return false;  // the rebuilt row is ascending
    for (int j = 0; j < R.n; ++j)
        if (map[x[j] & 63] != j) return false;
    return true;
}
}  // namespace

// Walks every chain of a .binpack with the device-side walker, keeping the HalfKP row up to date with
// halfkp_apply_move exactly as warp_emit_chains_halfkp does, and compares it with a rebuilt row after
// every ply. Returns the number of mismatching positions (or -1 on a malformed file); *updated counts
// the plies handled incrementally, *rebuilt those outside halfkp_apply_move's domain.
long long sim_halfkp_chains(const unsigned char* in, size_t n, uint64_t* updated, uint64_t* rebuilt)
{
    size_t pos = 0;
    long long bad = 0;
    *updated = *rebuilt = 0;
    while (pos < n) {
        if (n - pos < 8 || std::memcmp(in + pos, "BINP", 4) != 0) return -1;
        u32 size;
        std::memcpy(&size, in + pos + 4, 4);
        if (n - pos - 8 < size) return -1;
        const unsigned char* chunk = in + pos + 8;
        u32 cur = 0;
        while ((unsigned long long)cur + 34 <= size) {
            ChainCursor cc;
            chain_open(chunk + cur, cc);
            BitReader r;
            r.init(chunk + cur + 34, (u64)(size - cur - 34));
            HalfKpRow R;
            int x[HALFKP_STAGE];
            unsigned char map[64];
            halfkp_rebuild<true>(cc.pos, R, x, map, 1);
            for (u32 k = 0; k < cc.num_plies; ++k) {
                if (cc.mv.from > 63 || cc.mv.to > 63) return -1;
                const int moved = pos_piece_at(cc.pos, cc.mv.from);
                const bool inc = halfkp_apply_move(cc.pos, cc.mv, moved, R, x, map, 1);
                if (!chain_step(cc, r, false, moved)) return -1;
                const int pieces = popc64(pos_all(cc.pos) & ~pos_type_bb(cc.pos, PT_KING));
                if (!inc || pieces != R.n) {
                    halfkp_rebuild<true>(cc.pos, R, x, map, 1);
                    ++*rebuilt;
                } else {
                    ++*updated;
                    if (!row_matches(cc.pos, R, x, map)) ++bad;
                }
            }
            cur += 34 + ((r.pos + 7) >> 3);
        }
        pos += 8 + (size_t)size;
    }
    return bad;
}

// Pseudo-random moves of all four types (mostly illegal) on the positions of a .bin: whatever
// halfkp_apply_move accepts must give the row of the position pos_do_move produces.
uint64_t sim_halfkp_fuzz(const unsigned char* bin, size_t n, int moves_per_pos, uint64_t seed, uint64_t* mismatch)
{
    uint64_t accepted = 0, xs = seed;
    *mismatch = 0;
    for (size_t i = 0; i < n; ++i) {
        const Rec r = load(bin, i);
        Pos p;
        pos_clear(p);
        if (!sfen_decode([&](int j) { return r.w[j]; }, p)) continue;
        for (int k = 0; k < moves_per_pos; ++k) {
            xs = xs * 6364136223846793005ull + 1442695040888963407ull;
            const u32 v = (u32)(xs >> 33);
            Move m;
            m.from = v & 63;
            m.to = (v >> 6) & 63;
            m.type = (v >> 12) & 3;
            m.promo = NO_PIECE;
            if ((v >> 20) & 1) {
                u64 all = pos_all(p);
                m.from = nth_set_bit(all, (v >> 21) % (u32)popc64(all));
            }
            if (m.type == MT_CASTLE && ((v >> 14) & 1)) {
                m.from = ((v >> 15) & 1) ? 4 : 60;
                m.to = (m.from & 56) + (((v >> 16) & 1) ? 7 : 0);
            }
            if (m.type == MT_PROMOTION) m.promo = ((PT_KNIGHT + (int)((v >> 17) & 3)) << 1) | (int)((v >> 19) & 1);
            HalfKpRow R;
            int x[HALFKP_STAGE];
            unsigned char map[64];
            halfkp_rebuild<true>(p, R, x, map, 1);
            const int moved = pos_piece_at(p, m.from);
            if (!halfkp_apply_move(p, m, moved, R, x, map, 1)) continue;
            Pos q = p;
            pos_do_move(q, m, moved);
            if (popc64(pos_all(q) & ~pos_type_bb(q, PT_KING)) != R.n) continue;  // the caller rebuilds
            ++accepted;
            if (!row_matches(q, R, x, map)) ++*mismatch;
        }
    }
    return accepted;
}

// HalfKP rows of every position of a .binpack, each rebuilt from the chain walker's position
// ((kind, square) order): the reference for inputs whose positions a .bin record cannot hold (corrupted
// movetext that captures a king), where rows made from the oracle's .bin differ by construction.
// Returns the number of rows, or -1 on a malformed file.
long long sim_halfkp_rows(const unsigned char* in, size_t n, int* white, int* black, size_t cap_rows)
{
    size_t pos = 0;
    unsigned long long rec = 0;
    auto emit = [&](const Pos& q) {
        if (rec < cap_rows) {
            HalfKpRow R;
            int x[HALFKP_STAGE];
            halfkp_rebuild<false>(q, R, x, nullptr, 0);
            for (int j = 0; j < HALFKP_ROW; ++j) {
                white[rec * HALFKP_ROW + j] = j < R.n ? R.wbase + x[j] : -1;
                black[rec * HALFKP_ROW + j] = j < R.n ? R.bbase + (x[j] ^ 127) : -1;
            }
        }
        ++rec;
    };
    while (pos < n) {
        if (n - pos < 8 || std::memcmp(in + pos, "BINP", 4) != 0) return -1;
        u32 size;
        std::memcpy(&size, in + pos + 4, 4);
        if (n - pos - 8 < size) return -1;
        const unsigned char* chunk = in + pos + 8;
        u32 cur = 0;
        while ((unsigned long long)cur + 34 <= size) {
            u32 consumed = 0;
            if (!walk_chain(chunk + cur, size - cur - 34, [&](const ChainCursor& cc, u32) { emit(cc.pos); }, consumed)) return -1;
            cur += consumed;
        }
        pos += 8 + (size_t)size;
    }
    return (long long)rec;
}

// The row k_bin_halfkp lists from the piece tokens of a record while decoding it, against the row
// rebuilt from the decoded position: same entries (the order differs). Returns the mismatches.
uint64_t sim_halfkp_tokens(const unsigned char* bin, size_t n)
{
    uint64_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const Rec r = load(bin, i);
        Pos p;
        int x[64], listed = 0;
        const bool ok = sfen_decode([&](int j) { return r.w[j]; }, p, [&](int sq, u32 tok) {
            if (listed < 64) x[listed] = 64 * (int)(((tok >> 1) & 7u) * 2u + ((tok >> 4) & 1u)) + sq;
            ++listed;
        });
        if (!ok) continue;
        HalfKpRow S;
        int y[HALFKP_STAGE];
        halfkp_rebuild<false>(p, S, y, nullptr, 0);
        if (listed > HALFKP_ROW) continue;  // the kernel rebuilds these
        std::vector<int> a(x, x + listed), b(y, y + S.n);
        std::sort(a.begin(), a.end());
        if (a != b) ++bad;
    }
    return bad;
}


// stem_to_record (the direct transcode of single-position chains) against the general route (stem_unpack +
// stream_from_pos + stream_with_tail + the record fields) on every stem of a .binpack and on `mutations`
// randomly damaged copies of each: whatever the direct route accepts must be the same 40 bytes.
uint64_t sim_stem_transcode_fuzz(const unsigned char* in, size_t n, int mutations, uint64_t seed, uint64_t* mismatch)
{
    uint64_t accepted = 0, xs = seed * 0x9E3779B97F4A7C15ull + 1;
    *mismatch = 0;
    size_t pos = 0;
    u32 col[8];
    auto check = [&](const unsigned char* stem) {
        alignas(8) unsigned char buf[48];
        for (int shift = 0; shift < 4; ++shift) {  // every alignment of the stem within a word
            std::memset(buf, 0xEE, sizeof buf);
            std::memcpy(buf + 4 + shift, stem, 34);
            const unsigned char* sp = buf + 4 + shift;
            u32 w[8], w8, w9;
            if (!stem_to_record(sp, col, 1, w, w8, w9)) continue;
            ++accepted;
            ChainCursor cc;
            chain_open(sp, cc);
            u32 W[8], g[8];
            const int end = stream_from_pos(cc.pos, col, 1, W);
            stream_with_tail(W, end, cc.pos, g);
            const u32 g8 = ((u32)cc.score & 0xFFFFu) | (move_to_sfmove(cc.mv) << 16);
            const u32 g9 = ((u32)cc.ply & 0xFFFFu) | (((u32)cc.result & 0xFFu) << 16) | 0xFF000000u;
            if (std::memcmp(w, g, 32) != 0 || w8 != g8 || w9 != g9) ++*mismatch;
        }
    };
    while (pos < n) {
        if (n - pos < 8 || std::memcmp(in + pos, "BINP", 4) != 0) break;
        u32 size;
        std::memcpy(&size, in + pos + 4, 4);
        if (n - pos - 8 < size) break;
        const unsigned char* chunk = in + pos + 8;
        u32 cur = 0;
        while ((unsigned long long)cur + 34 <= size) {
            check(chunk + cur);
            for (int m = 0; m < mutations; ++m) {
                unsigned char stem[34];
                std::memcpy(stem, chunk + cur, 34);
                for (int k = 0; k < 1 + (m & 3); ++k) {
                    xs = xs * 6364136223846793005ull + 1442695040888963407ull;
                    if (m & 4) {
                        // a nibble with a meaning of its own (kings, ep pawn, castling rooks) dropped on a random piece
                        unsigned char& b = stem[8 + (xs >> 33) % 16];
                        const unsigned v = 10u + (unsigned)((xs >> 50) % 6);
                        b = (xs >> 60) & 1 ? (unsigned char)((b & 0x0F) | (v << 4)) : (unsigned char)((b & 0xF0) | v);
                    } else {
                        stem[(xs >> 33) % 32] ^= (unsigned char)(1u << ((xs >> 60) & 7));
                    }
                }
                check(stem);
            }
            u32 consumed = 0;
            if (!walk_chain(chunk + cur, size - cur - 34, [](const ChainCursor&, u32) {}, consumed)) break;
            cur += consumed;
        }
        pos += 8 + (size_t)size;
    }
    return accepted;
}


// record_to_stem (the direct transcode of chain heads, heads.cuh) against the general route (sfen_decode +
// stem_pack) on every record of a .bin and on `mutations` randomly damaged copies of each: where the direct
// route answers, it must give the same 32 bytes, and it must call malformed exactly what the decoder does.
uint64_t sim_heads_transcode_fuzz(const unsigned char* bin, size_t n, int mutations, uint64_t seed, uint64_t* mismatch)
{
    uint64_t accepted = 0, xs = seed * 0x9E3779B97F4A7C15ull + 1;
    *mismatch = 0;
    auto check = [&](const Rec& r) {
        Pos p;
        pos_clear(p);
        const bool ok = sfen_decode([&](int j) { return r.w[j]; }, p);
        alignas(32) u32 g[8];
        if (ok) store_stem(p, record_fields(r.w[8], r.w[9]), g);
        u32 d[8];
        const int st = record_to_stem([&](int j) { return r.w[j]; }, d);
        if (st == HEADS_OTHER) return;
        if (st == HEADS_BAD) {
            if (ok) ++*mismatch;
            return;
        }
        ++accepted;
        if (!ok || std::memcmp(d, g, 32) != 0) ++*mismatch;
    };
    for (size_t i = 0; i < n; ++i) {
        Rec r = load(bin, i);
        check(r);
        for (int m = 0; m < mutations; ++m) {
            Rec q = r;
            for (int k = 0; k < 1 + (m & 3); ++k) {
                xs = xs * 6364136223846793005ull + 1442695040888963407ull;
                q.w[(xs >> 33) % 10] ^= 1u << ((xs >> 58) & 31);
            }
            check(q);
        }
    }
    return accepted;
}


// sfen_decode_flat (four tokens per step, planes gathered by multiplication) against the token-driven decoder on
// every record of a .bin and on `mutations` randomly damaged copies of each: the same verdict always, and the same
// position whenever the record decodes.
uint64_t sim_sfen_decode_fuzz(const unsigned char* bin, size_t n, int mutations, uint64_t seed, uint64_t* mismatch)
{
    uint64_t decoded = 0, xs = seed * 0x9E3779B97F4A7C15ull + 1;
    *mismatch = 0;
    auto check = [&](const Rec& r) {
        if (((r.w[0] >> 1) & 63u) == ((r.w[0] >> 7) & 63u)) return;  // the flat form is not asked
        Pos a, b;
        pos_clear(a);
        pos_clear(b);
        const bool oka = sfen_decode([&](int j) { return r.w[j]; }, a, [](int, u32) {});
        const bool okb = sfen_decode_flat([&](int j) { return r.w[j]; }, b);
        if (oka != okb) { ++*mismatch; return; }
        if (!oka) return;
        ++decoded;
        if (!pos_equal(a, b) || a.stm != b.stm || a.ep != b.ep || a.cr != b.cr || a.rule50 != b.rule50 || a.ply != b.ply) ++*mismatch;
    };
    for (size_t i = 0; i < n; ++i) {
        Rec r = load(bin, i);
        check(r);
        for (int m = 0; m < mutations; ++m) {
            Rec q = r;
            for (int k = 0; k < 1 + (m & 3); ++k) {
                xs = xs * 6364136223846793005ull + 1442695040888963407ull;
                q.w[(xs >> 33) % 8] ^= 1u << ((xs >> 58) & 31);
            }
            check(q);
        }
    }
    return decoded;
}

}  // extern "C"
