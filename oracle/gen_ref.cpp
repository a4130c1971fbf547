// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Synthetic .bin generator that drives the *reference's own* chess library and
// emitBinEntry (compress_file.cpp:1239). It is compiled by oracle/Makefile against a
// scratch copy of /root/reference/src and lands in oracle/_ref/gen_ref. The recipe is
// the one described in SURVEY.md section 8(d): random legal-move games from the start
// position, at most L plies per chain, mt19937_64(seed).
//
// Usage: gen_ref <out.bin> <num_positions> <max_plies_per_chain> <seed> [mode]
//   mode 0 (default): plain games
//   mode 1: "shuffled plies" -- gamePly is replaced by a constant so that no record
//           continues its predecessor (every record becomes a chain head)
//   mode 2: every 7th game restarts with the same result sign / ply+1 as the
//           previous record to probe isContinuation's short-circuit order
#define main reference_main
#include "compress_file.cpp"
#undef main

#include "chess/MoveGenerator.h"

#include <random>
#include <cstdlib>

int main(int argc, char** argv)
{
    if (argc < 5)
    {
        std::fprintf(stderr, "usage: %s out.bin num_positions max_plies seed [mode]\n", argv[0]);
        return 2;
    }
    const std::string outPath = argv[1];
    const std::size_t numPositions = std::strtoull(argv[2], nullptr, 10);
    const std::size_t maxPlies = std::strtoull(argv[3], nullptr, 10);
    const std::uint64_t seed = std::strtoull(argv[4], nullptr, 10);
    const int mode = argc > 5 ? std::atoi(argv[5]) : 0;

    std::mt19937_64 rng(seed);
    std::ofstream out(outPath, std::ios_base::binary | std::ios_base::trunc);
    std::vector<char> buffer;
    buffer.reserve(2 * MiB);

    std::size_t emitted = 0;
    std::size_t gameNo = 0;
    TrainingDataEntry last{};
    while (emitted < numPositions)
    {
        TrainingDataEntry e;
        e.pos = Position::startPosition();
        e.pos.setPly(0);
        e.ply = 0;
        e.score = static_cast<std::int16_t>(static_cast<int>(rng() % 200) - 100);
        e.result = static_cast<std::int16_t>(static_cast<int>(rng() % 3) - 1);
        if (mode == 2 && gameNo % 7 == 3 && emitted > 0)
        {
            e.ply = last.ply + 1;
            e.result = -last.result;
        }
        ++gameNo;

        for (std::size_t i = 0; i < maxPlies && emitted < numPositions; ++i)
        {
            const auto moves = movegen::generateLegalMoves(e.pos);
            if (moves.empty() || e.pos.rule50Counter() >= 63)
            {
                break;
            }
            e.move = moves[rng() % moves.size()];

            TrainingDataEntry w = e;
            if (mode == 1) w.ply = 77;
            emitBinEntry(buffer, w);
            ++emitted;
            last = e;

            e.pos.doMove(e.move);
            e.ply += 1;
            e.result = -e.result;
            int s = -static_cast<int>(e.score) + static_cast<int>(rng() % 61) - 30;
            if (s > 3000) s = 3000;
            if (s < -3000) s = -3000;
            e.score = static_cast<std::int16_t>(s);

            if (buffer.size() > MiB)
            {
                out.write(buffer.data(), buffer.size());
                buffer.clear();
            }
        }
    }
    if (!buffer.empty())
    {
        out.write(buffer.data(), buffer.size());
    }
    return 0;
}
