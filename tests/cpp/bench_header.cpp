// Throughput of the C++ host API itself: SmithWatermanSA<std::string,char,'-'>::getAlignmentsPacked
// on N pairs of 150 bp std::strings from the shared splitmix64 generator -- what a user of the
// reference's classes sees, pageable std::string / std::vector buffers included.  Prints one JSON line.
//   bench_header [pairs=1000000] [reps=3] [materialise=0|1]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "SequenceAlignment.h"

static uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static std::string sequence(uint64_t seed, uint64_t pair, int which, int len)
{
    const uint64_t key = splitmix64(seed ^ (2 * pair + (uint64_t)which));
    std::string s((size_t)len, 'A');
    for (int pos = 0; pos < len; pos++) s[(size_t)pos] = "ACGT"[(splitmix64(key + (uint64_t)(pos >> 5)) >> (2 * (pos & 31))) & 3];
    return s;
}

int main(int argc, char **argv)
{
    const size_t N = argc > 1 ? (size_t)atoll(argv[1]) : 1000000;
    const int reps = argc > 2 ? atoi(argv[2]) : 3;
    const bool materialise = argc > 3 && atoi(argv[3]) != 0;
    const int L = 150;
    std::vector<std::pair<std::string, std::string>> pairs(N);
    for (size_t p = 0; p < N; p++) pairs[p] = {sequence(20240607ull, p, 0, L), sequence(20240607ull, p, 1, L)};
    SmithWatermanSA<std::string, char, '-'> SW(ScoringSystem(-1, 1, -1));
    long long checksum = 0;
    size_t entries = 0;
    double best = 1e30;
    for (int r = 0; r < reps + 1; r++) { // first call warms the device contexts up
        const auto t0 = std::chrono::steady_clock::now();
        if (materialise) {
            auto all = SW.getAlignments(pairs);
            entries = 0;
            for (auto &a : all) entries += a.size();
        } else {
            seqa::PackedAlignments pk = SW.getAlignmentsPacked(pairs);
            checksum = 0;
            for (size_t p = 0; p < pk.size(); p++) checksum += pk.Score[p];
        }
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (r > 0 && s < best) best = s;
    }
    std::printf("{\"api\": \"%s\", \"pairs\": %zu, \"len\": %d, \"seconds\": %.6f, \"gcups\": %.2f, \"score_checksum\": %lld, \"entries\": %zu}\n",
                materialise ? "SmithWatermanSA::getAlignments (std::list materialised)" : "SmithWatermanSA::getAlignmentsPacked", N, L, best,
                (double)N * L * L / best / 1e9, checksum, entries);
    return 0;
}
