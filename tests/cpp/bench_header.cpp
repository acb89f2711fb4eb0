// Throughput of the C++ host API itself, the calls a user of the reference's classes makes:
//   SmithWatermanSA<std::string,char,'-'>::getAlignmentsPacked  on N pairs of 150 bp std::strings  (scores + op strings)
//   SmithWatermanSA<std::string,char,'-'>::getAlignments        on M pairs (std::list<Entry> materialised, the reference's
//                                                                own result type: what the CPU reference arm also builds)
// Inputs come from the shared splitmix64 generator; pageable std::string / std::vector buffers, packing,
// PCIe and result expansion are all inside the timed call.  Prints one JSON line.
//   bench_header [packed_pairs=1000000] [reps=3] [list_pairs=200000]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
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
    uint64_t word = 0;
    for (int pos = 0; pos < len; pos++) {
        if ((pos & 31) == 0) word = splitmix64(key + (uint64_t)(pos >> 5));
        s[(size_t)pos] = "ACGT"[(word >> (2 * (pos & 31))) & 3];
    }
    return s;
}

int main(int argc, char **argv)
{
    const size_t N = argc > 1 ? (size_t)atoll(argv[1]) : 1000000;
    const int reps = argc > 2 ? atoi(argv[2]) : 3;
    const size_t M = argc > 3 ? (size_t)atoll(argv[3]) : 200000;
    const int L = 150;
    std::vector<std::pair<std::string, std::string>> pairs(std::max(N, M));
    seqa::detail::parallelFor(pairs.size(), seqa::detail::hostThreads(pairs.size(), 4096), [&](size_t lo, size_t hi, size_t) {
        for (size_t p = lo; p < hi; p++) pairs[p] = {sequence(20240607ull, p, 0, L), sequence(20240607ull, p, 1, L)};
    });
    SmithWatermanSA<std::string, char, '-'> SW(ScoringSystem(-1, 1, -1));
    long long checksum = 0;
    size_t entries = 0;
    double best_packed = 1e30, best_list = 1e30;
    bool two_bit_in = false;
    if (N) {
        std::vector<std::pair<std::string, std::string>> batch(pairs.begin(), pairs.begin() + (long)N);
        for (int r = 0; r < reps + 1; r++) { // first call warms the device contexts up
            const auto t0 = std::chrono::steady_clock::now();
            seqa::PackedAlignments pk = SW.getAlignmentsPacked(batch);
            checksum = 0;
            for (size_t p = 0; p < pk.size(); p++) checksum += pk.Score[p];
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (r > 0 && s < best_packed) best_packed = s;
        }
        two_bit_in = SW.LastInputsTwoBit;
    }
    if (M) {
        std::vector<std::pair<std::string, std::string>> batch(pairs.begin(), pairs.begin() + (long)M);
        for (int r = 0; r < 2; r++) {
            const auto t0 = std::chrono::steady_clock::now();
            auto all = SW.getAlignments(batch);
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); // results built; their destruction is the caller's
            entries = 0;
            for (auto &a : all) entries += a.size();
            if (s < best_list) best_list = s;
        }
    }
    std::printf("{\"packed\": {\"api\": \"SmithWatermanSA<std::string,char,'-'>::getAlignmentsPacked\", \"pairs\": %zu, \"len\": %d, \"seconds\": %.6f, "
                "\"gcups\": %.2f, \"score_checksum\": %lld, \"inputs_2bit\": %s}, "
                "\"list\": {\"api\": \"SmithWatermanSA<std::string,char,'-'>::getAlignments (std::list<Entry> materialised)\", \"pairs\": %zu, \"len\": %d, "
                "\"seconds\": %.6f, \"gcups\": %.2f, \"entries\": %zu}, \"host_threads\": %u}\n",
                N, L, N ? best_packed : 0.0, N ? (double)N * L * L / best_packed / 1e9 : 0.0, checksum, two_bit_in ? "true" : "false",
                M, L, M ? best_list : 0.0, M ? (double)M * L * L / best_list / 1e9 : 0.0, entries, std::thread::hardware_concurrency());
    return 0;
}
