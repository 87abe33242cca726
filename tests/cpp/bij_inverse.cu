// Host-side check of the pairing bijection and its analytic inverse (scorus_b200/csrc/rng.cuh): compiled with nvcc, run
// on the CPU by tests/test_matching_cpu.py.  A rank of a sharded run finds the matching position of each of its walkers
// with bij_inv instead of scanning the whole matching (api_shard.cu: shard_pairs_kernel).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../scorus_b200/csrc/rng.cuh"

using namespace scorus;

int main()
{
    long bad = 0, checked = 0;
    const uint32_t sizes[] = {2u, 6u, 1026u, 4096u, 100000u, 524288u, 1u << 22, 3000001u};
    for (uint32_t n : sizes) {
        for (uint64_t step = 0; step < 3; ++step) {
            for (uint32_t blk = 0; blk < 3; ++blk) {
                const BijKeys K = bij_keys(77, step, ST_PERM, 3, 2 * blk);
                const uint32_t b = bij_bits(n);
                const uint32_t lim = n < 100000 ? n : 100000;
                std::vector<unsigned char> seen(n < 100000 ? n : 0, 0);
                for (uint32_t i = 0; i < lim; ++i) {
                    const uint32_t x = n < 100000 ? i : (uint32_t)(((uint64_t)i * 2654435761u) % n);
                    const uint32_t y = bij(x, n, b, K);
                    ++checked;
                    if (y >= n || bij_inv(y, n, b, K) != x) ++bad;
                    if (!seen.empty()) { if (seen[y]) ++bad; seen[y] = 1; }   // a permutation: no value twice
                }
            }
        }
    }
    // keys of different matching blocks differ (sub-counter 2c), block 0 is the whole-rung key
    const BijKeys k0 = bij_keys(5, 9, ST_PERM, 1), k0b = bij_keys(5, 9, ST_PERM, 1, 0), k1 = bij_keys(5, 9, ST_PERM, 1, 2);
    for (int q = 0; q < 8; ++q) if (k0.k[q] != k0b.k[q]) ++bad;
    int same = 0;
    for (int q = 0; q < 8; ++q) same += k0.k[q] == k1.k[q];
    if (same == 8) ++bad;
    printf("checked=%ld bad=%ld\n", checked, bad);
    return bad != 0;
}
