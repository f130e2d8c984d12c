/*
 * store_result_check.cpp -- TEST INFRASTRUCTURE.  bm::b200::detail::store_result (the binding's result materialisation: recycles the
 * target's blocks in place, several host threads over disjoint top-level sub-trees) against a fresh build of the same result:
 * random kind mixes, changing sizes / block ranges / shard offsets from call to call into ONE target.  Host code only (no GPU needed);
 * built by oracle/Makefile into oracle/_ref/ because it needs the reference headers.
 */
#include <cstdio>
#include <random>
#include <vector>
#include "bm.h"
#include "bmaggregator.h"
#include "bmb200_aggregator.hpp"
typedef bm::bvector<> bvect;
int main()
{
    std::mt19937 rng(7);
    bvect target; int fails = 0;
    for (int it = 0; it < 60; ++it) {
        const uint32_t n_cols = (it % 3 == 0) ? 2400 : (it % 3 == 1 ? 700 : 2048 + 256 * (it % 5));
        const uint32_t nb_off = (it % 4 == 3) ? 100 : 0;
        std::vector<uint8_t> kind(n_cols); std::vector<uint64_t> off(n_cols);
        std::vector<uint16_t> gaps; std::vector<uint32_t> bits;
        const int mode = it % 6;
        for (uint32_t c = 0; c < n_cols; ++c) {
            unsigned r = rng() % 100;
            unsigned k = mode == 5 ? BMB200_BLK_FULL : r < 25 ? BMB200_BLK_NULL : r < 35 ? BMB200_BLK_FULL : r < 55 ? BMB200_BLK_BIT : BMB200_BLK_GAP;
            if (mode == 4 && (c >> 8) % 2) k = BMB200_BLK_NULL;       // whole top-level blocks empty
            kind[c] = (uint8_t)k;
            if (k == BMB200_BLK_BIT) { off[c] = bits.size() / 2048; for (int w = 0; w < 2048; ++w) bits.push_back(rng()); }
            else if (k == BMB200_BLK_GAP) {
                off[c] = gaps.size();
                uint32_t runs = 1 + rng() % (it % 2 ? 1270 : 200); std::vector<uint16_t> g(runs + 1);
                uint32_t pos = rng() % 20;
                for (uint32_t q = 1; q < runs; ++q) { g[q] = (uint16_t)pos; pos += 1 + rng() % 45; }
                g[runs] = 65535; g[0] = (uint16_t)((runs << 3) | (rng() & 1));
                gaps.insert(gaps.end(), g.begin(), g.end()); while (gaps.size() % 8) gaps.push_back(0);
            }
        }
        const bvect::size_type sz = (bvect::size_type)((it % 7 == 6 ? 3000u : 2700u) * 65536u - 1);
        bvect fresh;
        bm::b200::detail::store_result(fresh, sz, n_cols, kind.data(), off.data(), bits.data(), gaps.data(), nb_off);
        bm::b200::detail::store_result(target, sz, n_cols, kind.data(), off.data(), bits.data(), gaps.data(), nb_off);
        bvect::statistics a, b; fresh.calc_stat(&a); target.calc_stat(&b);
        bool ok = fresh.compare(target) == 0 && a.bit_blocks == b.bit_blocks && a.gap_blocks == b.gap_blocks && fresh.count() == target.count() && fresh.size() == target.size();
        if (!ok) { ++fails; std::printf("iteration %d differs (%zu/%zu vs %zu/%zu)\n", it, (size_t)a.bit_blocks, (size_t)a.gap_blocks, (size_t)b.bit_blocks, (size_t)b.gap_blocks); }
    }
    std::printf("%s: %d failures\n", fails ? "FAILED" : "OK", fails);
    return fails != 0;
}
