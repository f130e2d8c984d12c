/*
 * slab_heap_check.cpp -- TEST INFRASTRUCTURE ONLY (host-only; tests/test_host_logic.py runs it).
 *
 * bm::b200::slab_bvector (bitmagic_b200/include/bmb200_alloc.hpp) under real bm::bvector<> traffic, next to a plain bm::bvector<>
 * that receives the same operations: set / clear / ranges, optimize, logical ops, copy, swap, destruction, all from several
 * threads at once.  Checked: equal contents (compare through a bit-by-bit enumerator: the two types do not compare directly),
 * every real block of a slab_bvector lies inside the heap's slabs, 64-byte aligned, inside the extent snapshot() reports, and
 * freed blocks are reused (the heap does not grow when the same work is repeated).  The slabs come from aligned_alloc here
 * (set_backing): this container has no GPU to page-lock against.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "bm.h"
#include "bmb200.h"
#include "bmb200_alloc.hpp"

typedef bm::bvector<> ref_bv;
typedef bm::b200::slab_bvector slab_bv;

static int failures = 0;
#define CHECK(c, ...) do { if (!(c)) { ++failures; fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

static void* plain_alloc(size_t b) { return aligned_alloc(4096, (b + 4095) & ~(size_t)4095); }
static void plain_free(void* p) { free(p); }

static uint64_t mix(uint64_t& s) { s += 0x9e3779b97f4a7c15ull; uint64_t z = s; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }

template<class A, class B> static bool same(const A& a, const B& b)
{
    if (a.count() != b.count()) return false;
    typename A::enumerator ea = a.first(); typename B::enumerator eb = b.first();
    for (; ea.valid() && eb.valid(); ++ea, ++eb) if (*ea != *eb) return false;
    return !ea.valid() && !eb.valid();
}

static bool blocks_inside(const slab_bv& bv, const std::vector<bmb200_host_slab>& slabs)
{
    const slab_bv::blocks_manager_type& bman = bv.get_blocks_manager();
    if (!bman.is_init()) return true;
    for (unsigned i = 0; i < bman.top_block_size(); ++i)
        for (unsigned j = 0; j < bm::set_sub_array_size; ++j) {
            const bm::word_t* blk = bman.get_block_ptr(i, j);
            if (!IS_VALID_ADDR(blk)) continue;
            const uint8_t* p; size_t len;
            if (BM_IS_GAP(blk)) { const bm::gap_word_t* g = BMGAP_PTR(blk); p = (const uint8_t*)g; len = (size_t)bm::gap_length(g) * 2; }
            else { p = (const uint8_t*)blk; len = 8192; }
            if ((uintptr_t)p & 63u) return false;
            bool in = false;
            for (const bmb200_host_slab& s : slabs) if (p >= (const uint8_t*)s.base && p + len <= (const uint8_t*)s.base + s.bytes) { in = true; break; }
            if (!in) return false;
        }
    return true;
}

template<class BV> static void mutate(BV& bv, uint64_t seed, unsigned rounds)
{
    uint64_t s = seed;
    for (unsigned r = 0; r < rounds; ++r) {
        const uint64_t x = mix(s);
        const typename BV::size_type pos = (typename BV::size_type)(mix(s) % 40000000u);
        switch (x % 7u) {
        case 0: for (unsigned k = 0; k < 3000; ++k) bv.set((typename BV::size_type)(pos + (mix(s) % 300000u))); break;
        case 1: bv.set_range(pos, pos + (typename BV::size_type)(mix(s) % 200000u)); break;
        case 2: bv.clear_range(pos, pos + (typename BV::size_type)(mix(s) % 150000u)); break;
        case 3: bv.optimize(); break;
        case 4: { BV t; t.set_range(pos, pos + 70000u); for (unsigned k = 0; k < 500; ++k) t.set((typename BV::size_type)(mix(s) % 40000000u)); t.optimize(); if (x & 8u) bv |= t; else bv ^= t; } break;
        case 5: { BV t; for (unsigned k = 0; k < 2000; ++k) t.set((typename BV::size_type)(pos + 64u * k)); bv -= t; } break;
        default: { BV c(bv); c.optimize(); bv.swap(c); } break;
        }
    }
}

int main()
{
    bm::b200::slab_heap& heap = bm::b200::slab_heap::instance();
    heap.set_backing(&plain_alloc, &plain_free);
    heap.set_slab_bytes(8u << 20);

    static_assert(bm::b200::detail::slab_backed<slab_bv>::value, "trait");
    static_assert(!bm::b200::detail::slab_backed<ref_bv>::value, "trait");

    const unsigned T = 8, per = 6;
    std::vector<std::vector<slab_bv>> sv(T, std::vector<slab_bv>(per));
    std::vector<std::vector<ref_bv>> rv(T, std::vector<ref_bv>(per));
    auto run = [&](unsigned rounds, uint64_t salt) {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t]() { for (unsigned k = 0; k < per; ++k) { mutate(sv[t][k], salt + 1000u * t + k, rounds); mutate(rv[t][k], salt + 1000u * t + k, rounds); } });
        for (auto& x : th) x.join();
    };
    run(40, 1);
    std::vector<bmb200_host_slab> slabs;
    heap.snapshot(slabs);
    CHECK(!slabs.empty(), "no slabs after allocations");
    unsigned checks = 0;
    for (unsigned t = 0; t < T; ++t) for (unsigned k = 0; k < per; ++k) {
        CHECK(same(sv[t][k], rv[t][k]), "contents differ t=%u k=%u", t, k);
        bm::bvector<>::statistics a; slab_bv::statistics b; rv[t][k].calc_stat(&a); sv[t][k].calc_stat(&b);
        CHECK(a.bit_blocks == b.bit_blocks && a.gap_blocks == b.gap_blocks, "block kinds differ t=%u k=%u", t, k);
        CHECK(blocks_inside(sv[t][k], slabs), "a block of t=%u k=%u lies outside the slabs / is misaligned", t, k);
        checks += 3;
    }
    // free everything, repeat the same work: the heap must serve it from its free lists (growth of at most a few bump chunks)
    const size_t reserved1 = heap.bytes_reserved();
    for (auto& v : sv) for (auto& b : v) b.clear(true);
    for (auto& v : rv) for (auto& b : v) b.clear(true);
    run(40, 1);
    const size_t reserved2 = heap.bytes_reserved();
    CHECK(reserved2 <= reserved1 + (32u << 20), "heap grew from %zu to %zu bytes on identical work (freed blocks not reused)", reserved1, reserved2);
    heap.snapshot(slabs);
    for (unsigned t = 0; t < T; ++t) for (unsigned k = 0; k < per; ++k) {
        CHECK(same(sv[t][k], rv[t][k]), "contents differ after reuse t=%u k=%u", t, k);
        CHECK(blocks_inside(sv[t][k], slabs), "a block lies outside the slabs after reuse t=%u k=%u", t, k);
        checks += 2;
    }
    ++checks;
    sv.clear(); rv.clear();
    heap.release_all();
    CHECK(heap.slab_count() == 0, "release_all left slabs");
    { slab_bv again; again.set(5); again.set(70000); CHECK(again.count() == 2, "allocation after release_all"); ++checks; }
    printf("%s: %u checks, %d failed; heap %zu -> %zu bytes\n", failures ? "FAILED" : "OK", checks + 1, failures, reserved1, reserved2);
    return failures ? 1 : 0;
}
