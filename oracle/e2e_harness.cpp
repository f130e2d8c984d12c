/*
 * e2e_harness.cpp -- BENCH / TEST INFRASTRUCTURE (built by oracle/Makefile into oracle/_ref/libbmb200_e2e.so, next to
 * test_cxx_binding, because it needs the reference headers from /root/reference/src to compile).
 *
 * What it measures: the product through the call a BitMagic user makes -- bm::b200::aggregator<bm::bvector<>> on REAL
 * bm::bvector<> objects (standard allocator, blocks wherever malloc put them), result materialised into a bm::bvector<>:
 *   cold : every step walks the block trees, packs the blocks, copies them H2D, aggregates, fetches the result
 *          (= aggregator::combine_*(target, src...) with nothing resident)
 *   warm : the sources are resident in a bm::b200::device_set (uploaded once, outside the timed region); every step is
 *          aggregator::combine_*(target, src...) -> kernel -> D2H of the result -> bvector
 * and e2e_check compares the b200 target with the reference aggregator's on the same vectors (compare() == 0, equal
 * calc_stat block kinds), the way tests/perf/perf.cpp:3990-4012 compares after every timed pair.
 * The oracle is not involved here; the reference headers only provide the container type the product binds to.
 */
#include <chrono>
#include <malloc.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "bm.h"
#include "bmaggregator.h"

#include "bmb200.h"
#include "bmb200_aggregator.hpp"

#ifdef E2E_SLAB       /* the same harness on bvectors whose blocks live in page-locked slabs (bmb200_alloc.hpp): libbmb200_e2e_slab.so */
typedef bm::b200::slab_bvector bvect;
#else
typedef bm::bvector<> bvect;
#endif

namespace {

struct Harness {
    std::unique_ptr<bm::b200::context> ctx;
    std::vector<std::unique_ptr<bvect>> vec;
    uint32_t n_blocks = 0;
    uint64_t set_bytes = 0;          // bytes a cold upload moves H2D (descriptors + bases + pools)
    bvect last;                      // target of the last timed call
};

/* blocks [nb_from, nb_from + s->n_blocks) of vector v, taken from a packed chunk whose column 0 is block nb_from */
void append_blocks(const bmb200_packed_set* s, uint32_t v, uint32_t nb_from, bvect& bv)
{
    bvect::blocks_manager_type& bman = bv.get_blocks_manager();
    BM_DECLARE_TEMP_BLOCK(tb)
    for (uint32_t c = 0; c < s->n_blocks; ++c)
    {
        const uint32_t d = s->desc[(size_t)c * s->n_vec + v], kind = d & 3u, rel = d >> 2;
        if (kind == BMB200_BLK_NULL) continue;
        const uint32_t nb = nb_from + c;
        unsigned i = nb >> 8, j = nb & 255u;
        bman.reserve_top_blocks(i + 1);
        bman.check_alloc_top_subblock(i);
        if (kind == BMB200_BLK_FULL) bman.set_block_ptr(i, j, FULL_BLOCK_FAKE_ADDR);
        else if (kind == BMB200_BLK_BIT)
        {
            std::memcpy(tb.begin(), s->bit_pool + (s->bit_base[c] + rel) * (size_t)BMB200_BLOCK_WORDS, BMB200_BLOCK_BYTES);
            bman.copy_bit_block(i, j, tb.begin());
        }
        else
        {
            const bm::gap_word_t* g = s->gap_pool + (s->gap_base[c] + (rel & BMB200_DESC_REL_MASK)) * (size_t)BMB200_GAP_UNIT_WORDS + (rel >> 29);
            unsigned len = bm::gap_length(g) - 1;
            int level = bm::gap_calc_level(len, bman.glen());
            bm::gap_word_t* gb = bman.allocate_gap_block(unsigned(level), g);
            bman.set_block_ptr(i, j, (bm::word_t*)BMPTR_SETBIT0(gb));
        }
    }
}

struct Groups { std::vector<const bvect*> g0, g1; };
Groups groups(Harness* h, const uint32_t* g0, uint32_t n0, const uint32_t* g1, uint32_t n1)
{
    Groups g;
    for (uint32_t k = 0; k < n0; ++k) g.g0.push_back(h->vec[g0[k]].get());
    for (uint32_t k = 0; k < n1; ++k) g.g1.push_back(h->vec[g1[k]].get());
    return g;
}

template<class AGG>
bool call(AGG& agg, int op, bool compress, bvect& target, const Groups& g)
{
    agg.set_optimization(compress ? bvect::opt_compress : bvect::opt_none);
    switch (op) {
    case BMB200_OP_OR:  agg.combine_or(target, g.g0.data(), g.g0.size()); return target.any();
    case BMB200_OP_AND: agg.combine_and(target, g.g0.data(), g.g0.size()); return target.any();
    case BMB200_OP_AND_SUB: return agg.combine_and_sub(target, g.g0.data(), g.g0.size(), g.g1.empty() ? 0 : g.g1.data(), g.g1.size(), false);
    default: return false;
    }
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

} // namespace

extern "C" {

/* n_vec empty bm::bvector<> objects of n_blocks * 65536 bits; numa != 0 pins this thread to the GPU's NUMA node first, so the
 * bvector blocks (first touch) and the pinned staging ring sit next to the GPU's PCIe root */
void* e2e_create_empty(uint32_t n_vec, uint32_t n_blocks, int device, int numa, int* numa_node)
{
    try {
        mallopt(M_TOP_PAD, 64 << 20);      /* many threads fill the bvectors through malloc: grow the arenas in big steps (mmap lock) */
        std::unique_ptr<Harness> h(new Harness());
        h->ctx.reset(new bm::b200::context(device));
        int node = -1;
        if (numa) bmb200_ctx_bind_host_numa(h->ctx->get(), &node);
        if (numa_node) *numa_node = node;
        h->n_blocks = n_blocks;
        h->vec.resize(n_vec);
        uint64_t bits = (uint64_t)n_blocks * 65536ull;
        if (bits > (uint64_t)bm::id_max) bits = bm::id_max;
        for (auto& p : h->vec) { p.reset(new bvect()); p->resize((bvect::size_type)bits); p->init(); }
        h->set_bytes = (uint64_t)n_vec * n_blocks * 4 + ((uint64_t)n_blocks + 1) * 16;
        return h.release();
    } catch (...) { return 0; }
}

/* fill block columns [nb_from, nb_from + chunk->n_blocks) of every vector from a host packed chunk (threads: one vector at a time each) */
int e2e_append(void* hv, const bmb200_packed_set* chunk, uint32_t nb_from, int threads)
{
    try {
        Harness* h = (Harness*)hv;
        if (chunk->n_vec != h->vec.size() || nb_from + chunk->n_blocks > h->n_blocks) return 2;
        if (threads < 1) threads = 1;
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t)
            th.emplace_back([h, chunk, nb_from, t, threads]() {
                for (uint32_t v = (uint32_t)t; v < chunk->n_vec; v += (uint32_t)threads) append_blocks(chunk, v, nb_from, *h->vec[v]); });
        for (auto& x : th) x.join();
        h->set_bytes += chunk->bit_base[chunk->n_blocks] * (uint64_t)BMB200_BLOCK_BYTES + chunk->gap_base[chunk->n_blocks] * 16ull;
        return 0;
    } catch (...) { return 1; }
}

/* cold: nothing resident; ms[k] = wall time of step k (tree walk + pack + H2D + kernel + D2H + bvector store) */
int e2e_cold(void* hv, int op, int compress, const uint32_t* g0, uint32_t n0, const uint32_t* g1, uint32_t n1, int steps, double* ms,
             uint64_t* count, uint64_t* h2d_bytes, uint64_t* d2h_bytes)
{
    try {
        Harness* h = (Harness*)hv;
        Groups g = groups(h, g0, n0, g1, n1);
        bm::b200::aggregator<bvect> agg(*h->ctx);
        for (int k = 0; k < steps; ++k) {
            const double t0 = now_ms();
            call(agg, op, compress != 0, h->last, g);
            ms[k] = now_ms() - t0;
        }
        if (count) *count = h->last.count();
        if (h2d_bytes) *h2d_bytes = h->set_bytes + 4ull * (n0 + n1);
        if (d2h_bytes) *d2h_bytes = agg.last_d2h_bytes();    /* counted by the binding from what bmb200_result_fetch_view handed back */
        return 0;
    } catch (...) { return 1; }
}

/* the same cold step in two halves: assign_ms = device_set::assign (walk + pack + H2D), agg_ms = the aggregation on the resident copy */
int e2e_cold_split(void* hv, int op, int compress, const uint32_t* g0, uint32_t n0, const uint32_t* g1, uint32_t n1, double* assign_ms, double* agg_ms)
{
    try {
        Harness* h = (Harness*)hv;
        Groups g = groups(h, g0, n0, g1, n1);
        std::vector<const bvect*> all(g.g0); all.insert(all.end(), g.g1.begin(), g.g1.end());
        bm::b200::device_set<bvect> ds(*h->ctx);
        bm::b200::aggregator<bvect> agg(*h->ctx);
        ds.assign(all.data(), all.size());                 // warm-up: pinned ring, parked arena and the aggregator's result buffers exist afterwards,
        agg.set_device_set(&ds);                           //   exactly the state the timed cold steps of e2e_cold run in
        call(agg, op, compress != 0, h->last, g);
        ds.release();
        double t0 = now_ms();
        ds.assign(all.data(), all.size());
        double t1 = now_ms();
        call(agg, op, compress != 0, h->last, g);
        double t2 = now_ms();
        *assign_ms = t1 - t0; *agg_ms = t2 - t1;
        return 0;
    } catch (...) { return 1; }
}

/* warm: sources resident in a device_set (uploaded before the clock starts); ms[k] = combine_* call -> result bvector */
int e2e_warm(void* hv, int op, int compress, const uint32_t* g0, uint32_t n0, const uint32_t* g1, uint32_t n1, int warmup, int steps, double* ms,
             uint64_t* count, uint64_t* d2h_bytes)
{
    try {
        Harness* h = (Harness*)hv;
        Groups g = groups(h, g0, n0, g1, n1);
        std::vector<const bvect*> all(g.g0); all.insert(all.end(), g.g1.begin(), g.g1.end());
        bm::b200::device_set<bvect> ds(*h->ctx);
        ds.assign(all.data(), all.size());
        bm::b200::aggregator<bvect> agg(*h->ctx);
        agg.set_device_set(&ds);
        for (int k = 0; k < warmup; ++k) call(agg, op, compress != 0, h->last, g);
        std::vector<char> thrash(getenv("E2E_THRASH") ? (size_t)atoi(getenv("E2E_THRASH")) << 20 : 0);   /* diagnostic: evict the CPU caches between calls (untimed) */
        for (int k = 0; k < steps; ++k) {
            if (!thrash.empty()) memset(thrash.data(), k, thrash.size());
            const double t0 = now_ms();
            call(agg, op, compress != 0, h->last, g);
            ms[k] = now_ms() - t0;
        }
        if (ds.stale()) return 3;
        if (count) *count = h->last.count();
        if (d2h_bytes) *d2h_bytes = agg.last_d2h_bytes();    /* counted by the binding from what bmb200_result_fetch_view handed back */
        return 0;
    } catch (...) { return 1; }
}

/* reference aggregator (threads workers over column ranges is not possible on whole bvectors: single-threaded, like the
 * reference itself) on the same vectors vs the last b200 target: *equal = compare() == 0 && equal calc_stat kinds */
int e2e_check(void* hv, int op, int compress, const uint32_t* g0, uint32_t n0, const uint32_t* g1, uint32_t n1, int* equal, uint64_t* ref_count, double* ref_ms)
{
    try {
        Harness* h = (Harness*)hv;
        Groups g = groups(h, g0, n0, g1, n1);
        bm::aggregator<bvect> agg;
        bvect t;
        const double t0 = now_ms();
        call(agg, op, compress != 0, t, g);
        if (ref_ms) *ref_ms = now_ms() - t0;
        bvect::statistics a, b; t.calc_stat(&a); h->last.calc_stat(&b);
        *equal = (t.compare(h->last) == 0) && a.bit_blocks == b.bit_blocks && a.gap_blocks == b.gap_blocks;
        if (ref_count) *ref_count = t.count();
        return 0;
    } catch (...) { return 1; }
}

void e2e_free(void* hv) { delete (Harness*)hv; }

/* slab build only: how many host slabs the vectors occupy and how many bytes of them are in use (= what a cold upload DMAs) */
int e2e_slab_info(uint64_t* n_slabs, uint64_t* bytes)
{
#ifdef E2E_SLAB
    *n_slabs = bm::b200::slab_heap::instance().slab_count(); *bytes = bm::b200::slab_heap::instance().bytes_handed_out();
    return 0;
#else
    *n_slabs = 0; *bytes = 0;
    return 1;
#endif
}

} // extern "C"
