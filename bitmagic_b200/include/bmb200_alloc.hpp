// bmb200_alloc.hpp -- a slab-backed block allocator for bm::bvector<> whose memory the GPU can take as it lies.
//
// The reference lets the application replace the block allocator (template parameter Alloc of bm::bvector<>, src/bm.h:121;
// bm::mem_alloc<BA, PA, APool>, src/bmalloc.h:288-420; samples/bvsample06/sample6.cpp:47-110 shows a custom BA).  Every
// bit-block and GAP block goes through BA::allocate / BA::deallocate (src/bmalloc.h:355-397).  bm::b200::slab_block_allocator
// serves them from a few large page-locked slabs (bmb200_host_slab_alloc -> cudaHostAlloc), so that
//
//     typedef bm::b200::slab_bvector bvect;          // instead of bm::bvector<>
//
// is all an application changes: bm::b200::device_set<bvect>::assign and bm::b200::aggregator<bvect> then upload with
// bmb200_set_upload_slabs -- the slabs cross PCIe by DMA as they lie, no host thread copies a block, the tree walk overlaps the
// DMA, and a device kernel gathers the blocks into the column-major arena.  Everything else (set algebra, serialization,
// the reference's own aggregator) works on slab_bvector unchanged: it is a bm::bvector<>.
//
// The heap is process-wide (BA's functions are static in the reference's allocator model): size-class free lists + per-thread
// bump chunks, 64-byte aligned blocks (BM_ALLOC_ALIGN is 32 for AVX2, 64 for AVX-512).  Slabs are never returned to the system
// before slab_heap::release_all(); freed blocks are reused by later allocations of the same size.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <new>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#include "bm.h"
#include "bmb200.h"

namespace bm { namespace b200 {

class slab_heap
{
public:
    typedef void* (*slab_alloc_fn)(size_t);
    typedef void (*slab_free_fn)(void*);

    static slab_heap& instance() { static slab_heap* h = new slab_heap(); return *h; }   // never destroyed: bvectors with static lifetime may outlive main()

    /// slab size for slabs created from now on (default 256 MB; page-locking costs ~0.1 ms / MB once)
    void set_slab_bytes(size_t b) { std::lock_guard<std::mutex> lk(mu_); slab_bytes_ = b < (1u << 20) ? (1u << 20) : b; }
    /// where slabs come from: default bmb200_host_slab_alloc / bmb200_host_slab_free (page-locked); a host-only build or test
    /// may plug aligned_alloc / free (the upload then runs at pageable-memory speed but stays correct)
    void set_backing(slab_alloc_fn a, slab_free_fn f) { std::lock_guard<std::mutex> lk(mu_); slab_alloc_ = a; slab_free_ = f; }

    bm::word_t* allocate(size_t n_words)
    {
        const size_t bytes = (n_words * sizeof(bm::word_t) + kAlign - 1) & ~(kAlign - 1);
        if (free_blocks_.load(std::memory_order_relaxed))
        {
            std::lock_guard<std::mutex> lk(mu_);
            auto it = free_.find(bytes);
            if (it != free_.end() && !it->second.empty())
            {
                void* p = it->second.back(); it->second.pop_back();
                free_blocks_.fetch_sub(1, std::memory_order_relaxed);
                return (bm::word_t*)p;
            }
        }
        local_chunk& lc = local();
        if (lc.gen != generation_.load(std::memory_order_acquire)) { lc.cur = lc.end = nullptr; lc.gen = generation_.load(std::memory_order_acquire); }
        if ((size_t)(lc.end - lc.cur) < bytes)
        {
            if (bytes >= kLocalChunk / 4) return (bm::word_t*)carve(bytes);        // big requests go straight to the slab
            give_back(lc.cur, lc.end);
            refill(lc, bytes);
        }
        uint8_t* p = lc.cur; lc.cur += bytes;
        return (bm::word_t*)p;
    }

    void deallocate(bm::word_t* p, size_t n_words) noexcept
    {
        if (!p) return;
        const size_t bytes = (n_words * sizeof(bm::word_t) + kAlign - 1) & ~(kAlign - 1);
        try {
            std::lock_guard<std::mutex> lk(mu_);
            free_[bytes].push_back(p);
            free_blocks_.fetch_add(1, std::memory_order_relaxed);
        } catch (...) {}                                                             // the block stays lost inside its slab
    }

    /// the slabs as bmb200_set_upload_slabs takes them (base + bytes handed out so far)
    void snapshot(std::vector<bmb200_host_slab>& out) const
    {
        std::lock_guard<std::mutex> lk(mu_);
        out.clear();
        for (const slab& s : slabs_) if (s.used) out.push_back(bmb200_host_slab{s.base, (uint64_t)s.used});
    }
    size_t slab_count() const { std::lock_guard<std::mutex> lk(mu_); return slabs_.size(); }
    size_t bytes_reserved() const { std::lock_guard<std::mutex> lk(mu_); size_t t = 0; for (const slab& s : slabs_) t += s.cap; return t; }
    size_t bytes_handed_out() const { std::lock_guard<std::mutex> lk(mu_); size_t t = 0; for (const slab& s : slabs_) t += s.used; return t; }

    /// give every slab back.  Only when no bvector of this allocator is alive any more.
    void release_all()
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (slab& s : slabs_) slab_free_(s.base);
        slabs_.clear(); free_.clear(); spare_.clear(); free_blocks_.store(0);
        generation_.fetch_add(1, std::memory_order_release);                         // per-thread bump chunks of the old slabs are void
    }

private:
    static constexpr size_t kAlign = 64;
    static constexpr size_t kLocalChunk = 1u << 20;
    struct slab { uint8_t* base; size_t cap, used; };
    // a thread's bump chunk; what is left of it when the thread ends goes back to the heap (short-lived worker threads that fill
    // bvectors would otherwise strand up to 1 MB each)
    struct local_chunk
    {
        uint8_t* cur = nullptr; uint8_t* end = nullptr; uint64_t gen = 0;
        ~local_chunk() { if (cur != end && gen == slab_heap::instance().generation_.load(std::memory_order_acquire)) slab_heap::instance().give_back(cur, end); }
    };
    void give_back(uint8_t* cur, uint8_t* end) noexcept
    {
        if ((size_t)(end - cur) < 4096) return;
        try { std::lock_guard<std::mutex> lk(mu_); spare_.push_back(std::make_pair(cur, end)); } catch (...) {}
    }
    void refill(local_chunk& lc, size_t bytes)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t k = spare_.size(); k-- > 0; )
                if ((size_t)(spare_[k].second - spare_[k].first) >= bytes)
                { lc.cur = spare_[k].first; lc.end = spare_[k].second; spare_.erase(spare_.begin() + (long)k); return; }
        }
        lc.cur = (uint8_t*)carve(kLocalChunk); lc.end = lc.cur + kLocalChunk;
    }
    static local_chunk& local() { static thread_local local_chunk lc; return lc; }

    static void* pinned_alloc(size_t b) { void* p = nullptr; return bmb200_host_slab_alloc((uint64_t)b, &p) == BMB200_OK ? p : nullptr; }
    static void pinned_free(void* p) { bmb200_host_slab_free(p); }

    void* carve(size_t bytes)
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (slabs_.empty() || slabs_.back().cap - slabs_.back().used < bytes)
        {
            const size_t cap = bytes > slab_bytes_ ? bytes : slab_bytes_;
            uint8_t* b = (uint8_t*)slab_alloc_(cap);
            if (!b) throw std::bad_alloc();
            slabs_.push_back(slab{b, cap, 0});
        }
        slab& s = slabs_.back();
        void* p = s.base + s.used; s.used += bytes;
        return p;
    }

    slab_heap() : slab_alloc_(&pinned_alloc), slab_free_(&pinned_free) {}
    mutable std::mutex mu_;
    std::vector<slab> slabs_;
    std::unordered_map<size_t, std::vector<void*>> free_;
    std::vector<std::pair<uint8_t*, uint8_t*>> spare_;
    std::atomic<size_t> free_blocks_{0};
    std::atomic<uint64_t> generation_{1};
    size_t slab_bytes_ = 256u << 20;
    slab_alloc_fn slab_alloc_;
    slab_free_fn slab_free_;
};

/// BA of bm::mem_alloc<> (src/bmalloc.h:57-98): same two static functions
class slab_block_allocator
{
public:
    static bm::word_t* allocate(size_t n, const void*) { return slab_heap::instance().allocate(n); }
    static void deallocate(bm::word_t* p, size_t n) BMNOEXCEPT { slab_heap::instance().deallocate(p, n); }
};

typedef bm::alloc_pool<slab_block_allocator, bm::ptr_allocator> slab_alloc_pool;
typedef bm::mem_alloc<slab_block_allocator, bm::ptr_allocator, slab_alloc_pool> slab_allocator;
typedef bm::bvector<slab_allocator> slab_bvector;

namespace detail {
/// does BV keep its blocks in the slab heap?  (device_set / aggregator pick bmb200_set_upload_slabs then)
template<class BV> struct slab_backed
{ static const bool value = std::is_same<typename BV::allocator_type::block_allocator_type, slab_block_allocator>::value; };
}  // namespace detail

}}  // namespace bm::b200
