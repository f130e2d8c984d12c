/*
 * host_pack_check.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A host build of the product's upload packer (bitmagic_b200/csrc/host_pack.hpp: threaded layout pass + the packing pipeline
 * over a ring of staging slots) so that this GPU-less container can check that exact logic: the "DMA" of a finished chunk is a
 * memcpy from its slot into the output pools, everything else -- descriptor building, prefix sums, chunking, the worker /
 * release protocol -- is the code bmb200_set_upload_vectors runs.  tests/test_host_logic.py compares the outcome with
 * hostfmt.PackedSet.pack.  Nothing in bitmagic_b200/ uses this file.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../include/bmb200.h"
#include "../bitmagic_b200/csrc/host_pack.hpp"

using namespace bmb200;

extern "C" int host_pack_sizes(uint32_t n_vec, uint32_t n_blocks, const bmb200_vec_blocks* vecs, int threads, uint64_t* n_bit, uint64_t* n_gap)
{
    PackLayout L;
    pack_layout(n_vec, n_blocks, vecs, (unsigned)threads, L);
    if (L.rc) return L.rc;
    *n_bit = L.bb[n_blocks]; *n_gap = L.gb[n_blocks];
    return 0;
}

extern "C" int host_pack_check2(uint32_t n_vec, uint32_t n_blocks, const bmb200_vec_blocks* vecs, int threads, uint64_t slot_bytes,
                                uint32_t* desc, uint64_t* bb, uint64_t* gb, uint8_t* bit_pool, uint8_t* gap_pool, uint32_t* n_chunks, int poison)
{
    PackLayout L;
    pack_layout(n_vec, n_blocks, vecs, (unsigned)threads, L);
    if (L.rc) return L.rc;
    memcpy(desc, L.desc.data(), L.desc.size() * 4);
    memcpy(bb, L.bb.data(), L.bb.size() * 8); memcpy(gb, L.gb.data(), L.gb.size() * 8);
    const uint64_t maxcol = pack_max_column_bytes(L, n_blocks);
    if (slot_bytes < maxcol) slot_bytes = maxcol;
    if (!slot_bytes) slot_bytes = 64;
    std::vector<PackChunk> chunks;
    pack_chunks(L, n_blocks, slot_bytes, chunks);
    *n_chunks = (uint32_t)chunks.size();
    uint8_t* slot[kStageSlots];
    for (uint32_t k = 0; k < kStageSlots; ++k) { slot[k] = (uint8_t*)malloc(slot_bytes); memset(slot[k], 0xA5, slot_bytes); }   // stale bytes must not leak into the pools
    {
        PackPipeline pipe(n_vec, n_blocks, vecs, &L, &chunks, slot);
        pipe.start((unsigned)threads);
        for (uint32_t c = 0; c < chunks.size(); ++c) {
            pipe.wait_chunk(c);
            const PackChunk& ch = chunks[c];
            if (poison >= 0) {
            memcpy(bit_pool + L.bb[ch.c0] * (uint64_t)BMB200_BLOCK_BYTES, slot[c % kStageSlots], ch.bit_bytes);
            memcpy(gap_pool + L.gb[ch.c0] * 16ull, slot[c % kStageSlots] + ch.bit_bytes, ch.gap_bytes); }
            if (poison) memset(slot[c % kStageSlots], 0xA5, slot_bytes);     // the slot is recycled: whatever the next chunk does not write stays garbage
            if (c >= 1) pipe.release_through(c);                 // same protocol as the product: slot of chunk c-1 is free once copy c is queued
        }
        pipe.join();
    }
    for (uint32_t k = 0; k < kStageSlots; ++k) free(slot[k]);
    return 0;
}

extern "C" int host_pack_check(uint32_t n_vec, uint32_t n_blocks, const bmb200_vec_blocks* vecs, int threads, uint64_t slot_bytes,
                               uint32_t* desc, uint64_t* bb, uint64_t* gb, uint8_t* bit_pool, uint8_t* gap_pool, uint32_t* n_chunks)
{
    return host_pack_check2(n_vec, n_blocks, vecs, threads, slot_bytes, desc, bb, gb, bit_pool, gap_pool, n_chunks, 1);
}

/* bmb200_set_upload_slabs, host half: the mirror layout (slabs sorted by address, back to back at 256-byte steps) and the
 * per-block source table pack_sources builds.  dev_off[k] = mirror offset of INPUT slab k.  Returns 0, or 1 when a block lies
 * outside the slabs / is misaligned (the product then falls back to the packing path). */
extern "C" int host_pack_sources(uint32_t n_vec, uint32_t n_blocks, const bmb200_vec_blocks* vecs, int threads,
                                 const bmb200_host_slab* slabs, uint32_t n_slabs, uint32_t* src, uint64_t* dev_off, uint64_t* mirror_bytes)
{
    PackLayout L;
    pack_layout(n_vec, n_blocks, vecs, (unsigned)threads, L);
    if (L.rc) return L.rc;
    SlabMap M;
    std::vector<uint32_t> order;
    for (uint32_t k = 0; k < n_slabs; ++k) order.push_back(k);
    for (size_t a = 0; a < order.size(); ++a) for (size_t b = a + 1; b < order.size(); ++b)
        if ((uintptr_t)slabs[order[b]].base < (uintptr_t)slabs[order[a]].base) { uint32_t t = order[a]; order[a] = order[b]; order[b] = t; }
    uint64_t total = 0;
    for (uint32_t k : order) {
        M.base.push_back((uint64_t)(uintptr_t)slabs[k].base); M.end.push_back(M.base.back() + slabs[k].bytes);
        M.dev_off.push_back(total); dev_off[k] = total;
        total += (slabs[k].bytes + 255ull) & ~255ull;
    }
    *mirror_bytes = total;
    return pack_sources(n_vec, n_blocks, vecs, L, M, (unsigned)threads, src) ? 0 : 1;
}
