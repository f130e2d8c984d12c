/*
 * blob_host_check.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A HOST build of the product's deserialize-to-device decoder (bitmagic_b200/csrc/blob_entropy.cuh: token walk, entropy decoders,
 * bitmap assembly, run counting, bit -> GAP) so that this container -- which has the reference but no GPU -- can check that
 * exact logic against bm::deserialize (tests/test_oracle_vs_reference.py).  On the device the same functions run inside
 * blob_walk_kernel / blob_entropy_kernel with a warp as the team; here the team is one lane.  Nothing in bitmagic_b200/ uses
 * this file; the product has no CPU path.
 *
 * blob_host_check: one serialized vector -> kind[n_blocks], and for the blocks that came from entropy-coded tokens
 * (decoded[nb] = 1) the block bits (bit kinds) / the GAP words (GAP kinds) exactly as pass 2 stores them in the arena.
 * Blocks from explicit-length tokens are only walked (kind + GAP size), their payload is blob_decode_kernel's business.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../include/bmb200.h"
#include "../bitmagic_b200/csrc/blob_entropy.cuh"

using namespace bmb200;

extern "C" int blob_host_check(const uint8_t* blob, uint64_t size, uint32_t n_blocks, uint8_t* kind, uint8_t* decoded,
                               uint32_t* gap_words, uint32_t* blocks, uint16_t* gaps, uint32_t* n_entropy_tokens, uint32_t* n_segments)
{
    std::vector<uint8_t> stg(size + 64, 0);
    memcpy(stg.data(), blob, size);
    std::vector<uint32_t> bm(kEntWords);
    std::vector<uint8_t> scratch(kEntScratchBytes);
    EntCtx c; c.t.lane = 0; c.t.nl = 1; c.bm = bm.data();
    c.la = reinterpret_cast<uint16_t*>(scratch.data()); c.lb = c.la + kEntListCap; c.lc = c.lb + kEntListCap;
    c.wf = reinterpret_cast<uint32_t*>(c.lc + kEntListCap);
    // pass 1 the way bmb200_set_upload_blobs runs it: cut the BLOB at its bookmarks, walk every segment on its own, concatenate
    std::vector<EntSeg> segs;
    int rc = ent_find_segments(blob, size, 0u, 0u, segs);
    if (rc) return rc;
    const uint32_t cap = n_blocks + n_blocks / 256u + 2u;
    std::vector<BlobTok> toks(cap), seg_toks(cap);
    std::vector<uint8_t> full(n_blocks, 0);
    EntWalkOut o; o.toks = toks.data(); o.cap = cap; o.n = 0; o.full = full.data(); o.full_stride = 1;
    for (const EntSeg& sg : segs) {
        EntWalkOut so; so.toks = seg_toks.data(); so.cap = cap; so.n = 0; so.full = full.data(); so.full_stride = 1;
        rc = ent_walk_segment(c, stg.data(), sg, n_blocks, so);
        if (rc) return rc;
        for (uint32_t k = 0; k < so.n; ++k) { if (o.n >= cap) return BMB200_ERR_RANGE; toks[o.n++] = seg_toks[k]; }
    }
    if (n_segments) *n_segments = (uint32_t)segs.size();
    memset(kind, 0, n_blocks); memset(decoded, 0, n_blocks); memset(gap_words, 0, 4ull * n_blocks);
    for (uint32_t nb = 0; nb < n_blocks; ++nb) if (full[nb]) kind[nb] = BMB200_BLK_FULL;
    // single-vector arena: column nb holds at most one block (same descriptor encoding as bmb200_set_upload_blobs)
    std::vector<uint32_t> desc(n_blocks, 0); std::vector<uint64_t> bb(n_blocks + 1, 0), gb(n_blocks + 1, 0);
    std::vector<const BlobTok*> at(n_blocks, nullptr);
    // like the layout loop of bmb200_set_upload_blobs: records must arrive in strictly increasing block order (a bookmark chain that
    // lies about block indexes is a format error), super-block records aside
    int64_t last_nb = -1;
    for (uint32_t k = 0; k < o.n; ++k) {
        const BlobTok& t = toks[k];
        if ((t.type & 0xffu) == 68u && (t.type & kTokEntropy)) continue;
        if ((int64_t)t.nb <= last_nb || t.nb >= n_blocks) return BMB200_ERR_BADARG;
        last_nb = t.nb; at[t.nb] = &t;
    }
    for (uint32_t nb = 0; nb < n_blocks; ++nb) {
        uint64_t nbit = 0, ngap = 0;
        if (const BlobTok* t = at[nb]) {
            kind[nb] = (uint8_t)t->kind; gap_words[nb] = t->gap_words;
            if (t->kind == BMB200_BLK_BIT) { desc[nb] = BMB200_BLK_BIT; nbit = 1; }
            else { const uint32_t pad = t->first ? 0u : 1u; desc[nb] = BMB200_BLK_GAP | (pad ? BMB200_DESC_GAP_PAD : 0u) | BMB200_DESC_GAP_FLAT;
                   ngap = (t->gap_words + pad + BMB200_GAP_UNIT_WORDS - 1) / BMB200_GAP_UNIT_WORDS; }
        }
        bb[nb + 1] = bb[nb] + nbit; gb[nb + 1] = gb[nb] + ngap;
    }
    std::vector<uint32_t> bit_pool((size_t)(bb[n_blocks] + 1) * kEntWords, 0);
    std::vector<uint16_t> gap_pool((size_t)(gb[n_blocks] + 1) * BMB200_GAP_UNIT_WORDS + 64, 0);
    EntSetView sv{1u, n_blocks, desc.data(), bb.data(), gb.data()};
    uint32_t n_ent = 0;
    for (uint32_t k = 0; k < o.n; ++k) {
        const BlobTok& t = toks[k];
        if (!(t.type & kTokEntropy)) { if (t.type == kTokSbMember && t.nb < n_blocks) decoded[t.nb] = 1; continue; }
        ++n_ent;
        const uint32_t code = t.type & 0xffu;
        uint64_t dst; uint32_t aux2 = 0;
        if (code == 68u) dst = t.aux;
        else if (t.kind == BMB200_BLK_BIT) dst = bb[t.nb];
        else { dst = gb[t.nb]; aux2 = (t.first ? 0u : 1u) | (t.first << 1); }
        rc = ent_emit(c, stg.data(), t.off, size, code, 0u, dst, t.kind, aux2, sv, bit_pool.data(), gap_pool.data());
        if (rc) return rc;
        if (code != 68u && t.nb < n_blocks) decoded[t.nb] = 1;
    }
    for (uint32_t nb = 0; nb < n_blocks; ++nb) {
        if (!decoded[nb]) continue;
        if (kind[nb] == BMB200_BLK_BIT) memcpy(blocks + (size_t)nb * kEntWords, bit_pool.data() + bb[nb] * kEntWords, BMB200_BLOCK_BYTES);
        else if (kind[nb] == BMB200_BLK_GAP) {
            const uint16_t* u = gap_pool.data() + gb[nb] * BMB200_GAP_UNIT_WORDS;
            const uint32_t pad = (desc[nb] & BMB200_DESC_GAP_PAD) ? 1u : 0u;
            if (pad && u[0] != 0xffffu) return 1000;                       /* the lead pad of the flat form */
            const uint32_t len = u[pad] >> 3;
            if (len + 1u != gap_words[nb]) return 1001;                    /* pass 1 and pass 2 disagree on the GAP size */
            memcpy(gaps + (size_t)nb * BMB200_GAP_MAX_WORDS, u + pad, 2ull * (len + 1u));
        }
    }
    if (n_entropy_tokens) *n_entropy_tokens = n_ent;
    return 0;
}
