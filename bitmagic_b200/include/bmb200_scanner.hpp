// bmb200_scanner.hpp -- reference-side binding: bm::b200::scanner<SV>, the GPU counterpart of
// bm::sparse_vector_scanner<SV> (src/bmsparsevec_algo.h:1040-1182) for unsigned bm::sparse_vector<>.
//
// The scanner binds to one sparse vector (like sparse_vector_scanner::bind, :1060): its bit-planes
// (sparse_vector::get_slice(j), j < effective_slices()) and its searchable universe (the NOT-NULL plane of a nullable
// vector, else [0, size())) are uploaded ONCE as vectors of a bmb200 set; every search is then one bmb200_scan launch
// (csrc/scan_kernel.cuh: one pass over the planes per block column and search value).  Same method names and
// argument meaning as the reference: find_eq / find_gt / find_ge / find_lt / find_le / find_range / find_zero /
// find_nonzero, result written into the caller's bvector (replaced, like bv_out.clear() + search).
// Batched forms take arrays of values and return one vector / count per value (the scanner's pipeline mode, :1399-1431).
#ifndef BMB200_SCANNER_HPP_INCLUDED
#define BMB200_SCANNER_HPP_INCLUDED

#include "bmsparsevec.h"
#include "bmsparsevec_algo.h"
#include "bmb200_aggregator.hpp"

namespace bm { namespace b200 {

template<class SV>
class scanner
{
public:
    typedef typename SV::bvector_type bvector_type;
    typedef typename SV::value_type value_type;
    typedef typename bvector_type::size_type size_type;
    static_assert(!std::is_signed<value_type>::value, "bm::b200::scanner: unsigned sparse vectors only");

    scanner(context& c, const SV& sv) : ctx_(c) { bind(sv); }
    ~scanner() { if (set_) bmb200_set_free(set_); }
    scanner(const scanner&) = delete; scanner& operator=(const scanner&) = delete;

    /// upload the planes + universe of `sv` (sparse_vector_scanner::bind, :1060)
    void bind(const SV& sv)
    {
        if (set_) { bmb200_set_free(set_); set_ = nullptr; }
        size_ = sv.size();
        n_planes_ = sv.effective_slices();
        while (n_planes_ > 1 && !sv.get_slice(n_planes_ - 1)) --n_planes_;
        n_blocks_ = (uint32_t)((uint64_t(size_) + 65535ull) >> 16); if (!n_blocks_) n_blocks_ = 1;
        const bvector_type* nn = sv.get_null_bvector();
        if (nn) universe_ = *nn;                                     // NOT-NULL plane: finalize_search_result(:2426)
        else { universe_.clear(true); if (size_) universe_.set_range(0, size_ - 1); universe_.optimize(); }   // invert_internal(:1686) range
        std::vector<detail::tree_view<bvector_type>> views(n_planes_ + 1);
        std::vector<bmb200_vec_blocks> vb(n_planes_ + 1);
        bvector_type empty;
        for (unsigned j = 0; j <= n_planes_; ++j) {
            const bvector_type* bv = j < n_planes_ ? sv.get_slice(j) : &universe_;
            views[j].build(bv ? *bv : empty, n_blocks_);
            vb[j].n_blocks = n_blocks_; vb[j].kind = views[j].kind.data(); vb[j].ptr = views[j].ptr.data();
        }
        check(bmb200_set_upload_vectors(ctx_.get(), n_planes_ + 1, n_blocks_, vb.data(), &set_), "bmb200_set_upload_vectors");
    }

    void find_eq(value_type v, bvector_type& bv_out) { one(BMB200_SCAN_EQ, v, 0, bv_out); }
    void find_gt(value_type v, bvector_type& bv_out) { one(BMB200_SCAN_GT, v, 0, bv_out); }
    void find_ge(value_type v, bvector_type& bv_out) { one(BMB200_SCAN_GE, v, 0, bv_out); }
    void find_lt(value_type v, bvector_type& bv_out) { one(BMB200_SCAN_LT, v, 0, bv_out); }
    void find_le(value_type v, bvector_type& bv_out) { one(BMB200_SCAN_LE, v, 0, bv_out); }
    void find_range(value_type from, value_type to, bvector_type& bv_out) { one(BMB200_SCAN_RANGE, from, to, bv_out); }
    void find_zero(bvector_type& bv_out)    { one(BMB200_SCAN_EQ, 0, 0, bv_out); }
    void find_nonzero(bvector_type& bv_out) { one(BMB200_SCAN_GT, 0, 0, bv_out); }

    /// batched searches: one launch, out[k] replaced by the result of values[k]
    void find_batch(int pred, const std::vector<uint64_t>& values, std::vector<bvector_type>& out)
    {
        const size_t nv = pred == BMB200_SCAN_RANGE ? values.size() / 2 : values.size();
        out.resize(nv);
        run(pred, values.data(), (uint32_t)nv, out.data(), nullptr);
    }
    /// cardinalities only (pipeline agg_opt_only_counts)
    void count_batch(int pred, const std::vector<uint64_t>& values, std::vector<size_type>& counts)
    {
        const size_t nv = pred == BMB200_SCAN_RANGE ? values.size() / 2 : values.size();
        counts.assign(nv, 0);
        run(pred, values.data(), (uint32_t)nv, nullptr, counts.data());
    }

private:
    void one(int pred, value_type a, value_type b, bvector_type& bv_out)
    {
        const uint64_t v[2] = {uint64_t(a), uint64_t(b)};
        run(pred, v, 1, &bv_out, nullptr);
    }
    void run(int pred, const uint64_t* values, uint32_t nv, bvector_type* out, size_type* counts)
    {
        if (!nv) return;
        bmb200_scan_args a{0u, n_planes_, n_planes_, pred, out ? BMB200_F_OPT_COMPRESS : BMB200_F_COUNT_ONLY, values, nv, 0u, 0u};
        bmb200_result* res = nullptr;
        int rc = bmb200_scan(ctx_.get(), set_, &a, &res);
        const size_t ncols = (size_t)nv * n_blocks_;
        std::vector<uint64_t> totals(nv);
        std::vector<uint8_t> kind(ncols); std::vector<uint64_t> off(ncols); std::vector<uint32_t> bits; std::vector<uint16_t> gaps;
        if (!rc) rc = bmb200_result_group_totals(res, totals.data(), nv);
        if (!rc && out) { uint64_t nb = 0, ng = 0; rc = bmb200_result_sizes(res, &nb, &ng);
                          if (!rc) { bits.resize(nb * BMB200_BLOCK_WORDS); gaps.resize(ng);
                                     rc = bmb200_result_fetch(res, kind.data(), off.data(), bits.data(), gaps.data()); } }
        if (res) bmb200_result_free(res);
        check(rc, "bmb200_scan");
        for (uint32_t k = 0; k < nv; ++k) {
            if (counts) counts[k] = (size_type)totals[k];
            if (out) detail::store_result(out[k], size_, n_blocks_, kind.data() + (size_t)k * n_blocks_, off.data() + (size_t)k * n_blocks_,
                                          bits.data(), gaps.data());
        }
    }

    context& ctx_;
    bmb200_set* set_ = nullptr;
    bvector_type universe_;
    size_type size_ = 0;
    unsigned n_planes_ = 0;
    uint32_t n_blocks_ = 0;
};

}} // namespace bm::b200

#endif
