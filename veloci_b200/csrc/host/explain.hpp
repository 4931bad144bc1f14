// `explain` on the device side: feeds the walk of host/explain_walk.hpp.  The request's bare parts are matched and scored on the
// device in a batch of their own (as for parts with a per-part top), posting_lookup_kernel finds the posting weight of every
// (matched term, returned anchor) pair.  Not for sharded handles (an anchor's postings live on one shard) and not for plans
// imported from another process (the request itself does not travel with the plan).
#pragma once
#include "engine.hpp"
#include "explain_walk.hpp"

namespace vexplain {

class Explainer {
  public:
    Explainer(vdev::Batch& b, uint32_t q) {
        const vplan::RequestPlan& rp = b.plan.requests[q];
        vdev::DeviceIndex& ix = *b.ix;
        if (rp.status != 0) throw vplan::InvalidRequest("request failed: " + rp.message);
        if (!rp.explain) throw vplan::Unsupported(rp.explain_asked ? "the explanations of a plan imported from another process are not carried" : "the request does not ask for explain");
        if (ix.n_shards > 1) throw vplan::Unsupported("explain on a sharded index is outside the accelerated path");
        uint64_t num_hits = 0;
        const uint32_t cap = (uint32_t)std::min<uint64_t>(rp.top, vdev::kMaxKLarge);
        std::vector<vdev::vgpu_hit_pod> hits((size_t)cap + 1);
        const uint32_t n = b.result(q, &num_hits, hits.data(), cap);
        std::vector<uint32_t> anchors;
        for (uint32_t i = 0; i < n; ++i) anchors.push_back(hits[i].id);
        walk_.reset(new Walk(*ix.host, *rp.explain, anchors));
        if (anchors.empty() || walk_->n_parts() == 0) return;
        std::vector<vhost::SearchPart> bare;
        for (size_t i = 0; i < walk_->n_parts(); ++i) bare.push_back(walk_->bare_part(i));
        vdev::Batch own;
        std::vector<uint32_t> ids;
        own.prepare_parts(&ix, bare, &ids);
        own.run_match();
        vdev::DevBuf<uint32_t> d_terms, d_anchors;
        vdev::DevBuf<float> d_weight;
        d_anchors.upload(anchors);
        for (size_t i = 0; i < walk_->n_parts(); ++i) {
            std::vector<vdev::TermHit> raw;
            own.download_part_hits(ids[i], raw);
            const std::vector<vdev::TermHit>& final_hits = walk_->set_hits(i, std::move(raw));
            std::string path = walk_->part(i).path;
            if (!vfmt::ends_with(path, ".textindex")) path += ".textindex";
            const vdev::PostingsDev& pd = ix.postings.at(path + ".to_anchor_id_score");
            std::vector<uint32_t> terms;
            uint64_t postings = 0;
            for (const vdev::TermHit& h : final_hits) {
                terms.push_back(h.id);
                if (h.id < pd.n_terms) postings += pd.h_off[h.id + 1] - pd.h_off[h.id];
            }
            std::vector<float> weight(terms.size() * anchors.size(), -1.0f);
            if (!terms.empty()) {
                d_terms.upload(terms);
                d_weight.alloc(weight.size());
                vdev::launch_posting_lookup(own.stream, pd.view(), d_terms.p, (uint32_t)terms.size(), d_anchors.p, (uint32_t)anchors.size(), d_weight.p);
                VDEV_CUDA(cudaStreamSynchronize(own.stream));
                VDEV_CUDA(cudaMemcpy(weight.data(), d_weight.p, weight.size() * sizeof(float), cudaMemcpyDeviceToHost));
            }
            walk_->set_weights(i, std::move(weight), postings);
        }
    }
    Walk& walk() { return *walk_; }

  private:
    std::unique_ptr<Walk> walk_;
};

}  // namespace vexplain
