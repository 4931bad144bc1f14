// extern "C" boundary of libveloci_b200.so (include/veloci_b200.h).  No exception
// crosses it: every entry point maps failures to a VGPU_ERR_* status and leaves the
// message in a thread-local buffer (vgpu_last_error).
#include <malloc.h>
#include "../../../include/veloci_b200.h"

#include <atomic>
#include <cstring>
#include <string>

#include "engine.hpp"
#include "plan_channel.hpp"
#include "query_generator.hpp"
#include "steps.hpp"

namespace vdev {
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
uint64_t launches_so_far() { return g_launches.load(std::memory_order_relaxed); }
}  // namespace vdev

struct vgpu_index {
    std::unique_ptr<vdev::DeviceIndex> ix;
};
struct vgpu_plan_channel {
    std::unique_ptr<vplan::PlanChannel> ch;
};
struct vgpu_batch {
    vdev::Batch b;
    std::vector<std::vector<vsteps::FacetGroups>> facets;  // per request, filled on demand
    std::vector<char> jsonl;                                // vgpu_batch_prepare_jsonl: the request lines
    std::string kernel_times;
};

static thread_local std::string t_error;

// An unchecked runtime call that failed harmlessly leaves its code in the thread's "last error" slot, where the next
// cudaGetLastError() of an unrelated call would find it: every entry point starts and ends with the slot empty.
static void clear_stale_cuda_error(const char* where) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && getenv("VELOCI_DEBUG")) fprintf(stderr, "[veloci] stale CUDA error %s: %s\n", where, cudaGetErrorString(e));
}

template <class F>
static int32_t guarded(F&& f) {
    try {
        clear_stale_cuda_error("found at entry");
        f();
        clear_stale_cuda_error("left by this call");
        return VGPU_OK;
    } catch (const vplan::InvalidRequest& e) {
        t_error = e.what();
        return VGPU_ERR_INVALID_REQUEST;
    } catch (const vhost::FstNotFound& e) {
        t_error = e.what();
        return VGPU_ERR_FIELD_NOT_FOUND;
    } catch (const vhost::PathNotFound& e) {
        t_error = e.what();
        return VGPU_ERR_PATH_NOT_FOUND;
    } catch (const vhost::IoError& e) {
        t_error = e.what();
        return VGPU_ERR_IO;
    } catch (const vhost::MissingTextId& e) {
        t_error = e.what();
        return VGPU_ERR_IO;
    } catch (const vhost::RequestError& e) {
        t_error = e.what();
        return VGPU_ERR_JSON;
    } catch (const vjson::ParseError& e) {
        t_error = e.what();
        return VGPU_ERR_JSON;
    } catch (const vdev::CudaError& e) {
        t_error = e.what();
        return VGPU_ERR_CUDA;
    } catch (const vdev::NcclError& e) {
        t_error = e.what();
        return VGPU_ERR_NCCL;
    } catch (const vplan::BlobError& e) {
        t_error = e.what();
        return VGPU_ERR_INVALID_REQUEST;
    } catch (const vplan::Unsupported& e) {
        t_error = e.what();
        return VGPU_ERR_UNSUPPORTED;
    } catch (const vquery::ParseError& e) {
        t_error = e.what();
        return VGPU_ERR_INVALID_REQUEST;
    } catch (const vquery::GeneratorError& e) {
        t_error = e.what();
        return t_error.compare(0, 6, "Field ") == 0 ? VGPU_ERR_FIELD_NOT_FOUND : VGPU_ERR_INVALID_REQUEST;
    } catch (const vquery::ParamsError& e) {
        t_error = e.what();
        return VGPU_ERR_JSON;
    } catch (const std::exception& e) {
        t_error = e.what();
        return VGPU_ERR_INTERNAL;
    } catch (...) {
        t_error = "unknown error";
        return VGPU_ERR_INTERNAL;
    }
}

struct vgpu_hitlist_dev {
    vdev::DeviceHitList list;
    int device = -1;  // where its buffer lives: kept apart from the index, which may already be closed when the handle is freed
};
template <class F>
static int32_t make_dev_list(vgpu_index* idx, vgpu_hitlist_dev** out, F&& fill) {
    *out = nullptr;
    std::unique_ptr<vgpu_hitlist_dev> h(new vgpu_hitlist_dev());
    h->device = idx->ix->device;
    const int32_t rc = guarded([&]() { fill(h->list); });
    if (rc == VGPU_OK) *out = h.release();
    else cudaSetDevice(h->device);  // (its buffers go back to that device's pool)
    return rc;
}

extern "C" {

const char* vgpu_last_error(void) { return t_error.c_str(); }

int32_t vgpu_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

// The planner allocates and frees tens of megabytes per batch: keep that memory in the process instead of returning it
// to the kernel (and page-faulting it back in) every time.
static void tune_allocator_once() {
    static bool done = false;
    if (done) return;
    done = true;
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_ARENA_MAX, 32);
}

int32_t vgpu_index_open(const char* dir, int32_t device, uint32_t shard_rank, uint32_t n_shards, vgpu_index** out) {
    return vgpu_index_open_ex(dir, device, shard_rank, n_shards, 0, out);
}

int32_t vgpu_index_open_ex(const char* dir, int32_t device, uint32_t shard_rank, uint32_t n_shards, uint32_t flags, vgpu_index** out) {
    tune_allocator_once();
    if (!dir || !out) {
        t_error = "null argument";
        return VGPU_ERR_INVALID_REQUEST;
    }
    *out = nullptr;
    return guarded([&]() {
        std::unique_ptr<vgpu_index> h(new vgpu_index());
        h->ix = vdev::DeviceIndex::open(dir, device, shard_rank, n_shards, flags);
        *out = h.release();
    });
}

void vgpu_index_close(vgpu_index* idx) {
    if (!idx) return;
    if (idx->ix) cudaSetDevice(idx->ix->device);
    delete idx;
    vdev::ScratchPool::instance().trim();
    clear_stale_cuda_error("left by vgpu_index_close");
}

int32_t vgpu_index_info(const vgpu_index* idx, uint64_t* num_docs, uint64_t* anchor_lo, uint64_t* anchor_hi, uint64_t* device_bytes) {
    if (!idx) return VGPU_ERR_INVALID_REQUEST;
    if (num_docs) *num_docs = idx->ix->num_docs;
    if (anchor_lo) *anchor_lo = idx->ix->anchor_lo;
    if (anchor_hi) *anchor_hi = idx->ix->anchor_hi;
    if (device_bytes) *device_bytes = idx->ix->device_bytes;
    return VGPU_OK;
}

int32_t vgpu_batch_prepare(vgpu_index* idx, const char* const* request_json, uint32_t n, vgpu_batch** out) {
    if (!idx || !out || (n && !request_json)) {
        t_error = "null argument";
        return VGPU_ERR_INVALID_REQUEST;
    }
    *out = nullptr;
    return guarded([&]() {
        std::unique_ptr<vgpu_batch> b(new vgpu_batch());
        b->b.prepare(idx->ix.get(), request_json, n);
        *out = b.release();
    });
}

static int32_t prepare_lines(vgpu_index* idx, const char* text, size_t len, uint32_t n, bool upload, vgpu_batch** out);

int32_t vgpu_batch_prepare_lines(vgpu_index* idx, const char* text, size_t len, uint32_t n, vgpu_batch** out) { return prepare_lines(idx, text, len, n, true, out); }

static int32_t prepare_lines(vgpu_index* idx, const char* text, size_t len, uint32_t n, bool upload, vgpu_batch** out) {
    if (!idx || !out || (len && !text)) {
        t_error = "null argument";
        return VGPU_ERR_INVALID_REQUEST;
    }
    *out = nullptr;
    size_t feeds = 0;
    for (const char* p = text; p && (p = static_cast<const char*>(memchr(p, '\n', (size_t)(text + len - p)))) != nullptr; ++p) ++feeds;
    if (feeds + 1 != (size_t)n && !(n == 0 && len == 0)) {
        t_error = "the buffer does not hold n requests separated by single line feeds";
        return VGPU_ERR_INVALID_REQUEST;
    }
    return guarded([&]() {
        std::unique_ptr<vgpu_batch> b(new vgpu_batch());
        b->jsonl.assign(text, text + len);
        b->jsonl.push_back('\0');
        std::vector<const char*> lines;
        lines.reserve(n);
        char* p = b->jsonl.data();
        char* end = p + len;
        for (uint32_t i = 0; i < n; ++i) {
            char* nl = static_cast<char*>(memchr(p, '\n', (size_t)(end - p)));
            if (!nl) nl = end;
            *nl = '\0';
            lines.push_back(p);
            p = nl < end ? nl + 1 : end;
        }
        b->b.prepare(idx->ix.get(), lines.data(), n, upload);
        *out = b.release();
    });
}

int32_t vgpu_batch_prepare_jsonl(vgpu_index* idx, const char* text, size_t len, uint32_t* n_out, vgpu_batch** out) {
    if (!idx || !out || (len && !text)) {
        t_error = "null argument";
        return VGPU_ERR_INVALID_REQUEST;
    }
    *out = nullptr;
    return guarded([&]() {
        // one request per line (JSON escapes line feeds inside strings); empty lines are skipped.  The lines are
        // terminated in a private copy so that every request is a C string.
        std::unique_ptr<vgpu_batch> b(new vgpu_batch());
        b->jsonl.assign(text, text + len);
        b->jsonl.push_back('\0');
        std::vector<const char*> lines;
        char* p = b->jsonl.data();
        char* end = p + len;
        while (p < end) {
            char* nl = static_cast<char*>(memchr(p, '\n', (size_t)(end - p)));
            if (!nl) nl = end;
            *nl = '\0';
            if (nl > p) lines.push_back(p);
            p = nl + 1;
        }
        if (lines.size() > 0xFFFFFFF0ull) throw std::runtime_error("too many requests in one batch");
        b->b.prepare(idx->ix.get(), lines.data(), (uint32_t)lines.size());
        if (n_out) *n_out = (uint32_t)lines.size();
        *out = b.release();
    });
}

// ---- communicator, plan export / import, plan channel
int32_t vgpu_comm_unique_id(uint8_t id[VGPU_COMM_ID_BYTES]) {
    if (!id) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        vdev::NcclApi& api = vdev::NcclApi::get();
        api.require();
        vdev::NcclUniqueId uid;
        api.check(api.GetUniqueId(&uid), "ncclGetUniqueId");
        static_assert(sizeof uid.internal == VGPU_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
        memcpy(id, uid.internal, VGPU_COMM_ID_BYTES);
    });
}

int32_t vgpu_comm_init(vgpu_index* idx, const uint8_t id[VGPU_COMM_ID_BYTES]) {
    if (!idx || !id) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        vdev::NcclApi& api = vdev::NcclApi::get();
        api.require();
        VDEV_CUDA(cudaSetDevice(idx->ix->device));
        std::unique_ptr<vdev::ShardComm> c(new vdev::ShardComm());
        c->rank = idx->ix->shard_rank, c->n_ranks = idx->ix->n_shards;
        vdev::NcclUniqueId uid;
        memcpy(uid.internal, id, VGPU_COMM_ID_BYTES);
        api.check(api.CommInitRank(&c->comm, (int)c->n_ranks, uid, (int)c->rank), "ncclCommInitRank");
        idx->ix->comm = std::move(c);
    });
}

int32_t vgpu_comm_destroy(vgpu_index* idx) {
    if (!idx) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        if (idx->ix->comm) VDEV_CUDA(cudaSetDevice(idx->ix->device));
        idx->ix->comm.reset();
    });
}

int32_t vgpu_batch_export_plan(const vgpu_batch* batch, void** blob, size_t* len) {
    if (!batch || !blob || !len) return VGPU_ERR_INVALID_REQUEST;
    *blob = nullptr, *len = 0;
    return guarded([&]() {
        if (batch->b.mode != vdev::Batch::kRequests) throw vplan::BlobError("only request batches have a plan to export");
        const std::vector<uint8_t> bytes = vplan::export_plan(batch->b.plan);
        void* p = malloc(std::max<size_t>(1, bytes.size()));
        if (!p) throw std::bad_alloc();
        memcpy(p, bytes.data(), bytes.size());
        *blob = p, *len = bytes.size();
    });
}

int32_t vgpu_batch_prepare_from_plan(vgpu_index* idx, const void* blob, size_t len, vgpu_batch** out) {
    if (!idx || !blob || !out) {
        t_error = "null argument";
        return VGPU_ERR_INVALID_REQUEST;
    }
    *out = nullptr;
    return guarded([&]() {
        std::unique_ptr<vgpu_batch> b(new vgpu_batch());
        b->b.prepare_from_blob(idx->ix.get(), blob, len);
        *out = b.release();
    });
}

int32_t vgpu_plan_channel_open(const char* name, uint32_t local_rank, uint32_t local_ranks, size_t capacity, vgpu_plan_channel** out) {
    if (!name || !out) return VGPU_ERR_INVALID_REQUEST;
    *out = nullptr;
    return guarded([&]() {
        std::unique_ptr<vgpu_plan_channel> h(new vgpu_plan_channel());
        h->ch.reset(new vplan::PlanChannel(name, local_rank, local_ranks, capacity));
        *out = h.release();
    });
}

void vgpu_plan_channel_close(vgpu_plan_channel* ch) { delete ch; }

uint64_t vgpu_plan_channel_ticket(vgpu_plan_channel* ch) { return ch ? ch->ch->ticket() : 0; }

int32_t vgpu_batch_prepare_shared(vgpu_index* idx, vgpu_plan_channel* ch, uint64_t ticket, const char* text, size_t len, uint32_t n, vgpu_batch** out) {
    if (!idx || !ch || !out || ticket == 0) {
        t_error = "null argument";
        return VGPU_ERR_INVALID_REQUEST;
    }
    *out = nullptr;
    if (ch->ch->rank() == 0) {
        // plan here, publish the plan -- or the failure, so that the other ranks fail with it instead of waiting
        vgpu_batch* b = nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        vdev::plan_for_all_ranks() = true;
        int32_t rc = prepare_lines(idx, text, len, n, false, &b);  // the other ranks get the plan before this one uploads its own copy
        vdev::plan_for_all_ranks() = false;
        const auto t1 = std::chrono::steady_clock::now();
        std::vector<uint8_t> bytes;
        if (rc == VGPU_OK)
            rc = guarded([&]() {
                bytes = vplan::export_plan(b->b.plan);
                if (bytes.size() > ch->ch->capacity()) throw vplan::ChannelError("the plan (" + std::to_string(bytes.size()) + " bytes) exceeds the channel capacity");
            });
        const auto t2 = std::chrono::steady_clock::now();
        const std::string failure = t_error;
        const int32_t pub = guarded([&]() {
            if (rc == VGPU_OK) ch->ch->publish(ticket, bytes.data(), bytes.size(), 0);
            else ch->ch->publish(ticket, failure.data(), std::min(failure.size(), ch->ch->capacity()), (uint64_t)rc);
        });
        if (getenv("VELOCI_DEBUG")) {
            auto ms = [](auto a, auto b2) { return std::chrono::duration<double, std::milli>(b2 - a).count(); };
            fprintf(stderr, "[veloci] prepare_shared rank 0: prepare %.2f ms, export %.2f ms (%zu bytes), publish %.2f ms\n", ms(t0, t1), ms(t1, t2), bytes.size(), ms(t2, std::chrono::steady_clock::now()));
        }
        if (rc == VGPU_OK) rc = pub;
        else t_error = failure;
        if (rc == VGPU_OK) rc = guarded([&]() { b->b.upload_plan(); });
        if (rc != VGPU_OK) {
            vgpu_batch_free(b);
            return rc;
        }
        *out = b;
        return VGPU_OK;
    }
    int32_t remote = VGPU_OK;
    std::string remote_msg;
    const auto t0 = std::chrono::steady_clock::now();
    auto t1 = t0;
    const int32_t rc = guarded([&]() {
        std::unique_ptr<vgpu_batch> b(new vgpu_batch());
        ch->ch->consume(ticket, [&](const uint8_t* data, size_t blob_len, uint64_t status) {
            t1 = std::chrono::steady_clock::now();
            if (status != 0) {
                remote = (int32_t)status, remote_msg.assign(reinterpret_cast<const char*>(data), blob_len);
                return;
            }
            b->b.prepare_from_blob(idx->ix.get(), data, blob_len);
        });
        if (remote == VGPU_OK) *out = b.release();
    });
    if (getenv("VELOCI_DEBUG")) {
        auto ms = [](auto a, auto b2) { return std::chrono::duration<double, std::milli>(b2 - a).count(); };
        fprintf(stderr, "[veloci] prepare_shared rank %u: waited %.2f ms, import + upload %.2f ms\n", ch->ch->rank(), ms(t0, t1), ms(t1, std::chrono::steady_clock::now()));
    }
    if (rc != VGPU_OK) return rc;
    if (remote != VGPU_OK) {
        t_error = "local rank 0 could not plan the batch: " + remote_msg;
        return remote;
    }
    return VGPU_OK;
}

int32_t vgpu_batch_execute(vgpu_batch* batch) {
    if (!batch) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        batch->facets.clear();
        batch->b.execute();
    });
}

int32_t vgpu_batch_execute_begin(vgpu_batch* batch) {
    if (!batch) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        batch->facets.clear();
        batch->b.execute_begin();
    });
}

int32_t vgpu_batch_thresholds(const vgpu_batch* batch, uint64_t** tau_dev, uint32_t* n) {
    if (!batch || !tau_dev || !n) return VGPU_ERR_INVALID_REQUEST;
    *tau_dev = reinterpret_cast<uint64_t*>(batch->b.d_tau.p);
    *n = batch->b.n;
    return VGPU_OK;
}

int32_t vgpu_batch_facet_histograms(const vgpu_batch* batch, uint32_t** hist_dev, uint64_t* n) {
    if (!batch || !hist_dev || !n) return VGPU_ERR_INVALID_REQUEST;
    *hist_dev = batch->b.n_facets ? batch->b.d_facet_hist.p : nullptr;
    *n = batch->b.n_facets ? batch->b.d_facet_hist.n : 0;
    return VGPU_OK;
}

int32_t vgpu_batch_execute_finish(vgpu_batch* batch) {
    if (!batch) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { batch->b.execute_finish(); });
}

void vgpu_batch_free(vgpu_batch* batch) {
    if (!batch) return;
    if (batch->b.device >= 0) cudaSetDevice(batch->b.device);  // (the batch's own copy: its index may be closed already)
    delete batch;
    clear_stale_cuda_error("left by vgpu_batch_free");
}

int32_t vgpu_batch_size(const vgpu_batch* batch, uint32_t* n) {
    if (!batch || !n) return VGPU_ERR_INVALID_REQUEST;
    *n = batch->b.n;
    return VGPU_OK;
}

int32_t vgpu_batch_status(const vgpu_batch* batch, uint32_t q) {
    if (!batch || q >= batch->b.n) return VGPU_ERR_INVALID_REQUEST;
    return batch->b.plan.requests[q].status;
}

const char* vgpu_batch_message(const vgpu_batch* batch, uint32_t q) {
    if (!batch || q >= batch->b.n) return "";
    return batch->b.plan.requests[q].message.c_str();
}

int32_t vgpu_batch_result(const vgpu_batch* batch_c, uint32_t q, uint64_t* num_hits, vgpu_hit* hits, uint32_t cap, uint32_t* n_hits) {
    vgpu_batch* batch = const_cast<vgpu_batch*>(batch_c);
    if (!batch || q >= batch->b.n) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        uint32_t w = batch->b.result(q, num_hits, reinterpret_cast<vdev::vgpu_hit_pod*>(hits), cap);
        if (n_hits) *n_hits = w;
    });
}

int32_t vgpu_batch_results_flat(const vgpu_batch* batch_c, uint32_t k, uint32_t* ids, float* scores, uint64_t* num_hits, int32_t* status) {
    vgpu_batch* batch = const_cast<vgpu_batch*>(batch_c);
    if (!batch) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        std::vector<vdev::vgpu_hit_pod> row(k ? k : 1);
        for (uint32_t q = 0; q < batch->b.n; ++q) {
            uint64_t nh = 0;
            uint32_t w = batch->b.result(q, &nh, row.data(), k);
            if (num_hits) num_hits[q] = nh;
            if (status) status[q] = batch->b.plan.requests[q].status;
            for (uint32_t i = 0; i < k; ++i) {
                if (ids) ids[(size_t)q * k + i] = i < w ? row[i].id : 0xFFFFFFFFu;
                if (scores) scores[(size_t)q * k + i] = i < w ? row[i].score : 0.0f;
            }
        }
    });
}

int32_t vgpu_search_batch(vgpu_index* idx, const char* const* request_json, uint32_t n, uint32_t k, uint32_t* ids, float* scores, uint64_t* num_hits, int32_t* status) {
    vgpu_batch* b = nullptr;
    int32_t rc = vgpu_batch_prepare(idx, request_json, n, &b);
    if (rc != VGPU_OK) return rc;
    rc = vgpu_batch_execute(b);
    if (rc == VGPU_OK) rc = vgpu_batch_results_flat(b, k, ids, scores, num_hits, status);
    vgpu_batch_free(b);
    return rc;
}

int32_t vgpu_batch_facet_count(const vgpu_batch* batch, uint32_t q, uint32_t* n_fields) {
    if (!batch || q >= batch->b.n || !n_fields) return VGPU_ERR_INVALID_REQUEST;
    *n_fields = (uint32_t)batch->b.plan.requests[q].facets.size();
    return VGPU_OK;
}

int32_t vgpu_batch_facet(const vgpu_batch* batch_c, uint32_t q, uint32_t field, const char** field_name, uint32_t* n_groups) {
    vgpu_batch* batch = const_cast<vgpu_batch*>(batch_c);
    if (!batch || q >= batch->b.n) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        vsteps::materialize_facets(batch->b, batch->facets);
        const auto& f = batch->facets.at(q).at(field);
        if (field_name) *field_name = f.field.c_str();
        if (n_groups) *n_groups = (uint32_t)f.groups.size();
    });
}

int32_t vgpu_batch_facet_group(const vgpu_batch* batch_c, uint32_t q, uint32_t field, uint32_t group, uint32_t* value_id, uint32_t* count, const char** text) {
    vgpu_batch* batch = const_cast<vgpu_batch*>(batch_c);
    if (!batch || q >= batch->b.n) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        vsteps::materialize_facets(batch->b, batch->facets);
        const auto& g = batch->facets.at(q).at(field).groups.at(group);
        if (value_id) *value_id = g.id;
        if (count) *count = g.count;
        if (text) *text = g.text.c_str();
    });
}

int32_t vgpu_batch_local_topk(const vgpu_batch* batch, uint64_t** keys_dev, uint64_t** num_hits_dev, uint32_t* stride) {
    if (!batch || !batch->b.executed) {
        t_error = "batch was not executed";
        return VGPU_ERR_INVALID_REQUEST;
    }
    if (keys_dev) *keys_dev = batch->b.d_out_keys.p;
    if (num_hits_dev) *num_hits_dev = batch->b.d_out_hits.p;
    if (stride) *stride = batch->b.stride;
    return VGPU_OK;
}

int32_t vgpu_batch_merge_gathered(vgpu_batch* batch, const uint64_t* gathered_keys_dev, const uint64_t* gathered_num_hits_dev, uint32_t n_shards) {
    if (!batch || !gathered_keys_dev || !gathered_num_hits_dev || n_shards == 0) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { batch->b.merge_gathered(gathered_keys_dev, gathered_num_hits_dev, n_shards); });
}

void vgpu_free(void* p) { free(p); }
void vgpu_hitlist_free(vgpu_hitlist* l) {
    if (!l) return;
    free(l->hits);
    free(l->ids);
    l->hits = nullptr, l->ids = nullptr, l->n_hits = 0, l->n_ids = 0;
}

int32_t vgpu_field_search(vgpu_index* idx, const char* part_json, int32_t get_scores, int32_t get_ids, vgpu_hitlist* out) {
    if (!idx || !part_json || !out) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::field_search(*idx->ix, part_json, get_scores != 0, get_ids != 0, *out); });
}
int32_t vgpu_resolve_to_anchor(vgpu_index* idx, const char* part_json, const vgpu_hitlist* in, vgpu_hitlist* out) {
    if (!idx || !part_json || !in || !out) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::resolve_to_anchor(*idx->ix, part_json, *in, *out); });
}
int32_t vgpu_union_hits_score(vgpu_index* idx, const vgpu_hitlist* inputs, const char* const* terms, uint32_t n, vgpu_hitlist* out) {
    if (!idx || (n && (!inputs || !terms)) || !out) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::set_op(*idx->ix, inputs, terms, n, true, *out); });
}
int32_t vgpu_intersect_hits_score(vgpu_index* idx, const vgpu_hitlist* inputs, uint32_t n, vgpu_hitlist* out) {
    if (!idx || (n && !inputs) || !out) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::set_op(*idx->ix, inputs, nullptr, n, false, *out); });
}
int32_t vgpu_resolve_to_anchor_filtered(vgpu_index* idx, const char* part_json, const vgpu_hitlist* in, const uint32_t* filter_ids, uint32_t n_filter, vgpu_hitlist* out) {
    if (!idx || !part_json || !in || !out || (n_filter && !filter_ids)) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::resolve_to_anchor_filtered(*idx->ix, part_json, *in, filter_ids, n_filter, *out); });
}
int32_t vgpu_union_hits_ids(vgpu_index* idx, const vgpu_hitlist* inputs, uint32_t n, vgpu_hitlist* out) {
    if (!idx || (n && !inputs) || !out) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::set_op_ids(*idx->ix, inputs, n, true, *out); });
}
int32_t vgpu_intersect_hits_ids(vgpu_index* idx, const vgpu_hitlist* inputs, uint32_t n, vgpu_hitlist* out) {
    if (!idx || (n && !inputs) || !out) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::set_op_ids(*idx->ix, inputs, n, false, *out); });
}
int32_t vgpu_intersect_scores_with_ids(vgpu_index* idx, const vgpu_hitlist* scores, const vgpu_hitlist* ids, vgpu_hitlist* out) {
    if (!idx || !scores || !ids || !out) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::intersect_scores_with_ids(*idx->ix, *scores, *ids, *out); });
}
static void fill_suggestions(const std::vector<vsteps::Suggestion>& v, vgpu_suggestions* out);
int32_t vgpu_phrase_pairs_to_anchor(vgpu_index* idx, const char* path, const uint32_t* ids1, uint32_t n1, const uint32_t* ids2, uint32_t n2, vgpu_hitlist* out) {
    if (!idx || !path || !out || (n1 && !ids1) || (n2 && !ids2)) return VGPU_ERR_INVALID_REQUEST;
    memset(out, 0, sizeof *out);
    return guarded([&]() { vsteps::phrase_pairs_to_anchor(*idx->ix, path, ids1, n1, ids2, n2, *out); });
}
int32_t vgpu_boost_anchor_from_phrase_results(vgpu_index* idx, const vgpu_hitlist* hits, const vgpu_hitlist* phrase_results, const uint32_t* group, uint32_t n, vgpu_hitlist* out) {
    if (!idx || !hits || !out || (n && (!phrase_results || !group))) return VGPU_ERR_INVALID_REQUEST;
    memset(out, 0, sizeof *out);
    return guarded([&]() { vsteps::boost_anchor_from_phrase_results(*idx->ix, *hits, phrase_results, group, n, *out); });
}
int32_t vgpu_boost_to_anchor(vgpu_index* idx, const char* part_json, const vgpu_hitlist* in, const char* boost_json, vgpu_hitlist* out) {
    if (!idx || !part_json || !in || !boost_json || !out) return VGPU_ERR_INVALID_REQUEST;
    memset(out, 0, sizeof *out);
    return guarded([&]() { vsteps::boost_to_anchor(*idx->ix, part_json, *in, boost_json, *out); });
}
int32_t vgpu_apply_anchor_boost(vgpu_index* idx, const char* boost_json, const vgpu_hitlist* hits, const vgpu_hitlist* boost_ids, vgpu_hitlist* out) {
    if (!idx || !boost_json || !hits || !boost_ids || !out) return VGPU_ERR_INVALID_REQUEST;
    memset(out, 0, sizeof *out);
    return guarded([&]() { vsteps::apply_anchor_boost(*idx->ix, boost_json, *hits, *boost_ids, *out); });
}
int32_t vgpu_text_locality(vgpu_index* idx, const char* path, const vgpu_hitlist* term_hits, uint32_t n_terms, vgpu_hitlist* out) {
    if (!idx || !path || !out || (n_terms && !term_hits)) return VGPU_ERR_INVALID_REQUEST;
    memset(out, 0, sizeof *out);
    return guarded([&]() { vsteps::text_locality(*idx->ix, path, term_hits, n_terms, *out); });
}
int32_t vgpu_facet(vgpu_index* idx, const char* facet_json, const uint32_t* ids, uint32_t n_ids, vgpu_suggestions* out) {
    if (!idx || !facet_json || !out || (n_ids && !ids)) return VGPU_ERR_INVALID_REQUEST;
    memset(out, 0, sizeof *out);
    return guarded([&]() {
        std::vector<vsteps::Suggestion> groups;
        for (auto& g : vsteps::facet(*idx->ix, facet_json, ids, n_ids)) groups.push_back(vsteps::Suggestion{g.text, (float)g.count, g.id});
        fill_suggestions(groups, out);
    });
}
int32_t vgpu_add_boost(vgpu_index* idx, const char* boost_json, vgpu_hitlist* inout) {
    if (!idx || !boost_json || !inout) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::add_boost(*idx->ix, boost_json, *inout); });
}
int32_t vgpu_top_n(vgpu_index* idx, const vgpu_hitlist* in, uint32_t top, uint32_t skip, vgpu_hitlist* out) {
    if (!idx || !in || !out) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::top_n(*idx->ix, *in, top, skip, *out); });
}

// ---- device-resident step seam
int32_t vgpu_dev_upload(vgpu_index* idx, const vgpu_hitlist* in, vgpu_hitlist_dev** out) {
    if (!idx || !in || !out) return VGPU_ERR_INVALID_REQUEST;
    return make_dev_list(idx, out, [&](vdev::DeviceHitList& l) { vsteps::dev_upload(*idx->ix, *in, l); });
}
int32_t vgpu_dev_download(const vgpu_hitlist_dev* list, vgpu_hitlist* out) {
    if (!list || !out) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::dev_download(list->list, *out); });
}
uint32_t vgpu_dev_len(const vgpu_hitlist_dev* list) { return list ? list->list.n : 0u; }
void vgpu_dev_free(vgpu_hitlist_dev* list) {
    if (!list) return;
    if (list->device >= 0) cudaSetDevice(list->device);
    delete list;
}
int32_t vgpu_dev_resolve_to_anchor(vgpu_index* idx, const char* part_json, const vgpu_hitlist* term_hits, vgpu_hitlist_dev** out) {
    if (!idx || !part_json || !term_hits || !out) return VGPU_ERR_INVALID_REQUEST;
    return make_dev_list(idx, out, [&](vdev::DeviceHitList& l) { vsteps::dev_resolve_to_anchor(*idx->ix, part_json, *term_hits, l); });
}
static bool all_lists_given(const vgpu_hitlist_dev* const* inputs, uint32_t n) {
    if (n && !inputs) return false;
    for (uint32_t i = 0; i < n; ++i)
        if (!inputs[i]) return false;
    return true;
}
static int32_t dev_set_op(vgpu_index* idx, const vgpu_hitlist_dev* const* inputs, const char* const* terms, uint32_t n, bool is_union, vgpu_hitlist_dev** out) {
    std::vector<const vdev::DeviceHitList*> lists;
    for (uint32_t i = 0; i < n; ++i) lists.push_back(&inputs[i]->list);
    return make_dev_list(idx, out, [&](vdev::DeviceHitList& l) { vsteps::dev_set_op(*idx->ix, lists.data(), terms, n, is_union, l); });
}
int32_t vgpu_dev_union_hits_score(vgpu_index* idx, const vgpu_hitlist_dev* const* inputs, const char* const* terms, uint32_t n, vgpu_hitlist_dev** out) {
    if (!idx || !out || !all_lists_given(inputs, n) || (n && !terms)) return VGPU_ERR_INVALID_REQUEST;
    return dev_set_op(idx, inputs, terms, n, true, out);
}
int32_t vgpu_dev_intersect_hits_score(vgpu_index* idx, const vgpu_hitlist_dev* const* inputs, uint32_t n, vgpu_hitlist_dev** out) {
    if (!idx || !out || !all_lists_given(inputs, n)) return VGPU_ERR_INVALID_REQUEST;
    return dev_set_op(idx, inputs, nullptr, n, false, out);
}
int32_t vgpu_dev_add_boost(vgpu_index* idx, const char* boost_json, const vgpu_hitlist_dev* in, vgpu_hitlist_dev** out) {
    if (!idx || !boost_json || !in || !out) return VGPU_ERR_INVALID_REQUEST;
    return make_dev_list(idx, out, [&](vdev::DeviceHitList& l) { vsteps::dev_add_boost(*idx->ix, boost_json, in->list, l); });
}
int32_t vgpu_dev_top_n(vgpu_index* idx, const vgpu_hitlist_dev* in, uint32_t top, uint32_t skip, vgpu_hitlist* out) {
    if (!idx || !in || !out) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() { vsteps::dev_top_n(*idx->ix, in->list, top, skip, *out); });
}

static void fill_suggestions(const std::vector<vsteps::Suggestion>& v, vgpu_suggestions* out) {
    size_t bytes = 0;
    for (auto& sg : v) bytes += sg.text.size() + 1;
    out->items = (vgpu_suggestion*)malloc(std::max<size_t>(1, v.size()) * sizeof(vgpu_suggestion));
    out->text_block = (char*)malloc(std::max<size_t>(1, bytes));
    out->n = (uint32_t)v.size();
    size_t at = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        memcpy(out->text_block + at, v[i].text.c_str(), v[i].text.size() + 1);
        out->items[i] = vgpu_suggestion{out->text_block + at, v[i].score, v[i].id};
        at += v[i].text.size() + 1;
    }
}
int32_t vgpu_suggest(vgpu_index* idx, const char* request_json, vgpu_suggestions* out) {
    if (!idx || !request_json || !out) return VGPU_ERR_INVALID_REQUEST;
    memset(out, 0, sizeof *out);
    return guarded([&]() { fill_suggestions(vsteps::suggest(*idx->ix, vhost::read_request_json(request_json, strlen(request_json))), out); });
}
int32_t vgpu_suggest_part(vgpu_index* idx, const char* part_json, vgpu_suggestions* out) {
    if (!idx || !part_json || !out) return VGPU_ERR_INVALID_REQUEST;
    memset(out, 0, sizeof *out);
    return guarded([&]() {
        vhost::Request req;
        const vhost::SearchPart part = vsteps::parse_part(part_json);
        req.suggest = std::vector<vhost::SearchPart>{part};
        req.top = part.top, req.skip = part.skip;  // search_field.rs:224-225
        fill_suggestions(vsteps::suggest(*idx->ix, req), out);
    });
}
void vgpu_suggestions_free(vgpu_suggestions* s) {
    if (!s) return;
    free(s->items), free(s->text_block);
    s->items = nullptr, s->text_block = nullptr, s->n = 0;
}

static char* c_string(const std::string& s) {
    char* p = static_cast<char*>(malloc(s.size() + 1));
    if (!p) throw std::bad_alloc();
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
int32_t vgpu_search_query(vgpu_index* idx, const char* params_json, char** request_json) {
    if (!idx || !params_json || !request_json) return VGPU_ERR_INVALID_REQUEST;
    *request_json = nullptr;
    return guarded([&]() { *request_json = c_string(vquery::search_query_json(vquery::FieldCatalog::of(*idx->ix->host), params_json, strlen(params_json))); });
}
int32_t vgpu_suggest_query(vgpu_index* idx, const char* params_json, char** request_json) {
    if (!idx || !params_json || !request_json) return VGPU_ERR_INVALID_REQUEST;
    *request_json = nullptr;
    return guarded([&]() { *request_json = c_string(vquery::suggest_query_json(vquery::FieldCatalog::of(*idx->ix->host), params_json, strlen(params_json))); });
}
int32_t vgpu_highlight(vgpu_index* idx, const char* part_json, vgpu_suggestions* out) {
    if (!idx || !part_json || !out) return VGPU_ERR_INVALID_REQUEST;
    memset(out, 0, sizeof *out);
    return guarded([&]() { fill_suggestions(vsteps::highlight(*idx->ix, part_json), out); });
}
int32_t vgpu_get_doc(vgpu_index* idx, uint32_t doc_id, char** doc_json) {
    if (!idx || !doc_json) return VGPU_ERR_INVALID_REQUEST;
    *doc_json = nullptr;
    return guarded([&]() { *doc_json = c_string(idx->ix->host->get_doc(doc_id)); });
}
int32_t vgpu_read_doc(vgpu_index* idx, uint32_t doc_id, const char* fields_json, char** doc_json) {
    if (!idx || !fields_json || !doc_json) return VGPU_ERR_INVALID_REQUEST;
    *doc_json = nullptr;
    return guarded([&]() {
        const vjson::Value f = vjson::parse(fields_json, strlen(fields_json));
        if (!f.is_array()) throw vhost::RequestError("select must be an array");
        std::vector<std::string> fields;
        for (auto& e : f.arr) {
            if (!e.is_string()) throw vhost::RequestError("select must be an array of strings");
            fields.push_back(e.str);
        }
        *doc_json = c_string(vjson::to_string(vhost::read_data(*idx->ix->host, doc_id, fields)));
    });
}
int32_t vgpu_batch_result_docs(vgpu_batch* batch, uint32_t q, char** result_json) {
    if (!batch || !result_json || q >= batch->b.n) return VGPU_ERR_INVALID_REQUEST;
    *result_json = nullptr;
    if (batch->b.plan.requests[q].status != 0) {
        t_error = batch->b.plan.requests[q].message;
        return batch->b.plan.requests[q].status;
    }
    return guarded([&]() { *result_json = c_string(vsteps::result_docs(batch->b, q)); });
}
int32_t vgpu_batch_explain(vgpu_batch* batch, uint32_t q, char** explain_json) {
    if (!batch || !explain_json || q >= batch->b.n) return VGPU_ERR_INVALID_REQUEST;
    *explain_json = nullptr;
    if (batch->b.plan.requests[q].status != 0) {
        t_error = batch->b.plan.requests[q].message;
        return batch->b.plan.requests[q].status;
    }
    return guarded([&]() { *explain_json = c_string(vexplain::Explainer(batch->b, q).walk().to_json()); });
}
int32_t vgpu_explain_plan(const char* request_json, char** dot) {
    if (!request_json || !dot) return VGPU_ERR_INVALID_REQUEST;
    *dot = nullptr;
    return guarded([&]() { *dot = c_string(vhost::explain_plan(vhost::read_request_json(request_json, strlen(request_json)))); });
}
int32_t vgpu_query_parse(const char* text, uint32_t options, char** tree_debug) {
    if (!text || !tree_debug) return VGPU_ERR_INVALID_REQUEST;
    *tree_debug = nullptr;
    return guarded([&]() {
        vquery::ParserOptions o;
        o.no_attributes = options & 1u, o.no_parentheses = options & 2u, o.no_levensthein = options & 4u;
        *tree_debug = c_string(vquery::parse(text, o).debug());
    });
}

uint64_t vgpu_launch_count(void) { return vdev::launches_so_far(); }

int32_t vgpu_batch_phase_ms(const vgpu_batch* batch, float* ms, uint32_t n_phases) {
    if (!batch || !ms) return VGPU_ERR_INVALID_REQUEST;
    for (uint32_t i = 0; i < n_phases; ++i) ms[i] = i < (uint32_t)vdev::kPhases ? batch->b.phase_ms[i] : 0.0f;
    return VGPU_OK;
}

int32_t vgpu_batch_traffic_model(const vgpu_batch* batch_c, uint64_t* posting_bytes, uint64_t* boost_bytes, uint64_t* postings, uint64_t* union_hits) {
    vgpu_batch* batch = const_cast<vgpu_batch*>(batch_c);
    if (!batch) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        batch->b.fetch();
        // 6 B per posting expanded (+ 8 B of offsets per matched term),
        // 4 B of boost column per unioned hit and boost step, 8 B per returned hit
        uint64_t boost_steps = 0;
        for (uint32_t q = 0; q < batch->b.n; ++q) boost_steps += (uint64_t)batch->b.plan.programs[q].n_boosts * batch->b.h_hits[q];
        if (posting_bytes) *posting_bytes = batch->b.stat_postings * 6ull + batch->b.stat_matches * 8ull;
        if (boost_bytes) *boost_bytes = boost_steps * 4ull;
        if (postings) *postings = batch->b.stat_postings;
        if (union_hits) *union_hits = batch->b.stat_union;
    });
}

int32_t vgpu_batch_path_stats(const vgpu_batch* batch_c, uint64_t* plane_items, uint64_t* general_items, uint64_t* plane_evaluated) {
    vgpu_batch* batch = const_cast<vgpu_batch*>(batch_c);
    if (!batch) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        batch->b.fetch();
        if (plane_items) *plane_items = batch->b.stat_fast_items;
        if (general_items) *general_items = batch->b.stat_general_items;
        if (plane_evaluated) *plane_evaluated = batch->b.stat_plane_evaluated;
    });
}

int32_t vgpu_batch_set_profiling(vgpu_batch* batch, int32_t on) {
    if (!batch) return VGPU_ERR_INVALID_REQUEST;
    batch->b.profiling = on != 0;
    return VGPU_OK;
}

const char* vgpu_batch_kernel_times_json(vgpu_batch* batch) {
    if (!batch) return "{}";
    try {
        batch->kernel_times = batch->b.kernel_times_json();
    } catch (...) {
        batch->kernel_times = "{}";
    }
    return batch->kernel_times.c_str();
}

int32_t vgpu_batch_work_stats(const vgpu_batch* batch_c, uint64_t* out, uint32_t n) {
    vgpu_batch* batch = const_cast<vgpu_batch*>(batch_c);
    if (!batch || !out) return VGPU_ERR_INVALID_REQUEST;
    return guarded([&]() {
        batch->b.fetch();
        const vdev::Batch& b = batch->b;
        const uint64_t v[VGPU_WORK_STATS] = {b.stat_matches,        b.stat_postings,          b.stat_union,           b.stat_sparse,          b.stat_fast_items, b.stat_general_items,
                                             b.stat_plane_evaluated, b.stat_plane_item_evals, b.stat_plane_unconverged, b.stat_plane_sweepless, b.n_tiles,         b.n_parts,
                                             b.ix->planes.n_planes,  b.ix->planes.words};
        for (uint32_t i = 0; i < n; ++i) out[i] = i < VGPU_WORK_STATS ? v[i] : 0;
    });
}

int32_t vgpu_batch_io_bytes(const vgpu_batch* batch, uint64_t* h2d, uint64_t* d2h) {
    if (!batch) return VGPU_ERR_INVALID_REQUEST;
    if (h2d) *h2d = batch->b.h2d_bytes;
    if (d2h) *d2h = batch->b.d2h_bytes;
    return VGPU_OK;
}

}  // extern "C"
