// Device (HBM) layout of a veloci index directory, built once at open.
//
// Replaces the mmap'd views behind the reference's index-access traits
// (src/persistence.rs:80-87,142-181) with flat arrays the kernels can stream:
//
//   term dictionary  (`*.textindex.fst`, search_field.rs:54-65)
//       terms in byte order; per term its Unicode scalars as u16 alphabet codes
//       (lower-cased and, if any term has upper case, raw), u32 symbol offsets,
//       term ids, and one 20-byte common-prefix record per tile of 32 terms.
//   token -> anchor postings (`*.to_anchor_id_score`, token_to_anchor_score_vint.rs:128-204)
//       CSR: u64 offsets per term id, then 8-byte postings (u32 anchor ascending, f32
//       weight = f16 score / 100, the f16 round trip of AnchorScore applied at load,
//       persistence_score/mod.rs:7-17): one 64-bit load per posting, no divide in the kernel.
//       With anchor-range shards only the postings of [anchor_lo, anchor_hi) are kept.
//   id -> ids stores (`Indirect`, `SingleArrayPacked`; indirect.rs:10-89, single_array.rs:93-147)
//       CSR u32 offsets + u32 values (vint decoded at load).
//   boost columns (`*.boost_valid_to_value`, boost.rs:490-494)
//       dense u32 per id: f32 bits of get_value(), 0xFFFFFFFF = no value.
//   phrase pairs (`*.phrase_pair_to_anchor`, persistence_data_binary_search.rs:126-203)
//       sorted u64 keys (t1<<32|t2) + CSR of anchors.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../cuda/device_types.cuh"
#include "../cuda/kernels.cuh"
#include "comm.hpp"
#include "persistence.hpp"

namespace vdev {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define VDEV_CUDA(expr)                                                                                          \
    do {                                                                                                         \
        cudaError_t e__ = (expr);                                                                                \
        if (e__ != cudaSuccess)                                                                                  \
            throw vdev::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(e__) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

// Device allocations are recycled through a process-wide pool (per device, size classes of a
// quarter octave): a batch's scratch buffers come back from the previous batch instead of
// cudaMalloc / cudaFree, which cost milliseconds and synchronise the device.
class ScratchPool {
   public:
    static ScratchPool& instance() {
        static ScratchPool pool;
        return pool;
    }
    static size_t size_class(size_t bytes) {
        if (bytes <= 512) return 512;
        int top = 63 - __builtin_clzll((unsigned long long)bytes);
        const size_t step = (size_t)1 << (top - 2);
        return (bytes + step - 1) / step * step;
    }
    void* take(size_t cls) {
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> g(mu_);
            auto it = free_.find(std::make_pair(dev, cls));
            if (it != free_.end() && !it->second.empty()) {
                void* p = it->second.back();
                it->second.pop_back();
                cached_ -= cls;
                return p;
            }
        }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, cls);
        if (e != cudaSuccess) {
            trim();
            e = cudaMalloc(&p, cls);
        }
        if (e != cudaSuccess) throw std::runtime_error(std::string("cudaMalloc of ") + std::to_string(cls) + " bytes failed: " + cudaGetErrorString(e));
        return p;
    }
    void give(void* p, size_t cls) {
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> g(mu_);
            if (cached_ + cls <= kMaxCached) {
                free_[std::make_pair(dev, cls)].push_back(p);
                cached_ += cls;
                return;
            }
        }
        cudaFree(p);
    }
    void trim() {  // releases everything cached (all devices)
        std::lock_guard<std::mutex> g(mu_);
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto& kv : free_) {
            cudaSetDevice(kv.first.first);
            for (void* p : kv.second) cudaFree(p);
        }
        cudaSetDevice(cur);
        free_.clear();
        cached_ = 0;
    }

   private:
    static constexpr size_t kMaxCached = (size_t)24 << 30;
    std::mutex mu_;
    std::map<std::pair<int, size_t>, std::vector<void*>> free_;
    size_t cached_ = 0;
};

// While an UploadScope is alive on a thread, DevBuf::upload stages its source through pinned host memory and copies
// asynchronously on the scope's stream (a batch uploads a dozen small plan tables: one pageable cudaMemcpy each costs more
// than the whole transfer).  The scope's destructor waits for the copies, so the staging buffer can be reused.
class UploadScope {
   public:
    explicit UploadScope(cudaStream_t st) : stream_(st) {
        buf_ = take_buffer();
        used_ = 0;
        current() = buf_ ? this : nullptr;
    }
    ~UploadScope() {
        if (current() == this) {
            cudaStreamSynchronize(stream_);
            current() = nullptr;
        }
        if (buf_) give_buffer(buf_);
    }
    UploadScope(const UploadScope&) = delete;
    UploadScope& operator=(const UploadScope&) = delete;
    static UploadScope*& current() {
        static thread_local UploadScope* cur = nullptr;
        return cur;
    }
    // copies `bytes` from pageable `src` to device `dst`; false = too large for the staging buffer
    bool copy(void* dst, const void* src, size_t bytes) {
        if (bytes > kCap) return false;
        if (used_ + bytes > kCap) {
            cudaStreamSynchronize(stream_);
            used_ = 0;
        }
        char* at = buf_ + used_;
        memcpy(at, src, bytes);
        used_ += (bytes + 255) & ~(size_t)255;
        return cudaMemcpyAsync(dst, at, bytes, cudaMemcpyHostToDevice, stream_) == cudaSuccess;
    }

   private:
    static constexpr size_t kCap = (size_t)32 << 20;
    // The pinned staging buffers are shared by all threads (cudaHostAlloc costs tens of milliseconds and stalls the
    // device: a planner thread that comes and goes must not pay it again); a few are ever alive at once.
    struct BufferPool {
        std::mutex mu;
        std::vector<char*> free;
    };
    static BufferPool& pool() {
        static BufferPool* p = new BufferPool();  // never destroyed: outlives the CUDA context teardown order
        return *p;
    }
    static char* take_buffer() {
        {
            std::lock_guard<std::mutex> g(pool().mu);
            if (!pool().free.empty()) {
                char* p = pool().free.back();
                pool().free.pop_back();
                return p;
            }
        }
        char* p = nullptr;
        if (cudaHostAlloc((void**)&p, kCap, cudaHostAllocDefault) != cudaSuccess) return nullptr;
        return p;
    }
    static void give_buffer(char* p) {
        std::lock_guard<std::mutex> g(pool().mu);
        pool().free.push_back(p);
    }
    cudaStream_t stream_;
    char* buf_ = nullptr;
    size_t used_ = 0;
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    size_t cls = 0;  // pool size class of the allocation
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), cls(o.cls) { o.p = nullptr, o.n = 0, o.cls = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p, n = o.n, cls = o.cls;
            o.p = nullptr, o.n = 0, o.cls = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) ScratchPool::instance().give(p, cls);
        p = nullptr, n = 0, cls = 0;
    }
    void alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cls = ScratchPool::size_class(count * sizeof(T));
        p = static_cast<T*>(ScratchPool::instance().take(cls));
        n = count;
    }
    // keeps the allocation when it is already large enough
    void reserve(size_t count) {
        if (count > n) alloc(count + count / 4);
    }
    void upload(const std::vector<T>& v) {
        alloc(v.size());
        if (v.empty()) return;
        UploadScope* scope = UploadScope::current();
        if (scope && scope->copy(p, v.data(), v.size() * sizeof(T))) return;
        // No scope (index build, step seam): a pageable cudaMemcpy may return before its DMA has landed, and the batches'
        // non-blocking streams do not order against the legacy stream, so wait for the copy itself.
        VDEV_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        VDEV_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    }
    size_t bytes() const { return n * sizeof(T); }
};

// f32 -> f16 bit pattern, round to nearest even (half::f16::from_f32)
inline uint16_t f32_to_f16_bits(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xFF);
    uint32_t man = x & 0x7FFFFFu;
    if (exp == 255) return (uint16_t)(sign | 0x7C00u | (man ? (0x200u | (man >> 13)) : 0u));
    int32_t e = exp - 127 + 15;
    if (e >= 31) return (uint16_t)(sign | 0x7C00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        man |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - e);
        uint32_t hm = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (hm & 1u))) hm++;
        return (uint16_t)(sign | hm);
    }
    uint32_t v = ((uint32_t)e << 10) | (man >> 13);
    uint32_t rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (v & 1u))) v++;
    return (uint16_t)(sign | v);
}

// f16 bit pattern -> f32 (half::f16::to_f32, exact)
inline float f16_bits_to_f32(uint16_t h) {
    uint32_t hs = (uint32_t)(h & 0x8000u) << 16, he = (h >> 10) & 0x1Fu, hm = h & 0x3FFu, out;
    if (he == 0) {
        if (hm == 0) out = hs;
        else {
            int e = -1;
            do {
                hm <<= 1;
                e++;
            } while (!(hm & 0x400u));
            out = hs | ((uint32_t)(127 - 15 - e) << 23) | ((hm & 0x3FFu) << 13);
        }
    } else if (he == 31) out = hs | 0x7F800000u | (hm << 13);
    else out = hs | ((he + 112) << 23) | (hm << 13);
    float r;
    memcpy(&r, &out, 4);
    return r;
}

struct SymbolSet {  // one case variant of a dictionary
    DevBuf<uint16_t> sym;
    DevBuf<uint32_t> off;  // n + 1
    DevBuf<TilePrefix> tiles;
};

struct DelIndexDev {  // deletion-neighbourhood index (see DelIndexView)
    DevBuf<uint32_t> off;
    DevBuf<DelEntry> ent;
    uint32_t mask = 0, max_del = 0;
    bool built = false;
    DelIndexView view() const { return built ? DelIndexView{off.p, ent.p, mask, max_del} : DelIndexView{nullptr, nullptr, 0u, 0u}; }
};

struct DictDev {
    size_t n = 0;
    DelIndexDev del[2];       // [0] one deletion (built at open), [1] two deletions (built on first use)
    uint64_t variants[2] = {0, 0};  // entries each of them holds
    std::vector<uint32_t> alphabet;  // sorted scalars; code = index
    DevBuf<uint32_t> ids;            // term id per slot
    DevBuf<uint16_t> lower_bytes;    // byte length of the lower-cased term (clamped)
    SymbolSet lower;
    SymbolSet raw;  // empty when identical to `lower`
    bool has_raw = false;
    DevBuf<uint32_t> exc_slot, exc_off;  // terms whose to_lowercase is not the scalar-by-scalar lowering (DictView::exc_*)
    DevBuf<uint16_t> exc_sym;
    uint32_t n_exc = 0;
    uint32_t n_tiles = 0;

    uint16_t code_of(uint32_t scalar) const {
        auto it = std::lower_bound(alphabet.begin(), alphabet.end(), scalar);
        if (it == alphabet.end() || *it != scalar) return 0xFFFFu;
        return (uint16_t)(it - alphabet.begin());
    }
    DictView view() const {
        DictView v;
        v.n = (uint32_t)n;
        v.n_tiles = n_tiles;
        v.del[0] = del[0].view(), v.del[1] = del[1].view();
        v.ids = ids.p;
        v.lower_bytes = lower_bytes.p;
        v.sym[0] = lower.sym.p, v.off[0] = lower.off.p, v.tiles[0] = lower.tiles.p;
        const SymbolSet& r = has_raw ? raw : lower;
        v.sym[1] = r.sym.p, v.off[1] = r.off.p, v.tiles[1] = r.tiles.p;
        v.exc_slot = exc_slot.p, v.exc_off = exc_off.p, v.exc_sym = exc_sym.p, v.n_exc = n_exc, v.exc_pad = 0;
        return v;
    }
};

struct PostingsDev {
    size_t n_terms = 0;
    uint64_t n_postings = 0;
    DevBuf<uint64_t> off;
    DevBuf<Posting> post;
    std::vector<uint64_t> h_off;    // host copy of the offsets (head-term selection)
    DevBuf<uint32_t> term_plane;    // per term id: head-term plane or kNoValue (empty: the store has no planes)
    PostingsView view() const { return PostingsView{post.p, off.p, (uint32_t)n_terms, term_plane.p}; }
};

struct PlaneSetDev {  // term planes of every postings store of the shard (head planes first, then mid planes)
    uint32_t n_planes = 0, n_head = 0, words = 0;
    DevBuf<uint32_t> bits;
    DevBuf<uint16_t> score;   // head planes only
    DevBuf<float> wmax;
    DevBuf<PlaneInfo> info;
    DevBuf<uint32_t> tcount;  // [n_planes][tiles]
    DevBuf<uint32_t> tprefix; // [n_planes][tiles + 1]
    DevBuf<uint32_t> const_rows;  // kConstRowWords ones, then as many zeros (rows the sweep ANDs unrestricted / excluded terms with)
    static const uint32_t kConstRowWords = 32u << (kPlaneTileLog2 - 5);
    PlaneSetView view() const { return PlaneSetView{bits.p, score.p, wmax.p, info.p, tcount.p, tprefix.p, n_planes, n_head, words, 0u}; }
};

struct CsrDev {  // id -> list<u32>
    size_t n_ids = 0;
    DevBuf<uint32_t> off;  // n_ids + 1
    DevBuf<uint32_t> val;
    std::vector<uint32_t> h_off, h_val;  // host copy (small stores are also walked on the host)
    uint32_t n_values = 0;               // largest value + 1
    CsrView view() const { return CsrView{off.p, val.p, (uint32_t)n_ids}; }
};

struct ColumnDev {  // boost column
    size_t n = 0;
    DevBuf<uint32_t> bits;
    bool non_negative = true;  // every stored value is a non-negative, non-NaN float
    float vmax = 0.0f;
    // nested "value >= threshold" bitmaps over the shard's anchors (plane path)
    DevBuf<uint32_t> level_bits;
    DevBuf<uint32_t> seed_anchor, seed_bits;  // the column's seed set and the planes restricted to it (ColumnLevels)
    DevBuf<ColumnLevels> level_hdr;
    ColumnLevels h_levels{};
};

struct PhraseDev {
    size_t n = 0;
    DevBuf<uint64_t> keys;
    DevBuf<uint32_t> off;
    DevBuf<uint32_t> anchors;
    PhraseView view() const { return PhraseView{keys.p, off.p, anchors.p, (uint32_t)n}; }
};

struct DeviceIndex {
    int device = 0;
    int n_sms = 148;
    uint32_t shard_rank = 0, n_shards = 1;
    uint32_t open_flags = 0;  // VGPU_OPEN_* (include/veloci_b200.h)
    static const uint32_t kOpenNoPlanes = 1u, kOpenNoDelIndex = 2u;
    uint64_t num_docs = 0, anchor_lo = 0, anchor_hi = 0;
    std::unique_ptr<vhost::Persistence> host;
    std::map<std::string, DictDev> dicts;          // "<field>.textindex"
    std::map<std::string, PostingsDev> postings;   // "<field>.textindex.to_anchor_id_score"
    std::map<std::string, CsrDev> stores;          // key/value stores
    std::map<std::string, ColumnDev> boosts;       // "<field>.boost_valid_to_value"
    std::map<std::string, PhraseDev> phrases;
    PlaneSetDev planes;
    std::unique_ptr<ShardComm> comm;  // set by vgpu_comm_init: this handle is shard `comm->rank` of `comm->n_ranks`
    // Base pointers of the structures plan tables refer to, in name order (the same for every handle of one directory),
    // and a fingerprint of their names and sizes: see plan_blob.hpp.
    std::vector<const void*> reloc_ptrs;
    std::unordered_map<const void*, uint32_t> reloc_index;
    uint64_t reloc_fingerprint = 0;
    uint64_t max_posting_list = 0;   // longest posting list of the shard
    uint64_t max_nonplane_list = 0;  // ... among the terms without a plane
    size_t device_bytes = 0;

    uint32_t shard_words() const {  // 32-anchor words of the shard, padded to whole plane tiles
        const uint64_t span = anchor_hi - anchor_lo;
        const uint64_t tiles = (span + (1ull << kPlaneTileLog2) - 1) >> kPlaneTileLog2;
        return (uint32_t)(tiles << (kPlaneTileLog2 - 5));
    }

    static void build_symbol_set(const vhost::TermDict& d, const std::vector<uint32_t>& alphabet, bool lower, SymbolSet& out, std::vector<uint16_t>* lower_bytes) {
        const size_t n = d.size();
        std::vector<uint16_t> sym;
        std::vector<uint32_t> off(n + 1, 0);
        sym.reserve(d.bytes.size());
        if (lower_bytes) lower_bytes->assign(n, 0);
        for (size_t i = 0; i < n; ++i) {
            const uint8_t* key = &d.bytes[d.offsets[i]];
            const size_t klen = d.offsets[i + 1] - d.offsets[i];
            size_t pos = 0, lb = 0;
            while (pos < klen) {
                uint32_t cp = vfmt::utf8_next(key, klen, pos);
                if (lower) cp = vfmt::lower_scalar(cp);
                lb += cp < 0x80 ? 1 : cp < 0x800 ? 2 : cp < 0x10000 ? 3 : 4;
                auto it = std::lower_bound(alphabet.begin(), alphabet.end(), cp);
                sym.push_back((uint16_t)(it - alphabet.begin()));
            }
            off[i + 1] = (uint32_t)sym.size();
            if (lower_bytes) (*lower_bytes)[i] = (uint16_t)std::min<size_t>(lb, 65535);
        }
        // common prefix of each tile of kDictTile consecutive terms (first vs last term of the tile)
        const size_t n_tiles = (n + kDictTile - 1) / kDictTile;
        std::vector<TilePrefix> tiles(n_tiles);
        for (size_t t = 0; t < n_tiles; ++t) {
            size_t a = t * kDictTile, b = std::min(n, a + kDictTile) - 1;
            uint32_t la = off[a + 1] - off[a], lb2 = off[b + 1] - off[b];
            uint32_t l = 0, lim = std::min<uint32_t>(std::min(la, lb2), kTilePrefixMax);
            while (l < lim && sym[off[a] + l] == sym[off[b] + l]) ++l;
            // byte order == scalar order in UTF-8, but lower-casing can reorder terms:
            // only the raw variant is sorted.  For the lowered variant verify every term.
            if (lower)
                for (size_t i = a + 1; i < b && l > 0; ++i) {
                    uint32_t li = off[i + 1] - off[i];
                    uint32_t k = 0;
                    while (k < l && k < li && sym[off[i] + k] == sym[off[a] + k]) ++k;
                    l = k;
                }
            tiles[t].len = (uint16_t)l;
            tiles[t].pad = 0;
            for (uint32_t k = 0; k < kTilePrefixMax; ++k) tiles[t].sym[k] = k < l ? sym[off[a] + k] : 0;
        }
        out.sym.upload(sym);
        out.off.upload(off);
        out.tiles.upload(tiles);
    }

    // Files every term of the dictionary under the hashes of its variants with at most level + 1 deletions.
    void build_del_index(DictDev& dd, int level) {
        DelIndexDev& di = dd.del[level];
        if (di.built || dd.n == 0) return;
        const uint64_t entries = dd.variants[level];
        if (entries >= (1ull << 31)) return;  // the scan kernel stays responsible
        uint32_t buckets = 1024;
        while (buckets < entries * (level == 0 ? 2 : 1) && buckets < (1u << 28)) buckets <<= 1;
        DevBuf<uint32_t> count;
        count.alloc((size_t)buckets + 1);
        di.off.alloc((size_t)buckets + 2);
        di.ent.alloc((size_t)entries + 1);
        di.mask = buckets - 1, di.max_del = (uint32_t)level + 1;
        DictView v = dd.view();
        VDEV_CUDA(cudaMemset(count.p, 0, count.bytes()));
        launch_del_index_pass(nullptr, v, di.max_del, di.mask, count.p, nullptr, nullptr);
        launch_scan_u32(nullptr, count.p, di.off.p, buckets);
        VDEV_CUDA(cudaMemset(count.p, 0, count.bytes()));
        launch_del_index_pass(nullptr, v, di.max_del, di.mask, count.p, di.off.p, di.ent.p);
        VDEV_CUDA(cudaDeviceSynchronize());
        di.built = true;
        device_bytes += di.off.bytes() + di.ent.bytes();
    }

    std::mutex build_mu;
    // two-deletion index of a dictionary, built when the first request needs it
    void ensure_del_index(const std::string& path, int level) {
        std::lock_guard<std::mutex> g(build_mu);
        auto it = dicts.find(path);
        if (it == dicts.end() || it->second.del[level].built || (open_flags & kOpenNoDelIndex)) return;
        VDEV_CUDA(cudaSetDevice(device));
        try {
            build_del_index(it->second, level);
        } catch (const std::exception&) {  // no room for the two-deletion index: the dictionary scan stays responsible
            it->second.del[level] = DelIndexDev();
            cudaGetLastError();
        }
    }
    // Views of a dictionary taken while no build is in flight (the planner thread of search_stream may be inside
    // ensure_del_index for the same dictionary).
    DictView dict_view(const std::string& path) {
        std::lock_guard<std::mutex> g(build_mu);
        return dicts.at(path).view();
    }
    bool del_index_built(const std::string& path, int level) {
        std::lock_guard<std::mutex> g(build_mu);
        return dicts.at(path).del[level].built;
    }

    void build_dict(const std::string& path, const vhost::TermDict& d) {
        DictDev dd;
        dd.n = d.size();
        dd.n_tiles = (uint32_t)((dd.n + kDictTile - 1) / kDictTile);
        bool any_upper = false, context_case = false;
        std::vector<uint32_t> scalars;
        {
            std::vector<uint8_t> seen_small(0x10000 / 8, 0);
            std::vector<uint32_t> big;
            auto mark = [&](uint32_t cp) {
                if (cp < 0x10000) seen_small[cp >> 3] |= (uint8_t)(1u << (cp & 7));
                else big.push_back(cp);
            };
            size_t pos = 0;
            const size_t total = d.bytes.size();
            // term boundaries never split a scalar, so the byte stream can be decoded in one go
            while (pos < total) {
                uint32_t cp = vfmt::utf8_next(d.bytes.data(), total, pos);
                uint32_t lc = vfmt::lower_scalar(cp);
                if (lc != cp) any_upper = true;
                mark(cp);
                mark(lc);
                if (cp == 0x130) mark(0x69), mark(0x307), context_case = true;  // to_lowercase expands it
                if (cp == 0x3A3) mark(0x3C2), context_case = true;             // ... and may give the final sigma
            }
            for (uint32_t cp = 0; cp < 0x10000; ++cp)
                if (seen_small[cp >> 3] & (1u << (cp & 7))) scalars.push_back(cp);
            std::sort(big.begin(), big.end());
            big.erase(std::unique(big.begin(), big.end()), big.end());
            scalars.insert(scalars.end(), big.begin(), big.end());
        }
        if (scalars.size() >= 0xFFFF) throw std::runtime_error("dictionary " + path + " uses more than 65534 distinct scalars");
        dd.alphabet = std::move(scalars);
        dd.ids.upload(d.ids);
        std::vector<uint16_t> lower_bytes;
        build_symbol_set(d, dd.alphabet, true, dd.lower, &lower_bytes);
        if (context_case) {  // Rust's to_lowercase of these terms, kept beside the scalar-wise lowering the matching runs on
            std::vector<uint32_t> slots, offs{0}, scalars, exact;
            std::vector<uint16_t> syms;
            for (size_t i = 0; i < d.size(); ++i) {
                const uint8_t* key = &d.bytes[d.offsets[i]];
                const size_t klen = d.offsets[i + 1] - d.offsets[i];
                bool candidate = false;
                for (size_t b = 0; b + 1 < klen; ++b) candidate = candidate || (key[b] == 0xC4 && key[b + 1] == 0xB0) || (key[b] == 0xCE && key[b + 1] == 0xA3);
                if (!candidate) continue;
                scalars.clear();
                for (size_t pos = 0; pos < klen;) scalars.push_back(vfmt::utf8_next(key, klen, pos));
                vfmt::lowercase_scalars(scalars, exact);
                bool same = exact.size() == scalars.size();
                for (size_t j = 0; same && j < exact.size(); ++j) same = exact[j] == vfmt::lower_scalar(scalars[j]);
                if (same) continue;
                size_t bytes = 0;
                for (uint32_t cp : exact) {
                    syms.push_back(dd.code_of(cp));
                    bytes += cp < 0x80 ? 1 : cp < 0x800 ? 2 : cp < 0x10000 ? 3 : 4;
                }
                lower_bytes[i] = (uint16_t)std::min<size_t>(bytes, 65535);
                slots.push_back((uint32_t)i), offs.push_back((uint32_t)syms.size());
            }
            dd.n_exc = (uint32_t)slots.size();
            if (dd.n_exc) dd.exc_slot.upload(slots), dd.exc_off.upload(offs), dd.exc_sym.upload(syms);
        }
        dd.lower_bytes.upload(lower_bytes);
        dd.has_raw = any_upper;
        if (any_upper) build_symbol_set(d, dd.alphabet, false, dd.raw, nullptr);
        {
            std::vector<uint32_t> off(dd.n + 1);
            if (dd.n) VDEV_CUDA(cudaMemcpy(off.data(), dd.lower.off.p, (dd.n + 1) * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < dd.n; ++i) {
                const uint64_t len = off[i + 1] - off[i];
                dd.variants[0] += 1 + len;
                dd.variants[1] += 1 + len + (len >= 2 ? len * (len - 1) / 2 : 0);
            }
        }
        if (!(open_flags & kOpenNoDelIndex)) build_del_index(dd, 0);
        device_bytes += dd.ids.bytes() + dd.lower_bytes.bytes() + dd.lower.sym.bytes() + dd.lower.off.bytes() + dd.lower.tiles.bytes() + dd.raw.sym.bytes() + dd.raw.off.bytes() + dd.raw.tiles.bytes();
        dicts.emplace(path, std::move(dd));
    }

    void build_postings(const std::string& path, const vfmt::AnchorScoreView& v) {
        const size_t n = v.num_ids();
        std::vector<uint64_t> off(n + 1, 0);
        const uint32_t lo = (uint32_t)anchor_lo, hi = (uint32_t)std::min<uint64_t>(anchor_hi, 0xFFFFFFFFull);
        // every structure of the shard is sized from metaData.json's num_docs: a posting outside [0, num_docs) (stale or
        // corrupt metadata) must fail the load, not index past the planes and tile buckets
        std::atomic<uint64_t> out_of_range{0};
        const uint64_t docs = num_docs;
        unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        auto parallel_terms = [&](auto&& fn) {
            std::vector<std::thread> pool;
            const size_t chunk = (n + hw - 1) / hw;
            for (unsigned t = 0; t < hw; ++t) {
                size_t a = (size_t)t * chunk, b = std::min(n, a + chunk);
                if (a >= b) break;
                pool.emplace_back([=, &fn]() { fn(a, b); });
            }
            for (auto& th : pool) th.join();
        };
        parallel_terms([&](size_t a, size_t b) {
            for (size_t id = a; id < b; ++id) {
                uint64_t c = 0;
                uint64_t bad = 0;
                v.for_each((uint32_t)id, [&](uint32_t anchor, uint32_t) {
                    c += (anchor >= lo && anchor < hi) ? 1 : 0;
                    bad += anchor >= docs ? 1 : 0;
                });
                off[id + 1] = c;
                if (bad) out_of_range.fetch_add(bad, std::memory_order_relaxed);
            }
        });
        if (out_of_range.load()) throw vhost::IoError(path + ": " + std::to_string(out_of_range.load()) + " postings have an anchor id >= num_docs (" + std::to_string(docs) + ") of metaData.json");
        for (size_t i = 0; i < n; ++i) off[i + 1] += off[i];
        const uint64_t total = off[n];
        std::vector<Posting> post(total);
        parallel_terms([&](size_t a, size_t b) {
            for (size_t id = a; id < b; ++id) {
                uint64_t at = off[id];
                v.for_each((uint32_t)id, [&](uint32_t anchor, uint32_t raw) {
                    if (anchor >= lo && anchor < hi) {
                        post[at].anchor = anchor;
                        // AnchorScore keeps the score as f16 (persistence_score/mod.rs:9-12); resolve_token_to_anchor
                        // uses `score.to_f32() / 100.0` (search_field.rs:426): divided here, once, in IEEE f32
                        post[at].weight = f16_bits_to_f32(f32_to_f16_bits((float)raw)) / 100.0f;
                        ++at;
                    }
                });
            }
        });
        PostingsDev pd;
        pd.n_terms = n;
        pd.n_postings = total;
        pd.off.upload(off);
        pd.post.upload(post);
        pd.h_off = std::move(off);
        device_bytes += pd.off.bytes() + pd.post.bytes();
        postings.emplace(path, std::move(pd));
    }

    void build_store(const std::string& path, const vhost::KeyValueStore& s) {
        CsrDev c;
        std::vector<uint32_t> tmp;
        if (s.kind == vhost::KeyValueStore::Indirect) {
            c.n_ids = s.ind.n_ids;
            c.h_off.assign(c.n_ids + 1, 0);
            for (size_t id = 0; id < c.n_ids; ++id) {
                s.ind.append_values(id, c.h_val);
                c.h_off[id + 1] = (uint32_t)c.h_val.size();
            }
        } else if (s.kind == vhost::KeyValueStore::Packed) {
            c.n_ids = s.packed.width ? s.packed.len / (size_t)s.packed.width : 0;
            if (s.packed.width && s.packed.len % (size_t)s.packed.width) c.n_ids += 1;
            c.h_off.assign(c.n_ids + 1, 0);
            for (size_t id = 0; id < c.n_ids; ++id) {
                uint32_t v;
                if (s.packed.get_value(id, v)) c.h_val.push_back(v);
                c.h_off[id + 1] = (uint32_t)c.h_val.size();
            }
        } else {
            c.n_ids = 0;
            c.h_off.assign(1, 0);
        }
        for (uint32_t v : c.h_val) c.n_values = std::max(c.n_values, v + 1);
        c.off.upload(c.h_off);
        c.val.upload(c.h_val);
        device_bytes += c.off.bytes() + c.val.bytes();
        stores.emplace(path, std::move(c));
    }

    void build_boost(const std::string& path, const vhost::KeyValueStore& s) {
        ColumnDev c;
        size_t n = 0;
        if (s.kind == vhost::KeyValueStore::Indirect) n = s.ind.n_ids;
        else if (s.kind == vhost::KeyValueStore::Packed) n = s.packed.width ? (s.packed.len + (size_t)s.packed.width - 1) / (size_t)s.packed.width : 0;
        std::vector<uint32_t> bits(n, kNoValue);
        for (size_t id = 0; id < n; ++id) {
            uint32_t v;
            if (s.get_value(id, v)) {
                bits[id] = v;
                float f;
                memcpy(&f, &v, 4);
                if (!(f >= 0.0f)) c.non_negative = false;
                else if (f > c.vmax) c.vmax = f;
            }
        }
        c.n = n;
        c.bits.upload(bits);
        device_bytes += c.bits.bytes();
        // level thresholds: the (1 - 2^-(j+1)) quantiles of a sample of the shard's values
        const uint64_t span = anchor_hi - anchor_lo;
        if (c.non_negative && span > 0 && span < 0xFFFFFFFFull && !(open_flags & kOpenNoPlanes)) {
            std::vector<float> sample;
            const size_t lo = (size_t)std::min<uint64_t>(anchor_lo, n), hi = (size_t)std::min<uint64_t>(anchor_hi, n);
            const size_t stride = std::max<size_t>(1, (hi - lo) >> 20);
            for (size_t id = lo; id < hi; id += stride)
                if (bits[id] != kNoValue) {
                    float f;
                    memcpy(&f, &bits[id], 4);
                    sample.push_back(f);
                }
            std::sort(sample.begin(), sample.end());
            const uint32_t words = shard_words();
            c.h_levels.words = words, c.h_levels.pad = 0;
            for (uint32_t j = 0; j < kBoostLevels; ++j) {
                if (sample.empty()) {
                    c.h_levels.thr[j] = 3.0e38f;
                    continue;
                }
                const double frac = 1.0 - std::ldexp(1.0, -(int)(j + 1));
                const size_t at = std::min(sample.size() - 1, (size_t)(frac * (double)sample.size()));
                c.h_levels.thr[j] = sample[at];
            }
            c.level_bits.alloc((size_t)kBoostLevels * words);
            launch_level_fill(nullptr, c.bits.p, (uint32_t)n, (uint32_t)anchor_lo, (uint32_t)span, c.h_levels.thr, c.level_bits.p, words);
            c.h_levels.bits = c.level_bits.p;
            // seed set: the shard's anchors with a value in the top 1/8 of the column, and every plane restricted to them
            c.h_levels.seed_anchor = nullptr, c.h_levels.seed_bits = nullptr, c.h_levels.seed_n = 0, c.h_levels.seed_words = 0;
            if (planes.n_planes > 0 && !sample.empty()) {
                std::vector<uint32_t> seed;
                const float thr = c.h_levels.thr[kSeedLevel];
                {
                    // in descending value order: the seed pass meets the best-boosted anchors first and may stop early
                    std::vector<std::pair<uint32_t, uint32_t>> by_value;  // (value bits: non-negative floats order like their bits, anchor)
                    for (size_t id = lo; id < hi; ++id)
                        if (bits[id] != kNoValue) {
                            float f;
                            memcpy(&f, &bits[id], 4);
                            if (f >= thr) by_value.emplace_back(bits[id], (uint32_t)(id - anchor_lo));
                        }
                    std::sort(by_value.begin(), by_value.end(), [](const std::pair<uint32_t, uint32_t>& x, const std::pair<uint32_t, uint32_t>& y) {
                        return x.first != y.first ? x.first > y.first : x.second < y.second;
                    });
                    seed.reserve(by_value.size());
                    for (auto& v : by_value) seed.push_back(v.second);
                }
                if (!seed.empty() && seed.size() <= (size_t)(span / 4 + 1024)) {  // (a column of few distinct values can put most anchors at its top: no use as a seed set)
                    const uint32_t seed_words = (uint32_t)(((seed.size() + 31) / 32 + 127) / 128 * 128);
                    c.seed_anchor.upload(seed);
                    c.seed_bits.alloc((size_t)planes.n_planes * seed_words);
                    launch_seed_compact(nullptr, planes.bits.p, planes.n_planes, planes.words, c.seed_anchor.p, (uint32_t)seed.size(), seed_words, c.seed_bits.p);
                    c.h_levels.seed_anchor = c.seed_anchor.p, c.h_levels.seed_bits = c.seed_bits.p, c.h_levels.seed_n = (uint32_t)seed.size(), c.h_levels.seed_words = seed_words;
                    device_bytes += c.seed_anchor.bytes() + c.seed_bits.bytes();
                }
            }
            c.level_hdr.upload(std::vector<ColumnLevels>{c.h_levels});
            device_bytes += c.level_bits.bytes();
        }
        boosts.emplace(path, std::move(c));
    }

    // Term planes: head planes for the terms with df >= span / 128 (at most kMaxHeadPlanes, bits + f16 scores), mid planes
    // for the terms with df >= span / 1024 (bits only), at most kMaxPlanes in total, by descending df over all stores.
    void build_planes() {
        const uint64_t span = anchor_hi - anchor_lo;
        if ((open_flags & kOpenNoPlanes) || span == 0 || span >= 0xFFFFFFFFull) return;
        uint64_t mid_div = 1024;
#ifdef VELOCI_PROBES
        if (const char* env = getenv("VELOCI_MID_DIV")) mid_div = (uint64_t)std::max(128, atoi(env));
#endif
        const uint64_t head_df = std::max<uint64_t>(1, span / 128), mid_df = std::max<uint64_t>(16, span / mid_div);
        struct Cand {
            uint64_t df;
            PostingsDev* store;
            uint32_t term;
        };
        std::vector<Cand> cands;
        for (auto& kv : postings) {
            PostingsDev& pd = kv.second;
            for (size_t t = 0; t < pd.n_terms; ++t) {
                const uint64_t df = pd.h_off[t + 1] - pd.h_off[t];
                if (df >= mid_df) cands.push_back(Cand{df, &pd, (uint32_t)t});
            }
        }
        if (cands.empty()) return;
        std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.df > b.df; });
        if (cands.size() > kMaxPlanes) cands.resize(kMaxPlanes);
        uint32_t n_head = 0;
        while (n_head < cands.size() && n_head < kMaxHeadPlanes && cands[n_head].df >= head_df) ++n_head;
        const uint32_t n = (uint32_t)cands.size(), words = shard_words();
        planes.n_planes = n, planes.n_head = n_head, planes.words = words;
        planes.bits.alloc((size_t)n * words);
        planes.score.alloc((size_t)std::max<uint32_t>(n_head, 1) * words * 32);
        planes.wmax.alloc(n);
        DevBuf<uint32_t> bad;
        bad.alloc(n);
        VDEV_CUDA(cudaMemset(planes.bits.p, 0, planes.bits.bytes()));
        VDEV_CUDA(cudaMemset(planes.score.p, 0, planes.score.bytes()));
        VDEV_CUDA(cudaMemset(planes.wmax.p, 0, planes.wmax.bytes()));
        VDEV_CUDA(cudaMemset(bad.p, 0, bad.bytes()));
        std::vector<PlaneInfo> info(n);
        for (uint32_t p = 0; p < n; ++p) {
            const Cand& c = cands[p];
            info[p] = PlaneInfo{c.store->post.p + c.store->h_off[c.term], (uint32_t)c.df, 0u};
            launch_plane_fill(nullptr, info[p].post, c.df, planes.bits.p + (size_t)p * words, p < n_head ? planes.score.p + (size_t)p * words * 32 : nullptr, planes.wmax.p + p, bad.p + p,
                              (uint32_t)anchor_lo, (uint32_t)span);
        }
        planes.info.upload(info);
        planes.tcount.alloc((size_t)n * (words >> (kPlaneTileLog2 - 5)));
        planes.tprefix.alloc((size_t)n * ((words >> (kPlaneTileLog2 - 5)) + 1));
        launch_plane_tile_counts(nullptr, planes.bits.p, n, words, planes.tcount.p, planes.tprefix.p);
        planes.const_rows.alloc((size_t)2 * PlaneSetDev::kConstRowWords);
        VDEV_CUDA(cudaMemset(planes.const_rows.p, 0xFF, (size_t)PlaneSetDev::kConstRowWords * 4));
        VDEV_CUDA(cudaMemset(planes.const_rows.p + PlaneSetDev::kConstRowWords, 0, (size_t)PlaneSetDev::kConstRowWords * 4));
        std::vector<uint32_t> h_bad(n);
        VDEV_CUDA(cudaMemcpy(h_bad.data(), bad.p, n * 4, cudaMemcpyDeviceToHost));
        std::map<PostingsDev*, std::vector<uint32_t>> maps;
        for (uint32_t p = 0; p < n; ++p) {
            if (h_bad[p]) continue;  // the list cannot be represented as a plane: the term stays on the posting path
            auto& m = maps[cands[p].store];
            if (m.empty()) m.assign(cands[p].store->n_terms, kNoValue);
            m[cands[p].term] = p;
        }
        for (auto& kv : maps) {
            kv.first->term_plane.upload(kv.second);
            device_bytes += kv.first->term_plane.bytes();
        }
        max_nonplane_list = 0;
        for (auto& kv : postings) {
            PostingsDev& pd = kv.second;
            auto it = maps.find(&pd);
            for (size_t t = 0; t < pd.n_terms; ++t)
                if (it == maps.end() || it->second[t] == kNoValue) max_nonplane_list = std::max<uint64_t>(max_nonplane_list, pd.h_off[t + 1] - pd.h_off[t]);
        }
        device_bytes += planes.bits.bytes() + planes.score.bytes() + planes.wmax.bytes() + planes.info.bytes() + planes.tcount.bytes() + planes.tprefix.bytes();
    }

    void build_phrase(const std::string& path, const vfmt::PhrasePairView& v) {
        PhraseDev p;
        p.n = v.n;
        std::vector<uint64_t> keys(v.n);
        std::vector<uint32_t> off(v.n + 1, 0), anchors;
        for (size_t i = 0; i < v.n; ++i) {
            uint32_t t1 = vfmt::load_u32(v.recs + i * 12), t2 = vfmt::load_u32(v.recs + i * 12 + 4), o = vfmt::load_u32(v.recs + i * 12 + 8);
            keys[i] = ((uint64_t)t1 << 32) | t2;
            if (o < v.data_len) {
                vfmt::VintArrayIter it(v.data + o, v.data + v.data_len);
                uint32_t a;
                while (it.next(a)) anchors.push_back(a);
            }
            off[i + 1] = (uint32_t)anchors.size();
        }
        p.keys.upload(keys);
        p.off.upload(off);
        p.anchors.upload(anchors);
        device_bytes += p.keys.bytes() + p.off.bytes() + p.anchors.bytes();
        phrases.emplace(path, std::move(p));
    }

    static std::unique_ptr<DeviceIndex> open(const std::string& dir, int device, uint32_t rank, uint32_t n_shards, uint32_t flags = 0) {
        if (n_shards == 0 || rank >= n_shards) throw std::runtime_error("invalid shard rank");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) throw CudaError(std::string("no usable CUDA device: ") + cudaGetErrorString(e));
        if (device < 0 || device >= count) throw CudaError("device ordinal out of range");
        VDEV_CUDA(cudaSetDevice(device));
        std::unique_ptr<DeviceIndex> ix(new DeviceIndex());
        ix->device = device;
        VDEV_CUDA(cudaDeviceGetAttribute(&ix->n_sms, cudaDevAttrMultiProcessorCount, device));
        ix->shard_rank = rank;
        ix->n_shards = n_shards;
        ix->open_flags = flags;
        ix->host = vhost::Persistence::load(dir);
        ix->num_docs = ix->host->metadata.num_docs;
        ix->anchor_lo = ix->num_docs * rank / n_shards;
        ix->anchor_hi = ix->num_docs * (rank + 1) / n_shards;
        for (auto& kv : ix->host->dict) ix->build_dict(kv.first, kv.second);
        for (auto& kv : ix->host->token_to_anchor_score) ix->build_postings(kv.first, kv.second);
        for (auto& kv : ix->postings)
            for (size_t t = 0; t < kv.second.n_terms; ++t) ix->max_posting_list = std::max<uint64_t>(ix->max_posting_list, kv.second.h_off[t + 1] - kv.second.h_off[t]);
        ix->max_nonplane_list = ix->max_posting_list;
        ix->build_planes();
        for (auto& kv : ix->host->key_value_stores) ix->build_store(kv.first, kv.second);
        for (auto& kv : ix->host->boost_valueid_to_value) {
            // token values are keyed by term id and applied to a part's few term hits on the host (engine.hpp: apply_token_value)
            if (vfmt::ends_with(kv.first, ".token_values.boost_valid_to_value")) continue;
            ix->build_boost(kv.first, kv.second);
        }
        for (auto& kv : ix->host->phrase_pair_to_anchor) ix->build_phrase(kv.first, kv.second);
        VDEV_CUDA(cudaDeviceSynchronize());
        ix->build_reloc_table();
        return ix;
    }

    void build_reloc_table() {
        uint64_t h = 0xCBF29CE484222325ull;
        auto mix = [&](const void* p, size_t n) {
            const uint8_t* b = static_cast<const uint8_t*>(p);
            for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 0x100000001B3ull;
        };
        auto add = [&](const std::string& name, const void* p, uint64_t size) {
            mix(name.data(), name.size()), mix(&size, 8);
            reloc_index.emplace(p, (uint32_t)reloc_ptrs.size());
            reloc_ptrs.push_back(p);
        };
        mix(&num_docs, 8);
        for (auto& kv : boosts) add(kv.first, kv.second.bits.p, kv.second.n), add(kv.first + "#levels", kv.second.level_hdr.p, kv.second.level_hdr.p ? 1 : 0);
        for (auto& kv : stores) add(kv.first, kv.second.off.p, kv.second.n_ids), add(kv.first + "#val", kv.second.val.p, kv.second.h_val.size());
        for (auto& kv : phrases) add(kv.first, kv.second.keys.p, kv.second.n), add(kv.first + "#off", kv.second.off.p, kv.second.n), add(kv.first + "#anchors", kv.second.anchors.p, 0);
        for (auto& kv : dicts) mix(kv.first.data(), kv.first.size()), mix(&kv.second.n, sizeof kv.second.n);
        reloc_fingerprint = h;
    }
};

}  // namespace vdev
