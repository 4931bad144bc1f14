// The compressed document store (file `data` of an index directory; doc_store/src/lib.rs).  Layout: blocks of documents,
// each block `lz4 size-prepended( VIntArray::serialize([first doc id, 0, end of doc 0, end of doc 1, ...]) || documents )`,
// then the block index -- (first doc id of the block, byte offset of the block + 1) as u32 LE pairs, closed by a sentinel
// pair (number of documents + 1, end of the last block + 1) -- and the index's byte length as the file's last u32.
// DocLoader::get_doc (lib.rs:26-62): binary search of the block, decompress, slice.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../format/lz4_block.hpp"
#include "../format/vint.hpp"

namespace vhost {

struct DocStoreError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class DocLoader {
   public:
    DocLoader() = default;
    DocLoader(const uint8_t* data, size_t len) : data_(data), len_(len) {  // DocLoader::open (lib.rs:15-23)
        if (len < 4) throw DocStoreError("doc store: file too short");
        const uint32_t index_size = u32_at(len - 4);
        if ((uint64_t)index_size + 4 > len || index_size % 8 != 0 || index_size < 16) throw DocStoreError("doc store: bad block index size");
        index_ = data + len - 4 - index_size;
        n_entries_ = index_size / 8;
    }
    bool is_open() const { return data_ != nullptr; }
    // every document id below this exists (the sentinel holds the count + 1)
    uint32_t num_docs() const { return n_entries_ ? entry(n_entries_ - 1).first - 1 : 0; }

    std::string get_doc(uint32_t doc_id) const {
        if (!is_open()) throw DocStoreError("doc store: not open");
        if (doc_id >= num_docs()) throw DocStoreError("doc store: document " + std::to_string(doc_id) + " does not exist (" + std::to_string(num_docs()) + " documents)");
        // the last block whose first id is <= doc_id (lib.rs:210-240)
        size_t lo = 0, hi = n_entries_ - 1;  // entries [0, n-1) are blocks, entry n-1 is the sentinel
        while (hi - lo > 1) {
            const size_t mid = (lo + hi) / 2;
            if (entry(mid).first <= doc_id) lo = mid;
            else hi = mid;
        }
        while (lo > 0 && entry(lo - 1).first == entry(lo).first) --lo;  // (the reference's writer can close with an empty block that repeats the previous first id)
        const uint64_t begin = (uint64_t)entry(lo).second - 1, end = (uint64_t)entry(lo + 1).second - 1;
        if (begin >= end || end > len_) throw DocStoreError("doc store: block offsets out of range");
        std::vector<uint8_t> block;
        try {
            block = vfmt::lz4_decompress_size_prepended(data_ + begin, (size_t)(end - begin));
        } catch (const vfmt::Lz4Error& e) {
            throw DocStoreError(std::string("doc store: ") + e.what());
        }
        // header: vint(byte length of the array) || vints: first id, then the documents' boundaries
        const uint8_t* p = block.data();
        const uint8_t* const block_end = p + block.size();
        uint32_t arr_bytes = 0;
        size_t used = vfmt::vint_decode(p, block_end, arr_bytes);
        if (!used || arr_bytes > (size_t)(block_end - p) - used) throw DocStoreError("doc store: bad block header");
        p += used;
        const uint8_t* const arr_end = p + arr_bytes;
        uint32_t first_id = 0;
        used = vfmt::vint_decode(p, arr_end, first_id);
        if (!used || doc_id < first_id) throw DocStoreError("doc store: bad block header");
        p += used;
        uint32_t from = 0, to = 0;
        for (uint32_t i = first_id;; ++i) {  // boundaries: offsets[k] .. offsets[k + 1] hold document first_id + k
            used = vfmt::vint_decode(p, arr_end, to);
            if (!used) throw DocStoreError("doc store: document not in its block");
            p += used;
            if (i == doc_id + 1) break;
            from = to;
        }
        const size_t docs_len = (size_t)(block_end - arr_end);
        if (from > to || to > docs_len) throw DocStoreError("doc store: document boundaries out of range");
        return std::string((const char*)arr_end + from, to - from);
    }

   private:
    uint32_t u32_at(size_t at) const { return (uint32_t)data_[at] | ((uint32_t)data_[at + 1] << 8) | ((uint32_t)data_[at + 2] << 16) | ((uint32_t)data_[at + 3] << 24); }
    std::pair<uint32_t, uint32_t> entry(size_t i) const {
        const size_t at = (size_t)(index_ - data_) + i * 8;
        return {u32_at(at), u32_at(at + 4)};
    }
    const uint8_t* data_ = nullptr;
    const uint8_t* index_ = nullptr;
    size_t len_ = 0, n_entries_ = 0;
};

// DocStoreWriter (lib.rs:83-170), for the index builder: blocks are flushed once they hold more than 16 KiB.
class DocStoreWriter {
   public:
    void add_doc(const std::string& doc, std::vector<uint8_t>& out) {
        if (block_.empty() && bounds_.empty()) first_id_ = next_id_, bounds_.push_back(0);
        block_.insert(block_.end(), doc.begin(), doc.end());
        bounds_.push_back((uint32_t)block_.size());
        if (block_.size() > 16384) flush(out);
        ++next_id_;
    }
    void finish(std::vector<uint8_t>& out) {
        if (!bounds_.empty() || index_.empty()) flush(out);  // (the reference also writes an empty trailing block when the last document closed one: skipped)
        index_.emplace_back(next_id_ + 1, (uint32_t)out_bytes_ + 1);
        for (auto& e : index_)
            for (uint32_t v : {e.first, e.second})
                for (int i = 0; i < 4; ++i) out.push_back((uint8_t)(v >> (8 * i)));
        const uint32_t index_size = (uint32_t)index_.size() * 8;
        for (int i = 0; i < 4; ++i) out.push_back((uint8_t)(index_size >> (8 * i)));
    }
    uint32_t num_docs() const { return next_id_; }

   private:
    void flush(std::vector<uint8_t>& out) {
        std::vector<uint32_t> header{first_id_};
        header.insert(header.end(), bounds_.begin(), bounds_.end());
        std::vector<uint8_t> raw;
        vfmt::vint_array_serialize(raw, header.data(), header.size());
        raw.insert(raw.end(), block_.begin(), block_.end());
        const size_t before = out.size();
        vfmt::lz4_compress_prepend_size(raw.data(), raw.size(), out);
        index_.emplace_back(first_id_, (uint32_t)out_bytes_ + 1);
        out_bytes_ += out.size() - before;
        block_.clear(), bounds_.clear();
    }
    uint32_t next_id_ = 0, first_id_ = 0;
    uint64_t out_bytes_ = 0;
    std::vector<uint8_t> block_;
    std::vector<uint32_t> bounds_;
    std::vector<std::pair<uint32_t, uint32_t>> index_;
};

}  // namespace vhost
