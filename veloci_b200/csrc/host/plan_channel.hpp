// Plans through POSIX shared memory between the processes of one box (vgpu_plan_channel_*): local rank 0 publishes
// the plan blob of each batch (plan_blob.hpp), the other local ranks import it.  Two slots, so the publisher may be one
// batch ahead of the slowest consumer, and two plans may be in the making at once (Index.search_stream keeps two
// planner threads busy while a batch is on the GPUs: batches are numbered by tickets taken in execution order).  Synchronisation is a sequence number per slot (release / acquire on lock-free 64-bit atomics in the
// mapping) and one "consumed up to" counter per (slot, rank); waiting is a bounded spin, then short sleeps.
#pragma once
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace vplan {

struct ChannelError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class PlanChannel {
   public:
    static const uint32_t kMaxRanks = 64;
    static const uint64_t kMagic = 0x314e4843504c4556ull;  // "VELPCHN1"
    struct Header {
        std::atomic<uint64_t> magic;
        uint64_t capacity, n_ranks, pad;
        std::atomic<uint64_t> ready_seq[2];  // sequence number of the plan the slot holds
        std::atomic<uint64_t> len[2];
        std::atomic<uint64_t> status[2];     // 0: a plan; else the VGPU_ERR_* of the publisher's failure, the slot holds its message
        std::atomic<uint64_t> consumed[2][kMaxRanks];
    };
    static_assert(std::atomic<uint64_t>::is_always_lock_free, "the channel needs lock-free 64-bit atomics");
    static const size_t kHeaderBytes = 4096;
    static_assert(sizeof(Header) <= kHeaderBytes, "header fits its page");

    PlanChannel(const std::string& name, uint32_t rank, uint32_t n_ranks, size_t capacity) : name_(name), rank_(rank), n_ranks_(n_ranks) {
        if (n_ranks == 0 || n_ranks > kMaxRanks || rank >= n_ranks) throw ChannelError("plan channel: invalid rank");
        if (name.empty() || name[0] != '/') name_ = "/" + name;
        capacity_ = (capacity + 4095) & ~(size_t)4095;
        map_len_ = kHeaderBytes + 2 * capacity_;
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(120);
        int fd = -1;
        if (rank == 0) {
            shm_unlink(name_.c_str());
            fd = shm_open(name_.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0) throw ChannelError("plan channel: shm_open(" + name_ + ") failed: " + strerror(errno));
            if (ftruncate(fd, (off_t)map_len_) != 0) {
                close(fd), shm_unlink(name_.c_str());
                throw ChannelError(std::string("plan channel: ftruncate failed: ") + strerror(errno));
            }
        } else {
            while (true) {  // until rank 0 has created and sized the segment
                fd = shm_open(name_.c_str(), O_RDWR, 0600);
                if (fd >= 0) {
                    struct stat st;
                    if (fstat(fd, &st) == 0 && (size_t)st.st_size == map_len_) break;
                    close(fd), fd = -1;
                }
                if (std::chrono::steady_clock::now() > deadline) throw ChannelError("plan channel: " + name_ + " did not appear (is local rank 0 running?)");
                usleep(1000);
            }
        }
        void* m = mmap(nullptr, map_len_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (m == MAP_FAILED) throw ChannelError(std::string("plan channel: mmap failed: ") + strerror(errno));
        base_ = static_cast<uint8_t*>(m);
        hdr_ = reinterpret_cast<Header*>(base_);
        if (rank == 0) {
            hdr_->capacity = capacity_, hdr_->n_ranks = n_ranks;
            hdr_->magic.store(kMagic, std::memory_order_release);  // the mapping starts zeroed: sequence numbers and counters are 0
        } else {
            while (hdr_->magic.load(std::memory_order_acquire) != kMagic) {
                if (std::chrono::steady_clock::now() > deadline) throw ChannelError("plan channel: " + name_ + " was never initialised");
                usleep(200);
            }
            if (hdr_->capacity != capacity_ || hdr_->n_ranks != n_ranks) throw ChannelError("plan channel: capacity / rank count differ from local rank 0's");
        }
    }
    ~PlanChannel() {
        if (base_) munmap(base_, map_len_);
        if (rank_ == 0) shm_unlink(name_.c_str());
    }
    PlanChannel(const PlanChannel&) = delete;
    PlanChannel& operator=(const PlanChannel&) = delete;

    uint32_t rank() const { return rank_; }
    size_t capacity() const { return capacity_; }

    // Sequence number of the caller's next batch (1, 2, ...).  Every rank takes its tickets in the same order (the order
    // the batches will be executed in); the prepares themselves may then run concurrently, two at a time.
    uint64_t ticket() { return next_ticket_.fetch_add(1, std::memory_order_relaxed) + 1; }

    // rank 0: plan number `seq` (status 0) or the failure the consumers must see instead of waiting for ever
    void publish(uint64_t seq, const void* data, size_t len, uint64_t status) {
        const uint32_t slot = (uint32_t)(seq & 1u);
        if (len > capacity_) throw ChannelError("plan channel: blob of " + std::to_string(len) + " bytes exceeds the capacity of " + std::to_string(capacity_));
        for (uint32_t r = 1; r < n_ranks_; ++r)  // the slot's previous plan (seq - 2) must have been taken by everyone
            wait_until([&]() { return hdr_->consumed[slot][r].load(std::memory_order_acquire) + 2 >= seq; }, "a consumer to take the previous plan");
        memcpy(base_ + kHeaderBytes + slot * capacity_, data, len);
        hdr_->len[slot].store(len, std::memory_order_relaxed);
        hdr_->status[slot].store(status, std::memory_order_relaxed);
        hdr_->ready_seq[slot].store(seq, std::memory_order_release);
    }

    // ranks > 0: waits for plan number `seq`; `use(data, len, status)` runs while the slot is held
    template <class F>
    void consume(uint64_t seq, F&& use) {
        const uint32_t slot = (uint32_t)(seq & 1u);
        wait_until([&]() { return hdr_->ready_seq[slot].load(std::memory_order_acquire) == seq; }, "local rank 0 to publish the plan");
        const size_t len = (size_t)hdr_->len[slot].load(std::memory_order_relaxed);
        const uint64_t status = hdr_->status[slot].load(std::memory_order_relaxed);
        struct Release {
            Header* h;
            uint32_t slot, rank;
            uint64_t seq;
            ~Release() { h->consumed[slot][rank].store(seq, std::memory_order_release); }
        } release{hdr_, slot, rank_, seq};
        use(base_ + kHeaderBytes + slot * capacity_, len, status);
    }

   private:
    template <class P>
    void wait_until(P&& ready, const char* what) {
        for (int i = 0; i < 2000; ++i) {
            if (ready()) return;
            sched_yield();
        }
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(300);
        while (!ready()) {
            if (std::chrono::steady_clock::now() > deadline) throw ChannelError(std::string("plan channel: timed out waiting for ") + what);
            usleep(50);
        }
    }
    std::string name_;
    uint32_t rank_, n_ranks_;
    size_t capacity_ = 0, map_len_ = 0;
    uint8_t* base_ = nullptr;
    Header* hdr_ = nullptr;
    std::atomic<uint64_t> next_ticket_{0};
};

}  // namespace vplan
