// Multi-GPU exchange of the anchor-range shards: NCCL over NVLink / NVSwitch, driven from inside the
// library so that a host in any language gets the sharded search from one call (vgpu_comm_init, then
// vgpu_batch_execute on every rank).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library loads on a CPU-only box, a process that already
// carries an NCCL (torch's) shares it, and nothing here needs NCCL's headers.  Only the stable C entry points below are
// used; enum values are NCCL's (nccl.h: ncclUint32 = 3, ncclUint64 = 5, ncclSum = 0, ncclMax = 2).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>

namespace vdev {

struct NcclError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct NcclUniqueId {
    char internal[128];
};
typedef struct ncclComm* NcclCommHandle;

class NcclApi {
   public:
    static NcclApi& get() {
        static NcclApi api;
        return api;
    }
    enum : int { kUint32 = 3, kUint64 = 5, kSum = 0, kMax = 2 };
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclCommHandle*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclCommHandle) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclCommHandle, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclCommHandle, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;

    void require() {
        std::lock_guard<std::mutex> g(mu_);
        if (handle_) return;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        std::string tried;
        for (const char* n : names) {
            handle_ = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle_) break;
            tried += std::string(tried.empty() ? "" : "; ") + dlerror();
        }
        if (!handle_) throw NcclError("NCCL is not available: " + tried);
        load(GetUniqueId, "ncclGetUniqueId"), load(CommInitRank, "ncclCommInitRank"), load(CommDestroy, "ncclCommDestroy");
        load(AllReduce, "ncclAllReduce"), load(AllGather, "ncclAllGather"), load(GroupStart, "ncclGroupStart"), load(GroupEnd, "ncclGroupEnd");
        load(GetErrorString, "ncclGetErrorString"), load(GetVersion, "ncclGetVersion");
    }
    void check(int rc, const char* what) {
        if (rc != 0) throw NcclError(std::string(what) + " failed: " + (GetErrorString ? GetErrorString(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
    }

   private:
    template <class F>
    void load(F& fn, const char* name) {
        fn = reinterpret_cast<F>(dlsym(handle_, name));
        if (!fn) {
            handle_ = nullptr;
            throw NcclError(std::string("NCCL symbol missing: ") + name);
        }
    }
    std::mutex mu_;
    void* handle_ = nullptr;
};

// One communicator per index handle: rank r of n holds anchor-range shard r of n.
struct ShardComm {
    NcclCommHandle comm = nullptr;
    uint32_t rank = 0, n_ranks = 1;
    ~ShardComm() {
        if (comm) NcclApi::get().CommDestroy(comm);
    }
    void all_reduce_max_u64(void* buf, size_t count, cudaStream_t st) {
        NcclApi& api = NcclApi::get();
        api.check(api.AllReduce(buf, buf, count, NcclApi::kUint64, NcclApi::kMax, comm, st), "ncclAllReduce(max)");
    }
    void all_reduce_sum_u32(void* buf, size_t count, cudaStream_t st) {
        NcclApi& api = NcclApi::get();
        api.check(api.AllReduce(buf, buf, count, NcclApi::kUint32, NcclApi::kSum, comm, st), "ncclAllReduce(sum)");
    }
    // recv = [n_ranks][count] in rank order
    void all_gather_u64(const void* send, void* recv, size_t count, cudaStream_t st) {
        NcclApi& api = NcclApi::get();
        api.check(api.AllGather(send, recv, count, NcclApi::kUint64, comm, st), "ncclAllGather");
    }
    void group_start() { NcclApi::get().check(NcclApi::get().GroupStart(), "ncclGroupStart"); }
    void group_end() { NcclApi::get().check(NcclApi::get().GroupEnd(), "ncclGroupEnd"); }
};

}  // namespace vdev
