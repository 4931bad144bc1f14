// Device-resident hit lists of the step seam ("device-resident handles between steps").
//
// A hit list that stays on the device between two plan steps is an anchor-sorted array of SparseEntry
// (anchor id, score key) -- the very form in which the tile path reads the postings of rarely matched
// terms, so a list handle is consumed by the same kernels as any other leaf.  Two things are needed
// around them: the per-tile offsets of a sorted list (where the host builds them for host lists), and
// turning the unordered hits a step emits (tiles.cu: emit, one atomic append per hit) into such a list.
#include <cub/device/device_radix_sort.cuh>

#include "kernels.cuh"

namespace vdev {

// row[t] = first entry of the list with anchor >= anchor_lo + t * tile, t = 0 .. n_tiles (row[n_tiles] = n):
// the bucket row prepare_lists builds on the host for a host list.  One thread per tile boundary.
__global__ void list_bucket_kernel(const SparseEntry* __restrict__ entries, uint32_t n, uint32_t anchor_lo, uint32_t tile_log2, uint32_t n_tiles, uint32_t* __restrict__ row) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if (t == n_tiles) {
        row[t] = n;
        return;
    }
    const uint64_t first = (uint64_t)anchor_lo + ((uint64_t)t << tile_log2);
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if ((uint64_t)entries[mid].anchor < first) lo = mid + 1;
        else hi = mid;
    }
    row[t] = lo;
}
void launch_list_bucket(cudaStream_t st, const SparseEntry* entries, uint32_t n, uint32_t anchor_lo, uint32_t tile_log2, uint32_t n_tiles, uint32_t* row) {
    list_bucket_kernel<<<(n_tiles + 1 + 127) / 128, 128, 0, st>>>(entries, n, anchor_lo, tile_log2, n_tiles, row);
    count_launch();
}

// emitted hit ((score key << 32) | anchor) -> sortable ((anchor << 32) | key); a key of 0 (score +0.0) becomes 1 as
// for host lists: 0 marks an empty slot of the tile arrays
__global__ void emitted_to_sortable_kernel(unsigned long long* __restrict__ buf, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long v = buf[i];
    const uint32_t anchor = (uint32_t)v;
    uint32_t key = (uint32_t)(v >> 32);
    if (key == 0) key = 1;
    buf[i] = ((unsigned long long)anchor << 32) | key;
}
// sortable -> SparseEntry {anchor, key} (little endian: anchor in the low word)
__global__ void sortable_to_entry_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long v = in[i];
    out[i] = (v << 32) | (v >> 32);
}

size_t emitted_sort_temp_bytes(uint32_t n) {
    size_t bytes = 0;
    cub::DoubleBuffer<unsigned long long> keys(nullptr, nullptr);
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, keys, (int)n, 32, 64);
    return bytes;
}
// The `n` hits a step emitted into `buf`, as an anchor-sorted SparseEntry list written to `out` (anchors are unique: the
// sort runs over the anchor bits only).  `alt` is scratch of the same size; `buf` is overwritten.
cudaError_t launch_emitted_to_list(cudaStream_t st, unsigned long long* buf, unsigned long long* alt, unsigned long long* out, uint32_t n, void* temp, size_t temp_bytes) {
    if (!n) return cudaSuccess;
    emitted_to_sortable_kernel<<<(n + 255) / 256, 256, 0, st>>>(buf, n);
    count_launch();
    cub::DoubleBuffer<unsigned long long> keys(buf, alt);
    cudaError_t e = cub::DeviceRadixSort::SortKeys(temp, temp_bytes, keys, (int)n, 32, 64, st);
    if (e != cudaSuccess) return e;
    sortable_to_entry_kernel<<<(n + 255) / 256, 256, 0, st>>>(keys.Current(), out, n);
    count_launch();
    return cudaGetLastError();
}

}  // namespace vdev
