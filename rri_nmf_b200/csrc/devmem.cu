#include "devmem.h"

#include <map>
#include <mutex>
#include <unordered_map>

namespace rri {

namespace {
constexpr size_t BLOCK_MAX = (size_t)256 << 20;        // do not keep blocks above 256 MB
constexpr size_t TOTAL_MAX = (size_t)2 << 30;          // nor more than 2 GB per device in total
struct DevCache {
    std::multimap<size_t, void*> free_blocks;
    std::unordered_map<void*, size_t> live;            // blocks handed out -> size
    size_t cached_bytes = 0;
};
std::mutex g_mu;
std::map<int, DevCache> g_cache;
}  // namespace

cudaError_t cached_malloc(void** p, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_mu);
    DevCache& c = g_cache[dev];
    auto it = c.free_blocks.find(bytes);
    if (it != c.free_blocks.end()) {
        *p = it->second;
        c.free_blocks.erase(it);
        c.cached_bytes -= bytes;
        c.live[*p] = bytes;
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess && !c.free_blocks.empty()) {
        // out of memory with blocks parked in the cache: release them and retry once
        cudaGetLastError();
        for (auto& kv : c.free_blocks) cudaFree(kv.second);
        c.free_blocks.clear();
        c.cached_bytes = 0;
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) c.live[*p] = bytes;
    return e;
}

void cached_free(void* p)
{
    if (!p) return;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_mu);
    DevCache& c = g_cache[dev];
    auto it = c.live.find(p);
    if (it == c.live.end()) { cudaFree(p); return; }
    const size_t bytes = it->second;
    c.live.erase(it);
    if (bytes > BLOCK_MAX || c.cached_bytes + bytes > TOTAL_MAX) { cudaFree(p); return; }
    c.free_blocks.emplace(bytes, p);
    c.cached_bytes += bytes;
}

void cache_trim()
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_mu);
    DevCache& c = g_cache[dev];
    for (auto& kv : c.free_blocks) cudaFree(kv.second);
    c.free_blocks.clear();
    c.cached_bytes = 0;
}

}  // namespace rri
