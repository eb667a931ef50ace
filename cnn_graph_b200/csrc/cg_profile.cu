// Launch accounting and optional per-kernel CUDA-event timing (used by bench.py to time the
// dominant kernels live on their launch stream; off by default -- zero device-side cost).
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "cg_common.cuh"

namespace {
struct Pending {
    const char *name;
    cudaEvent_t a, b;
};
std::mutex g_mu;
bool g_enabled = false;
std::vector<Pending> g_pending;
std::map<std::string, std::pair<double, int64_t>> g_totals;
std::atomic<int64_t> g_launches{0};
}  // namespace

CgProfScope::CgProfScope(const char *name, cudaStream_t stream) : name_(name), stream_(stream), active_(false) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_enabled) return;
    if (cudaEventCreate(&a_) != cudaSuccess || cudaEventCreate(&b_) != cudaSuccess) return;
    active_ = true;
    cudaEventRecord(a_, stream_);
}

CgProfScope::~CgProfScope() {
    if (!active_) return;
    cudaEventRecord(b_, stream_);
    std::lock_guard<std::mutex> lock(g_mu);
    g_pending.push_back({name_, a_, b_});
}

static void collect_locked() {
    for (const Pending &p : g_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            auto &slot = g_totals[p.name];
            slot.first += ms;
            slot.second += 1;
        }
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    g_pending.clear();
}

extern "C" int64_t cg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int cg_profile_enable(int on) {
    std::lock_guard<std::mutex> lock(g_mu);
    g_enabled = on != 0;
    return CG_OK;
}

extern "C" int cg_profile_reset(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    collect_locked();
    g_totals.clear();
    return CG_OK;
}

extern "C" int cg_profile_query(int index, char *name, int name_cap, double *total_ms, int64_t *count) {
    std::lock_guard<std::mutex> lock(g_mu);
    collect_locked();
    const int n = (int)g_totals.size();
    if (index < 0 || index >= n || !name || name_cap <= 0) return n;
    auto it = g_totals.begin();
    std::advance(it, index);
    snprintf(name, (size_t)name_cap, "%s", it->first.c_str());
    if (total_ms) *total_ms = it->second.first;
    if (count) *count = it->second.second;
    return n;
}
