// Error plumbing and device queries behind the C ABI.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace mb {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what, bool debug_sync, cudaStream_t s) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && debug_sync) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MB_ERR_CUDA;
    }
    return MB_OK;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}


// ---- per-kernel event timing ---------------------------------------------------------------------------
struct TimedLaunch { const char *name; cudaEvent_t start, stop; };
static bool g_profile = false;
static std::mutex g_profile_mu;
static std::vector<TimedLaunch> g_launches;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_event_pool;

KernelTimer::KernelTimer(const char *name, cudaStream_t s) : slot(-1), stream(s) {
    if (!g_profile) return;
    std::lock_guard<std::mutex> lk(g_profile_mu);
    TimedLaunch t;
    t.name = name;
    if (!g_event_pool.empty()) {
        t.start = g_event_pool.back().first;
        t.stop = g_event_pool.back().second;
        g_event_pool.pop_back();
    } else if (cudaEventCreate(&t.start) != cudaSuccess || cudaEventCreate(&t.stop) != cudaSuccess) {
        return;
    }
    cudaEventRecord(t.start, s);
    g_launches.push_back(t);
    slot = (int)g_launches.size() - 1;
}

KernelTimer::~KernelTimer() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_profile_mu);
    if (slot < (int)g_launches.size()) cudaEventRecord(g_launches[slot].stop, stream);
}

}  // namespace mb

extern "C" void mb_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(mb::g_profile_mu);
    mb::g_profile = on != 0;
}

// Synchronises the device, writes one line "name launches total_ms" per kernel into buf, clears the record.
extern "C" int mb_profile_report(char *buf, size_t cap) {
    if (cudaDeviceSynchronize() != cudaSuccess) return MB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(mb::g_profile_mu);
    std::map<std::string, std::pair<int, double>> acc;
    for (auto &t : mb::g_launches) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.start, t.stop) == cudaSuccess) {
            auto &a = acc[t.name];
            a.first += 1;
            a.second += ms;
        }
        mb::g_event_pool.push_back({t.start, t.stop});
    }
    mb::g_launches.clear();
    size_t off = 0;
    if (buf && cap) buf[0] = 0;
    for (auto &kv : acc) {
        int n = snprintf(buf ? buf + off : nullptr, buf && cap > off ? cap - off : 0, "%s %d %.6f\n", kv.first.c_str(), kv.second.first,
                         kv.second.second);
        if (n < 0) break;
        off += (size_t)n;
        if (off >= cap) break;
    }
    return MB_OK;
}

extern "C" int mb_version(void) { return 100; }

extern "C" const char *mb_last_error(void) { return mb::g_error; }

extern "C" int mb_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}
