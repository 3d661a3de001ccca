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
struct TimedLaunch { const char *name; cudaEvent_t start, stop; unsigned flags; cudaStream_t stream; };
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
    // inside a stream capture the records must become event-record NODES of the graph (cudaEventRecordExternal): then every
    // replay re-records them and mb_profile_timeline shows when each kernel of the replayed graph ran
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    t.flags = cap == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault;
    t.stream = s;
    cudaEventRecordWithFlags(t.start, s, t.flags);
    g_launches.push_back(t);
    slot = (int)g_launches.size() - 1;
}

KernelTimer::~KernelTimer() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_profile_mu);
    if (slot < (int)g_launches.size()) cudaEventRecordWithFlags(g_launches[slot].stop, stream, g_launches[slot].flags);
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

// Synchronises the device and writes one line "name stream start_us duration_us" per recorded launch (times relative to the
// earliest start), WITHOUT clearing the record: launches recorded during a graph capture are re-recorded by every replay.
extern "C" int mb_profile_timeline(char *buf, size_t cap) {
    if (cudaDeviceSynchronize() != cudaSuccess) return MB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(mb::g_profile_mu);
    if (buf && cap) buf[0] = 0;
    if (mb::g_launches.empty()) return MB_OK;
    const cudaEvent_t ref = mb::g_launches[0].start;
    std::vector<float> st(mb::g_launches.size(), 0.f), du(mb::g_launches.size(), -1.f);
    float t_min = 0.f;
    for (size_t i = 0; i < mb::g_launches.size(); ++i) {
        float a = 0.f, d = 0.f;
        if (cudaEventElapsedTime(&a, ref, mb::g_launches[i].start) != cudaSuccess ||
            cudaEventElapsedTime(&d, mb::g_launches[i].start, mb::g_launches[i].stop) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        st[i] = a; du[i] = d;
        if (a < t_min) t_min = a;
    }
    std::map<cudaStream_t, int> ids;
    size_t off = 0;
    for (size_t i = 0; i < mb::g_launches.size(); ++i) {
        if (du[i] < 0.f) continue;
        const int sid = ids.emplace(mb::g_launches[i].stream, (int)ids.size()).first->second;
        int n = snprintf(buf ? buf + off : nullptr, buf && cap > off ? cap - off : 0, "%s %d %.3f %.3f\n", mb::g_launches[i].name, sid,
                         (st[i] - t_min) * 1e3f, du[i] * 1e3f);
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
