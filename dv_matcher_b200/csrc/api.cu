// Library-level entry points: version, per-thread error string, device check.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace dvm {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(); }

// CUDA-event brackets around the dominant (candidate-pass) kernel, for bench.py's live roofline.
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_prof_ev;      // start/stop pairs
static size_t g_prof_used = 0;
static bool g_prof_open = false;
constexpr size_t kProfMaxPairs = 4096;
void prof_begin(cudaStream_t st) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof_used / 2 >= kProfMaxPairs) return;
    if (g_prof_ev.size() < g_prof_used + 2) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        g_prof_ev.push_back(a); g_prof_ev.push_back(b);
    }
    cudaEventRecord(g_prof_ev[g_prof_used], st);
    g_prof_open = true;
}
void prof_end(cudaStream_t st) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_open) return;
    g_prof_open = false;
    cudaEventRecord(g_prof_ev[g_prof_used + 1], st);
    g_prof_used += 2;
}
}  // namespace dvm

extern "C" long long dvm_launch_count(void) { return dvm::launches(); }
extern "C" int dvm_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(dvm::g_prof_mu);
    dvm::g_prof_on = on != 0;
    dvm::g_prof_used = 0;
    return 0;
}
extern "C" int dvm_profile_read(double* total_ms, int* brackets) {
    std::lock_guard<std::mutex> lk(dvm::g_prof_mu);
    double tot = 0.0;
    int n = 0;
    for (size_t i = 0; i + 1 < dvm::g_prof_used; i += 2) {
        float ms = 0.f;
        DVM_CUDA(cudaEventSynchronize(dvm::g_prof_ev[i + 1]));
        DVM_CUDA(cudaEventElapsedTime(&ms, dvm::g_prof_ev[i], dvm::g_prof_ev[i + 1]));
        tot += ms; ++n;
    }
    if (total_ms) *total_ms = tot;
    if (brackets) *brackets = n;
    return 0;
}

extern "C" int dvm_version(void) { return DVM_VERSION; }
extern "C" const char* dvm_last_error_string(void) { return dvm::g_err; }
extern "C" int dvm_device_check(void) {
    int dev = 0;
    DVM_CUDA(cudaGetDevice(&dev));
    int major = 0;
    DVM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        dvm::set_error("libdvm_b200 is built for sm_100a only; current device has compute capability %d.x", major);
        return DVM_ERR_DEVICE;
    }
    return 0;
}
