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

// CUDA-event brackets for bench.py's live rooflines.  Channel 0 = the dominant (candidate-pass) kernels,
// channel 1 = the whole fused op dvm_softmap_fwd (prep + prime + sweep + finalize + rescue); the brackets of different
// channels nest.  Events are recorded on the stream the kernels are launched on.
constexpr int kProfChannels = 2;
static std::mutex g_prof_mu;
static bool g_prof_on = false;
struct ProfChannel { std::vector<cudaEvent_t> ev; size_t used = 0; bool open = false; };
static ProfChannel g_prof[kProfChannels];
constexpr size_t kProfMaxPairs = 4096;
void prof_begin(cudaStream_t st, int ch) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfChannel& c = g_prof[ch];
    if (c.used / 2 >= kProfMaxPairs) return;
    if (c.ev.size() < c.used + 2) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        c.ev.push_back(a); c.ev.push_back(b);
    }
    cudaEventRecord(c.ev[c.used], st);
    c.open = true;
}
void prof_end(cudaStream_t st, int ch) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfChannel& c = g_prof[ch];
    if (!c.open) return;
    c.open = false;
    cudaEventRecord(c.ev[c.used + 1], st);
    c.used += 2;
}
static int prof_read(int ch, double* total_ms, int* brackets) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfChannel& c = g_prof[ch];
    double tot = 0.0;
    int n = 0;
    for (size_t i = 0; i + 1 < c.used; i += 2) {
        float ms = 0.f;
        DVM_CUDA(cudaEventSynchronize(c.ev[i + 1]));
        DVM_CUDA(cudaEventElapsedTime(&ms, c.ev[i], c.ev[i + 1]));
        tot += ms; ++n;
    }
    if (total_ms) *total_ms = tot;
    if (brackets) *brackets = n;
    return 0;
}
}  // namespace dvm

extern "C" long long dvm_launch_count(void) { return dvm::launches(); }
extern "C" int dvm_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(dvm::g_prof_mu);
    dvm::g_prof_on = on != 0;
    for (auto& c : dvm::g_prof) { c.used = 0; c.open = false; }
    return 0;
}
extern "C" int dvm_profile_read(double* total_ms, int* brackets) { return dvm::prof_read(0, total_ms, brackets); }
extern "C" int dvm_profile_read_channel(int channel, double* total_ms, int* brackets) {
    if (channel < 0 || channel >= dvm::kProfChannels) { dvm::set_error("dvm_profile_read_channel: bad channel %d", channel); return DVM_ERR_INVALID_ARG; }
    return dvm::prof_read(channel, total_ms, brackets);
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
