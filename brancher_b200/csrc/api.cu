// Error / variant reporting of the C ABI (include/brancher_cuda.h).
#include "common.cuh"
#include <vector>
#include <utility>

namespace brn {
static thread_local char g_err[512] = "";
static thread_local char g_variant[64] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void set_variant(const char* name) {
    strncpy(g_variant, name, sizeof(g_variant) - 1);
    g_variant[sizeof(g_variant) - 1] = 0;
}

// ---- profiling: event pairs per named stage, summed on collect --------------------------------
struct StageRec {
    char name[48];
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pairs;
    size_t used = 0;
    double total_ms = 0.0;
    long long calls = 0;
};
static bool g_prof = false;
static std::vector<StageRec*> g_stages;
static long long g_launches = 0;

void count_launch(int n) { g_launches += n; }

// ---- "data ready" event: lets the host overlap the host->device copy of a minibatch with the sampling stage --------
static thread_local cudaEvent_t g_data_ready = nullptr;
int wait_data_ready(cudaStream_t stream) {
    if (!g_data_ready) return 0;
    cudaError_t e = cudaStreamWaitEvent(stream, g_data_ready, 0);
    g_data_ready = nullptr;
    if (e != cudaSuccess) { set_error("cudaStreamWaitEvent(data ready): %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

static int stage_slot(const char* name) {
    for (size_t i = 0; i < g_stages.size(); ++i)
        if (strcmp(g_stages[i]->name, name) == 0) return (int)i;
    StageRec* r = new StageRec();
    strncpy(r->name, name, sizeof(r->name) - 1);
    r->name[sizeof(r->name) - 1] = 0;
    g_stages.push_back(r);
    return (int)g_stages.size() - 1;
}

StageTimer::StageTimer(const char* name, cudaStream_t s) : slot(-1), stream(s) {
    if (!g_prof) return;
    slot = stage_slot(name);
    StageRec* r = g_stages[slot];
    if (r->used == r->pairs.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        r->pairs.push_back({a, b});
    }
    cudaEventRecord(r->pairs[r->used].first, stream);
}
StageTimer::~StageTimer() {
    if (slot < 0) return;
    StageRec* r = g_stages[slot];
    cudaEventRecord(r->pairs[r->used].second, stream);
    r->used++;
}
}  // namespace brn

extern "C" void brn_profile_enable(int on) { brn::g_prof = on != 0; }

extern "C" int brn_profile_collect(void) {
    for (auto* r : brn::g_stages) {
        for (size_t i = 0; i < r->used; ++i) {
            cudaError_t e = cudaEventSynchronize(r->pairs[i].second);
            if (e != cudaSuccess) { brn::set_error("brn_profile_collect: %s", cudaGetErrorString(e)); return -2; }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, r->pairs[i].first, r->pairs[i].second);
            r->total_ms += ms;
            r->calls++;
        }
        r->used = 0;
    }
    return 0;
}
extern "C" int brn_profile_num_stages(void) { return (int)brn::g_stages.size(); }
extern "C" const char* brn_profile_stage(int i, double* total_ms, long long* calls) {
    if (i < 0 || i >= (int)brn::g_stages.size()) return nullptr;
    if (total_ms) *total_ms = brn::g_stages[i]->total_ms;
    if (calls) *calls = brn::g_stages[i]->calls;
    return brn::g_stages[i]->name;
}
extern "C" void brn_profile_reset(void) {
    for (auto* r : brn::g_stages) { r->total_ms = 0.0; r->calls = 0; r->used = 0; }
}
extern "C" long long brn_launch_count(void) { return brn::g_launches; }

extern "C" void brn_set_data_ready_event(void* event) { brn::g_data_ready = (cudaEvent_t)event; }

extern "C" int brn_abi_version(void) { return BRN_ABI_VERSION; }
extern "C" const char* brn_last_error(void) { return brn::g_err; }
extern "C" const char* brn_last_variant(void) { return brn::g_variant; }
