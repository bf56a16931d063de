// Launch counter, per-category event profiler and library/device queries.
#include <vector>

#include "common.cuh"
#include "prof.h"

unsigned long long g_bhs_launches = 0;

namespace {
struct Rec {
    cudaEvent_t e0, e1;
    double work;
    int cat;
};
bool g_on = false;
std::vector<Rec> g_recs;
cudaEvent_t g_open[BHS_PROF_NCAT];
bool g_is_open[BHS_PROF_NCAT];
}  // namespace

void bhs_prof_begin(int cat, cudaStream_t st) {
    if (!g_on || cat < 0 || cat >= BHS_PROF_NCAT || g_is_open[cat]) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_open[cat] = e;
    g_is_open[cat] = true;
}

void bhs_prof_end(int cat, double work, cudaStream_t st) {
    if (!g_on || cat < 0 || cat >= BHS_PROF_NCAT || !g_is_open[cat]) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_recs.push_back(Rec{g_open[cat], e, work, cat});
    g_is_open[cat] = false;
}

static void prof_clear() {
    for (Rec& r : g_recs) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_recs.clear();
    for (int c = 0; c < BHS_PROF_NCAT; ++c) {
        if (g_is_open[c]) cudaEventDestroy(g_open[c]);
        g_is_open[c] = false;
    }
}

extern "C" int bhs_profile(int enable) {
    prof_clear();
    g_on = enable != 0;
    return BHS_OK;
}

extern "C" int bhs_profile_read(int cat, double* ms, double* work, int64_t* count) {
    if (cat < 0 || cat >= BHS_PROF_NCAT) return BHS_ERR_INVALID;
    double t = 0.0, w = 0.0;
    int64_t n = 0;
    for (Rec& r : g_recs) {
        if (r.cat != cat) continue;
        cudaError_t e = cudaEventSynchronize(r.e1);
        if (e != cudaSuccess) return (int)e;
        float f = 0.f;
        e = cudaEventElapsedTime(&f, r.e0, r.e1);
        if (e != cudaSuccess) return (int)e;
        t += f;
        w += r.work;
        ++n;
    }
    if (ms) *ms = t;
    if (work) *work = w;
    if (count) *count = n;
    return BHS_OK;
}

extern "C" int64_t bhs_launch_count(int reset) {
    unsigned long long v = __sync_fetch_and_add(&g_bhs_launches, 0ULL);
    if (reset) __sync_fetch_and_and(&g_bhs_launches, 0ULL);
    return (int64_t)v;
}

int bhs_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

extern "C" int bhs_version(void) { return 200; }
extern "C" int bhs_device_sm_count(int* out) {
    if (!out) return BHS_ERR_INVALID;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev);
    return e == cudaSuccess ? BHS_OK : (int)e;
}
