// Error reporting and device queries shared by every rb200 entry point.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";
unsigned long long g_rb_launches = 0;

int rb_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

struct DevInfo { int sm = 0, smem = 0; bool ok = false; };
static DevInfo& dev_info() {
    static thread_local DevInfo d[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { static thread_local DevInfo none; return none; }
    if (!d[dev].ok) {
        cudaDeviceGetAttribute(&d[dev].sm, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d[dev].smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        d[dev].ok = d[dev].sm > 0;
    }
    return d[dev];
}
int rb_sm_count() { int s = dev_info().sm; return s > 0 ? s : 148; }
int rb_max_smem_optin() { int s = dev_info().smem; return s > 0 ? s : 232448; }

extern "C" int rb200_version(void) { return RB200_VERSION; }
extern "C" const char* rb200_last_error(void) { return g_err; }
extern "C" int rb200_sm_count(void) { return rb_sm_count(); }
extern "C" uint64_t rb200_launch_count(void) { return g_rb_launches; }

// one thread: out[idx] = %globaltimer (ns).  A capturable timestamp: phase boundaries INSIDE a CUDA-graph replay, where events cannot go
__global__ void stamp_kernel(unsigned long long* out, int idx) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    out[idx] = t;
}
extern "C" int rb200_stamp(uint64_t* out, int idx, void* stream) {
    RB_REQUIRE(out && idx >= 0, "stamp: bad arguments");
    stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)out, idx);
    RB_LAUNCH_CHECK("stamp_kernel");
    return RB200_OK;
}

// sizeof() of the ABI structs, so that foreign-language bindings can verify their mirrors.
extern "C" size_t rb200_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(rb200_tower_job);
        case 1: return sizeof(rb200_tower_bwd_job);
        case 2: return sizeof(rb200_opt_state);
        case 3: return sizeof(rb200_step_params);
        case 4: return sizeof(rb200_step_views);
        case 5: return sizeof(rb200_sumsq_seg);
        case 6: return sizeof(rb200_sampler);
        default: return 0;
    }
}
