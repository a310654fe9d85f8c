// Shared helpers for the rb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rb200.h"

int rb_set_error(int code, const char* fmt, ...);

#define RB_REQUIRE(cond, ...)                                             \
    do {                                                                  \
        if (!(cond)) return rb_set_error(RB200_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define RB_CUDA(call)                                                                             \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return rb_set_error(RB200_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, \
                                cudaGetErrorString(e__));                                         \
    } while (0)

extern unsigned long long g_rb_launches;   // kernels of this library launched by this process (not cub's)

#define RB_LAUNCH_CHECK(name)                                                                     \
    do {                                                                                          \
        ++g_rb_launches;                                                                          \
        cudaError_t e__ = cudaGetLastError();                                                     \
        if (e__ != cudaSuccess)                                                                   \
            return rb_set_error(RB200_ERR_CUDA, "launch of %s failed: %s", name,                  \
                                cudaGetErrorString(e__));                                         \
    } while (0)

int rb_sm_count();
int rb_max_smem_optin();

static inline size_t rb_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (all blocks 256-byte aligned).
struct RbArena {
    char* base;
    size_t size, off;
    RbArena(void* p, size_t n) : base((char*)p), size(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t o = rb_align_up(off, 256);
        off = o + count * sizeof(T);
        return (T*)(base + o);
    }
    bool ok() const { return off <= size; }
};

#ifdef __CUDACC__

#define RB_FULL_MASK 0xffffffffu

__device__ __forceinline__ float rb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double rb_warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RB_FULL_MASK, v, o);
    return v;
}

// leading dimension (in floats) for a shared-memory operand read with 128-bit loads by lanes
// that sit on different rows: ld/4 must be odd so that 8 consecutive rows cover all 32 banks.
__host__ __device__ __forceinline__ int rb_ld_odd4(int k) { return ((k / 4) & 1) ? k : k + 4; }

// softplus(x) = max(x,0) + log1p(exp(-|x|))   (torch's -logsigmoid(-x), numerically stable)
__device__ __forceinline__ float rb_softplus(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float rb_sigmoid(float x) {
    float e = expf(-fabsf(x));
    float r = 1.f / (1.f + e);
    return x >= 0.f ? r : e * r;
}

// Philox4x32-10 counter-based RNG (Salmon et al. 2011), used for in-kernel dropout masks.
__device__ __forceinline__ uint4 rb_philox4x32(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float rb_u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }

#endif  // __CUDACC__

// per-step constants of Adam (bias corrections, step size) and reset of the grad-norm accumulator; one thread
__device__ __forceinline__ void rb_opt_begin_step_dev(rb200_opt_state* st) {
    const long long step = st->step + 1;
    st->step = step;
    const double bc1 = 1.0 - pow(st->beta1, (double)step);
    const double bc2 = 1.0 - pow(st->beta2, (double)step);
    st->step_size = (float)(st->lr / bc1);
    st->bias_corr2_sqrt = (float)sqrt(bc2);
    st->sumsq = 0.0;
    st->ticket = 0u;
}

// split-K partials of a tower's weight gradients left unreduced by rb_tower_bwd (tensor-core modes): nsplit blocks of P
// floats, the first H·Din of each holding W1 transposed ([k][h]); nsplit == 0 → the gradient is all zeros
struct RbPartials { const float* part; int nsplit; int P; int H; int Din; };
