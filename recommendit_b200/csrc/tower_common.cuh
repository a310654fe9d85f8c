// Job descriptors shared by the FFMA (tower.cu) and tcgen05 (tower_tc.cu) tower kernels.
#pragma once
#include "common.cuh"

constexpr int MAX_JOBS = 3;
constexpr float NORM_EPS = 1e-12f;

struct FwdJob {
    const float* table; const int64_t* ids; const float* extra;
    const float* W1; const float* b1; const float* W2; const float* b2;
    float* out; float* hid; float* denom; const uint8_t* keep_mask; const unsigned char* img;
    long long n_rows; int B; int E; int extra_by_id; int cta_begin; int cta_count;
};
struct FwdParams {
    FwdJob job[MAX_JOBS];
    int n_jobs; float drop_p; unsigned long long seed, offset; const long long* offset_dev; int* err_flag;
};


struct BwdJob {
    const float* table; const int64_t* ids; const float* extra; long long n_rows; int B; int E;
    const float* W1; const float* W2;
    const float* dY; const float* y; const float* denom; const float* hid;
    float* dpre; float* dact; float* dRows; const unsigned char* img;
    int extra_by_id; int cta_begin; int cta_count;
};
struct BwdParams {
    BwdJob job[MAX_JOBS];
    int n_jobs; float keep_scale;
    float* part; int nsplit; int P;
};


#ifdef __CUDACC__
// Dropout keep-decision for hidden unit (row, col): Philox4x32-10 counter (row, col/4, offset), word col%4.
// Both tower implementations use this mapping, so the precision modes draw identical masks.
__device__ __forceinline__ bool rb_dropout_keep(unsigned long long seed, unsigned long long off, int row, int col, float p) {
    const uint4 o = rb_philox4x32(make_uint4((uint32_t)row, (uint32_t)(col >> 2), (uint32_t)off, (uint32_t)(off >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t w = (col & 3) == 0 ? o.x : (col & 3) == 1 ? o.y : (col & 3) == 2 ? o.z : o.w;
    return rb_u01(w) >= p;
}
#endif
