// Pieces shared by the two tcgen05 tower implementations: tower_tc.cu (both operands in shared memory, D = 64) and
// tower_ts.cu (activations as the TMEM A operand, D = 64 or 128).
#pragma once
#include "common.cuh"
#include "tower_common.cuh"
#include "umma.cuh"

namespace towertc {

constexpr int ROWS = 128;      // samples per tile = TMEM lanes
constexpr int SUBR = 64;       // rows of one sub-image of the N-split operands (W2, W1[:, :D]ᵀ)

// ---- weight images --------------------------------------------------------------------------------------------- //
// [W1 hi|lo : H × Kp] [W2 : D/64 × (hi|lo : 64 × H)] [W2ᵀ hi|lo : H × D] [W1[:, :D]ᵀ : D/64 × (hi|lo : 64 × H)], each block in the
// K-major UMMA layout.  The operands whose rows are the D output columns are cut into sub-images of 64 rows so that a kernel
// can keep one resident and stream the other (D = 128: all four do not fit 227 KB); at D = 64 there is one sub-image and the
// layout is that of a plain [D × H] operand.
struct ImgLayout {
    size_t w1, w2, w2t, w1t, total;
    int Kp;
};
__host__ __device__ inline ImgLayout img_layout(int D, int H, int E) {
    ImgLayout L;
    L.Kp = (D + E + 7) & ~7;
    L.w1 = 0;
    L.w2 = L.w1 + (size_t)2 * H * L.Kp * 4;
    L.w2t = L.w2 + (size_t)2 * D * H * 4;
    L.w1t = L.w2t + (size_t)2 * H * D * 4;
    L.total = L.w1t + (size_t)2 * D * H * 4;
    return L;
}
constexpr size_t SUB_BYTES = (size_t)2 * SUBR * 128 * 4;     // one sub-image (hi|lo) at H = 128: 65 536 B

#ifdef __CUDACC__
// store 4 consecutive-k values of row r (hi and optionally lo) into an [R × K] K-major operand
template <int MODE>
__device__ __forceinline__ void put4(unsigned char* hi_base, unsigned char* lo_base, int R, int r, int k, const float4& v) {
    const uint32_t off = umma::kmajor_offset(R, r, k);
    float4 hi, lo;
    umma::split4(v, hi, lo);
    *reinterpret_cast<float4*>(hi_base + off) = hi;
    if (MODE == 2) *reinterpret_cast<float4*>(lo_base + off) = lo;
}

struct Bar {           // mbarrier with bounded waits and a sticky failure flag
    uint64_t* bar;
    uint32_t phase;
    int* dead;
    int* err_flag;
    __device__ void wait() {
        if (!*dead && !umma::mbar_wait(bar, phase)) { *dead = 1; if (err_flag) atomicOr(err_flag, 2); }
        phase ^= 1;
    }
};

__device__ __forceinline__ int find_job(const int* begin, int n_jobs) {
    int j = 0;
#pragma unroll
    for (int t = 1; t < MAX_JOBS; ++t)
        if (t < n_jobs && (int)blockIdx.x >= begin[t]) j = t;
    return j;
}

__device__ __forceinline__ float sel4(const float4& v, int e) { return e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w; }

// Transposing store of a 4(sample) × 4(row) block into an [R × KC] K-major operand: thread (sq, mq) holds
// v[i] = 4 consecutive operand rows (4·mq … +3) of sample 4·sq+i and writes, for each of its rows, the 16-byte unit of
// the 4 samples.  The row handled in store t is rotated with the lane (e = (t + lane/2) & 3) so that the 8 lanes of a
// shared-memory phase hit 8 different 16-byte slots: conflict-free.
template <int MODE>
__device__ __forceinline__ void put_block_t(unsigned char* hi, unsigned char* lo, int R, int mq, int sq, int lane, const float4 (&v)[4]) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int e = (t + (lane >> 1)) & 3;
        put4<MODE>(hi, lo, R, mq * 4 + e, sq * 4, make_float4(sel4(v[0], e), sel4(v[1], e), sel4(v[2], e), sel4(v[3], e)));
    }
}
#endif

template <typename JobT>
inline int assign_tiles(JobT* jobs, int n_jobs) {
    int begin = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const int tiles = (jobs[j].B + ROWS - 1) / ROWS;
        jobs[j].cta_begin = begin;
        jobs[j].cta_count = tiles;
        begin += tiles;
    }
    return begin;
}

template <typename K>
inline int set_smem(K kernel, size_t bytes) {
    RB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return RB200_OK;
}

}  // namespace towertc
