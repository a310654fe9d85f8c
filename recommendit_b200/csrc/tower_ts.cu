// Two-tower MLP on tcgen05 with the ACTIVATIONS AS THE TMEM A OPERAND (TS-mode MMAs) — modes 1 (TF32) and 2 (3xTF32) of
// rb200_tower_fwd / rb200_tower_bwd at D = 128 (BASELINE config C4) and, optionally, D = 64.  Same math as tower.cu / tower_tc.cu
// (reference src/models/two_tower.py:39-42, 68-72 and their autograd).
//
// Why a second tensor-core implementation: at D = 128 the shared-memory-operand kernels of tower_tc.cu would need 311 KB per
// 128-sample tile (X hi/lo 152 KB + W1 hi/lo 152 KB).  Here the activation tile lives in TMEM (thread = sample = TMEM lane, one
// 32-bit column per k, hi and lo images side by side) and only the WEIGHT image is in shared memory:
//
//   TMEM (512 columns)   [0,128) accumulator (GEMM1: N = H = 128; GEMM2: N = D) | A hi | A lo  (X, then the hidden tile)
//   shared memory        R0: the K-wide weight image (forward W1 [H × Kp], backward W2ᵀ [H × D])  ≤ 152 KB
//                        R1: sub-image 0 of the N-split operand (forward W2 rows 0-63, backward W1[:, :D]ᵀ rows 0-63)   64 KB
//                        D = 64 : R2 = 64 KB staging tile (hidden / dact leave the SM as 512-byte row segments)
//                        D = 128: once GEMM1 has completed R0 is dead → sub-image 1 streams into R0[0, 64 KB) and the staging
//                                 tile is R0[64 KB, 128 KB)
//
// so that every weight byte arrives by bulk asynchronous copies (TMA engine) issued before they are needed, the gathered rows
// never touch shared memory, and GEMM2 of D = 128 runs as two N = 64 halves, the second of which lands while epilogue 1 runs.
//
//   bwd wts   split over the batch AND over the two products: blockIdx.y = 0 computes dW2ᵀ = hidᵀ·dpre (+ db2), blockIdx.y = 1
//             dW1 = dactᵀ·X (+ db1).  A = hidᵀ / dactᵀ needs NO transposition in TS mode: TMEM lane = hidden unit h, column =
//             sample, and a warp's 32 lanes read 32 consecutive h of one sample (128-byte coalesced loads).  B = dpreᵀ / Xᵀ is
//             transposed in registers while staged (put_block_t).  Each 32-sample chunk is a fresh TMEM accumulation flushed into
//             fp32 registers (tower_tc.cu explains why: the tensor core's accumulation is not round-to-nearest).
#include "tower_tc_common.cuh"

using namespace towertc;

namespace {

constexpr int NTD = 512;                 // forward / backward-data: 16 warps = 4 TMEM lane quadrants × 4 column parts
constexpr int NTW = 256;                 // weight gradients: 8 warps = 4 lane quadrants × 2 column halves
constexpr int HH = 128;                  // hidden width
constexpr uint32_t TM_COLS = 512, ACC = 0, A_HI = 128;
constexpr size_t STAGE_BYTES = (size_t)ROWS * HH * 4;
constexpr int KC = 32;                   // samples per weight-gradient chunk

template <int N> struct TmemLd;
template <> struct TmemLd<16> { static __device__ __forceinline__ void ld(uint32_t a, float (&v)[16]) { umma::tmem_ld16(a, v); } };
template <> struct TmemLd<32> { static __device__ __forceinline__ void ld(uint32_t a, float (&v)[32]) { umma::tmem_ld32(a, v); } };

// D[tmem] = A[tmem: 128 lanes × K columns, hi at a_hi / lo at a_lo] · B[N × K]ᵀ (shared memory, K-major, RB rows); one thread
template <int MODE>
__device__ __forceinline__ void issue_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, const unsigned char* b_hi, const unsigned char* b_lo,
                                         int RB, int N, int K) {
    const uint32_t idesc = umma::idesc_tf32(128, N);
    const uint32_t lbo_b = (uint32_t)(RB / 8) * 128;
    const uint64_t dbh = umma::smem_desc(umma::smem_u32(b_hi), lbo_b, 128), dbl = umma::smem_desc(umma::smem_u32(b_lo), lbo_b, 128);
    const uint64_t sb = (uint64_t)((2 * lbo_b) >> 4);
    uint64_t ob = 0;
    for (int j = 0; j < K / 8; ++j, ob += sb) {
        if (MODE == 2) {
            umma::mma_tf32_ts(d, a_lo + 8 * j, dbh + ob, idesc, j > 0);
            umma::mma_tf32_ts(d, a_hi + 8 * j, dbl + ob, idesc, true);
            umma::mma_tf32_ts(d, a_hi + 8 * j, dbh + ob, idesc, true);
        } else {
            umma::mma_tf32_ts(d, a_hi + 8 * j, dbh + ob, idesc, j > 0);
        }
    }
}

// 16 fp32 values → hi (and lo) images, 16 consecutive TMEM columns of this thread's lane
template <int MODE>
__device__ __forceinline__ void st16_split(uint32_t hi_addr, uint32_t lo_addr, const float* x) {
    float hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { hi[i] = umma::tf32_hi(x[i]); lo[i] = x[i] - hi[i]; }
    umma::tmem_st16(hi_addr, hi);
    if (MODE == 2) umma::tmem_st16(lo_addr, lo);
}

// staging tile [128 rows × 128 floats]: 16-byte chunk c of row r sits at chunk (c ^ (r & 7)) — row-per-thread writes and
// row-per-warp reads are both conflict-free
__device__ __forceinline__ float4* stage_at(unsigned char* stage, int r, int c) {
    return reinterpret_cast<float4*>(stage + (size_t)r * 512 + ((c ^ (r & 7)) << 4));
}

// the staging tile → global rows (512-byte segments), 8 rows per warp, all loads issued first
__device__ __forceinline__ void stage_copy_out(unsigned char* stage, float* dst, int row0, int B, int warp, int lane) {
    float4 x[ROWS / (NTD / 32)];
#pragma unroll
    for (int i = 0; i < ROWS / (NTD / 32); ++i) x[i] = *stage_at(stage, warp + i * (NTD / 32), lane);
#pragma unroll
    for (int i = 0; i < ROWS / (NTD / 32); ++i) {
        const int r = warp + i * (NTD / 32);
        if (row0 + r < B) *reinterpret_cast<float4*>(dst + (long long)(row0 + r) * HH + lane * 4) = x[i];
    }
}

template <int D> __host__ __device__ constexpr size_t ts_r0(size_t big) { return D == 128 ? (big > 2 * SUB_BYTES ? big : 2 * SUB_BYTES) : big; }
template <int D> __host__ __device__ constexpr size_t ts_smem(size_t big) { return ts_r0<D>(big) + SUB_BYTES + (D == 128 ? 0 : STAGE_BYTES); }

#define RB_TS_PROLOGUE(err_ptr)                                                                                   \
    __shared__ __align__(8) uint64_t bars[4];        /* 0: big image, 1: sub-image 0, 2: sub-image 1, 3: MMAs */ \
    __shared__ uint32_t tmem_slot;                                                                                \
    __shared__ int dead;                                                                                          \
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                                                \
    if (warp == 0) umma::tmem_alloc(&tmem_slot, TM_COLS);                                                         \
    if (tid == 0) {                                                                                               \
        for (int i = 0; i < 4; ++i) umma::mbar_init(&bars[i], 1);                                                 \
        umma::fence_mbar_init();                                                                                  \
        dead = 0;                                                                                                 \
    }                                                                                                             \
    umma::fence_before_sync();                                                                                    \
    __syncthreads();                                                                                              \
    umma::fence_after_sync();                                                                                     \
    const uint32_t tmem = tmem_slot;                                                                              \
    Bar big_bar{&bars[0], 0u, &dead, err_ptr}, s0_bar{&bars[1], 0u, &dead, err_ptr}, mma_bar{&bars[3], 0u, &dead, err_ptr}; \
    const int r_own = ((warp & 3) << 5) + lane, cp = warp >> 2;      /* TMEM lane (= tile row) and column part */   \
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;

// ------------------------------------------------------------------------------------------------------------ //
// forward
// ------------------------------------------------------------------------------------------------------------ //
template <int D, int MODE>
__global__ void __launch_bounds__(NTD, 1) tower_fwd_ts_kernel(const FwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ float ss_part[4][ROWS];
    __shared__ long long ids_s[ROWS];
    constexpr int NH = D / SUBR;
    constexpr uint32_t KA = D == 128 ? 160 : 128, A_LO = A_HI + KA;
    static_assert(ACC + 128 <= A_HI && A_LO + KA <= TM_COLS, "TMEM budget");
    int begins[MAX_JOBS];
#pragma unroll
    for (int t = 0; t < MAX_JOBS; ++t) begins[t] = p.job[t].cta_begin;
    const int j = find_job(begins, p.n_jobs);
    const FwdJob J = p.job[j];
    const int E = J.E, Din = D + E;
    const ImgLayout L = img_layout(D, HH, E);
    const int Kp = L.Kp;
    const int row0 = ((int)blockIdx.x - J.cta_begin) * ROWS;
    if (threadIdx.x < ROWS) {
        const int row = row0 + (int)threadIdx.x;
        long long id = row < J.B ? J.ids[row] : 0;
        if ((unsigned long long)id >= (unsigned long long)J.n_rows) { if (p.err_flag) atomicOr(p.err_flag, 1); id = 0; }
        ids_s[threadIdx.x] = id;
    }
    RB_TS_PROLOGUE(p.err_flag)

    const size_t w1_bytes = (size_t)2 * HH * Kp * 4;
    unsigned char* w1_img = smem;
    unsigned char* sub0 = smem + ts_r0<D>(w1_bytes);
    unsigned char* sub1 = smem;                                   // D = 128: lands in R0 after GEMM1
    unsigned char* stage = D == 128 ? smem + SUB_BYTES : sub0 + SUB_BYTES;
    constexpr uint32_t sub_copy = (uint32_t)((MODE == 2 ? 2 : 1) * SUBR * HH * 4);
    if (tid == 0) {
        const uint32_t b1 = (uint32_t)((MODE == 2 ? 2 : 1) * HH * Kp * 4);
        umma::mbar_expect_tx(&bars[0], b1);
        umma::bulk_g2s(w1_img, J.img + L.w1, b1, &bars[0]);
        umma::mbar_expect_tx(&bars[1], sub_copy);
        umma::bulk_g2s(sub0, J.img + L.w2, sub_copy, &bars[1]);
    }

    // ---- stage 1: the gathered rows [128 × Kp] → TMEM (thread = row; column part cp takes the 16-column groups cp, cp+4, …) -- //
    {
        constexpr int GI = D == 128 ? 3 : 2;
        const int G = (Kp + 15) >> 4;
        const bool rv = row0 + r_own < J.B;
        const long long id = ids_s[r_own];
        float4 v[GI][4];
#pragma unroll
        for (int gi = 0; gi < GI; ++gi) {
            const int g = cp + 4 * gi;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = g * 16 + i * 4;
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g < G && rv && k < Kp) {
                    if (k < D) {
                        x = __ldg(reinterpret_cast<const float4*>(J.table + id * D + k));
                    } else {
                        const float* ex = J.extra + (J.extra_by_id ? id : (long long)(row0 + r_own)) * E + (k - D);
                        x.x = k + 0 < Din ? __ldg(ex + 0) : 0.f;
                        x.y = k + 1 < Din ? __ldg(ex + 1) : 0.f;
                        x.z = k + 2 < Din ? __ldg(ex + 2) : 0.f;
                        x.w = k + 3 < Din ? __ldg(ex + 3) : 0.f;
                    }
                }
                v[gi][i] = x;
            }
        }
#pragma unroll
        for (int gi = 0; gi < GI; ++gi) {
            const int g = cp + 4 * gi;
            if (g < G)                                            // warp-uniform
                st16_split<MODE>(tmem + lane_off + A_HI + g * 16, tmem + lane_off + A_LO + g * 16, &v[gi][0].x);
        }
        umma::tmem_st_wait();
    }
    umma::fence_before_sync();
    big_bar.wait();
    __syncthreads();
    if (!dead && umma::elect_issuer(tid)) {
        umma::fence_after_sync();
        issue_ts<MODE>(tmem + ACC, tmem + A_HI, tmem + A_LO, w1_img, w1_img + (size_t)HH * Kp * 4, HH, HH, Kp);
        umma::commit(&bars[3]);
    }
    float4 bias1[8];                       // this thread's 32 hidden-unit biases, fetched while GEMM1 runs
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) bias1[i4] = __ldg(reinterpret_cast<const float4*>(J.b1 + cp * 32) + i4);
    mma_bar.wait();
    umma::fence_after_sync();
    if (NH == 2 && tid == 0) {             // GEMM1 has completed (observed above): R0 is free for the second half of W2
        umma::mbar_expect_tx(&bars[2], sub_copy);
        umma::bulk_g2s(sub1, J.img + L.w2 + SUB_BYTES, sub_copy, &bars[2]);
    }

    // ---- epilogue 1: bias, ReLU, dropout; hidden → TMEM (A operand of GEMM2) and → staging tile (saved for backward) ------- //
    const int row = row0 + r_own;
    const bool valid = row < J.B;
    {
        const bool do_drop = p.drop_p > 0.f;
        const float keep_scale = do_drop ? 1.f / (1.f - p.drop_p) : 1.f;
        const unsigned long long drop_off = p.offset + (unsigned long long)j +
                                            (p.offset_dev ? (unsigned long long)__ldg(p.offset_dev) * MAX_JOBS : 0ull);
        const int col0 = cp * 32;
        float v[32];
        umma::tmem_ld32(tmem + lane_off + ACC + col0, v);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
            const int col = col0 + i4 * 4;
            uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
            if (do_drop && J.keep_mask == nullptr)
                rnd = rb_philox4x32(make_uint4((uint32_t)row, (uint32_t)(col >> 2), (uint32_t)drop_off, (uint32_t)(drop_off >> 32)),
                                    make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
            const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
            const float bv[4] = {bias1[i4].x, bias1[i4].y, bias1[i4].z, bias1[i4].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float x = fmaxf(v[i4 * 4 + e] + bv[e], 0.f);
                if (do_drop) {
                    bool keep;
                    if (J.keep_mask) keep = valid ? (J.keep_mask[(long long)row * HH + col + e] != 0) : true;
                    else keep = rb_u01(rw[e]) >= p.drop_p;
                    x = keep ? x * keep_scale : 0.f;
                }
                v[i4 * 4 + e] = x;
            }
            if (J.hid) *stage_at(stage, r_own, cp * 8 + i4) = make_float4(v[i4 * 4], v[i4 * 4 + 1], v[i4 * 4 + 2], v[i4 * 4 + 3]);
        }
        st16_split<MODE>(tmem + lane_off + A_HI + col0, tmem + lane_off + A_LO + col0, v);
        st16_split<MODE>(tmem + lane_off + A_HI + col0 + 16, tmem + lane_off + A_LO + col0 + 16, v + 16);
        umma::tmem_st_wait();
    }
    umma::fence_before_sync();
    s0_bar.wait();
    __syncthreads();
    if (!dead && umma::elect_issuer(tid)) {
        umma::fence_after_sync();
        issue_ts<MODE>(tmem + ACC, tmem + A_HI, tmem + A_LO, sub0, sub0 + (size_t)SUBR * HH * 4, SUBR, SUBR, HH);
        if (NH == 2) {
            if (umma::mbar_wait(&bars[2], 0u))
                issue_ts<MODE>(tmem + ACC + SUBR, tmem + A_HI, tmem + A_LO, sub1, sub1 + (size_t)SUBR * HH * 4, SUBR, SUBR, HH);
            else { dead = 1; if (p.err_flag) atomicOr(p.err_flag, 2); }
        }
        umma::commit(&bars[3]);
    }
    if (J.hid) stage_copy_out(stage, J.hid, row0, J.B, warp, lane);       // while GEMM2 runs
    constexpr int QC = D / 4;
    float4 bias2[QC / 4];
#pragma unroll
    for (int i4 = 0; i4 < QC / 4; ++i4) bias2[i4] = __ldg(reinterpret_cast<const float4*>(J.b2 + cp * QC) + i4);
    mma_bar.wait();
    umma::fence_after_sync();

    // ---- epilogue 2: bias, L2 normalise (four threads per row: partial sums combined in fixed order) ----------------- //
    {
        float y[QC];
        float ss = 0.f;
        TmemLd<QC>::ld(tmem + lane_off + ACC + cp * QC, y);
#pragma unroll
        for (int i4 = 0; i4 < QC / 4; ++i4) {
            const float4 b = bias2[i4];
            y[i4 * 4] += b.x; y[i4 * 4 + 1] += b.y; y[i4 * 4 + 2] += b.z; y[i4 * 4 + 3] += b.w;
        }
#pragma unroll
        for (int i = 0; i < QC; ++i) ss = fmaf(y[i], y[i], ss);
        ss_part[cp][r_own] = ss;
        __syncthreads();
        const float den = fmaxf(sqrtf((ss_part[0][r_own] + ss_part[1][r_own]) + (ss_part[2][r_own] + ss_part[3][r_own])), NORM_EPS);
        const float inv_den = 1.f / den;
        if (valid) {
#pragma unroll
            for (int c4 = 0; c4 < QC / 4; ++c4)
                *reinterpret_cast<float4*>(J.out + (long long)row * D + cp * QC + c4 * 4) =
                    make_float4(y[c4 * 4] * inv_den, y[c4 * 4 + 1] * inv_den, y[c4 * 4 + 2] * inv_den, y[c4 * 4 + 3] * inv_den);
            if (J.denom && cp == 0) J.denom[row] = den;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, TM_COLS);
}

// ------------------------------------------------------------------------------------------------------------ //
// backward, data part
// ------------------------------------------------------------------------------------------------------------ //
template <int D, int MODE>
__global__ void __launch_bounds__(NTD, 1) tower_bwd_data_ts_kernel(const BwdParams p, int* err_flag) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ float dot_part[4][ROWS];
    constexpr int NH = D / SUBR;
    constexpr uint32_t KA = 128, A_LO = A_HI + KA;               // A = dpre [128 × D], then dact [128 × H]
    RB_TS_PROLOGUE(err_flag)
    int begins[MAX_JOBS];
#pragma unroll
    for (int t = 0; t < MAX_JOBS; ++t) begins[t] = p.job[t].cta_begin;
    const int j = find_job(begins, p.n_jobs);
    const BwdJob J = p.job[j];
    const ImgLayout L = img_layout(D, HH, J.E);
    const int row0 = ((int)blockIdx.x - J.cta_begin) * ROWS;
    const int row = row0 + r_own;
    const bool valid = row < J.B;

    constexpr size_t w2t_bytes = (size_t)2 * HH * D * 4;
    unsigned char* w2t_img = smem;
    unsigned char* sub0 = smem + ts_r0<D>(w2t_bytes);
    unsigned char* sub1 = smem;
    unsigned char* stage = D == 128 ? smem + SUB_BYTES : sub0 + SUB_BYTES;
    constexpr uint32_t sub_copy = (uint32_t)((MODE == 2 ? 2 : 1) * SUBR * HH * 4);
    if (tid == 0) {
        constexpr uint32_t b1 = (uint32_t)((MODE == 2 ? 2 : 1) * HH * D * 4);
        umma::mbar_expect_tx(&bars[0], b1);
        umma::bulk_g2s(w2t_img, J.img + L.w2t, b1, &bars[0]);
        umma::mbar_expect_tx(&bars[1], sub_copy);
        umma::bulk_g2s(sub0, J.img + L.w1t, sub_copy, &bars[1]);
    }
    // ---- stage 1: normalise-backward per row → dpre (global, and the TMEM A operand) --------------------------------- //
    {
        constexpr int HC = D / 4;            // columns per thread
        float g[HC], yv[HC];
        float dot = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < HC / 4; ++c4) {
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), y4 = gv;
            if (valid) {
                gv = __ldg(reinterpret_cast<const float4*>(J.dY + (long long)row * D + cp * HC) + c4);
                y4 = __ldg(reinterpret_cast<const float4*>(J.y + (long long)row * D + cp * HC) + c4);
            }
            g[c4 * 4] = gv.x; g[c4 * 4 + 1] = gv.y; g[c4 * 4 + 2] = gv.z; g[c4 * 4 + 3] = gv.w;
            yv[c4 * 4] = y4.x; yv[c4 * 4 + 1] = y4.y; yv[c4 * 4 + 2] = y4.z; yv[c4 * 4 + 3] = y4.w;
            dot = fmaf(gv.x, y4.x, dot); dot = fmaf(gv.y, y4.y, dot); dot = fmaf(gv.z, y4.z, dot); dot = fmaf(gv.w, y4.w, dot);
        }
        dot_part[cp][r_own] = dot;
        __syncthreads();
        dot = (dot_part[0][r_own] + dot_part[1][r_own]) + (dot_part[2][r_own] + dot_part[3][r_own]);
        const float den = valid ? __ldg(J.denom + row) : 1.f;
        const bool clamped = den <= NORM_EPS;
        const float inv_den = 1.f / den;
#pragma unroll
        for (int i = 0; i < HC; ++i) g[i] = clamped ? g[i] * inv_den : (g[i] - yv[i] * dot) * inv_den;
        if (valid) {
#pragma unroll
            for (int c4 = 0; c4 < HC / 4; ++c4)
                *reinterpret_cast<float4*>(J.dpre + (long long)row * D + cp * HC + c4 * 4) =
                    make_float4(g[c4 * 4], g[c4 * 4 + 1], g[c4 * 4 + 2], g[c4 * 4 + 3]);
        }
#pragma unroll
        for (int c = 0; c < HC / 16; ++c)
            st16_split<MODE>(tmem + lane_off + A_HI + cp * HC + c * 16, tmem + lane_off + A_LO + cp * HC + c * 16, g + c * 16);
        umma::tmem_st_wait();
    }
    umma::fence_before_sync();
    big_bar.wait();
    __syncthreads();
    if (!dead && umma::elect_issuer(tid)) {
        umma::fence_after_sync();
        issue_ts<MODE>(tmem + ACC, tmem + A_HI, tmem + A_LO, w2t_img, w2t_img + (size_t)HH * D * 4, HH, HH, D);
        umma::commit(&bars[3]);
    }
    // while GEMM1 runs: this thread's 32 saved hidden activations → one 32-bit "unit was active and kept" mask
    uint32_t act_mask = 0u;
    if (valid) {
        float4 hv[8];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) hv[i4] = __ldg(reinterpret_cast<const float4*>(J.hid + (long long)row * HH + cp * 32) + i4);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4)
            act_mask |= ((hv[i4].x > 0.f ? 1u : 0u) | (hv[i4].y > 0.f ? 2u : 0u) | (hv[i4].z > 0.f ? 4u : 0u) | (hv[i4].w > 0.f ? 8u : 0u)) << (4 * i4);
    }
    mma_bar.wait();
    umma::fence_after_sync();
    if (NH == 2 && tid == 0) {
        umma::mbar_expect_tx(&bars[2], sub_copy);
        umma::bulk_g2s(sub1, J.img + L.w1t + SUB_BYTES, sub_copy, &bars[2]);
    }
    // ---- epilogue 1: ReLU / dropout mask → dact (staging tile → global, and the TMEM A operand of GEMM2) --------------- //
    {
        const int col0 = cp * 32;
        float v[32];
        umma::tmem_ld32(tmem + lane_off + ACC + col0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ((act_mask >> i) & 1u) ? v[i] * p.keep_scale : 0.f;
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4)
            *stage_at(stage, r_own, cp * 8 + i4) = make_float4(v[i4 * 4], v[i4 * 4 + 1], v[i4 * 4 + 2], v[i4 * 4 + 3]);
        st16_split<MODE>(tmem + lane_off + A_HI + col0, tmem + lane_off + A_LO + col0, v);
        st16_split<MODE>(tmem + lane_off + A_HI + col0 + 16, tmem + lane_off + A_LO + col0 + 16, v + 16);
        umma::tmem_st_wait();
    }
    umma::fence_before_sync();
    s0_bar.wait();
    __syncthreads();
    if (!dead && umma::elect_issuer(tid)) {
        umma::fence_after_sync();
        issue_ts<MODE>(tmem + ACC, tmem + A_HI, tmem + A_LO, sub0, sub0 + (size_t)SUBR * HH * 4, SUBR, SUBR, HH);
        if (NH == 2) {
            if (umma::mbar_wait(&bars[2], 0u))
                issue_ts<MODE>(tmem + ACC + SUBR, tmem + A_HI, tmem + A_LO, sub1, sub1 + (size_t)SUBR * HH * 4, SUBR, SUBR, HH);
            else { dead = 1; if (err_flag) atomicOr(err_flag, 2); }
        }
        umma::commit(&bars[3]);
    }
    stage_copy_out(stage, J.dact, row0, J.B, warp, lane);                 // while GEMM2 runs
    mma_bar.wait();
    umma::fence_after_sync();
    {
        constexpr int QC = D / 4;
        float v[QC];
        TmemLd<QC>::ld(tmem + lane_off + ACC + cp * QC, v);
        if (valid) {
#pragma unroll
            for (int i4 = 0; i4 < QC / 4; ++i4)
                *reinterpret_cast<float4*>(J.dRows + (long long)row * D + cp * QC + i4 * 4) =
                    make_float4(v[i4 * 4], v[i4 * 4 + 1], v[i4 * 4 + 2], v[i4 * 4 + 3]);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, TM_COLS);
}

// ------------------------------------------------------------------------------------------------------------ //
// backward, weight part: grid (nsplit, 2)
// ------------------------------------------------------------------------------------------------------------ //
template <int D, int NK, int MODE>      // NK = padded Din (multiple of 32) = N of the dW1 product
__global__ void __launch_bounds__(NTW, 1) tower_bwd_weights_ts_kernel(const BwdParams p, int* err_flag) {
    static_assert(NK % 32 == 0 && NK >= D && NK <= 160 && (D == 64 || D == 128), "operand widths");
    extern __shared__ __align__(1024) unsigned char smem[];       // the B operand chunk [NB × KC], hi then lo
    __shared__ long long ids_s[2][KC];
    __shared__ float bias_part[KC / 4][HH];
    __shared__ __align__(8) uint64_t mma_bar_s;
    __shared__ uint32_t tmem_slot;
    __shared__ int dead;
    constexpr uint32_t WT_COLS = 256, WA_HI = 160, WA_LO = 192;   // accumulator [0, ≤160) | A hi (32 samples) | A lo
    constexpr int NCOL_MAX = NK / 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, WT_COLS);
    if (tid == 0) { umma::mbar_init(&mma_bar_s, 1); umma::fence_mbar_init(); dead = 0; }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    Bar mma_bar{&mma_bar_s, 0u, &dead, err_flag};
    const int r_own = ((warp & 3) << 5) + lane, half = warp >> 2;   // TMEM lane = hidden unit h; column half
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int role = blockIdx.y;             // 0: dW2ᵀ[h][d] = Σ hid[r][h]·dpre[r][d] (+ db2);  1: dW1[h][k] = Σ dact[r][h]·X[r][k] (+ db1)
    const int s = blockIdx.x;
    const int E = p.job[0].E, Din = D + E;
    const int NB = role ? NK : D;
    unsigned char* b_hi = smem;
    unsigned char* b_lo = smem + (size_t)NB * KC * 4;
    const int sq = warp, mq = lane;          // B staging: this thread's 4 samples (4·sq…) × 4 operand rows (4·mq…)
    const int sq2 = tid >> 3, mq2 = 32 + (tid & 7);               // NK = 160: the operand rows 128..159 (threads 0..63)
    const bool extra_b = role == 1 && NK > 128 && tid < 64;

    int jb[2] = {0, 0}, je[2] = {0, 0}, nc[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (j < p.n_jobs) {
            const int B = p.job[j].B;
            const int chunk = (((B + p.nsplit - 1) / p.nsplit) + KC - 1) / KC * KC;   // rows per CTA, multiple of KC
            jb[j] = min(B, s * chunk); je[j] = min(B, jb[j] + chunk);
            nc[j] = (je[j] - jb[j] + KC - 1) / KC;
        }
    }
    const int n_chunks = nc[0] + nc[1];
    auto chunk_at = [&](int t, int& j, int& r0, int& nr) {
        j = t < nc[0] ? 0 : 1;
        r0 = jb[j] + (t - (j ? nc[0] : 0)) * KC;
        nr = min(KC, je[j] - r0);
    };
    auto load_ids = [&](int t) {
        if (role == 1 && tid < KC && t < n_chunks) {
            int j, r0, nr;
            chunk_at(t, j, r0, nr);
            long long id = tid < nr ? p.job[j].ids[r0 + tid] : 0;
            if ((unsigned long long)id >= (unsigned long long)p.job[j].n_rows) id = 0;
            ids_s[t & 1][tid] = id;
        }
    };
    float av[16];                            // A: this lane's hidden unit over 16 samples of the chunk (half·16 …)
    float4 bv[4], bv2[4];                    // B: 4 samples × 4 operand rows (and the extra rows at NK = 160)
    auto x_quad = [&](const BwdJob& J, int t, int r, long long gr, int q4) -> float4 {      // X[sample][4·q4 … +3]
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q4 < D / 4) {
            x = __ldg(reinterpret_cast<const float4*>(J.table + ids_s[t & 1][r] * D) + q4);
        } else if (E > 0) {
            const int k = q4 * 4 - D;          // genre column
            const float* ex = J.extra + (J.extra_by_id ? ids_s[t & 1][r] : gr) * E;
            x.x = k + 0 < E ? __ldg(ex + k + 0) : 0.f;
            x.y = k + 1 < E ? __ldg(ex + k + 1) : 0.f;
            x.z = k + 2 < E ? __ldg(ex + k + 2) : 0.f;
            x.w = k + 3 < E ? __ldg(ex + k + 3) : 0.f;
        }
        return x;
    };
    auto load_chunk = [&](int t) {
        int j, r0, nr;
        chunk_at(t, j, r0, nr);
        const BwdJob& J = p.job[j];
        const float* asrc = role ? J.dact : J.hid;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int r = half * 16 + i;
            av[i] = r < nr ? __ldg(asrc + (long long)(r0 + r) * HH + r_own) : 0.f;
        }
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = sq * 4 + i;
            const bool ok = r < nr;
            const long long gr = r0 + r;
            if (role == 0) bv[i] = (ok && mq < D / 4) ? __ldg(reinterpret_cast<const float4*>(J.dpre + gr * D) + mq) : z;
            else bv[i] = (ok && mq < NK / 4) ? x_quad(J, t, r, gr, mq) : z;
        }
        if (NK > 128) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = sq2 * 4 + i;
                bv2[i] = (extra_b && r < nr) ? x_quad(J, t, r, (long long)r0 + r, mq2) : z;
            }
        }
    };
    float acc[NCOL_MAX];
#pragma unroll
    for (int i = 0; i < NCOL_MAX; ++i) acc[i] = 0.f;
    float db_a = 0.f;                                            // role 1: Σ_samples dact[.][h] over this thread's samples
    float4 db_b = make_float4(0.f, 0.f, 0.f, 0.f);              // role 0: Σ_samples dpre[.][4·mq…] over this thread's samples
    bool first = true;

    load_ids(0);
    load_ids(1);
    __syncthreads();
    if (n_chunks > 0) load_chunk(0);
    for (int t = 0; t < n_chunks; ++t) {
#pragma unroll
        for (int i = 0; i < 16; ++i) db_a += av[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) { db_b.x += bv[i].x; db_b.y += bv[i].y; db_b.z += bv[i].z; db_b.w += bv[i].w; }
        st16_split<MODE>(tmem + lane_off + WA_HI + half * 16, tmem + lane_off + WA_LO + half * 16, av);
        if (mq < NB / 4) put_block_t<MODE>(b_hi, b_lo, NB, mq, sq, lane, bv);
        if (NK > 128 && extra_b) put_block_t<MODE>(b_hi, b_lo, NB, mq2, sq2, lane, bv2);
        umma::tmem_st_wait();
        umma::fence_proxy_async();
        umma::fence_before_sync();
        __syncthreads();                         // operands staged; ids of chunk t+1 (stored one iteration ago) visible
        if (!dead && umma::elect_issuer(tid)) {
            umma::fence_after_sync();
            issue_ts<MODE>(tmem, tmem + WA_HI, tmem + WA_LO, b_hi, b_lo, NB, NB, KC);
            umma::commit(&mma_bar_s);
        }
        first = false;
        if (t + 1 < n_chunks) load_chunk(t + 1);  // in flight while the tensor core works on chunk t
        mma_bar.wait();
        umma::fence_after_sync();
        if (!dead) {
            if (role == 0) {
#pragma unroll
                for (int c = 0; c < D / 32; ++c) {
                    float v[16];
                    umma::tmem_ld16(tmem + lane_off + half * (D / 2) + c * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[c * 16 + i] += v[i];
                }
            } else {
#pragma unroll
                for (int c = 0; c < NK / 32; ++c) {
                    float v[16];
                    umma::tmem_ld16(tmem + lane_off + half * (NK / 2) + c * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[c * 16 + i] += v[i];
                }
            }
        }
        load_ids(t + 2);
        umma::fence_before_sync();
        __syncthreads();
    }
    // partial block layout: [W1ᵀ (Din*H, transposed) | b1 (H) | W2 (D*H) | b2 (D)]; this CTA writes its role's two blocks
    float* part = p.part + (long long)s * p.P;
    float* w1o = part, *b1o = part + HH * Din, *w2o = part + HH * Din + HH, *b2o = part + HH * Din + HH + D * HH;
    if (role == 0) {
        if (mq < D / 4) *reinterpret_cast<float4*>(&bias_part[sq][mq * 4]) = db_b;
    } else {
        bias_part[half][r_own] = db_a;
    }
    __syncthreads();
    if (first) {          // this CTA had no rows: its blocks are all zeros
        if (role == 0) for (int i = tid; i < D * HH + D; i += NTW) w2o[i] = 0.f;
        else for (int i = tid; i < HH * Din + HH; i += NTW) w1o[i] = 0.f;
    } else if (role == 0) {
        const int h = r_own;
#pragma unroll
        for (int i = 0; i < D / 2; ++i) w2o[(long long)(half * (D / 2) + i) * HH + h] = acc[i];          // dW2[d][h] = acc[h][d]
        if (tid < D) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < KC / 4; ++q) t += bias_part[q][tid];      // fixed order
            b2o[tid] = t;
        }
    } else {
        const int h = r_own;
#pragma unroll
        for (int i = 0; i < NK / 2; ++i) {
            const int k = half * (NK / 2) + i;
            if (k < Din) w1o[(long long)k * HH + h] = acc[i];          // dW1[h][k], stored transposed: 128-byte warp stores
        }
        if (tid < HH) b1o[tid] = bias_part[0][tid] + bias_part[1][tid];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, WT_COLS);
}

template <int D, int MODE>
int launch_fwd_ts(FwdParams& p, int grid, size_t smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        int rc = set_smem(tower_fwd_ts_kernel<D, MODE>, ts_smem<D>((size_t)2 * HH * (D + 24) * 4));
        if (rc) return rc;
        attr = true;
    }
    tower_fwd_ts_kernel<D, MODE><<<grid, NTD, smem, st>>>(p);
    RB_LAUNCH_CHECK("tower_fwd_ts_kernel");
    return RB200_OK;
}

template <int D, int MODE>
int launch_bwd_data_ts(BwdParams& p, int grid, cudaStream_t st) {
    static bool attr = false;
    constexpr size_t smem = ts_smem<D>((size_t)2 * HH * D * 4);
    if (!attr) {
        int rc = set_smem(tower_bwd_data_ts_kernel<D, MODE>, smem);
        if (rc) return rc;
        attr = true;
    }
    tower_bwd_data_ts_kernel<D, MODE><<<grid, NTD, smem, st>>>(p, nullptr);
    RB_LAUNCH_CHECK("tower_bwd_data_ts_kernel");
    return RB200_OK;
}

template <int D, int NK, int MODE>
int launch_bwd_weights_ts(BwdParams& p, cudaStream_t st) {
    static bool attr = false;
    constexpr size_t smem = (size_t)2 * NK * KC * 4;
    if (!attr) {
        int rc = set_smem(tower_bwd_weights_ts_kernel<D, NK, MODE>, smem);
        if (rc) return rc;
        attr = true;
    }
    tower_bwd_weights_ts_kernel<D, NK, MODE><<<dim3(p.nsplit, 2), NTW, smem, st>>>(p, nullptr);
    RB_LAUNCH_CHECK("tower_bwd_weights_ts_kernel");
    return RB200_OK;
}

}  // namespace

bool rb_tower_ts_supported(int D, int H, int E) { return (D == 64 || D == 128) && H == 128 && E >= 0 && E <= 24; }

// jobs carry their weight image (rb_tower_fwd_tc builds missing ones)
int rb_tower_fwd_ts(FwdParams& p, int D, int mode, cudaStream_t st) {
    const int grid = assign_tiles(p.job, p.n_jobs);
    int kp = 0;
    for (int j = 0; j < p.n_jobs; ++j) { const int k = (D + p.job[j].E + 7) & ~7; if (k > kp) kp = k; }
    const size_t big = (size_t)2 * HH * kp * 4;
    const size_t smem = D == 128 ? ts_smem<128>(big) : ts_smem<64>(big);
    RB_REQUIRE(kp <= D + 24, "tower_fwd (tcgen05, TMEM operand): extra_dim <= 24");
    if (D == 128) return mode == 1 ? launch_fwd_ts<128, 1>(p, grid, smem, st) : launch_fwd_ts<128, 2>(p, grid, smem, st);
    return mode == 1 ? launch_fwd_ts<64, 1>(p, grid, smem, st) : launch_fwd_ts<64, 2>(p, grid, smem, st);
}

// data-gradient kernel + weight-gradient kernel (split-K partials left in p.part)
int rb_tower_bwd_ts(BwdParams& p, int D, int mode, cudaStream_t st) {
    const int E = p.job[0].E;
    const int grid = assign_tiles(p.job, p.n_jobs);
    int rc;
    if (D == 128) rc = mode == 1 ? launch_bwd_data_ts<128, 1>(p, grid, st) : launch_bwd_data_ts<128, 2>(p, grid, st);
    else rc = mode == 1 ? launch_bwd_data_ts<64, 1>(p, grid, st) : launch_bwd_data_ts<64, 2>(p, grid, st);
    if (rc) return rc;
    if (D == 128) {
        if (E == 0) return mode == 1 ? launch_bwd_weights_ts<128, 128, 1>(p, st) : launch_bwd_weights_ts<128, 128, 2>(p, st);
        return mode == 1 ? launch_bwd_weights_ts<128, 160, 1>(p, st) : launch_bwd_weights_ts<128, 160, 2>(p, st);
    }
    if (E == 0) return mode == 1 ? launch_bwd_weights_ts<64, 64, 1>(p, st) : launch_bwd_weights_ts<64, 64, 2>(p, st);
    return mode == 1 ? launch_bwd_weights_ts<64, 96, 1>(p, st) : launch_bwd_weights_ts<64, 96, 2>(p, st);
}
