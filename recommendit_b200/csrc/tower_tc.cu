// Two-tower MLP on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM) — modes 1 (TF32) and
// 2 (3xTF32, fp32-grade) of rb200_tower_fwd / rb200_tower_bwd.  Same math as tower.cu (reference
// src/models/two_tower.py:39-42, 68-72 and their autograd), different engine:
//
//   forward   one CTA (128 threads) per tile of 128 samples.  Stage 1: gathered rows [128 × Kp] and W1 [H × Kp] are
//             split hi/lo and stored in shared memory in the canonical K-major UMMA layout; GEMM1 → TMEM.  Epilogue 1
//             (thread = sample = TMEM lane): bias, ReLU, dropout, hidden saved for backward and re-staged (hi/lo) as the A
//             operand of GEMM2 together with W2; GEMM2 → TMEM.  Epilogue 2: bias, row L2 norm (per-thread, fixed order),
//             normalised output.  Activations never leave the SM between the two GEMMs.
//   bwd data  normalise-backward per row → dpre (A) × W2ᵀ (B) → TMEM → ReLU/dropout mask → dact (A) × W1[:, :D]ᵀ (B) → dRows.
//   bwd wts   split-K over the batch: per CTA, chunks of 32 samples; operands are transposed while being staged
//             (K = sample index): dW2ᵀ[h][d] = Σ_r hid[r][h]·dpre[r][d], dW1[h][k] = Σ_r dact[r][h]·X[r][k]; accumulators
//             stay in TMEM across chunks; per-CTA partials reduced in fixed order (deterministic).
#include <stdlib.h>

#include "tower_tc_common.cuh"

using namespace towertc;

namespace {

constexpr int NT = 256;        // weight-gradient kernel: 8 warps; warps w and w+4 share TMEM lane quadrant w%4 and split the columns
constexpr int NTD = 512;       // forward / backward-data kernels: 16 warps, four column quarters per TMEM lane quadrant
constexpr int TMEM_COLS = 256;

struct PrepSet { const float* W1; const float* W2; unsigned char* img; int E; };
struct PrepParams { PrepSet set[2]; int n_sets, D, H; rb200_opt_state* opt; };   // opt != NULL: also begin the optimizer step

__device__ __forceinline__ void put4_both(unsigned char* hi_base, size_t lo_delta, int R, int r, int k, const float4& v) {
    const uint32_t off = umma::kmajor_offset(R, r, k);
    float4 hi, lo;
    umma::split4(v, hi, lo);
    *reinterpret_cast<float4*>(hi_base + off) = hi;
    *reinterpret_cast<float4*>(hi_base + lo_delta + off) = lo;
}

__global__ void __launch_bounds__(256) tower_prep_kernel(const PrepParams p) {
    if (p.opt && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) rb_opt_begin_step_dev(p.opt);
    const PrepSet S = p.set[blockIdx.y];
    const int D = p.D, H = p.H, Din = D + S.E;
    const ImgLayout L = img_layout(D, H, S.E);
    const int Kp = L.Kp;
    const int n1 = H * (Kp / 4), n2 = D * (H / 4), n3 = H * (D / 4), n4 = D * (H / 4);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n1 + n2 + n3 + n4; idx += gridDim.x * blockDim.x) {
        if (idx < n1) {                       // W1 [H × Kp]: consecutive threads → consecutive rows (conflict-free image writes)
            const int c4 = idx / H, h = idx - c4 * H, k = c4 * 4;
            float4 v;
            v.x = k + 0 < Din ? __ldg(S.W1 + (long long)h * Din + k + 0) : 0.f;
            v.y = k + 1 < Din ? __ldg(S.W1 + (long long)h * Din + k + 1) : 0.f;
            v.z = k + 2 < Din ? __ldg(S.W1 + (long long)h * Din + k + 2) : 0.f;
            v.w = k + 3 < Din ? __ldg(S.W1 + (long long)h * Din + k + 3) : 0.f;
            put4_both(S.img + L.w1, (size_t)H * Kp * 4, H, h, k, v);
        } else if (idx < n1 + n2) {           // W2 [D × H], in sub-images of 64 rows
            const int i = idx - n1, c4 = i / D, d = i - c4 * D;
            put4_both(S.img + L.w2 + (size_t)(d / SUBR) * 2 * SUBR * H * 4, (size_t)SUBR * H * 4, SUBR, d % SUBR, c4 * 4,
                      __ldg(reinterpret_cast<const float4*>(S.W2 + (long long)d * H) + c4));
        } else if (idx < n1 + n2 + n3) {      // W2ᵀ [H × D]: (h, d) = W2[d][h]
            const int i = idx - n1 - n2, c4 = i / H, h = i - c4 * H;
            const float* wp = S.W2 + (long long)(c4 * 4) * H + h;
            put4_both(S.img + L.w2t, (size_t)H * D * 4, H, h, c4 * 4, make_float4(__ldg(wp), __ldg(wp + H), __ldg(wp + 2 * H), __ldg(wp + 3 * H)));
        } else {                              // W1[:, :D]ᵀ [D × H]: (d, h) = W1[h][d]
            const int i = idx - n1 - n2 - n3, c4 = i / D, d = i - c4 * D;
            const float* wp = S.W1 + (long long)(c4 * 4) * Din + d;
            put4_both(S.img + L.w1t + (size_t)(d / SUBR) * 2 * SUBR * H * 4, (size_t)SUBR * H * 4, SUBR, d % SUBR, c4 * 4,
                      make_float4(__ldg(wp), __ldg(wp + Din), __ldg(wp + 2 * Din), __ldg(wp + 3 * Din)));
        }
    }
}

// issue the MMAs of one GEMM: D[tmem] = A[M=128 × K] · B[N × K]ᵀ, operands fully resident (K-major, rows RA / RB)
template <int MODE>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, unsigned char* a_hi, unsigned char* a_lo, int RA, unsigned char* b_hi,
                                           unsigned char* b_lo, int RB, int N, int K, bool accumulate_first) {
    const uint32_t idesc = umma::idesc_tf32(128, N);
    const uint32_t lbo_a = (RA / 8) * 128, lbo_b = (RB / 8) * 128;
    // descriptors are built once; a k-step (8 floats = two 16-byte K chunks) only advances the 14-bit start-address field
    const uint64_t dah = umma::smem_desc(umma::smem_u32(a_hi), lbo_a, 128), dal = umma::smem_desc(umma::smem_u32(a_lo), lbo_a, 128);
    const uint64_t dbh = umma::smem_desc(umma::smem_u32(b_hi), lbo_b, 128), dbl = umma::smem_desc(umma::smem_u32(b_lo), lbo_b, 128);
    const uint64_t sa = (uint64_t)((2 * lbo_a) >> 4), sb = (uint64_t)((2 * lbo_b) >> 4);
    uint64_t oa = 0, ob = 0;
    for (int j = 0; j < K / 8; ++j, oa += sa, ob += sb) {
        const bool acc = accumulate_first || j > 0;
        if (MODE == 2) {
            umma::mma_tf32(tmem_d, dal + oa, dbh + ob, idesc, acc);
            umma::mma_tf32(tmem_d, dah + oa, dbl + ob, idesc, true);
            umma::mma_tf32(tmem_d, dah + oa, dbh + ob, idesc, true);
        } else {
            umma::mma_tf32(tmem_d, dah + oa, dbh + ob, idesc, acc);
        }
    }
}

#define RB_TC_PROLOGUE(err_ptr)                                                                                   \
    __shared__ __align__(8) uint64_t mma_bar_s, w_bar_s;                                                          \
    __shared__ uint32_t tmem_slot;                                                                                \
    __shared__ int dead;                                                                                          \
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                                                \
    if (warp == 0) umma::tmem_alloc(&tmem_slot, TMEM_COLS);                                                       \
    if (tid == 0) { umma::mbar_init(&mma_bar_s, 1); umma::mbar_init(&w_bar_s, 1); umma::fence_mbar_init(); dead = 0; } \
    umma::fence_before_sync();                                                                                    \
    __syncthreads();                                                                                              \
    umma::fence_after_sync();                                                                                     \
    const uint32_t tmem = tmem_slot;                                                                              \
    Bar mma_bar{&mma_bar_s, 0u, &dead, err_ptr}, w_bar{&w_bar_s, 0u, &dead, err_ptr};                              \
    const int r_own = ((warp & 3) << 5) + lane, half = warp >> 2;     /* TMEM lane (= tile row) and column part  */ \
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;

#define RB_TC_EPILOGUE()                                                               \
    umma::fence_before_sync();                                                         \
    __syncthreads();                                                                   \
    if (warp == 0) umma::tmem_free(tmem, TMEM_COLS);

// ------------------------------------------------------------------------------------------------------------ //
// forward
// ------------------------------------------------------------------------------------------------------------ //
template <int D, int H, int MODE>
__global__ void __launch_bounds__(NTD, 1) tower_fwd_tc_kernel(const FwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ float ss_part[4][ROWS];
    __shared__ long long ids_s[ROWS];
    int begins[MAX_JOBS];
#pragma unroll
    for (int t = 0; t < MAX_JOBS; ++t) begins[t] = p.job[t].cta_begin;
    const int j = find_job(begins, p.n_jobs);
    const FwdJob J = p.job[j];
    const int E = J.E, Din = D + E;
    const ImgLayout L = img_layout(D, H, E);
    const int Kp = L.Kp;
    const int row0 = ((int)blockIdx.x - J.cta_begin) * ROWS;
    // the ids are fetched first: their latency hides under the TMEM allocation / barrier setup of the prologue, whose
    // __syncthreads also publishes ids_s
    if (threadIdx.x < ROWS) {
        const int row = row0 + (int)threadIdx.x;
        long long id = row < J.B ? J.ids[row] : 0;
        if ((unsigned long long)id >= (unsigned long long)J.n_rows) { if (p.err_flag) atomicOr(p.err_flag, 1); id = 0; }
        ids_s[threadIdx.x] = id;
    }
    RB_TC_PROLOGUE(p.err_flag)

    // ---- stage 1: X [128 × Kp] gathered by the threads, W1 image pulled in by one bulk asynchronous copy --------- //
    unsigned char* x_hi = smem;
    unsigned char* x_lo = x_hi + (size_t)ROWS * Kp * 4;
    unsigned char* w1_hi = x_lo + (size_t)ROWS * Kp * 4;
    unsigned char* w1_lo = w1_hi + (size_t)H * Kp * 4;
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)((MODE == 2 ? 2 : 1) * H * Kp * 4);
        umma::mbar_expect_tx(&w_bar_s, bytes);
        umma::bulk_g2s(w1_hi, J.img + L.w1, bytes, &w_bar_s);
    }
    {
        // lane → (row%8 = lane&7, 16-byte chunk = 4·cq + lane>>3): 64-byte global segments, conflict-free 16-byte stores;
        // several units' loads are issued before their stores so that the gather latencies overlap
        const int K4 = Kp / 4, CQ = (K4 + 3) / 4, n_units = (ROWS / 8) * CQ;
        constexpr int UF = 6;       // units in flight per warp: the whole tile's gather is one round of loads at D+E <= 96
        for (int u0 = warp; u0 < n_units; u0 += UF * (NTD / 32)) {
            float4 v[UF];
            int rr[UF], kk[UF];
#pragma unroll
            for (int q = 0; q < UF; ++q) {
                const int u = u0 + q * (NTD / 32);
                const int rg = u / CQ, cq = u - rg * CQ;
                const int r = rg * 8 + (lane & 7), c4 = cq * 4 + (lane >> 3);
                rr[q] = r; kk[q] = (u < n_units && c4 < K4) ? c4 * 4 : -1;
                v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kk[q] >= 0 && row0 + r < J.B) {
                    const int k = kk[q];
                    const long long id = ids_s[r];
                    if (k < D) {
                        v[q] = __ldg(reinterpret_cast<const float4*>(J.table + id * D + k));
                    } else {
                        const float* ex = J.extra + (J.extra_by_id ? id : (long long)(row0 + r)) * E + (k - D);
                        v[q].x = k + 0 < Din ? __ldg(ex + 0) : 0.f;
                        v[q].y = k + 1 < Din ? __ldg(ex + 1) : 0.f;
                        v[q].z = k + 2 < Din ? __ldg(ex + 2) : 0.f;
                        v[q].w = k + 3 < Din ? __ldg(ex + 3) : 0.f;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < UF; ++q)
                if (kk[q] >= 0) put4<MODE>(x_hi, x_lo, ROWS, rr[q], kk[q], v[q]);
        }
    }
    umma::fence_proxy_async();
    w_bar.wait();
    __syncthreads();
    if (!dead && umma::elect_issuer(tid)) {
        umma::fence_after_sync();
        issue_gemm<MODE>(tmem, x_hi, x_lo, ROWS, w1_hi, w1_lo, H, H, Kp, false);
        umma::commit(&mma_bar_s);
    }
    float4 bias1[8];                       // this thread's 32 hidden-unit biases, fetched while GEMM1 runs
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) bias1[i4] = __ldg(reinterpret_cast<const float4*>(J.b1 + half * 32) + i4);
    mma_bar.wait();
    umma::fence_after_sync();

    // ---- epilogue 1 + stage 2: hidden [128 × H] staged by the threads, W2 image by bulk copy --------------------- //
    unsigned char* h_hi = smem;
    unsigned char* h_lo = h_hi + (size_t)ROWS * H * 4;
    unsigned char* w2_hi = h_lo + (size_t)ROWS * H * 4;
    unsigned char* w2_lo = w2_hi + (size_t)D * H * 4;
    if (tid == 0) {      // GEMM1 has completed (observed above): the stage-1 buffers this copy overwrites are free
        const uint32_t bytes = (uint32_t)((MODE == 2 ? 2 : 1) * D * H * 4);
        umma::mbar_expect_tx(&w_bar_s, bytes);
        umma::bulk_g2s(w2_hi, J.img + L.w2, bytes, &w_bar_s);
    }
    const int row = row0 + r_own;
    const bool valid = row < J.B;
    {
        const bool do_drop = p.drop_p > 0.f;
        const float keep_scale = do_drop ? 1.f / (1.f - p.drop_p) : 1.f;
        const unsigned long long drop_off = p.offset + (unsigned long long)j +
                                            (p.offset_dev ? (unsigned long long)__ldg(p.offset_dev) * MAX_JOBS : 0ull);
        static_assert(H == 128, "one 32-column TMEM load per thread");
        {
            const int col0 = half * 32;          // `half` is the column quarter here (16 warps)
            float v[32];
            umma::tmem_ld32(tmem + lane_off + col0, v);
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const int col = col0 + i4 * 4;
                float4 o;
                float* op = &o.x;
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
                        if (J.keep_mask) keep = valid ? (J.keep_mask[(long long)row * H + col + e] != 0) : true;
                        else keep = rb_u01(rw[e]) >= p.drop_p;
                        x = keep ? x * keep_scale : 0.f;
                    }
                    op[e] = x;
                }
                // (3xTF32: the saved hidden tile leaves through shared memory below, with 64-byte row segments — one 16-byte
                //  store per thread-row here costs 32 LSU wavefronts per instruction and dominated this epilogue)
                if (MODE != 2 && J.hid && valid) *reinterpret_cast<float4*>(J.hid + (long long)row * H + col) = o;
                put4<MODE>(h_hi, h_lo, ROWS, r_own, col, o);
            }
        }
    }
    umma::fence_before_sync();
    umma::fence_proxy_async();
    w_bar.wait();
    __syncthreads();
    if (!dead && umma::elect_issuer(tid)) {
        umma::fence_after_sync();
        issue_gemm<MODE>(tmem + H, h_hi, h_lo, ROWS, w2_hi, w2_lo, D, D, H, false);
        umma::commit(&mma_bar_s);
    }
    if (MODE == 2 && J.hid) {
        // hidden activations for the backward pass, copied out of the operand tile (hi + lo = the fp32 value, exactly) while
        // GEMM2 runs: lane → (row%8, 16-byte chunk): conflict-free shared-memory phases, 64-byte global segments
        constexpr int CQ = H / 16, UNITS = (ROWS / 8) * CQ;
        for (int u = warp; u < UNITS; u += NTD / 32) {
            const int rg = u / CQ, cq = u - rg * CQ;
            const int r = rg * 8 + (lane & 7), c4 = cq * 4 + (lane >> 3);
            if (row0 + r < J.B) {
                const uint32_t off = umma::kmajor_offset(ROWS, r, c4 * 4);
                const float4 a = *reinterpret_cast<const float4*>(h_hi + off), b = *reinterpret_cast<const float4*>(h_lo + off);
                *reinterpret_cast<float4*>(J.hid + (long long)(row0 + r) * H + c4 * 4) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
            }
        }
    }
    float4 bias2[D / 16];
#pragma unroll
    for (int i4 = 0; i4 < D / 16; ++i4) bias2[i4] = __ldg(reinterpret_cast<const float4*>(J.b2 + half * (D / 4)) + i4);
    mma_bar.wait();
    umma::fence_after_sync();

    // ---- epilogue 2: bias, L2 normalise (four threads per row: partial sums combined in fixed order) ----------------- //
    {
        static_assert(D == 64, "16 output columns per thread");
        constexpr int QC = D / 4;
        float y[QC];
        float ss = 0.f;
        umma::tmem_ld16(tmem + lane_off + H + half * QC, y);
#pragma unroll
        for (int i4 = 0; i4 < QC / 4; ++i4) {
            const float4 b = bias2[i4];
            y[i4 * 4] += b.x; y[i4 * 4 + 1] += b.y; y[i4 * 4 + 2] += b.z; y[i4 * 4 + 3] += b.w;
        }
#pragma unroll
        for (int i = 0; i < QC; ++i) ss = fmaf(y[i], y[i], ss);
        ss_part[half][r_own] = ss;
        __syncthreads();
        const float den = fmaxf(sqrtf((ss_part[0][r_own] + ss_part[1][r_own]) + (ss_part[2][r_own] + ss_part[3][r_own])), NORM_EPS);
        const float inv_den = 1.f / den;         // (x · (1/den) is within 1 ulp of x / den; the forward bound is 2e-6)
        if (valid) {
#pragma unroll
            for (int c4 = 0; c4 < QC / 4; ++c4)
                *reinterpret_cast<float4*>(J.out + (long long)row * D + half * QC + c4 * 4) =
                    make_float4(y[c4 * 4] * inv_den, y[c4 * 4 + 1] * inv_den, y[c4 * 4 + 2] * inv_den, y[c4 * 4 + 3] * inv_den);
            if (J.denom && half == 0) J.denom[row] = den;
        }
    }
    RB_TC_EPILOGUE()
}

// ------------------------------------------------------------------------------------------------------------ //
// backward, data part
// ------------------------------------------------------------------------------------------------------------ //
template <int D, int H, int MODE>
__global__ void __launch_bounds__(NTD, 1) tower_bwd_data_tc_kernel(const BwdParams p, int* err_flag) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ float dot_part[4][ROWS];
    RB_TC_PROLOGUE(err_flag)
    int begins[MAX_JOBS];
#pragma unroll
    for (int t = 0; t < MAX_JOBS; ++t) begins[t] = p.job[t].cta_begin;
    const int j = find_job(begins, p.n_jobs);
    const BwdJob J = p.job[j];
    const ImgLayout L = img_layout(D, H, J.E);
    const int row = ((int)blockIdx.x - J.cta_begin) * ROWS + r_own;
    const bool valid = row < J.B;

    // ---- stage 1: dpre [128 × D] (A), W2ᵀ image [H × D] (B) -------------------------------------------------------- //
    unsigned char* g_hi = smem;
    unsigned char* g_lo = g_hi + (size_t)ROWS * D * 4;
    unsigned char* w2t_hi = g_lo + (size_t)ROWS * D * 4;
    unsigned char* w2t_lo = w2t_hi + (size_t)H * D * 4;
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)((MODE == 2 ? 2 : 1) * H * D * 4);
        umma::mbar_expect_tx(&w_bar_s, bytes);
        umma::bulk_g2s(w2t_hi, J.img + L.w2t, bytes, &w_bar_s);
    }
    {
        constexpr int HC = D / 4;            // columns per thread (`half` is the column quarter: 16 warps)
        float g[HC], yv[HC];
        float dot = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < HC / 4; ++c4) {
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), y4 = gv;
            if (valid) {
                gv = __ldg(reinterpret_cast<const float4*>(J.dY + (long long)row * D + half * HC) + c4);
                y4 = __ldg(reinterpret_cast<const float4*>(J.y + (long long)row * D + half * HC) + c4);
            }
            g[c4 * 4] = gv.x; g[c4 * 4 + 1] = gv.y; g[c4 * 4 + 2] = gv.z; g[c4 * 4 + 3] = gv.w;
            yv[c4 * 4] = y4.x; yv[c4 * 4 + 1] = y4.y; yv[c4 * 4 + 2] = y4.z; yv[c4 * 4 + 3] = y4.w;
            dot = fmaf(gv.x, y4.x, dot); dot = fmaf(gv.y, y4.y, dot); dot = fmaf(gv.z, y4.z, dot); dot = fmaf(gv.w, y4.w, dot);
        }
        dot_part[half][r_own] = dot;
        __syncthreads();
        dot = (dot_part[0][r_own] + dot_part[1][r_own]) + (dot_part[2][r_own] + dot_part[3][r_own]);
        const float den = valid ? __ldg(J.denom + row) : 1.f;
        const bool clamped = den <= NORM_EPS;
        const float inv_den = 1.f / den;         // one IEEE reciprocal per row-quarter instead of 16 divisions (≤ 1 ulp apart)
#pragma unroll
        for (int c4 = 0; c4 < HC / 4; ++c4) {
            float4 o;
            float* op = &o.x;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float gg = g[c4 * 4 + e];
                op[e] = clamped ? gg * inv_den : (gg - yv[c4 * 4 + e] * dot) * inv_den;
            }
            if (valid) *reinterpret_cast<float4*>(J.dpre + (long long)row * D + half * HC + c4 * 4) = o;
            put4<MODE>(g_hi, g_lo, ROWS, r_own, half * HC + c4 * 4, o);
        }
    }
    umma::fence_proxy_async();
    w_bar.wait();
    __syncthreads();
    if (!dead && umma::elect_issuer(tid)) {
        umma::fence_after_sync();
        issue_gemm<MODE>(tmem, g_hi, g_lo, ROWS, w2t_hi, w2t_lo, H, H, D, false);
        umma::commit(&mma_bar_s);
    }
    // while GEMM1 runs: this thread's 32 saved hidden activations → one 32-bit "unit was active and kept" mask
    uint32_t act_mask = 0u;
    if (valid) {
        float4 hv[8];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) hv[i4] = __ldg(reinterpret_cast<const float4*>(J.hid + (long long)row * H + half * 32) + i4);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4)
            act_mask |= ((hv[i4].x > 0.f ? 1u : 0u) | (hv[i4].y > 0.f ? 2u : 0u) | (hv[i4].z > 0.f ? 4u : 0u) | (hv[i4].w > 0.f ? 8u : 0u)) << (4 * i4);
    }
    mma_bar.wait();
    umma::fence_after_sync();

    // ---- epilogue 1 + stage 2: dact [128 × H] (A), W1[:, :D]ᵀ image [D × H] (B) -------------------------------------- //
    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + (size_t)ROWS * H * 4;
    unsigned char* w1t_hi = a_lo + (size_t)ROWS * H * 4;
    unsigned char* w1t_lo = w1t_hi + (size_t)D * H * 4;
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)((MODE == 2 ? 2 : 1) * D * H * 4);
        umma::mbar_expect_tx(&w_bar_s, bytes);
        umma::bulk_g2s(w1t_hi, J.img + L.w1t, bytes, &w_bar_s);
    }
    static_assert(H == 128, "one 32-column TMEM load per thread");
    {
        const int col0 = half * 32;
        float v[32];
        umma::tmem_ld32(tmem + lane_off + col0, v);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
            const int col = col0 + i4 * 4;
            const uint32_t m4 = act_mask >> (4 * i4);
            float4 o;
            o.x = (m4 & 1u) ? v[i4 * 4] * p.keep_scale : 0.f;
            o.y = (m4 & 2u) ? v[i4 * 4 + 1] * p.keep_scale : 0.f;
            o.z = (m4 & 4u) ? v[i4 * 4 + 2] * p.keep_scale : 0.f;
            o.w = (m4 & 8u) ? v[i4 * 4 + 3] * p.keep_scale : 0.f;
            if (valid) *reinterpret_cast<float4*>(J.dact + (long long)row * H + col) = o;
            put4<MODE>(a_hi, a_lo, ROWS, r_own, col, o);
        }
    }
    umma::fence_before_sync();
    umma::fence_proxy_async();
    w_bar.wait();
    __syncthreads();
    if (!dead && umma::elect_issuer(tid)) {
        umma::fence_after_sync();
        issue_gemm<MODE>(tmem + H, a_hi, a_lo, ROWS, w1t_hi, w1t_lo, D, D, H, false);
        umma::commit(&mma_bar_s);
    }
    mma_bar.wait();
    umma::fence_after_sync();
    {
        constexpr int QC = D / 4;
        float v[QC];
        umma::tmem_ld16(tmem + lane_off + H + half * QC, v);
        if (valid) {
#pragma unroll
            for (int i4 = 0; i4 < QC / 4; ++i4)
                *reinterpret_cast<float4*>(J.dRows + (long long)row * D + half * QC + i4 * 4) =
                    make_float4(v[i4 * 4], v[i4 * 4 + 1], v[i4 * 4 + 2], v[i4 * 4 + 3]);
        }
    }
    RB_TC_EPILOGUE()
}

// ------------------------------------------------------------------------------------------------------------ //
// backward, weight part (split over the batch)
// ------------------------------------------------------------------------------------------------------------ //
constexpr int KC = 32;    // samples per staged chunk

template <int D, int H, int NK, int MODE>      // NK = padded Din (multiple of 32) = N of the dW1 product
__global__ void __launch_bounds__(NT, 1) tower_bwd_weights_tc_kernel(const BwdParams p, int* err_flag) {
    static_assert(H == 128 && D == 64 && KC == 32 && NT == 256, "thread mapping assumes 8 warps = 8 sample quads, 32 lanes = 32 row quads");
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ long long ids_s[2][KC];
    __shared__ float bias1_part[KC / 4][H];      // per sample-quad partial column sums of dact
    __shared__ float bias2_part[KC / 4][D];      // … of dpre
    RB_TC_PROLOGUE(err_flag)
    (void)w_bar;
    const int s = blockIdx.x;
    const int E = p.job[0].E, Din = D + E;
    // chunk buffers: A1 = hidᵀ [H × KC], B1 = dpreᵀ [D × KC], A2 = dactᵀ [H × KC], B2 = Xᵀ [NK × KC]
    unsigned char* a1_hi = smem;
    unsigned char* a1_lo = a1_hi + H * KC * 4;
    unsigned char* b1_hi = a1_lo + H * KC * 4;
    unsigned char* b1_lo = b1_hi + D * KC * 4;
    unsigned char* a2_hi = b1_lo + D * KC * 4;
    unsigned char* a2_lo = a2_hi + H * KC * 4;
    unsigned char* b2_hi = a2_lo + H * KC * 4;
    unsigned char* b2_lo = b2_hi + NK * KC * 4;
    const int sq = warp, mq = lane;              // this thread's 4 samples (4·sq…) × 4 operand rows (4·mq…)
    float4 db1 = make_float4(0.f, 0.f, 0.f, 0.f), db2 = db1;
    bool first = true;
    // The tensor core's fp32 accumulation is not round-to-nearest: over a long chain the error grows linearly and is
    // amplified when large partial sums cancel later (positive vs negative items in dW2).  Each chunk therefore starts a
    // fresh TMEM accumulation (12 MMAs) that is flushed into these registers with ordinary fp32 adds.
    constexpr int HC = D / 2, NB = (NK / 32 + 1) / 2;
    float acc_a[HC];             // dW2ᵀ[h = r_own][d = half·HC + i]
    float acc_b[NB][32];         // dW1[h = r_own][k = (half + 2·b)·32 + i]
#pragma unroll
    for (int i = 0; i < HC; ++i) acc_a[i] = 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc_b[b][i] = 0.f;
    // ---- this CTA's chunks: up to two jobs (positive / negative items share the weights), rows [r_begin, r_end) of each -- //
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
    auto load_ids = [&](int t) {                 // ids of chunk t → ids_s[t & 1] (visible after the next __syncthreads)
        if (tid < KC && t < n_chunks) {
            int j, r0, nr;
            chunk_at(t, j, r0, nr);
            long long id = tid < nr ? p.job[j].ids[r0 + tid] : 0;
            if ((unsigned long long)id >= (unsigned long long)p.job[j].n_rows) id = 0;
            ids_s[t & 1][tid] = id;
        }
    };
    // all global loads of a chunk (16-byte, coalesced along the operand-row dimension); issued one chunk ahead so that their
    // latency overlaps the previous chunk's MMAs and accumulator flush
    float4 hv[4], av[4], gv[4], xv[4];
    auto load_chunk = [&](int t) {
        int j, r0, nr;
        chunk_at(t, j, r0, nr);
        const BwdJob& J = p.job[j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = sq * 4 + i;
            const bool ok = r < nr;
            const long long gr = r0 + r;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            hv[i] = ok ? __ldg(reinterpret_cast<const float4*>(J.hid + gr * H) + mq) : z;
            av[i] = ok ? __ldg(reinterpret_cast<const float4*>(J.dact + gr * H) + mq) : z;
            gv[i] = (ok && mq < D / 4) ? __ldg(reinterpret_cast<const float4*>(J.dpre + gr * D) + mq) : z;
            float4 x = z;
            if (ok && mq < D / 4) {
                x = __ldg(reinterpret_cast<const float4*>(J.table + ids_s[t & 1][r] * D) + mq);
            } else if (ok && mq < NK / 4 && E > 0) {
                const int k = mq * 4 - D;          // genre column
                const float* ex = J.extra + (J.extra_by_id ? ids_s[t & 1][r] : gr) * E;
                x.x = k + 0 < E ? __ldg(ex + k + 0) : 0.f;
                x.y = k + 1 < E ? __ldg(ex + k + 1) : 0.f;
                x.z = k + 2 < E ? __ldg(ex + k + 2) : 0.f;
                x.w = k + 3 < E ? __ldg(ex + k + 3) : 0.f;
            }
            xv[i] = x;
        }
    };
    load_ids(0);
    load_ids(1);
    __syncthreads();
    if (n_chunks > 0) load_chunk(0);
    for (int t = 0; t < n_chunks; ++t) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            db1.x += av[i].x; db1.y += av[i].y; db1.z += av[i].z; db1.w += av[i].w;
            db2.x += gv[i].x; db2.y += gv[i].y; db2.z += gv[i].z; db2.w += gv[i].w;
        }
        put_block_t<MODE>(a1_hi, a1_lo, H, mq, sq, lane, hv);
        put_block_t<MODE>(a2_hi, a2_lo, H, mq, sq, lane, av);
        if (mq < D / 4) put_block_t<MODE>(b1_hi, b1_lo, D, mq, sq, lane, gv);
        if (mq < NK / 4) put_block_t<MODE>(b2_hi, b2_lo, NK, mq, sq, lane, xv);
        umma::fence_proxy_async();
        __syncthreads();                         // operands staged; ids of chunk t+1 (stored one iteration ago) visible
        if (!dead && umma::elect_issuer(tid)) {
            umma::fence_after_sync();
            issue_gemm<MODE>(tmem, a1_hi, a1_lo, H, b1_hi, b1_lo, D, D, KC, false);        // dW2ᵀ [H × D]
            issue_gemm<MODE>(tmem + D, a2_hi, a2_lo, H, b2_hi, b2_lo, NK, NK, KC, false);  // dW1  [H × NK]
            umma::commit(&mma_bar_s);
        }
        first = false;
        if (t + 1 < n_chunks) load_chunk(t + 1);  // in flight while the tensor core works on chunk t
        mma_bar.wait();                          // chunk buffers are reused
        umma::fence_after_sync();
        if (!dead) {
#pragma unroll
            for (int cb = 0; cb < HC / 32; ++cb) {
                float v[32];
                umma::tmem_ld32(tmem + lane_off + half * HC + cb * 32, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) acc_a[cb * 32 + i] += v[i];
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int cb = half + 2 * b;
                if (cb < NK / 32) {
                    float v[32];
                    umma::tmem_ld32(tmem + lane_off + D + cb * 32, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc_b[b][i] += v[i];
                }
            }
        }
        load_ids(t + 2);                         // ids_s[t & 1] was last read by load_chunk(t), long done
        umma::fence_before_sync();
        __syncthreads();
    }
    // partial block layout: [W1ᵀ (Din*H, transposed) | b1 (H) | W2 (D*H) | b2 (D)]; reduce_partials_tc_kernel undoes the transpose
    float* part = p.part + (long long)s * p.P;
    float* w1o = part, *b1o = part + H * Din, *w2o = part + H * Din + H, *b2o = part + H * Din + H + D * H;
    *reinterpret_cast<float4*>(&bias1_part[sq][mq * 4]) = db1;
    if (mq < D / 4) *reinterpret_cast<float4*>(&bias2_part[sq][mq * 4]) = db2;
    __syncthreads();
    if (first) {          // this CTA had no rows: its partial is all zeros
        for (int i = tid; i < p.P; i += NT) part[i] = 0.f;
    } else {
        const int h = r_own;          // accumulator row
#pragma unroll
        for (int i = 0; i < HC; ++i) w2o[(long long)(half * HC + i) * H + h] = acc_a[i];          // dW2[d][h] = acc[h][d]
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int cb = half + 2 * b;
            if (cb < NK / 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i)          // the W1 block of a partial is stored TRANSPOSED ([k][h]): 128-byte warp stores
                    if (cb * 32 + i < Din) w1o[(long long)(cb * 32 + i) * H + h] = acc_b[b][i];  // dW1[h][k]
            }
        }
        if (tid < H) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < KC / 4; ++q) t += bias1_part[q][tid];      // fixed order
            b1o[tid] = t;
        }
        if (tid < D) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < KC / 4; ++q) t += bias2_part[q][tid];
            b2o[tid] = t;
        }
    }
    RB_TC_EPILOGUE()
}

// out[i] = Σ_k part[k][i] in a fixed order.  Block = 32 partial elements × 8 k-groups: group y adds the partials k ≡ y (mod 8)
// (coalesced 128-byte rows), the 8 group sums are combined through shared memory in index order.  The first H·Din elements
// of a partial hold W1 transposed ([k][h]); they are written back to out as [h][k].
__global__ void __launch_bounds__(256) reduce_partials_tc_kernel(const float* __restrict__ part, int nsplit, int P, int H, int Din,
                                                                 float* __restrict__ out, int accumulate) {
    __shared__ float sm[8][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + x;
    float a = 0.f;
    if (i < P) {
        for (int k0 = y; k0 < nsplit; k0 += 48) {          // six independent loads in flight; additions in the order k = y, y+8, …
            float v[6];
#pragma unroll
            for (int u = 0; u < 6; ++u) { const int k = k0 + 8 * u; v[u] = k < nsplit ? __ldg(part + (long long)k * P + i) : 0.f; }
#pragma unroll
            for (int u = 0; u < 6; ++u) a += v[u];
        }
    }
    sm[y][x] = a;
    __syncthreads();
    if (y == 0 && i < P) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) s += sm[g][x];
        int o = i;
        if (i < H * Din) { const int k = i / H, h = i - k * H; o = h * Din + k; }
        out[o] = accumulate ? out[o] + s : s;
    }
}

}  // namespace

int rb_tower_fwd_ts(FwdParams& p, int D, int mode, cudaStream_t st);
int rb_tower_bwd_ts(BwdParams& p, int D, int mode, cudaStream_t st);

bool rb_tower_tc_supported(int D, int H, int E) { return (D == 64 || D == 128) && H == 128 && E >= 0 && E <= 24; }

// which tensor-core implementation serves width D: the TMEM-operand kernels (tower_ts.cu) — always at D = 128 (the kernels of this
// file do not fit it), and at D = 64 unless RB200_TOWER_TS=0 selects the shared-memory-operand kernels of this file (measured on
// the C2 step: 0.128 ms with the TMEM-operand kernels, 0.137 ms with these)
bool rb_tower_use_ts(int D) {
    static int ts64 = -1;
    if (ts64 < 0) { const char* e = getenv("RB200_TOWER_TS"); ts64 = (e && atoi(e) == 0) ? 0 : 1; }
    return D == 128 || ts64 == 1;
}

size_t rb_tower_img_bytes(int D, int H, int E) { return rb_align_up(img_layout(D, H, E).total, 256); }

// builds the weight images of up to 2 weight sets in one launch
int rb_tower_prep_tc(int n_sets, const float* const W1[], const float* const W2[], const int E[], int D, int H,
                     unsigned char* const img[], rb200_opt_state* opt, cudaStream_t st) {
    RB_REQUIRE(n_sets >= 1 && n_sets <= 2, "tower_prep: 1 or 2 weight sets");
    PrepParams p{};
    p.n_sets = n_sets; p.D = D; p.H = H; p.opt = opt;
    for (int i = 0; i < n_sets; ++i) {
        RB_REQUIRE(W1[i] && W2[i] && img[i], "tower_prep: NULL pointer");
        RB_REQUIRE((reinterpret_cast<uintptr_t>(img[i]) & 15) == 0, "tower_prep: image must be 16-byte aligned");
        p.set[i].W1 = W1[i]; p.set[i].W2 = W2[i]; p.set[i].img = img[i]; p.set[i].E = E[i];
    }
    const int items = H * (((D + 24 + 7) & ~7) / 4) + 3 * D * (H / 4);      // float4 units of the largest image
    tower_prep_kernel<<<dim3((items + 255) / 256, n_sets), 256, 0, st>>>(p);
    RB_LAUNCH_CHECK("tower_prep_kernel");
    return RB200_OK;
}

extern "C" size_t rb200_tower_img_bytes(int D, int H, int extra_dim) { return rb_tower_img_bytes(D, H, extra_dim); }

extern "C" int rb200_tower_prep(const float* W1, const float* W2, int D, int H, int extra_dim, void* img, void* stream) {
    RB_REQUIRE(rb_tower_tc_supported(D, H, extra_dim), "tower_prep: tensor-core modes cover D in {64,128}, H=128, extra_dim<=24");
    const float* w1[1] = {W1}; const float* w2[1] = {W2}; const int e[1] = {extra_dim};
    unsigned char* im[1] = {(unsigned char*)img};
    return rb_tower_prep_tc(1, w1, w2, e, D, H, im, nullptr, (cudaStream_t)stream);
}

// jobs whose img is NULL get their image built into the workspace (jobs sharing W1 share the image)
int rb_tower_fwd_tc(FwdParams& p, int D, int H, int mode, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    RbArena ar(workspace, workspace_bytes);
    const float* w1[2]; const float* w2[2]; int es[2]; unsigned char* im[2];
    int n_sets = 0;
    for (int j = 0; j < p.n_jobs; ++j) {
        if (p.job[j].img) continue;
        int found = -1;
        for (int k = 0; k < n_sets; ++k) if (w1[k] == p.job[j].W1 && es[k] == p.job[j].E) found = k;
        if (found < 0) {
            RB_REQUIRE(n_sets < 2, "tower_fwd (tcgen05): at most two distinct weight sets per call");
            w1[n_sets] = p.job[j].W1; w2[n_sets] = p.job[j].W2; es[n_sets] = p.job[j].E;
            im[n_sets] = ar.take<unsigned char>(rb_tower_img_bytes(D, H, p.job[j].E));
            found = n_sets++;
        }
        p.job[j].img = im[found];
    }
    if (n_sets) {
        if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "tower_fwd (tcgen05): workspace too small for the weight images");
        int rc = rb_tower_prep_tc(n_sets, w1, w2, es, D, H, im, nullptr, st);
        if (rc) return rc;
    }
    if (rb_tower_use_ts(D)) return rb_tower_fwd_ts(p, D, mode, st);
    const int grid = assign_tiles(p.job, p.n_jobs);
    int kp = 0;
    for (int j = 0; j < p.n_jobs; ++j) { const int k = (64 + p.job[j].E + 7) & ~7; if (k > kp) kp = k; }
    const size_t s1 = (size_t)4 * 128 * kp * 4, s2 = (size_t)2 * 128 * 128 * 4 + (size_t)2 * 64 * 128 * 4;
    const size_t smem = s1 > s2 ? s1 : s2;
    static bool attr = false;
    if (!attr) {
        int rc;
        if ((rc = set_smem(tower_fwd_tc_kernel<64, 128, 1>, 200 * 1024))) return rc;
        if ((rc = set_smem(tower_fwd_tc_kernel<64, 128, 2>, 200 * 1024))) return rc;
        attr = true;
    }
    RB_REQUIRE(smem <= 200 * 1024, "tower_fwd (tcgen05): shared memory budget exceeded");
    if (mode == 1) tower_fwd_tc_kernel<64, 128, 1><<<grid, NTD, smem, st>>>(p);
    else tower_fwd_tc_kernel<64, 128, 2><<<grid, NTD, smem, st>>>(p);
    RB_LAUNCH_CHECK("tower_fwd_tc_kernel");
    return RB200_OK;
}

// p.part holds the split-K partials; `img_ws` (rb_tower_img_bytes) is used when the jobs carry no image
// grads_out == NULL: the split-K partials are left unreduced for the caller (csrc/step.cu fuses the reduction into its
// gradient-finish kernel)
int rb_tower_bwd_tc(BwdParams& p, int D, int H, int mode, float* grads_out, int accumulate, unsigned char* img_ws, cudaStream_t st) {
    const int E = p.job[0].E;
    if (!p.job[0].img) {
        RB_REQUIRE(img_ws, "tower_bwd (tcgen05): no weight image and no workspace for one");
        const float* w1[1] = {p.job[0].W1}; const float* w2[1] = {p.job[0].W2}; const int e[1] = {E};
        unsigned char* im[1] = {img_ws};
        int rc = rb_tower_prep_tc(1, w1, w2, e, D, H, im, nullptr, st);
        if (rc) return rc;
        for (int j = 0; j < p.n_jobs; ++j) p.job[j].img = img_ws;
    }
    if (rb_tower_use_ts(D)) {
        int rc = rb_tower_bwd_ts(p, D, mode, st);
        if (rc || !grads_out) return rc;
        reduce_partials_tc_kernel<<<(p.P + 31) / 32, 256, 0, st>>>(p.part, p.nsplit, p.P, H, D + E, grads_out, accumulate);
        RB_LAUNCH_CHECK("reduce_partials_tc_kernel");
        return RB200_OK;
    }
    const int grid = assign_tiles(p.job, p.n_jobs);
    const size_t smem_d = (size_t)2 * 128 * 128 * 4 + (size_t)2 * 64 * 128 * 4;
    const int NK = 96;
    const size_t smem_w = (size_t)4 * (128 + 128 + 64 + NK) * KC * 2;
    static bool attr = false;
    if (!attr) {
        int rc;
        if ((rc = set_smem(tower_bwd_data_tc_kernel<64, 128, 1>, 200 * 1024))) return rc;
        if ((rc = set_smem(tower_bwd_data_tc_kernel<64, 128, 2>, 200 * 1024))) return rc;
        if ((rc = set_smem(tower_bwd_weights_tc_kernel<64, 128, 64, 1>, 128 * 1024))) return rc;
        if ((rc = set_smem(tower_bwd_weights_tc_kernel<64, 128, 64, 2>, 128 * 1024))) return rc;
        if ((rc = set_smem(tower_bwd_weights_tc_kernel<64, 128, 96, 1>, 128 * 1024))) return rc;
        if ((rc = set_smem(tower_bwd_weights_tc_kernel<64, 128, 96, 2>, 128 * 1024))) return rc;
        attr = true;
    }
    if (mode == 1) tower_bwd_data_tc_kernel<64, 128, 1><<<grid, NTD, smem_d, st>>>(p, nullptr);
    else tower_bwd_data_tc_kernel<64, 128, 2><<<grid, NTD, smem_d, st>>>(p, nullptr);
    RB_LAUNCH_CHECK("tower_bwd_data_tc_kernel");
    if (E == 0) {
        if (mode == 1) tower_bwd_weights_tc_kernel<64, 128, 64, 1><<<p.nsplit, NT, smem_w, st>>>(p, nullptr);
        else tower_bwd_weights_tc_kernel<64, 128, 64, 2><<<p.nsplit, NT, smem_w, st>>>(p, nullptr);
    } else {
        if (mode == 1) tower_bwd_weights_tc_kernel<64, 128, 96, 1><<<p.nsplit, NT, smem_w, st>>>(p, nullptr);
        else tower_bwd_weights_tc_kernel<64, 128, 96, 2><<<p.nsplit, NT, smem_w, st>>>(p, nullptr);
    }
    RB_LAUNCH_CHECK("tower_bwd_weights_tc_kernel");
    if (!grads_out) return RB200_OK;
    reduce_partials_tc_kernel<<<(p.P + 31) / 32, 256, 0, st>>>(p.part, p.nsplit, p.P, H, 64 + E, grads_out, accumulate);
    RB_LAUNCH_CHECK("reduce_partials_tc_kernel");
    return RB200_OK;
}
