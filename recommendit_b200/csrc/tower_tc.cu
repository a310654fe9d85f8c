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
#include "common.cuh"
#include "tower_common.cuh"
#include "umma.cuh"

namespace {

constexpr int NT = 128;        // threads per CTA = rows per tile = TMEM lanes
constexpr int TMEM_COLS = 256;

struct Smem {
    unsigned char* base;
    __device__ unsigned char* at(size_t off) const { return base + off; }
};

// store 4 consecutive-k values of row r (hi and optionally lo) into an [R × K] K-major operand
template <int MODE>
__device__ __forceinline__ void put4(unsigned char* hi_base, unsigned char* lo_base, int R, int r, int k, const float4& v) {
    const uint32_t off = umma::kmajor_offset(R, r, k);
    float4 hi, lo;
    umma::split4(v, hi, lo);
    *reinterpret_cast<float4*>(hi_base + off) = hi;
    if (MODE == 2) *reinterpret_cast<float4*>(lo_base + off) = lo;
}

// issue the MMAs of one GEMM: D[tmem] = A[M=128 × K] · B[N × K]ᵀ, operands fully resident (K-major, rows RA / RB)
template <int MODE>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, unsigned char* a_hi, unsigned char* a_lo, int RA, unsigned char* b_hi,
                                           unsigned char* b_lo, int RB, int N, int K, bool accumulate_first) {
    const uint32_t idesc = umma::idesc_tf32(128, N);
    const uint32_t lbo_a = (RA / 8) * 128, lbo_b = (RB / 8) * 128;
    const uint32_t ah = umma::smem_u32(a_hi), al = umma::smem_u32(a_lo), bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo);
    for (int j = 0; j < K / 8; ++j) {
        const uint32_t oa = 2 * j * lbo_a, ob = 2 * j * lbo_b;
        const bool acc = accumulate_first || j > 0;
        if (MODE == 2) {
            umma::mma_tf32(tmem_d, umma::smem_desc(al + oa, lbo_a, 128), umma::smem_desc(bh + ob, lbo_b, 128), idesc, acc);
            umma::mma_tf32(tmem_d, umma::smem_desc(ah + oa, lbo_a, 128), umma::smem_desc(bl + ob, lbo_b, 128), idesc, true);
            umma::mma_tf32(tmem_d, umma::smem_desc(ah + oa, lbo_a, 128), umma::smem_desc(bh + ob, lbo_b, 128), idesc, true);
        } else {
            umma::mma_tf32(tmem_d, umma::smem_desc(ah + oa, lbo_a, 128), umma::smem_desc(bh + ob, lbo_b, 128), idesc, acc);
        }
    }
}

struct Pipe {          // TMEM + one mbarrier, with bounded waits
    uint32_t tmem;
    uint64_t* bar;
    uint32_t phase;
    int* dead;
    int* err_flag;
    __device__ void wait() {
        if (!*dead && !umma::mbar_wait(bar, phase)) { *dead = 1; if (err_flag) atomicOr(err_flag, 2); }
        phase ^= 1;
        umma::fence_after_sync();
    }
};

#define RB_TC_PROLOGUE(err_ptr)                                                        \
    __shared__ __align__(8) uint64_t mbar;                                             \
    __shared__ uint32_t tmem_slot;                                                     \
    __shared__ int dead;                                                               \
    const int tid = threadIdx.x, warp = tid >> 5;                                      \
    if (warp == 0) umma::tmem_alloc(&tmem_slot, TMEM_COLS);                            \
    if (tid == 0) { umma::mbar_init(&mbar, 1); umma::fence_mbar_init(); dead = 0; }    \
    umma::fence_before_sync();                                                         \
    __syncthreads();                                                                   \
    umma::fence_after_sync();                                                          \
    Pipe pipe{tmem_slot, &mbar, 0u, &dead, err_ptr};                                   \
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

#define RB_TC_EPILOGUE()                                                               \
    umma::fence_before_sync();                                                         \
    __syncthreads();                                                                   \
    if (warp == 0) umma::tmem_free(pipe.tmem, TMEM_COLS);

__device__ __forceinline__ int find_job(const int* begin, int n_jobs) {
    int j = 0;
#pragma unroll
    for (int t = 1; t < MAX_JOBS; ++t)
        if (t < n_jobs && (int)blockIdx.x >= begin[t]) j = t;
    return j;
}

// ------------------------------------------------------------------------------------------------------------ //
// forward
// ------------------------------------------------------------------------------------------------------------ //
template <int D, int H, int MODE>
__global__ void __launch_bounds__(NT, 1) tower_fwd_tc_kernel(const FwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    RB_TC_PROLOGUE(p.err_flag)
    int begins[MAX_JOBS];
#pragma unroll
    for (int t = 0; t < MAX_JOBS; ++t) begins[t] = p.job[t].cta_begin;
    const int j = find_job(begins, p.n_jobs);
    const FwdJob J = p.job[j];
    const int E = J.E, Din = D + E, Kp = (Din + 7) & ~7;
    const int row = ((int)blockIdx.x - J.cta_begin) * NT + tid;
    const bool valid = row < J.B;

    // ---- stage 1: X [128 × Kp], W1 [H × Kp] ------------------------------------------------------------------ //
    unsigned char* x_hi = smem;
    unsigned char* x_lo = x_hi + (size_t)NT * Kp * 4;
    unsigned char* w1_hi = x_lo + (size_t)NT * Kp * 4;
    unsigned char* w1_lo = w1_hi + (size_t)H * Kp * 4;
    {
        long long id = valid ? J.ids[row] : 0;
        if ((unsigned long long)id >= (unsigned long long)J.n_rows) { if (p.err_flag) atomicOr(p.err_flag, 1); id = 0; }
        const float4* src = reinterpret_cast<const float4*>(J.table + id * D);
#pragma unroll 4
        for (int c4 = 0; c4 < D / 4; ++c4)
            put4<MODE>(x_hi, x_lo, NT, tid, c4 * 4, valid ? __ldg(src + c4) : make_float4(0.f, 0.f, 0.f, 0.f));
        const float* ex = J.extra ? J.extra + (J.extra_by_id ? id : (long long)row) * E : nullptr;
        for (int k = D; k < Kp; k += 4) {
            float4 v;
            v.x = (valid && k + 0 < Din) ? __ldg(ex + k + 0 - D) : 0.f;
            v.y = (valid && k + 1 < Din) ? __ldg(ex + k + 1 - D) : 0.f;
            v.z = (valid && k + 2 < Din) ? __ldg(ex + k + 2 - D) : 0.f;
            v.w = (valid && k + 3 < Din) ? __ldg(ex + k + 3 - D) : 0.f;
            put4<MODE>(x_hi, x_lo, NT, tid, k, v);
        }
        if (tid < H) {
            const float* wr = J.W1 + (long long)tid * Din;
            for (int k = 0; k < Kp; k += 4) {
                float4 v;
                v.x = k + 0 < Din ? __ldg(wr + k + 0) : 0.f;
                v.y = k + 1 < Din ? __ldg(wr + k + 1) : 0.f;
                v.z = k + 2 < Din ? __ldg(wr + k + 2) : 0.f;
                v.w = k + 3 < Din ? __ldg(wr + k + 3) : 0.f;
                put4<MODE>(w1_hi, w1_lo, H, tid, k, v);
            }
        }
    }
    umma::fence_proxy_async();
    __syncthreads();
    if (tid == 0 && !dead) {
        umma::fence_after_sync();
        issue_gemm<MODE>(pipe.tmem, x_hi, x_lo, NT, w1_hi, w1_lo, H, H, Kp, false);
        umma::commit(&mbar);
    }
    pipe.wait();
    __syncthreads();          // everyone is past GEMM1: stage-1 buffers are free

    // ---- epilogue 1 + stage 2: hidden [128 × H], W2 [D × H] ---------------------------------------------------- //
    unsigned char* h_hi = smem;
    unsigned char* h_lo = h_hi + (size_t)NT * H * 4;
    unsigned char* w2_hi = h_lo + (size_t)NT * H * 4;
    unsigned char* w2_lo = w2_hi + (size_t)D * H * 4;
    {
        const bool do_drop = p.drop_p > 0.f;
        const float keep_scale = do_drop ? 1.f / (1.f - p.drop_p) : 1.f;
        const unsigned long long drop_off = p.offset + (unsigned long long)j +
                                            (p.offset_dev ? (unsigned long long)__ldg(p.offset_dev) * MAX_JOBS : 0ull);
#pragma unroll 1
        for (int cb = 0; cb < H / 32; ++cb) {
            float v[32];
            umma::tmem_ld32(pipe.tmem + lane_off + cb * 32, v);
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const int col = cb * 32 + i4 * 4;
                float4 o;
                float* op = &o.x;
                uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
                if (do_drop && J.keep_mask == nullptr)
                    rnd = rb_philox4x32(make_uint4((uint32_t)row, (uint32_t)(col >> 2), (uint32_t)drop_off, (uint32_t)(drop_off >> 32)),
                                        make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
                const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float x = fmaxf(v[i4 * 4 + e] + __ldg(J.b1 + col + e), 0.f);
                    if (do_drop) {
                        bool keep;
                        if (J.keep_mask) keep = valid ? (J.keep_mask[(long long)row * H + col + e] != 0) : true;
                        else keep = rb_u01(rw[e]) >= p.drop_p;
                        x = keep ? x * keep_scale : 0.f;
                    }
                    op[e] = x;
                }
                if (J.hid && valid) *reinterpret_cast<float4*>(J.hid + (long long)row * H + col) = o;
                put4<MODE>(h_hi, h_lo, NT, tid, col, o);
            }
        }
        if (tid < D) {
            const float4* wr = reinterpret_cast<const float4*>(J.W2 + (long long)tid * H);
#pragma unroll 4
            for (int c4 = 0; c4 < H / 4; ++c4) put4<MODE>(w2_hi, w2_lo, D, tid, c4 * 4, __ldg(wr + c4));
        }
    }
    umma::fence_before_sync();
    umma::fence_proxy_async();
    __syncthreads();
    if (tid == 0 && !dead) {
        umma::fence_after_sync();
        issue_gemm<MODE>(pipe.tmem + H, h_hi, h_lo, NT, w2_hi, w2_lo, D, D, H, false);
        umma::commit(&mbar);
    }
    pipe.wait();

    // ---- epilogue 2: bias, L2 normalise -------------------------------------------------------------------------- //
    {
        float y[D];
        float ss = 0.f;
#pragma unroll
        for (int cb = 0; cb < D / 32; ++cb) {
            float v[32];
            umma::tmem_ld32(pipe.tmem + lane_off + H + cb * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float t = v[i] + __ldg(J.b2 + cb * 32 + i);
                y[cb * 32 + i] = t;
                ss = fmaf(t, t, ss);
            }
        }
        const float den = fmaxf(sqrtf(ss), NORM_EPS);
        if (valid) {
#pragma unroll
            for (int c4 = 0; c4 < D / 4; ++c4)
                *reinterpret_cast<float4*>(J.out + (long long)row * D + c4 * 4) =
                    make_float4(y[c4 * 4] / den, y[c4 * 4 + 1] / den, y[c4 * 4 + 2] / den, y[c4 * 4 + 3] / den);
            if (J.denom) J.denom[row] = den;
        }
    }
    RB_TC_EPILOGUE()
}

// ------------------------------------------------------------------------------------------------------------ //
// backward, data part
// ------------------------------------------------------------------------------------------------------------ //
template <int D, int H, int MODE>
__global__ void __launch_bounds__(NT, 1) tower_bwd_data_tc_kernel(const BwdParams p, int* err_flag) {
    extern __shared__ __align__(1024) unsigned char smem[];
    RB_TC_PROLOGUE(err_flag)
    int begins[MAX_JOBS];
#pragma unroll
    for (int t = 0; t < MAX_JOBS; ++t) begins[t] = p.job[t].cta_begin;
    const int j = find_job(begins, p.n_jobs);
    const BwdJob J = p.job[j];
    const int Din = D + J.E;
    const int row = ((int)blockIdx.x - J.cta_begin) * NT + tid;
    const bool valid = row < J.B;

    // ---- stage 1: dpre [128 × D] (A), W2ᵀ [H × D] (B) ------------------------------------------------------------ //
    unsigned char* g_hi = smem;
    unsigned char* g_lo = g_hi + (size_t)NT * D * 4;
    unsigned char* w2t_hi = g_lo + (size_t)NT * D * 4;
    unsigned char* w2t_lo = w2t_hi + (size_t)H * D * 4;
    {
        float g[D];
        float dot = 0.f;
        const float den = valid ? __ldg(J.denom + row) : 1.f;
#pragma unroll
        for (int c4 = 0; c4 < D / 4; ++c4) {
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), yv = gv;
            if (valid) {
                gv = __ldg(reinterpret_cast<const float4*>(J.dY + (long long)row * D) + c4);
                yv = __ldg(reinterpret_cast<const float4*>(J.y + (long long)row * D) + c4);
            }
            g[c4 * 4] = gv.x; g[c4 * 4 + 1] = gv.y; g[c4 * 4 + 2] = gv.z; g[c4 * 4 + 3] = gv.w;
            dot = fmaf(gv.x, yv.x, dot); dot = fmaf(gv.y, yv.y, dot); dot = fmaf(gv.z, yv.z, dot); dot = fmaf(gv.w, yv.w, dot);
        }
        const bool clamped = den <= NORM_EPS;
#pragma unroll
        for (int c4 = 0; c4 < D / 4; ++c4) {
            float4 yv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) yv = __ldg(reinterpret_cast<const float4*>(J.y + (long long)row * D) + c4);
            float4 o;
            o.x = clamped ? g[c4 * 4] / den : (g[c4 * 4] - yv.x * dot) / den;
            o.y = clamped ? g[c4 * 4 + 1] / den : (g[c4 * 4 + 1] - yv.y * dot) / den;
            o.z = clamped ? g[c4 * 4 + 2] / den : (g[c4 * 4 + 2] - yv.z * dot) / den;
            o.w = clamped ? g[c4 * 4 + 3] / den : (g[c4 * 4 + 3] - yv.w * dot) / den;
            if (valid) *reinterpret_cast<float4*>(J.dpre + (long long)row * D + c4 * 4) = o;
            put4<MODE>(g_hi, g_lo, NT, tid, c4 * 4, o);
        }
        if (tid < H) {       // B row n = h, K = d : W2ᵀ[h][d] = W2[d][h]
#pragma unroll 4
            for (int c4 = 0; c4 < D / 4; ++c4) {
                const float* wp = J.W2 + (long long)(c4 * 4) * H + tid;
                put4<MODE>(w2t_hi, w2t_lo, H, tid, c4 * 4, make_float4(__ldg(wp), __ldg(wp + H), __ldg(wp + 2 * H), __ldg(wp + 3 * H)));
            }
        }
    }
    umma::fence_proxy_async();
    __syncthreads();
    if (tid == 0 && !dead) {
        umma::fence_after_sync();
        issue_gemm<MODE>(pipe.tmem, g_hi, g_lo, NT, w2t_hi, w2t_lo, H, H, D, false);
        umma::commit(&mbar);
    }
    pipe.wait();
    __syncthreads();

    // ---- epilogue 1 + stage 2: dact [128 × H] (A), W1[:, :D]ᵀ [D × H] (B) ---------------------------------------- //
    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + (size_t)NT * H * 4;
    unsigned char* w1t_hi = a_lo + (size_t)NT * H * 4;
    unsigned char* w1t_lo = w1t_hi + (size_t)D * H * 4;
    {
#pragma unroll 1
        for (int cb = 0; cb < H / 32; ++cb) {
            float v[32];
            umma::tmem_ld32(pipe.tmem + lane_off + cb * 32, v);
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const int col = cb * 32 + i4 * 4;
                float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) hv = __ldg(reinterpret_cast<const float4*>(J.hid + (long long)row * H + col));
                float4 o;
                o.x = hv.x > 0.f ? v[i4 * 4] * p.keep_scale : 0.f;
                o.y = hv.y > 0.f ? v[i4 * 4 + 1] * p.keep_scale : 0.f;
                o.z = hv.z > 0.f ? v[i4 * 4 + 2] * p.keep_scale : 0.f;
                o.w = hv.w > 0.f ? v[i4 * 4 + 3] * p.keep_scale : 0.f;
                if (valid) *reinterpret_cast<float4*>(J.dact + (long long)row * H + col) = o;
                put4<MODE>(a_hi, a_lo, NT, tid, col, o);
            }
        }
        if (tid < D) {       // B row n = d, K = h : W1ᵀ[d][h] = W1[h][d]
#pragma unroll 4
            for (int c4 = 0; c4 < H / 4; ++c4) {
                const float* wp = J.W1 + (long long)(c4 * 4) * Din + tid;
                put4<MODE>(w1t_hi, w1t_lo, D, tid, c4 * 4,
                           make_float4(__ldg(wp), __ldg(wp + Din), __ldg(wp + 2 * Din), __ldg(wp + 3 * Din)));
            }
        }
    }
    umma::fence_before_sync();
    umma::fence_proxy_async();
    __syncthreads();
    if (tid == 0 && !dead) {
        umma::fence_after_sync();
        issue_gemm<MODE>(pipe.tmem + H, a_hi, a_lo, NT, w1t_hi, w1t_lo, D, D, H, false);
        umma::commit(&mbar);
    }
    pipe.wait();
#pragma unroll
    for (int cb = 0; cb < D / 32; ++cb) {
        float v[32];
        umma::tmem_ld32(pipe.tmem + lane_off + H + cb * 32, v);
        if (valid) {
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4)
                *reinterpret_cast<float4*>(J.dRows + (long long)row * D + cb * 32 + i4 * 4) =
                    make_float4(v[i4 * 4], v[i4 * 4 + 1], v[i4 * 4 + 2], v[i4 * 4 + 3]);
        }
    }
    RB_TC_EPILOGUE()
}

// ------------------------------------------------------------------------------------------------------------ //
// backward, weight part (split over the batch)
// ------------------------------------------------------------------------------------------------------------ //
constexpr int KC = 32;    // samples per staged chunk

template <int D, int H, int NK, int MODE>      // NK = padded Din (multiple of 16) = N of the dW1 product
__global__ void __launch_bounds__(NT, 1) tower_bwd_weights_tc_kernel(const BwdParams p, int* err_flag) {
    extern __shared__ __align__(1024) unsigned char smem[];
    RB_TC_PROLOGUE(err_flag)
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
    __shared__ long long ids_s[KC];
    float db1 = 0.f, db2 = 0.f;
    bool first = true;
    for (int j = 0; j < p.n_jobs; ++j) {
        const BwdJob& J = p.job[j];
        const int chunk = (((J.B + p.nsplit - 1) / p.nsplit) + KC - 1) / KC * KC;   // rows per CTA, multiple of KC
        const int r_begin = min(J.B, s * chunk), r_end = min(J.B, r_begin + chunk);
        for (int r0 = r_begin; r0 < r_end; r0 += KC) {
            const int nr = min(KC, r_end - r0);
            if (tid < KC) {
                long long id = tid < nr ? J.ids[r0 + tid] : 0;
                if ((unsigned long long)id >= (unsigned long long)J.n_rows) id = 0;
                ids_s[tid] = id;
            }
            __syncthreads();
            // thread t: operand row t, K = sample (4 samples per 16-byte chunk)
#pragma unroll 2
            for (int c4 = 0; c4 < KC / 4; ++c4) {
                float hv[4], av[4], gv[4], xv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int r = c4 * 4 + e;
                    const bool ok = r < nr;
                    const long long gr = r0 + r;
                    hv[e] = (ok && tid < H) ? __ldg(J.hid + gr * H + tid) : 0.f;
                    av[e] = (ok && tid < H) ? __ldg(J.dact + gr * H + tid) : 0.f;
                    gv[e] = (ok && tid < D) ? __ldg(J.dpre + gr * D + tid) : 0.f;
                    float x = 0.f;
                    if (ok && tid < Din) {
                        if (tid < D) x = __ldg(J.table + ids_s[r] * D + tid);
                        else x = __ldg(J.extra + (J.extra_by_id ? ids_s[r] : gr) * E + (tid - D));
                    }
                    xv[e] = x;
                    db1 += av[e];
                    db2 += gv[e];
                }
                if (tid < H) {
                    put4<MODE>(a1_hi, a1_lo, H, tid, c4 * 4, make_float4(hv[0], hv[1], hv[2], hv[3]));
                    put4<MODE>(a2_hi, a2_lo, H, tid, c4 * 4, make_float4(av[0], av[1], av[2], av[3]));
                }
                if (tid < D) put4<MODE>(b1_hi, b1_lo, D, tid, c4 * 4, make_float4(gv[0], gv[1], gv[2], gv[3]));
                if (tid < NK) put4<MODE>(b2_hi, b2_lo, NK, tid, c4 * 4, make_float4(xv[0], xv[1], xv[2], xv[3]));
            }
            umma::fence_proxy_async();
            __syncthreads();
            if (tid == 0 && !dead) {
                umma::fence_after_sync();
                issue_gemm<MODE>(pipe.tmem, a1_hi, a1_lo, H, b1_hi, b1_lo, D, D, KC, !first);        // dW2ᵀ [H × D]
                issue_gemm<MODE>(pipe.tmem + D, a2_hi, a2_lo, H, b2_hi, b2_lo, NK, NK, KC, !first);  // dW1  [H × NK]
                umma::commit(&mbar);
            }
            first = false;
            pipe.wait();          // chunk buffers are reused
            __syncthreads();
        }
    }
    // partial block layout: [W1 (H*Din) | b1 (H) | W2 (D*H) | b2 (D)]
    float* part = p.part + (long long)s * p.P;
    float* w1o = part, *b1o = part + H * Din, *w2o = part + H * Din + H, *b2o = part + H * Din + H + D * H;
    if (first) {          // this CTA had no rows: its partial is all zeros
        for (int i = tid; i < p.P; i += NT) part[i] = 0.f;
    } else if (tid < H) {
#pragma unroll
        for (int cb = 0; cb < D / 32; ++cb) {
            float v[32];
            umma::tmem_ld32(pipe.tmem + lane_off + cb * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) w2o[(long long)(cb * 32 + i) * H + tid] = v[i];       // dW2[d][h] = acc[h][d]
        }
#pragma unroll
        for (int cb = 0; cb < NK / 32; ++cb) {
            float v[32];
            umma::tmem_ld32(pipe.tmem + lane_off + D + cb * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (cb * 32 + i < Din) w1o[(long long)tid * Din + cb * 32 + i] = v[i];
        }
        b1o[tid] = db1;
        if (tid < D) b2o[tid] = db2;
    }
    RB_TC_EPILOGUE()
}

__global__ void reduce_partials_tc_kernel(const float* __restrict__ part, int nsplit, int P, float* __restrict__ out, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < nsplit; ++k) s += __ldg(part + (long long)k * P + i);   // fixed order
    out[i] = accumulate ? out[i] + s : s;
}

template <typename JobT>
int assign_tiles(JobT* jobs, int n_jobs) {
    int begin = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const int tiles = (jobs[j].B + NT - 1) / NT;
        jobs[j].cta_begin = begin;
        jobs[j].cta_count = tiles;
        begin += tiles;
    }
    return begin;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    RB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return RB200_OK;
}

}  // namespace

bool rb_tower_tc_supported(int D, int H, int E) { return D == 64 && H == 128 && E >= 0 && E <= 24; }

int rb_tower_fwd_tc(FwdParams& p, int D, int H, int mode, cudaStream_t st) {
    (void)D; (void)H;
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
    if (mode == 1) tower_fwd_tc_kernel<64, 128, 1><<<grid, NT, smem, st>>>(p);
    else tower_fwd_tc_kernel<64, 128, 2><<<grid, NT, smem, st>>>(p);
    RB_LAUNCH_CHECK("tower_fwd_tc_kernel");
    return RB200_OK;
}

int rb_tower_bwd_tc(BwdParams& p, int D, int H, int mode, float* grads_out, int accumulate, cudaStream_t st) {
    (void)D; (void)H;
    const int grid = assign_tiles(p.job, p.n_jobs);
    const int E = p.job[0].E;
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
    if (mode == 1) tower_bwd_data_tc_kernel<64, 128, 1><<<grid, NT, smem_d, st>>>(p, nullptr);
    else tower_bwd_data_tc_kernel<64, 128, 2><<<grid, NT, smem_d, st>>>(p, nullptr);
    RB_LAUNCH_CHECK("tower_bwd_data_tc_kernel");
    if (E == 0) {
        if (mode == 1) tower_bwd_weights_tc_kernel<64, 128, 64, 1><<<p.nsplit, NT, smem_w, st>>>(p, nullptr);
        else tower_bwd_weights_tc_kernel<64, 128, 64, 2><<<p.nsplit, NT, smem_w, st>>>(p, nullptr);
    } else {
        if (mode == 1) tower_bwd_weights_tc_kernel<64, 128, 96, 1><<<p.nsplit, NT, smem_w, st>>>(p, nullptr);
        else tower_bwd_weights_tc_kernel<64, 128, 96, 2><<<p.nsplit, NT, smem_w, st>>>(p, nullptr);
    }
    RB_LAUNCH_CHECK("tower_bwd_weights_tc_kernel");
    reduce_partials_tc_kernel<<<(p.P + 255) / 256, 256, 0, st>>>(p.part, p.nsplit, p.P, grads_out, accumulate);
    RB_LAUNCH_CHECK("reduce_partials_tc_kernel");
    return RB200_OK;
}
