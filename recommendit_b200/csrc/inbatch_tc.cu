// In-batch BPR (src/models/two_tower.py:132-160 of the reference) on the 5th-generation tensor cores.
//
//   S = U·Iᵀ,   loss = Σ_{i≠j} softplus(S_ij − S_ii) / (B(B−1)),
//   G_ij = g·σ(S_ij − S_ii) (i≠j; g = grad_scale / (B(B−1))),  r_i = Σ_j G_ij,
//   dU_i = Σ_j G_ij I_j − r_i I_i,      dI_j = Σ_i G_ij U_i − r_j U_j.
//
// The B×B matrices never exist in memory.  One templated kernel is run twice ("flash" style, scores recomputed):
//   pass U: X = U (128-row tile per CTA), Y = I streamed in 64-row tiles.  S-MMA: S = X·Yᵀ (tcgen05, TMEM).  Epilogue:
//           every thread owns one S row (TMEM lane) × 16 columns → softplus (loss), G; G goes to shared memory as the
//           K-major A operand of the second MMA  acc += G·Y  (B operand = Yᵀ, transposed while staged).
//   pass I: X = I, Y = U: the same code computes Sᵀ tiles; the diagonal S_ii now varies along the columns.
// The Y range is split over `split` CTAs per X tile so that the grid fills the SMs; partial sums (fp32 tiles, fp64 row
// sums) go to the workspace and a finishing kernel adds them in a fixed order and subtracts r·(other side): deterministic.
//
// The second product is centred: Σ_j G_ij (Y_j − c) − r_i (other_i − c) [+ (q_j − r_j) c in pass I, q = column sums of G]
// with c = column mean of Y.  The identity holds
// for any c; with c = mean the terms are small when the embeddings are alike (untrained towers), so the fp32-grade
// product error is not amplified by the cancellation between Σ_j G_ij Y_j and r_i·other_i.
//
// mode 2 = 3xTF32 (hi/lo split of both operands of both MMAs, fp32-grade), mode 1 = single TF32 (stated fast mode).
// The tensor core's fp32 accumulation is not round-to-nearest over long chains (see tower_tc.cu), so every Y tile starts
// a fresh TMEM accumulation that is flushed into fp32 registers.  D = 64 only (the production width).
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int NT_IB = 512;            // worker threads
constexpr int NT_ALL = NT_IB + 32;     // + the MMA-issuing warp
constexpr int XT = 128, YT = 64, DD = 64;
constexpr int X_BYTES = XT * DD * 4;      // 32 KB: X tile [128 × 64] K-major over d
constexpr int Y_BYTES = YT * DD * 4;      // 16 KB: Y tile [64 × 64] K-major over d
constexpr int YT_BYTES = DD * YT * 4;     // 16 KB: Yᵀ tile [64 d × 64 y] K-major over y
constexpr int G_BYTES = XT * YT * 4;      // 32 KB: G tile [128 × 64 y] K-major over y
constexpr int TMEM_COLS_IB = 256;         // S: 2 × 64 columns (double-buffered), acc: 64 columns

template <int MODE>
__device__ __forceinline__ void put4(unsigned char* hi_base, unsigned char* lo_base, int R, int r, int k, const float4& v) {
    const uint32_t off = umma::kmajor_offset(R, r, k);
    float4 hi, lo;
    umma::split4(v, hi, lo);
    *reinterpret_cast<float4*>(hi_base + off) = hi;
    if (MODE == 2) *reinterpret_cast<float4*>(lo_base + off) = lo;
}
__device__ __forceinline__ float sel4(const float4& v, int e) { return e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w; }

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// (hi, lo) += x, error-free (Knuth two-sum): fp32 pair instead of an fp64 accumulator (fp64 adds are slow on this part)
__device__ __forceinline__ void two_sum(float& hi, float& lo, float x) {
    const float t = hi + x;
    const float z = t - hi;
    lo += (hi - (t - z)) + (x - z);
    hi = t;
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

__global__ void __launch_bounds__(256) rowdot64_kernel(const float* __restrict__ U, const float* __restrict__ I, int B, int Bd,
                                                       float* __restrict__ diag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= Bd) return;
    if (row >= B) { if (lane == 0) diag[row] = 0.f; return; }
    const float2 a = __ldg(reinterpret_cast<const float2*>(U + (long long)row * DD) + lane);
    const float2 b = __ldg(reinterpret_cast<const float2*>(I + (long long)row * DD) + lane);
    float s = fmaf(a.x, b.x, a.y * b.y);
    s = rb_warp_sum(s);
    if (lane == 0) diag[row] = s;
}

struct IbParams {
    const float* X; const float* Y;      // row-side / streamed-side embeddings [B × 64]
    const float* diag;                   // S_ii
    const float* ycen;                   // [64] column means of Y (centring of the second product, see header)
    int B, split, tiles_per_split;       // Y tiles (of 64) per CTA
    float g;                             // grad_scale / (B(B−1))
    float* acc_part;                     // [split][Bp][64]
    double* r_part;                      // [split][NQ][Bp]  (pass U only)
    double* loss_part;                   // [gridDim.x]      (pass U only)
    int Bp;
    int* err_flag;
};

constexpr int NQ = NT_IB / 128;          // column quarters of a tile: 16 warps = 4 TMEM lane groups × 4 quarters
constexpr int QC = YT / NQ;              // 16 columns per thread

// descriptors of one resident GEMM (computed once; the k-step only moves the 14-bit start address field)
struct GemmDesc { uint64_t a_hi, a_lo, b_hi, b_lo; uint32_t step_a, step_b, idesc; };
__device__ __forceinline__ GemmDesc make_gemm(unsigned char* a_hi, unsigned char* a_lo, int RA, unsigned char* b_hi, unsigned char* b_lo,
                                              int RB, int N) {
    GemmDesc g;
    const uint32_t lbo_a = (RA / 8) * 128, lbo_b = (RB / 8) * 128;
    g.a_hi = umma::smem_desc(umma::smem_u32(a_hi), lbo_a, 128); g.a_lo = umma::smem_desc(umma::smem_u32(a_lo), lbo_a, 128);
    g.b_hi = umma::smem_desc(umma::smem_u32(b_hi), lbo_b, 128); g.b_lo = umma::smem_desc(umma::smem_u32(b_lo), lbo_b, 128);
    g.step_a = (2 * lbo_a) >> 4; g.step_b = (2 * lbo_b) >> 4;
    g.idesc = umma::idesc_tf32(128, N);
    return g;
}
template <int MODE, int KSTEPS>
__device__ __forceinline__ void issue(const GemmDesc& g, uint32_t tmem_d) {
#pragma unroll
    for (int j = 0; j < KSTEPS; ++j) {
        const uint64_t oa = (uint64_t)(j * g.step_a), ob = (uint64_t)(j * g.step_b);
        if (MODE == 2) {
            umma::mma_tf32(tmem_d, g.a_lo + oa, g.b_hi + ob, g.idesc, j > 0);
            umma::mma_tf32(tmem_d, g.a_hi + oa, g.b_lo + ob, g.idesc, true);
            umma::mma_tf32(tmem_d, g.a_hi + oa, g.b_hi + ob, g.idesc, true);
        } else {
            umma::mma_tf32(tmem_d, g.a_hi + oa, g.b_hi + ob, g.idesc, j > 0);
        }
    }
}

__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// PASS_I = false: X = U, diagonal indexed by the row;  true: X = I, diagonal indexed by the column.
//
// 16 worker warps + 1 MMA-issuing warp.  Software pipeline over the Y tiles (S double-buffered in TMEM): while the
// workers run the epilogue of tile t the tensor core computes S(t+1), and the loads of tile t+2 (K-major copy) /
// t+1 (transposed copy) are in flight.  Hand-offs: workers → issuer through named barriers 1 ("Y(t+1) staged", S(t)
// drained) and 2 ("G(t), Yᵀ(t) staged"); issuer → workers through tcgen05.commit on two mbarriers.
template <int MODE, bool PASS_I, bool WITH_GRAD>
__global__ void __launch_bounds__(NT_ALL, 1) inbatch_tc_kernel(const IbParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* x_hi = smem;
    unsigned char* x_lo = x_hi + X_BYTES;
    unsigned char* y_hi = x_lo + X_BYTES;
    unsigned char* y_lo = y_hi + Y_BYTES;
    unsigned char* yt_hi = y_lo + Y_BYTES;
    unsigned char* yt_lo = yt_hi + YT_BYTES;
    unsigned char* g_hi = yt_lo + YT_BYTES;
    unsigned char* g_lo = g_hi + G_BYTES;
    __shared__ __align__(8) uint64_t bar_s_mem, bar_g_mem;
    __shared__ uint32_t tmem_slot;
    __shared__ int dead;
    __shared__ double red[NT_IB / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, TMEM_COLS_IB);
    if (tid == 0) { umma::mbar_init(&bar_s_mem, 1); umma::mbar_init(&bar_g_mem, 1); umma::fence_mbar_init(); dead = 0; }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;                 // columns: S buffer 0 | S buffer 1 | acc
    const uint32_t tmem_acc = tmem + 2 * YT;
    const int B = p.B;
    const int xt = blockIdx.x / p.split, part = blockIdx.x - xt * p.split;
    const int x0 = xt * XT;
    const int t_begin = part * p.tiles_per_split;
    const int n_ytiles = (B + YT - 1) / YT;
    int t_end = t_begin + p.tiles_per_split;
    if (t_end > n_ytiles) t_end = n_ytiles;

    if (warp == NT_IB / 32) {
        // ================================ MMA-issuing warp ================================================== //
        GemmDesc gs = make_gemm(x_hi, x_lo, XT, y_hi, y_lo, YT, YT);        // S   [128 × 64 y] = X · Yᵀ
        GemmDesc gg = make_gemm(g_hi, g_lo, XT, yt_hi, yt_lo, DD, DD);      // acc [128 × 64 d] = G · Y
        bar_sync(1, NT_ALL);                                                // X and Y(t_begin) staged
        if (umma::elect_one()) {
            umma::fence_after_sync();
            issue<MODE, DD / 8>(gs, tmem);
            umma::commit(&bar_s_mem);
        }
        for (int t = t_begin; t < t_end; ++t) {
            bar_sync(1, NT_ALL);                                            // Y(t+1) staged, S(t−1) drained
            if (t + 1 < t_end && umma::elect_one()) {
                umma::fence_after_sync();
                issue<MODE, DD / 8>(gs, tmem + (uint32_t)((((t - t_begin) & 1) ^ 1) * YT));
                umma::commit(&bar_s_mem);
            }
            if (WITH_GRAD) {
                bar_sync(2, NT_ALL);                                        // G(t), Yᵀ(t) staged, acc(t−1) drained
                if (umma::elect_one()) {
                    umma::fence_after_sync();
                    issue<MODE, YT / 8>(gg, tmem_acc);
                    umma::commit(&bar_g_mem);
                }
            }
        }
    } else {
        // ================================ worker warps ====================================================== //
        Bar bar_s{&bar_s_mem, 0u, &dead, p.err_flag}, bar_g{&bar_g_mem, 0u, &dead, p.err_flag};
        const int r_own = ((warp & 3) << 5) + lane, quarter = warp >> 2;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const int gi = x0 + r_own;                   // this thread's global X row
        {   // stage the X tile once: thread = (row, quarter) → 4 float4
            const long long row = x0 + r_own;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = (quarter * 4 + q) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < B) v = __ldg(reinterpret_cast<const float4*>(p.X + row * DD + k));
                put4<MODE>(x_hi, x_lo, XT, r_own, k, v);
            }
        }
        const float d_own = (!PASS_I && gi < B) ? __ldg(p.diag + gi) : 0.f;
        float acc[QC];
#pragma unroll
        for (int i = 0; i < QC; ++i) acc[i] = 0.f;
        float r_hi = 0.f, r_lo = 0.f, loss_hi = 0.f, loss_lo = 0.f;
        // Staging roles: threads 256…511 copy the Y tile K-major (row yr, float4 columns 4·yq…), threads 0…255 copy it
        // transposed in 4(y) × 4(d) blocks (sq, mq); every thread carries 4 float4 of the tile it stages next.
        const bool role_t = tid < 256;
        const int yr = tid & 63, yq = (tid >> 6) & 3;
        const int mq = tid & 15, sq = (tid >> 4) & 15;
        const float4 cen = (WITH_GRAD && role_t) ? __ldg(reinterpret_cast<const float4*>(p.ycen) + mq) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v[4];
        auto load_tile = [&](int t) {                // role_t: transposed copy of tile t; else: K-major copy of tile t
            const int y0 = t * YT;
            if (role_t) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const long long rt = y0 + sq * 4 + i;
                    v[i] = (WITH_GRAD && t < t_end && rt < B) ? __ldg(reinterpret_cast<const float4*>(p.Y + rt * DD) + mq) : cen;
                    v[i].x -= cen.x; v[i].y -= cen.y; v[i].z -= cen.z; v[i].w -= cen.w;
                }
            } else {
                const long long row = y0 + yr;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    v[e] = (t < t_end && row < B) ? __ldg(reinterpret_cast<const float4*>(p.Y + row * DD) + yq * 4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto store_y = [&]() {                       // K-major copy (threads 256…511)
#pragma unroll
            for (int e = 0; e < 4; ++e) put4<MODE>(y_hi, y_lo, YT, yr, (yq * 4 + e) * 4, v[e]);
        };
        auto store_yt = [&]() {                      // transposed copy (threads 0…255), lane-rotated: conflict-free
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int e = (s + (lane >> 1)) & 3;
                put4<MODE>(yt_hi, yt_lo, DD, mq * 4 + e, sq * 4, make_float4(sel4(v[0], e), sel4(v[1], e), sel4(v[2], e), sel4(v[3], e)));
            }
        };
        // ---- prologue: Y(t_begin) staged → S(t_begin) in flight; registers hold Yᵀ(t_begin) / Y(t_begin+1) ---- //
        if (!role_t) { load_tile(t_begin); store_y(); }
        umma::fence_proxy_async();
        bar_arrive(1, NT_ALL);
        load_tile(role_t ? t_begin : t_begin + 1);
        bool pending_g = false;
        for (int t = t_begin; t < t_end; ++t) {
            const int y0 = t * YT;
            const uint32_t s_cur = tmem + (uint32_t)(((t - t_begin) & 1) * YT);
            float4 dq[QC / 4];                       // S_ii of this thread's columns (pass I; diag is padded with zeros)
            if (PASS_I) {
#pragma unroll
                for (int q = 0; q < QC / 4; ++q) dq[q] = __ldg(reinterpret_cast<const float4*>(p.diag + y0 + quarter * QC) + q);
            }
            bar_s.wait();                            // S(t) ready (issued one tile ago); the Y buffer is free
            umma::fence_after_sync();
            if (!role_t) store_y();                  // Y(t+1) (zeros past the end)
            umma::fence_proxy_async();
            umma::fence_before_sync();               // (this thread's tcgen05.ld of S(t−1) completed in the last iteration)
            bar_arrive(1, NT_ALL);
            if (!role_t) load_tile(t + 2);           // in flight during the epilogue
            // ---- epilogue of tile t: this thread's row × 16 columns, kept in registers ---------------------- //
            float s[QC], gq[QC];
            if (!dead) umma::tmem_ld16(s_cur + lane_off + quarter * QC, s);
            float tile_loss = 0.f, tile_r = 0.f;
            // tiles that touch neither the diagonal nor the ragged end need no masking (uniform per CTA)
            const bool clean = y0 + YT <= B && x0 + XT <= B && (y0 + YT <= x0 || y0 >= x0 + XT);
            if (clean) {
#pragma unroll
                for (int c = 0; c < QC; ++c) {
                    const float x = s[c] - (PASS_I ? sel4(dq[c >> 2], c & 3) : d_own);
                    const float ex = __expf(-fabsf(x));
                    const float inv = rcp_approx(1.f + ex);
                    gq[c] = p.g * (x >= 0.f ? inv : ex * inv);
                    tile_r += gq[c];
                    if (!PASS_I) tile_loss += fmaxf(x, 0.f) + __logf(1.f + ex);
                }
            } else {
#pragma unroll
                for (int c = 0; c < QC; ++c) {
                    const int gj = y0 + quarter * QC + c;
                    const float x = s[c] - (PASS_I ? sel4(dq[c >> 2], c & 3) : d_own);
                    const bool valid = gi < B && gj < B && gi != gj;
                    const float ex = __expf(-fabsf(x));
                    const float inv = rcp_approx(1.f + ex);
                    gq[c] = valid ? p.g * (x >= 0.f ? inv : ex * inv) : 0.f;
                    tile_r += gq[c];
                    if (!PASS_I) tile_loss += valid ? fmaxf(x, 0.f) + __logf(1.f + ex) : 0.f;
                }
            }
            two_sum(r_hi, r_lo, tile_r);             // row sums of this pass's G tile: r_i (pass U) / column sums q_j (pass I)
            if (!PASS_I) two_sum(loss_hi, loss_lo, tile_loss);
            if (WITH_GRAD) {
                if (pending_g) {                     // G-MMA(t−1) finished under the epilogue: G / Yᵀ buffers free, result in TMEM
                    bar_g.wait();
                    umma::fence_after_sync();
                    if (!dead) {
                        float f[QC];
                        umma::tmem_ld16(tmem_acc + lane_off + quarter * QC, f);
#pragma unroll
                        for (int i = 0; i < QC; ++i) acc[i] += f[i];
                    }
                }
                if (role_t) store_yt();              // Yᵀ(t)
#pragma unroll
                for (int q = 0; q < QC / 4; ++q)
                    put4<MODE>(g_hi, g_lo, XT, r_own, quarter * QC + q * 4, make_float4(gq[q * 4], gq[q * 4 + 1], gq[q * 4 + 2], gq[q * 4 + 3]));
                umma::fence_proxy_async();
                umma::fence_before_sync();
                bar_arrive(2, NT_ALL);
                pending_g = true;
                if (role_t) load_tile(t + 1);        // Yᵀ(t+1): consumed after the next epilogue
            }
        }
        if (WITH_GRAD && pending_g) {
            bar_g.wait();
            umma::fence_after_sync();
            if (!dead) {
                float f[QC];
                umma::tmem_ld16(tmem_acc + lane_off + quarter * QC, f);
#pragma unroll
                for (int i = 0; i < QC; ++i) acc[i] += f[i];
            }
        }
        // ---- partial results -------------------------------------------------------------------------------- //
        if (WITH_GRAD) {
            float* dst = p.acc_part + ((long long)part * p.Bp + gi) * DD + quarter * QC;
#pragma unroll
            for (int q = 0; q < QC / 4; ++q) *reinterpret_cast<float4*>(dst + q * 4) = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
        }
        if (WITH_GRAD) p.r_part[((long long)part * NQ + quarter) * p.Bp + gi] = (double)r_hi + (double)r_lo;
        if (!PASS_I) {
            const double loss_sum = rb_warp_sum_d((double)loss_hi + (double)loss_lo);
            if (lane == 0) red[warp] = loss_sum;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (!PASS_I && tid == 0) {
        double tsum = 0.0;
        for (int w = 0; w < NT_IB / 32; ++w) tsum += red[w];
        p.loss_part[blockIdx.x] = tsum;
    }
    if (warp == 0) umma::tmem_free(tmem, TMEM_COLS_IB);
}

// out_i = Σ_parts acc_part[part][i] − r_i · (other_i − c) + (q_i − r_i) · c   (q = r in pass U, which also stores r_i)
template <bool PASS_I>
__global__ void __launch_bounds__(256) inbatch_finish_kernel(const float* __restrict__ acc_part, const double* __restrict__ r_part,
                                                             float* __restrict__ r_total, const float* __restrict__ other,
                                                             const float* __restrict__ cen, float* __restrict__ out, int B, int Bp,
                                                             int split, int nq) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int row = idx >> 4, c4 = idx & 15;
    if (row >= B) return;
    // rs = row sum of this pass's G: r_i in pass U; the column sum q_j = Σ_i G_ij in pass I (whose diagonal term uses r_j)
    double rs = 0.0;
    for (int s = 0; s < split * nq; ++s) rs += r_part[(long long)s * Bp + row];
    const float q = (float)rs;
    float r = q;
    if (!PASS_I) { if (c4 == 0) r_total[row] = r; }
    else r = r_total[row];
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < split; ++s) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(acc_part + ((long long)s * Bp + row) * DD) + c4);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    float4 o = __ldg(reinterpret_cast<const float4*>(other + (long long)row * DD) + c4);
    const float4 cc = __ldg(reinterpret_cast<const float4*>(cen) + c4);
    o.x -= cc.x; o.y -= cc.y; o.z -= cc.z; o.w -= cc.w;
    a.x = fmaf(-r, o.x, a.x); a.y = fmaf(-r, o.y, a.y); a.z = fmaf(-r, o.z, a.z); a.w = fmaf(-r, o.w, a.w);
    if (PASS_I) {                  // Σ_i G_ij c = q_j c, while the diagonal term took r_j c
        const float dq = q - r;
        a.x = fmaf(dq, cc.x, a.x); a.y = fmaf(dq, cc.y, a.y); a.z = fmaf(dq, cc.z, a.z); a.w = fmaf(dq, cc.w, a.w);
    }
    reinterpret_cast<float4*>(out + (long long)row * DD)[c4] = a;
}

// err: the tensor-core kernels' time-out flag — a pipeline that gave up (wrong descriptor, heavy preemption) must not pass garbage
// on as a result: the loss becomes NaN
__global__ void inbatch_loss_kernel(const double* __restrict__ partials, int n, double scale, float* __restrict__ out, const int* __restrict__ err) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) t += partials[i];
    t = rb_warp_sum_d(t);
    if (threadIdx.x == 0) out[0] = (err && *err) ? __int_as_float(0x7fc00000) : (float)(t * scale);
}
// after the gradient pass: a time-out there also poisons the loss
__global__ void inbatch_poison_kernel(const int* __restrict__ err, float* __restrict__ loss) {
    if (*err) loss[0] = __int_as_float(0x7fc00000);
}

// column means of U (blockIdx.y = 0) and I (1): fixed-order partial sums over CM_BLOCKS row ranges, then one small block
constexpr int CM_BLOCKS = 64;
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ U, const float* __restrict__ I, int B,
                                                             float* __restrict__ part) {
    __shared__ float4 red[16][16];
    const float* X = blockIdx.y ? I : U;
    const int c4 = threadIdx.x & 15, rg = threadIdx.x >> 4;
    const int per = (B + CM_BLOCKS - 1) / CM_BLOCKS;
    const int r0 = blockIdx.x * per, r1 = min(B, r0 + per);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = r0 + rg; r < r1; r += 16) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(X + (long long)r * DD) + c4);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    red[rg][c4] = a;
    __syncthreads();
    if (rg == 0) {
        for (int k = 1; k < 16; ++k) { const float4 v = red[k][c4]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
        reinterpret_cast<float4*>(part + ((long long)blockIdx.y * CM_BLOCKS + blockIdx.x) * DD)[c4] = a;
    }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int B, float* __restrict__ cen) {   // <<<1, 128>>>
    const int which = threadIdx.x >> 6, d = threadIdx.x & 63;
    float a = 0.f;
    for (int k = 0; k < CM_BLOCKS; ++k) a += part[((long long)which * CM_BLOCKS + k) * DD + d];
    cen[which * DD + d] = a / (float)B;
}

// ================================================================================================================ //
// TMEM-operand, TMA-fed form of the same two passes (default; RB200_INBATCH_TS=0 selects the kernel above)
//
//   * the X tile [128 × 64] and the G tile [128 × 64] are the A operands of the two products and live in TMEM (thread = row =
//     TMEM lane: the epilogue writes G where the tensor core reads it — no shared-memory staging, no transposes);
//   * the streamed side comes from operand IMAGES built once per call by inbatch_images_kernel — per 64-row tile [Y hi | Y lo |
//     (Y−c)ᵀ hi | (Y−c)ᵀ lo], 64 KB, already split and in the K-major UMMA layout — fetched by bulk asynchronous copies (TMA
//     engine) into two 3-slot rings; no thread ever touches a Y element;
//   * warp-specialised: warp 8 issues copies and MMAs, warps 0-7 run the epilogue (row × 32 columns per thread); every hand-off
//     is an mbarrier.  Tensor queue: S(0) S(1) | GY(0) S(2) | GY(1) S(3) | … — the epilogue of tile t runs under GY(t−1) and
//     S(t+1), so the tensor pipe (48 MMAs = 1536 issue cycles per tile) is the critical path, not the B² epilogue.
//   TMEM (512 columns): X hi|lo [0,128) · S double-buffered [128,256) · G hi|lo [256,384) · acc [384,448).
// ================================================================================================================ //
constexpr int NT_EPI2 = 256, NT_TS = NT_EPI2 + 32;
constexpr int IMG_HALF = YT * DD * 4;            // 16 KB: one of hi / lo of a [64 × 64] operand
constexpr int IMG_PAIR = 2 * IMG_HALF;           // hi | lo
constexpr int IMG_TILE = 2 * IMG_PAIR;           // Y pair, then centred-transposed pair: 64 KB per 64-row tile
constexpr int NSLOT = 3;
constexpr uint32_t TS_X_HI = 0, TS_X_LO = 64, TS_S = 128, TS_G_HI = 256, TS_G_LO = 320, TS_ACC = 384, TS_COLS = 512;
constexpr int NQ2 = NT_EPI2 / 128;               // column halves of a tile
constexpr size_t IB_TS_SMEM = (size_t)2 * NSLOT * IMG_PAIR;      // 192 KB

// images of U (blockIdx.y = 0) and I (1): tile t → img[y][t]; rows past B are zero; cen == NULL: no centring (loss-only calls)
__global__ void __launch_bounds__(256) inbatch_images_kernel(const float* __restrict__ U, const float* __restrict__ I, int B,
                                                             const float* __restrict__ cen, unsigned char* __restrict__ img_u,
                                                             unsigned char* __restrict__ img_i) {
    const float* X = blockIdx.y ? I : U;
    unsigned char* img = (blockIdx.y ? img_i : img_u) + (size_t)blockIdx.x * IMG_TILE;
    const int tid = threadIdx.x, lane = tid & 31;
    const int mq = tid & 15, sq = tid >> 4;          // 4 d-columns (4·mq…) × 4 y-rows (4·sq…)
    const int y0 = blockIdx.x * YT;
    const float4 c = cen ? __ldg(reinterpret_cast<const float4*>(cen + blockIdx.y * DD) + mq) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v[4], w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = y0 + sq * 4 + i;
        const bool ok = r < B;
        v[i] = ok ? __ldg(reinterpret_cast<const float4*>(X + r * DD) + mq) : make_float4(0.f, 0.f, 0.f, 0.f);
        w[i] = ok ? make_float4(v[i].x - c.x, v[i].y - c.y, v[i].z - c.z, v[i].w - c.w) : make_float4(0.f, 0.f, 0.f, 0.f);
        put4<2>(img, img + IMG_HALF, YT, sq * 4 + i, mq * 4, v[i]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {                    // transposed, lane-rotated (conflict-free in shared memory; harmless here)
        const int e = (t + (lane >> 1)) & 3;
        put4<2>(img + IMG_PAIR, img + IMG_PAIR + IMG_HALF, DD, mq * 4 + e, sq * 4,
                make_float4(sel4(w[0], e), sel4(w[1], e), sel4(w[2], e), sel4(w[3], e)));
    }
}

template <int MODE>
__device__ __forceinline__ void st16_split(uint32_t hi_addr, uint32_t lo_addr, const float* x) {
    float hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { hi[i] = umma::tf32_hi(x[i]); lo[i] = x[i] - hi[i]; }
    umma::tmem_st16(hi_addr, hi);
    if (MODE == 2) umma::tmem_st16(lo_addr, lo);
}

// D[tmem] = A[tmem, 128 lanes × 64 columns: hi at a_hi, lo at a_lo] · B[64 × 64]ᵀ (shared-memory image: hi, then lo); one thread
template <int MODE>
__device__ __forceinline__ void issue_ts64(uint32_t d, uint32_t a_hi, uint32_t a_lo, const unsigned char* b_img) {
    const uint32_t idesc = umma::idesc_tf32(128, 64);
    constexpr uint32_t lbo = (64 / 8) * 128;
    const uint64_t dbh = umma::smem_desc(umma::smem_u32(b_img), lbo, 128), dbl = umma::smem_desc(umma::smem_u32(b_img + IMG_HALF), lbo, 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint64_t ob = (uint64_t)((2 * j * lbo) >> 4);
        if (MODE == 2) {
            umma::mma_tf32_ts(d, a_lo + 8 * j, dbh + ob, idesc, j > 0);
            umma::mma_tf32_ts(d, a_hi + 8 * j, dbl + ob, idesc, true);
            umma::mma_tf32_ts(d, a_hi + 8 * j, dbh + ob, idesc, true);
        } else {
            umma::mma_tf32_ts(d, a_hi + 8 * j, dbh + ob, idesc, j > 0);
        }
    }
}

struct IbTsParams {
    const float* X;                      // row-side embeddings [B × 64]
    const unsigned char* yimg;           // images of the streamed side (IMG_TILE per 64-row tile)
    const float* diag;
    int B, split, tiles_per_split, Bp;
    float g;
    float* acc_part; double* r_part; double* loss_part;
    int* err_flag;
};

template <int MODE, bool PASS_I, bool WITH_GRAD>
__global__ void __launch_bounds__(NT_TS, 1) inbatch_ts_kernel(const IbTsParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ybuf = smem;                              // [NSLOT][IMG_PAIR]   Y images (B operand of S = X·Yᵀ)
    unsigned char* ytbuf = smem + NSLOT * IMG_PAIR;          // [NSLOT][IMG_PAIR]   (Y−c)ᵀ images (B operand of acc = G·(Y−c))
    __shared__ __align__(8) uint64_t bar_yfull[NSLOT], bar_ytfull[NSLOT], bar_s[2], bar_gy, bar_e, bar_x;
    __shared__ uint32_t tmem_slot;
    __shared__ int dead;
    __shared__ double red[NT_EPI2 / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, TS_COLS);
    if (tid == NT_EPI2) {
        for (int i = 0; i < NSLOT; ++i) { umma::mbar_init(&bar_yfull[i], 1); umma::mbar_init(&bar_ytfull[i], 1); }
        umma::mbar_init(&bar_s[0], 1); umma::mbar_init(&bar_s[1], 1); umma::mbar_init(&bar_gy, 1);
        umma::mbar_init(&bar_e, NT_EPI2); umma::mbar_init(&bar_x, NT_EPI2);
        umma::fence_mbar_init();
        dead = 0;
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int B = p.B;
    const int xt = blockIdx.x / p.split, part = blockIdx.x - xt * p.split;
    const int x0 = xt * XT;
    const int t_begin = part * p.tiles_per_split;
    const int n_ytiles = (B + YT - 1) / YT;
    int t_end = t_begin + p.tiles_per_split;
    if (t_end > n_ytiles) t_end = n_ytiles;
    const int nt = t_end - t_begin;
    constexpr uint32_t Y_COPY = MODE == 2 ? IMG_PAIR : IMG_HALF;

    if (warp == NT_EPI2 / 32) {
        // ================================ copy + MMA issuer ================================================= //
        const unsigned char* img0 = p.yimg + (size_t)t_begin * IMG_TILE;
        if (umma::elect_one()) {
            for (int j = 0; j < NSLOT && j < nt; ++j) {
                umma::mbar_expect_tx(&bar_yfull[j], Y_COPY);
                umma::bulk_g2s(ybuf + j * IMG_PAIR, img0 + (size_t)j * IMG_TILE, Y_COPY, &bar_yfull[j]);
            }
            if (WITH_GRAD) {
                for (int j = 0; j < 2 && j < nt; ++j) {
                    umma::mbar_expect_tx(&bar_ytfull[j], Y_COPY);
                    umma::bulk_g2s(ytbuf + j * IMG_PAIR, img0 + (size_t)j * IMG_TILE + IMG_PAIR, Y_COPY, &bar_ytfull[j]);
                }
            }
        }
        __syncwarp();
        bool ok = umma::mbar_wait(&bar_x, 0u);                 // X tile staged in TMEM
        umma::fence_after_sync();
        for (int j = 0; j < 2 && j < nt && ok; ++j) {
            ok = umma::mbar_wait(&bar_yfull[j], 0u);
            if (ok && umma::elect_one()) {
                issue_ts64<MODE>(tmem + TS_S + j * YT, tmem + TS_X_HI, tmem + TS_X_LO, ybuf + j * IMG_PAIR);
                umma::commit(&bar_s[j]);
            }
            __syncwarp();
        }
        for (int i = 0; i < nt && ok; ++i) {
            // tile i consumed: S(i) read (and, with gradients, GY(i−1) drained and G(i) stored) — everything this iteration
            // overwrites or reads is implied by it
            ok = umma::mbar_wait(&bar_e, (uint32_t)(i & 1));
            if (!ok) break;
            umma::fence_after_sync();
            if (WITH_GRAD) {
                if (i + 2 < nt && umma::elect_one()) {
                    const int sl = (i + 2) % NSLOT;
                    umma::mbar_expect_tx(&bar_ytfull[sl], Y_COPY);
                    umma::bulk_g2s(ytbuf + sl * IMG_PAIR, img0 + (size_t)(i + 2) * IMG_TILE + IMG_PAIR, Y_COPY, &bar_ytfull[sl]);
                }
                __syncwarp();
                ok = umma::mbar_wait(&bar_ytfull[i % NSLOT], (uint32_t)((i / NSLOT) & 1));
                if (!ok) break;
                if (umma::elect_one()) {
                    issue_ts64<MODE>(tmem + TS_ACC, tmem + TS_G_HI, tmem + TS_G_LO, ytbuf + (i % NSLOT) * IMG_PAIR);
                    umma::commit(&bar_gy);
                }
                __syncwarp();
            }
            if (i + 2 < nt) {
                ok = umma::mbar_wait(&bar_yfull[(i + 2) % NSLOT], (uint32_t)(((i + 2) / NSLOT) & 1));
                if (!ok) break;
                if (umma::elect_one()) {
                    issue_ts64<MODE>(tmem + TS_S + (i & 1) * YT, tmem + TS_X_HI, tmem + TS_X_LO, ybuf + ((i + 2) % NSLOT) * IMG_PAIR);
                    umma::commit(&bar_s[i & 1]);
                }
                __syncwarp();
            }
            if (i + 3 < nt && umma::elect_one()) {
                const int sl = i % NSLOT;                      // held Y(i): S(i) completed before the epilogue of tile i read it
                umma::mbar_expect_tx(&bar_yfull[sl], Y_COPY);
                umma::bulk_g2s(ybuf + sl * IMG_PAIR, img0 + (size_t)(i + 3) * IMG_TILE, Y_COPY, &bar_yfull[sl]);
            }
            __syncwarp();
        }
        if (!ok && lane == 0) { dead = 1; if (p.err_flag) atomicOr(p.err_flag, 2); }
    } else {
        // ================================ epilogue warps ==================================================== //
        const int r_own = ((warp & 3) << 5) + lane, half = warp >> 2;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const int gi = x0 + r_own;                   // this thread's global X row
        {   // the X tile → TMEM (A operand of every S product of this CTA): this thread's row, 32 of its 64 columns
            float xv[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gi < B) v = __ldg(reinterpret_cast<const float4*>(p.X + (long long)gi * DD + half * 32) + q);
                xv[q * 4] = v.x; xv[q * 4 + 1] = v.y; xv[q * 4 + 2] = v.z; xv[q * 4 + 3] = v.w;
            }
            st16_split<MODE>(tmem + lane_off + TS_X_HI + half * 32, tmem + lane_off + TS_X_LO + half * 32, xv);
            st16_split<MODE>(tmem + lane_off + TS_X_HI + half * 32 + 16, tmem + lane_off + TS_X_LO + half * 32 + 16, xv + 16);
            umma::tmem_st_wait();
            umma::fence_before_sync();
            umma::mbar_arrive(&bar_x);
        }
        Bar w_s0{&bar_s[0], 0u, &dead, p.err_flag}, w_s1{&bar_s[1], 0u, &dead, p.err_flag}, w_gy{&bar_gy, 0u, &dead, p.err_flag};
        const float d_own = (!PASS_I && gi < B) ? __ldg(p.diag + gi) : 0.f;
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        float r_hi = 0.f, r_lo = 0.f, loss_hi = 0.f, loss_lo = 0.f;
        for (int i = 0; i < nt; ++i) {
            const int y0 = (t_begin + i) * YT;
            float dq[32];                            // S_jj of this thread's columns (pass I; diag is padded with zeros)
            if (PASS_I) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(p.diag + y0 + half * 32) + q);
                    dq[q * 4] = v.x; dq[q * 4 + 1] = v.y; dq[q * 4 + 2] = v.z; dq[q * 4 + 3] = v.w;
                }
            }
            if (i & 1) w_s1.wait(); else w_s0.wait();          // S(i) complete
            umma::fence_after_sync();
            float s[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) s[c] = 0.f;
            if (!dead) umma::tmem_ld32(tmem + lane_off + TS_S + (i & 1) * YT + half * 32, s);
            if (!WITH_GRAD) { umma::fence_before_sync(); umma::mbar_arrive(&bar_e); }
            float tile_loss = 0.f, tile_r = 0.f;
            const bool clean = y0 + YT <= B && x0 + XT <= B && (y0 + YT <= x0 || y0 >= x0 + XT);
            if (clean) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float prod = 1.f;
#pragma unroll
                    for (int c = h * 16; c < h * 16 + 16; ++c) {
                        const float x = s[c] - (PASS_I ? dq[c] : d_own);
                        const float ex = __expf(-fabsf(x));
                        const float inv = rcp_approx(1.f + ex);
                        s[c] = p.g * (x >= 0.f ? inv : ex * inv);
                        tile_r += s[c];
                        if (!PASS_I) { tile_loss += fmaxf(x, 0.f); prod *= 1.f + ex; }      // Π(1+e) ≤ 2¹⁶: one log per 16 scores
                    }
                    if (!PASS_I) tile_loss += __logf(prod);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int gj = y0 + half * 32 + c;
                    const float x = s[c] - (PASS_I ? dq[c] : d_own);
                    const bool valid = gi < B && gj < B && gi != gj;
                    const float ex = __expf(-fabsf(x));
                    const float inv = rcp_approx(1.f + ex);
                    s[c] = valid ? p.g * (x >= 0.f ? inv : ex * inv) : 0.f;
                    tile_r += s[c];
                    if (!PASS_I) tile_loss += valid ? fmaxf(x, 0.f) + __logf(1.f + ex) : 0.f;
                }
            }
            two_sum(r_hi, r_lo, tile_r);
            if (!PASS_I) two_sum(loss_hi, loss_lo, tile_loss);
            if (WITH_GRAD) {
                if (i > 0) {                         // GY(i−1) finished under this epilogue: drain it; the G columns are free again
                    w_gy.wait();
                    umma::fence_after_sync();
                    if (!dead) {
                        float f[32];
                        umma::tmem_ld32(tmem + lane_off + TS_ACC + half * 32, f);
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc[c] += f[c];
                    }
                }
                st16_split<MODE>(tmem + lane_off + TS_G_HI + half * 32, tmem + lane_off + TS_G_LO + half * 32, s);
                st16_split<MODE>(tmem + lane_off + TS_G_HI + half * 32 + 16, tmem + lane_off + TS_G_LO + half * 32 + 16, s + 16);
                umma::tmem_st_wait();
                umma::fence_before_sync();
                umma::mbar_arrive(&bar_e);           // G(i) ready → GY(i), S(i+2)
            }
        }
        if (WITH_GRAD && nt > 0) {
            w_gy.wait();
            umma::fence_after_sync();
            if (!dead) {
                float f[32];
                umma::tmem_ld32(tmem + lane_off + TS_ACC + half * 32, f);
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[c] += f[c];
            }
        }
        if (WITH_GRAD) {
            float* dst = p.acc_part + ((long long)part * p.Bp + gi) * DD + half * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(dst + q * 4) = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
            p.r_part[((long long)part * NQ2 + half) * p.Bp + gi] = (double)r_hi + (double)r_lo;
        }
        if (!PASS_I) {
            const double loss_sum = rb_warp_sum_d((double)loss_hi + (double)loss_lo);
            if (lane == 0) red[warp] = loss_sum;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (!PASS_I && tid == 0) {
        double tsum = 0.0;
        for (int w = 0; w < NT_EPI2 / 32; ++w) tsum += red[w];
        p.loss_part[blockIdx.x] = tsum;
    }
    if (warp == 0) umma::tmem_free(tmem, TS_COLS);
}

template <int MODE, bool PASS_I, bool WITH_GRAD>
int launch_ts(const IbTsParams& p, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(inbatch_ts_kernel<MODE, PASS_I, WITH_GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IB_TS_SMEM));
        attr_set = true;
    }
    inbatch_ts_kernel<MODE, PASS_I, WITH_GRAD><<<grid, NT_TS, IB_TS_SMEM, st>>>(p);
    RB_LAUNCH_CHECK("inbatch_ts_kernel");
    return RB200_OK;
}

struct IbPlan { int n_xt, n_yt, split, tps, Bp, grid; };
IbPlan plan(int B) {
    IbPlan pl;
    pl.n_xt = (B + XT - 1) / XT;
    pl.n_yt = (B + YT - 1) / YT;
    int split = rb_sm_count() / pl.n_xt;
    if (split < 1) split = 1;
    if (split > pl.n_yt) split = pl.n_yt;
    pl.tps = (pl.n_yt + split - 1) / split;
    pl.split = (pl.n_yt + pl.tps - 1) / pl.tps;      // no empty parts
    pl.Bp = pl.n_xt * XT;
    pl.grid = pl.n_xt * pl.split;
    return pl;
}

constexpr size_t IB_SMEM = 2 * X_BYTES + 2 * Y_BYTES + 2 * YT_BYTES + 2 * G_BYTES;     // 192 KB

template <int MODE, bool PASS_I, bool WITH_GRAD>
int launch(const IbParams& p, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(inbatch_tc_kernel<MODE, PASS_I, WITH_GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IB_SMEM));
        attr_set = true;
    }
    inbatch_tc_kernel<MODE, PASS_I, WITH_GRAD><<<grid, NT_ALL, IB_SMEM, st>>>(p);
    RB_LAUNCH_CHECK("inbatch_tc_kernel");
    return RB200_OK;
}

}  // namespace

size_t rb_inbatch_tc_workspace_bytes(int B, int D) {
    if (D != DD || B < 1) return 0;
    const IbPlan pl = plan(B);
    return 256 * 12 + sizeof(float) * ((size_t)2 * B + YT + 2 * DD + 2 * CM_BLOCKS * DD) + sizeof(int) + sizeof(float) * (size_t)pl.split * pl.Bp * DD +
           sizeof(double) * ((size_t)pl.split * NQ * pl.Bp + pl.grid) + (size_t)2 * pl.n_yt * IMG_TILE;
}

int rb_inbatch_tc(const float* U, const float* I, int B, int D, int mode, float* loss, float* dU, float* dI,
                  float grad_scale, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    RB_REQUIRE(mode == 1 || mode == 2, "bpr_inbatch: tensor-core modes are 1 (TF32) and 2 (3xTF32)");
    RB_REQUIRE(D == DD, "bpr_inbatch: the tcgen05 kernel is built for D=64 (got %d); use mode 0", D);
    RB_REQUIRE(B >= 2, "bpr_inbatch: B must be >= 2");
    const IbPlan pl = plan(B);
    RbArena ar(workspace, workspace_bytes);
    const int Bd = pl.n_yt * YT;                     // diag padded to whole Y tiles (zeros past B)
    float* diag = ar.take<float>(Bd);
    float* r_total = ar.take<float>(B);
    int* err = ar.take<int>(1);
    float* cen = ar.take<float>(2 * DD);              // column means: [0] of U, [1] of I
    float* cen_part = ar.take<float>((size_t)2 * CM_BLOCKS * DD);
    float* acc_part = ar.take<float>((size_t)pl.split * pl.Bp * DD);
    double* r_part = ar.take<double>((size_t)pl.split * NQ * pl.Bp);
    double* loss_part = ar.take<double>(pl.grid);
    unsigned char* img_u = ar.take<unsigned char>((size_t)pl.n_yt * IMG_TILE);
    unsigned char* img_i = ar.take<unsigned char>((size_t)pl.n_yt * IMG_TILE);
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "bpr_inbatch: workspace too small (%zu given)", workspace_bytes);
    static int use_ts = -1;                 // RB200_INBATCH_TS=0: the shared-memory-operand kernel of round 1
    if (use_ts < 0) { const char* e = getenv("RB200_INBATCH_TS"); use_ts = (e && atoi(e) == 0) ? 0 : 1; }
    rowdot64_kernel<<<(Bd + 7) / 8, 256, 0, st>>>(U, I, B, Bd, diag);
    RB_LAUNCH_CHECK("rowdot64_kernel");
    const double denom = (double)B * (double)(B - 1);
    IbParams p{};
    p.diag = diag; p.B = B; p.split = pl.split; p.tiles_per_split = pl.tps; p.g = (float)((double)grad_scale / denom);
    p.acc_part = acc_part; p.r_part = r_part; p.loss_part = loss_part; p.Bp = pl.Bp; p.err_flag = err;
    const bool grad = dU != nullptr;
    int rc;
    if (grad) {
        colsum_partial_kernel<<<dim3(CM_BLOCKS, 2), 256, 0, st>>>(U, I, B, cen_part);
        RB_LAUNCH_CHECK("colsum_partial_kernel");
        colsum_final_kernel<<<1, 128, 0, st>>>(cen_part, B, cen);
        RB_LAUNCH_CHECK("colsum_final_kernel");
    }
    const int fgrid = (B * 16 + 255) / 256;
    RB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
    if (use_ts) {
        inbatch_images_kernel<<<dim3(pl.n_yt, 2), 256, 0, st>>>(U, I, B, grad ? cen : nullptr, img_u, img_i);
        RB_LAUNCH_CHECK("inbatch_images_kernel");
        IbTsParams q{};
        q.diag = diag; q.B = B; q.split = pl.split; q.tiles_per_split = pl.tps; q.Bp = pl.Bp; q.g = p.g;
        q.acc_part = acc_part; q.r_part = r_part; q.loss_part = loss_part; q.err_flag = err;
        q.X = U; q.yimg = img_i;
        if (mode == 2) rc = grad ? launch_ts<2, false, true>(q, pl.grid, st) : launch_ts<2, false, false>(q, pl.grid, st);
        else rc = grad ? launch_ts<1, false, true>(q, pl.grid, st) : launch_ts<1, false, false>(q, pl.grid, st);
        if (rc) return rc;
        inbatch_loss_kernel<<<1, 32, 0, st>>>(loss_part, pl.grid, 1.0 / denom, loss, err);
        RB_LAUNCH_CHECK("inbatch_loss_kernel");
        if (!grad) return RB200_OK;
        inbatch_finish_kernel<false><<<fgrid, 256, 0, st>>>(acc_part, r_part, r_total, I, cen + DD, dU, B, pl.Bp, pl.split, NQ2);
        RB_LAUNCH_CHECK("inbatch_finish_kernel");
        q.X = I; q.yimg = img_u;
        rc = mode == 2 ? launch_ts<2, true, true>(q, pl.grid, st) : launch_ts<1, true, true>(q, pl.grid, st);
        if (rc) return rc;
        inbatch_finish_kernel<true><<<fgrid, 256, 0, st>>>(acc_part, r_part, r_total, U, cen, dI, B, pl.Bp, pl.split, NQ2);
        RB_LAUNCH_CHECK("inbatch_finish_kernel");
        inbatch_poison_kernel<<<1, 1, 0, st>>>(err, loss);
        RB_LAUNCH_CHECK("inbatch_poison_kernel");
        return RB200_OK;
    }
    p.X = U; p.Y = I; p.ycen = cen + DD;
    if (mode == 2) rc = grad ? launch<2, false, true>(p, pl.grid, st) : launch<2, false, false>(p, pl.grid, st);
    else rc = grad ? launch<1, false, true>(p, pl.grid, st) : launch<1, false, false>(p, pl.grid, st);
    if (rc) return rc;
    inbatch_loss_kernel<<<1, 32, 0, st>>>(loss_part, pl.grid, 1.0 / denom, loss, err);
    RB_LAUNCH_CHECK("inbatch_loss_kernel");
    if (!grad) return RB200_OK;
    inbatch_finish_kernel<false><<<fgrid, 256, 0, st>>>(acc_part, r_part, r_total, I, cen + DD, dU, B, pl.Bp, pl.split, NQ);
    RB_LAUNCH_CHECK("inbatch_finish_kernel");
    p.X = I; p.Y = U; p.ycen = cen;
    rc = mode == 2 ? launch<2, true, true>(p, pl.grid, st) : launch<1, true, true>(p, pl.grid, st);
    if (rc) return rc;
    inbatch_finish_kernel<true><<<fgrid, 256, 0, st>>>(acc_part, r_part, r_total, U, cen, dI, B, pl.Bp, pl.split, NQ);
    RB_LAUNCH_CHECK("inbatch_finish_kernel");
    inbatch_poison_kernel<<<1, 1, 0, st>>>(err, loss);
    RB_LAUNCH_CHECK("inbatch_poison_kernel");
    return RB200_OK;
}
