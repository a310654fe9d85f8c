// Exhaustive inner-product scan for MANY queries (more than 128; BASELINE cfg 5: 4096 queries against 12.5 M rows per GPU) — the round
// kernel of rb200_flat_search where the scan is bound by the tensor pipe.  One-pass TF32 FILTER + exact re-score of the survivors:
//
//   * only the hi images of rows and queries are multiplied (1 MMA per logical MMA instead of the 3 of 3xTF32): the accumulator holds
//     S̃ = Σ hi(q_i)·hi(x_i) with |S − S̃| ≤ (2ε + ε²)·Σ|q_i||x_i| ≤ 2⁻¹⁰·(1 + 2⁻¹²)·‖q‖‖x‖ (ε = 2⁻¹¹: round-to-nearest tf32 split; the fp32
//     accumulation error is four orders of magnitude smaller).  A score survives when S̃ > thr[q] − qmarg[q]·max‖x‖ with
//     qmarg[q] = 1.5·2⁻¹⁰·‖q‖ (flat_qimage_kernel) and max‖x‖ over the CTA's rows: no row that could beat the threshold is lost, a few per
//     cent more survive, and the host loop re-scores every survivor in fp32 (flat_rescore_kernel) before the select — same ids;
//   * database-stationary: a CTA owns 2 × 128 rows, written once into TMEM as the A operands (hi image, lane = row, column = k; TS-mode
//     MMAs: the tensor core reads only the query operand from shared memory), and the whole query set streams past in blocks of 128
//     queries — M128 × N128 × K8 MMAs, 8 per (tile, block) unit, from a pre-split image (2 MB at nq = 4096, L2-resident) by one bulk
//     copy per block into a 4-slot ring;
//   * THREE special warps.  ncu on the previous form (one warp issuing copies and the MMAs of both tiles, N = 64) showed the tensor pipe
//     30 % active with the epilogue warps waiting on the MMAs 41 % of the time and the issuing warp's samples spread evenly over its
//     ≈ 250 instructions per 64-query chunk: a single in-order warp needs ≈ 7 cycles per instruction, so the ISSUER bounded the kernel.
//     Now: warp 16 = copy producer, warps 17 and 18 = one MMA issuer per tile (≈ 100 instructions per 128-query unit each, half the MMAs
//     per query at N = 128), warps 0-15 = epilogue (8 per tile: TMEM lane quadrant × 64-query half);
//   * a ring of three accumulators [128 × 128] in TMEM (384 columns + 2 × 64 for the rows = 512) shared by the units in issue order;
//     hand-offs through mbarriers only (copy → MMA: expect_tx; MMA → epilogue and MMA → slot reuse: tcgen05.commit; epilogue → MMA: arrive);
//   * THE COMPARISON RUNS ON THE TENSOR CORE.  With the thresholds in shared memory the epilogue needed a load and a predicate-chained
//     compare per score (ncu: 40 % of the kernel's instructions, 1 200 issue cycles per 128-query block against 1 024 of MMAs).  Instead
//     every unit gets a ninth K step, an SS-mode MMA of a constant [128 × 8] A tile (n, 1, 1, 0 …; n = this CTA's largest row norm rounded
//     UP to tf32) with the k = 64 … 71 rows of the query image (m_q, −t_hi, −t_lo, 0 …; m_q = the query's margin rounded up, t_hi + t_lo =
//     its threshold, rewritten every round by flat_filter_ext_kernel): the accumulator holds S̃ + m_q·n − thr_q and a score survives iff
//     it is POSITIVE — a 3-input integer max tree over the raw bits (16 instructions per 32 scores, no dependency chain, no loads);
//   * survivors through per-warp shared-memory buffers (survivors.cuh).
// D = 64, 129 … 65 536 queries per call.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "survivors.cuh"
#include "umma.cuh"

namespace {

constexpr int VT = 128, QB = 128, DD = 64, TILES = 2, NSLOT = 4, NACC = 3, EW = 16;
// Two operand formats (template BF16).  TF32: rows and queries as round-to-nearest tf32 (ε = 2⁻¹¹), K = 8 per MMA.  BF16 (default): both
// rounded to bf16 (ε = 2⁻⁹: |S − S̃| ≤ 2⁻⁸·(1 + 2⁻¹⁰)·‖q‖‖x‖, margin 1.5·2⁻⁸·‖q‖·max‖x‖), K = 16 per MMA at twice the tf32 rate: half the
// tensor time per score for ≈ 20 % more survivors to re-score (the score density at the k-th best of 12.5 M unit rows is ≈ 16 k per unit
// score per query) — the RESULT is the same either way, every survivor is re-scored in fp32.
template <bool BF16> struct Fmt {
    static constexpr int ELT = BF16 ? 2 : 4;                     // bytes per operand element
    static constexpr int KS = BF16 ? DD / 16 : DD / 8;           // K steps per unit (then one more for the threshold)
    static constexpr int A_COLS = BF16 ? DD / 2 : DD;            // TMEM columns of one tile's rows
    static constexpr int Q_HI = QB * DD * ELT;                   // image of one 128-query block: 16 / 32 KB
    static constexpr int Q_SLOT = Q_HI + QB * 32;                // + its threshold K step (8 tf32 / 16 bf16 per query = 32 B): 4 KB
};
constexpr int Q_EXT = QB * 32;                   // threshold rows of a block, contiguous behind its image (K-major core matrices)
constexpr int Q_BLK = QB * DD * 8;               // stride of a block in the image workspace (64 KB, as flat_qimage_kernel's [hi | lo])
constexpr int MAX_Q = 65536;                     // survivors.cuh keeps the query in 16 bits
constexpr int CAPW = 256;
constexpr int NT_F = (EW + 3) * 32;
template <bool BF16> constexpr size_t smem_f() { return (size_t)NSLOT * Fmt<BF16>::Q_SLOT + (size_t)EW * WarpSurvivors<CAPW, false>::BYTES + Q_EXT; }

__device__ __forceinline__ uint32_t bf16_bits(float x) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint32_t bf16_pack(float lo, float hi) { return bf16_bits(lo) | (bf16_bits(hi) << 16); }
// round a non-negative float UP to a bf16-representable value → its 16 bits
__device__ __forceinline__ uint32_t bf16_up_bits(float x) { return (__float_as_uint(x) + 0xFFFFu) >> 16; }
constexpr float FILTER_C_BF16 = 1.5f / 256.f;

// bf16 image of the queries: block b (128 queries) = [128 × 64] K-major core matrices (a 16-byte chunk = 8 consecutive k), rows past nq
// zero; qmarg[r] = FILTER_C_BF16·‖q_r‖.  Thread = (row, chunk of 8 k); grid covers nq_pad·8 exactly.
__global__ void __launch_bounds__(256) flat_filter_image_bf16_kernel(const float* __restrict__ q, int nq, unsigned char* __restrict__ qimg,
                                                                     float* __restrict__ qmarg) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int r = i >> 3, g = i & 7;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (r < nq) {
        a = __ldg(reinterpret_cast<const float4*>(q + (long long)r * DD) + 2 * g);
        b = __ldg(reinterpret_cast<const float4*>(q + (long long)r * DD) + 2 * g + 1);
    }
    float ss = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, a.w * a.w))) + fmaf(b.x, b.x, fmaf(b.y, b.y, fmaf(b.z, b.z, b.w * b.w)));
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (g == 0) qmarg[r] = FILTER_C_BF16 * sqrtf(ss) * 1.0001f;
    const int rr = r % QB;
    unsigned char* dst = qimg + (size_t)(r / QB) * Q_BLK + (uint32_t)(((g * (QB / 8) + (rr >> 3)) << 7) + ((rr & 7) << 4));
    *reinterpret_cast<uint4*>(dst) = make_uint4(bf16_pack(a.x, a.y), bf16_pack(a.z, a.w), bf16_pack(b.x, b.y), bf16_pack(b.z, b.w));
}

// round a non-negative float UP to a tf32-representable value (the tensor core ignores the low 13 mantissa bits)
__device__ __forceinline__ float tf32_up(float x) { return __uint_as_float((__float_as_uint(x) + 0x1FFFu) & 0xFFFFE000u); }

// per round: the threshold K step of every query's image = (margin rounded up, −thr in 2 tf32 / 3 bf16 pieces, 0 …).  Padding queries
// carry thr = FLT_MAX: −1e30 keeps their accumulators negative without an infinity inside the tensor core.
template <bool BF16>
__global__ void __launch_bounds__(256) flat_filter_ext_kernel(const float* __restrict__ thr, const float* __restrict__ qmarg, int nq_pad,
                                                              unsigned char* __restrict__ qimg) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= nq_pad) return;
    const float t = fminf(__ldg(thr + r), 1e30f);
    unsigned char* ext = qimg + (size_t)(r / QB) * Q_BLK + Fmt<BF16>::Q_HI;
    const int rr = r % QB;
    const uint32_t off = (uint32_t)(((rr >> 3) << 7) + ((rr & 7) << 4));
    if (BF16) {
        const float t1 = __bfloat162float(__float2bfloat16_rn(t)), t2 = __bfloat162float(__float2bfloat16_rn(t - t1));
        const float t3 = (t - t1) - t2;
        *reinterpret_cast<uint4*>(ext + off) = make_uint4(bf16_up_bits(__ldg(qmarg + r)) | (bf16_bits(-t1) << 16), bf16_pack(-t2, -t3), 0u, 0u);
        *reinterpret_cast<uint4*>(ext + (QB / 8) * 128 + off) = make_uint4(0u, 0u, 0u, 0u);
    } else {
        const float th = umma::tf32_hi(t), tl = umma::tf32_hi(t - th);
        *reinterpret_cast<float4*>(ext + off) = make_float4(tf32_up(__ldg(qmarg + r)), -th, -tl, 0.f);
        *reinterpret_cast<float4*>(ext + (QB / 8) * 128 + off) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <bool BF16>
__global__ void __launch_bounds__(NT_F, 1)
flat_filter_tc_kernel(const float* __restrict__ x, long long n_rows, const unsigned char* __restrict__ qimg, int n_blocks,
                      const float* __restrict__ thr, const float* __restrict__ qmarg, int* __restrict__ count,
                      float* __restrict__ cand_s, long long stride, int kprev, int* __restrict__ cand_r, int cap, int* __restrict__ flags) {
    using F = Fmt<BF16>;
    constexpr int Q_SLOT = F::Q_SLOT, A_COLS = F::A_COLS;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* qbuf = smem;                                  // [NSLOT][Q_SLOT]
    unsigned char* surv_mem = qbuf + NSLOT * Q_SLOT;             // [EW][WarpSurvivors<CAPW, false>::BYTES]
    unsigned char* a_ext = surv_mem + EW * WarpSurvivors<CAPW, false>::BYTES;      // [128 × 8] A tile of the threshold K step (K-major core matrices)
    __shared__ __align__(8) uint64_t bar_qfull[NSLOT], bar_qfree[NSLOT], bar_done[TILES][NACC], bar_free[TILES][NACC];
    __shared__ uint32_t tmem_slot;
    __shared__ float nrm_s[TILES * 4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long row0 = (long long)blockIdx.x * (TILES * VT);
    constexpr uint32_t ACC_COLS = NACC * QB, TM_COLS = 512;      // then per tile the image of its rows, 64 (tf32) / 32 (bf16 pairs) columns
    static_assert(ACC_COLS + TILES * A_COLS <= TM_COLS, "TMEM budget");

    if (warp == 0) umma::tmem_alloc(&tmem_slot, TM_COLS);
    if (tid == EW * 32) {
        for (int i = 0; i < NSLOT; ++i) { umma::mbar_init(&bar_qfull[i], 1); umma::mbar_init(&bar_qfree[i], TILES); }
        for (int i = 0; i < NACC; ++i) {
            umma::mbar_init(&bar_done[0][i], 1); umma::mbar_init(&bar_done[1][i], 1);
            umma::mbar_init(&bar_free[0][i], 256); umma::mbar_init(&bar_free[1][i], 256);
        }
        umma::fence_mbar_init();
        for (int i = 0; i < NSLOT && i < n_blocks; ++i) {        // the first query blocks are on their way while the rows are staged
            umma::mbar_expect_tx(&bar_qfull[i], Q_SLOT);
            umma::bulk_g2s(qbuf + i * Q_SLOT, qimg + (size_t)i * Q_BLK, Q_SLOT, &bar_qfull[i]);
        }
    }
    // ---- rows → TMEM (warps 0-7: thread = one row, 16 independent 16-byte loads, hi image, 4 tcgen05.st of 16 columns) ---------- //
    const bool stager = warp < TILES * 4;
    const long long srow = row0 + (warp >> 2) * VT + ((warp & 3) << 5) + lane;
    float4 vv[DD / 4];
    if (stager) {
#pragma unroll
        for (int i = 0; i < DD / 4; ++i)
            vv[i] = srow < n_rows ? __ldcs(reinterpret_cast<const float4*>(x + srow * DD) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float ss[4] = {0.f, 0.f, 0.f, 0.f};                      // max ‖x‖ over the CTA's rows: the unit of the queries' margins
#pragma unroll
        for (int i = 0; i < DD / 4; ++i)
            ss[i & 3] = fmaf(vv[i].x, vv[i].x, fmaf(vv[i].y, vv[i].y, fmaf(vv[i].z, vv[i].z, fmaf(vv[i].w, vv[i].w, ss[i & 3]))));
        const uint32_t mx = __reduce_max_sync(0xffffffffu, __float_as_uint((ss[0] + ss[1]) + (ss[2] + ss[3])));     // ≥ 0: bits order like values
        if (lane == 0) nrm_s[warp] = sqrtf(__uint_as_float(mx)) * 1.0001f;
    }
    umma::fence_before_sync();
    __syncthreads();                                             // TMEM base address, barriers and norms published
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (stager) {
        const uint32_t a_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ACC_COLS + (uint32_t)(warp >> 2) * A_COLS;
        if (BF16) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {                        // 16 columns = 32 consecutive k, two per column
                float pk[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v = vv[g * 8 + i];
                    pk[2 * i] = __uint_as_float(bf16_pack(v.x, v.y));
                    pk[2 * i + 1] = __uint_as_float(bf16_pack(v.z, v.w));
                }
                umma::tmem_st16(a_base + g * 16, pk);
            }
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float hi[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 h, l;
                    umma::split4(vv[g * 4 + i], h, l);
                    hi[4 * i] = h.x; hi[4 * i + 1] = h.y; hi[4 * i + 2] = h.z; hi[4 * i + 3] = h.w;
                }
                umma::tmem_st16(a_base + g * 16, hi);
            }
        }
        umma::tmem_st_wait();
    }
    if (tid < VT) {                                              // A tile of the threshold K step: row r = (max ‖x‖ of this CTA, 1, 1 (, 1), 0 …)
        float nxmax = 0.f;
#pragma unroll
        for (int i = 0; i < TILES * 4; ++i) nxmax = fmaxf(nxmax, nrm_s[i]);
        const uint32_t off = (uint32_t)(((tid >> 3) << 7) + ((tid & 7) << 4));
        if (BF16) {
            *reinterpret_cast<uint4*>(a_ext + off) = make_uint4(bf16_up_bits(nxmax) | (0x3F80u << 16), 0x3F803F80u, 0u, 0u);
            *reinterpret_cast<uint4*>(a_ext + (VT / 8) * 128 + off) = make_uint4(0u, 0u, 0u, 0u);
        } else {
            *reinterpret_cast<float4*>(a_ext + off) = make_float4(tf32_up(nxmax), 1.f, 1.f, 0.f);
            *reinterpret_cast<float4*>(a_ext + (VT / 8) * 128 + off) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();                                             // rows in TMEM, threshold tile staged
    umma::fence_after_sync();

    if (warp == EW) {
        // ================================ copy producer ================================ //
        bool ok = true;
        for (int c = NSLOT; c < n_blocks && ok; ++c) {
            const int sl = c % NSLOT;
            ok = umma::mbar_wait(&bar_qfree[sl], ((c / NSLOT) - 1) & 1);
            if (!ok && lane == 0) atomicCAS(flags + 2, 0, (1 << 24) | (warp << 16) | c);
            if (ok && umma::elect_one()) {
                umma::mbar_expect_tx(&bar_qfull[sl], Q_SLOT);
                umma::bulk_g2s(qbuf + sl * Q_SLOT, qimg + (size_t)c * Q_BLK, Q_SLOT, &bar_qfull[sl]);
            }
            __syncwarp();
        }
        if (!ok && lane == 0) atomicOr(flags + 1, 4);
    } else if (warp > EW) {
        // ================================ MMA issuer of tile t (a converged warp, one elected lane) ================================ //
        const int t = warp - EW - 1;
        const uint32_t idesc = BF16 ? umma::idesc_bf16(VT, QB) : umma::idesc_tf32(VT, QB);
        constexpr uint32_t lbo_b = (QB / 8) * 128;
        const uint32_t q_s = umma::smem_u32(qbuf), a_t = tmem + ACC_COLS + (uint32_t)t * A_COLS;
        const uint64_t da_ext = umma::smem_desc(umma::smem_u32(a_ext), (VT / 8) * 128, 128);
        // Unit u = 2c + t uses accumulator a = u % 3 for the k-th time, k = u / 3; the uses of an accumulator alternate between the two
        // tiles.  Its previous use (k − 1, the OTHER tile's) is released on bar_free[1 − t][a], a barrier that counts only that tile's
        // releases: the ((k − 1) / 2)-th; completions go to bar_done[t][a], this tile's (k / 2)-th.  With ONE barrier per accumulator
        // for both tiles every waiter would see only every other phase, and a one-bit parity cannot tell phase k from phase k − 2:
        // a warp running ahead passes the wait two uses early (seen as wrong scores and dead-locks while this kernel was written).
        int a = t, k = 0;
        bool ok = true;
        for (int c = 0; c < n_blocks; ++c) {
            const int sl = c % NSLOT;
            ok = umma::mbar_wait(&bar_qfull[sl], (c / NSLOT) & 1);
            if (!ok && lane == 0) atomicCAS(flags + 2, 0, (2 << 24) | (warp << 16) | c);
            if (ok && k >= 1) {
                ok = umma::mbar_wait(&bar_free[1 - t][a], ((k - 1) >> 1) & 1);
                if (!ok && lane == 0) atomicCAS(flags + 2, 0, (3 << 24) | (warp << 16) | c);
            }
            if (!ok) break;
            umma::fence_after_sync();
            if (umma::elect_one()) {
                const uint64_t db = umma::smem_desc(q_s + sl * Q_SLOT, lbo_b, 128);
                const uint32_t acc = tmem + (uint32_t)a * QB;
                // a K step (8 tf32 / 16 bf16) = 8 TMEM columns of the rows and two 16-byte core-matrix columns of the image
#pragma unroll
                for (int j = 0; j < F::KS; ++j) {
                    if (BF16) umma::mma_bf16_ts(acc, a_t + 8 * j, db + (uint64_t)((2 * j * lbo_b) >> 4), idesc, j > 0);
                    else umma::mma_tf32_ts(acc, a_t + 8 * j, db + (uint64_t)((2 * j * lbo_b) >> 4), idesc, j > 0);
                }
                if (BF16) umma::mma_bf16(acc, da_ext, db + (uint64_t)((2 * F::KS * lbo_b) >> 4), idesc, true);     // + m_q·n − thr_q
                else umma::mma_tf32(acc, da_ext, db + (uint64_t)((2 * F::KS * lbo_b) >> 4), idesc, true);
                umma::commit(&bar_done[t][a]);                   // → epilogue of this unit
                umma::commit(&bar_qfree[sl]);                    // → (with the other tile's) slot sl may be refilled
            }
            __syncwarp();
            a += 2;
            if (a >= NACC) { a -= NACC; ++k; }
        }
        if (!ok && lane == 0) atomicOr(flags + 1, 1);
    } else {
        // ================================ epilogue warps (8 per tile) ================================ //
        const int t = warp >> 3, wq = warp & 7;
        const int r_own = ((wq & 3) << 5) + lane, half = wq >> 2;
        const uint32_t lane_off = (uint32_t)((wq & 3) * 32) << 16;
        const long long row = row0 + t * VT + r_own;
        WarpSurvivors<CAPW, false> surv;
        surv.init(surv_mem + warp * WarpSurvivors<CAPW, false>::BYTES, lane);
        int a = t, k = 0;
        for (int c = 0; c < n_blocks; ++c) {
            if (!umma::mbar_wait(&bar_done[t][a], (k >> 1) & 1)) {
                atomicOr(flags + 1, 2);
                if (lane == 0) atomicCAS(flags + 2, 0, (4 << 24) | (warp << 16) | c);
                break;
            }
            umma::fence_after_sync();
            // both 32-query blocks into registers, then the accumulator goes straight back to the issuer: everything below works on
            // registers (survivors keep (query, row) only), so the MMAs of the unit three later run under this epilogue
            const uint32_t acc = tmem + lane_off + (uint32_t)a * QB + half * 64;
            float s0[32], s1[32];
            umma::tmem_ld32(acc, s0);
            umma::tmem_ld32(acc + 32, s1);
            umma::fence_before_sync();
            umma::mbar_arrive(&bar_free[t][a]);                  // the accumulator may be overwritten by the unit three later (the other tile's)
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                const int q0 = c * QB + half * 64 + blk * 32;
                const float (&s)[32] = blk ? s1 : s0;
                // survive ⇔ accumulator > 0 ⇔ its bit pattern is a positive integer: a 3-input max tree decides whether the warp has any
                // survivor at all; only then is the per-score mask built and handed to the warp's survivor buffer
                int mx[11];
#pragma unroll
                for (int i = 0; i < 10; ++i) mx[i] = __vimax3_s32(__float_as_int(s[3 * i]), __float_as_int(s[3 * i + 1]), __float_as_int(s[3 * i + 2]));
                mx[10] = max(__float_as_int(s[30]), __float_as_int(s[31]));
                const int m3a = __vimax3_s32(mx[0], mx[1], mx[2]), m3b = __vimax3_s32(mx[3], mx[4], mx[5]), m3c = __vimax3_s32(mx[6], mx[7], mx[8]);
                const int top = __vimax3_s32(__vimax3_s32(m3a, m3b, m3c), mx[9], mx[10]);
                if (__any_sync(0xffffffffu, top > 0 && row < n_rows)) {
                    uint32_t m = 0;                              // bit j = score j positive: the sign bit of −bits, shifted in from the top (two
#pragma unroll                                                   // instructions per score; −0.0 counts as positive: one more row to re-score)
                    for (int j = 31; j >= 0; --j) m = __funnelshift_l((uint32_t)(-__float_as_int(s[j])), m, 1);
                    if (row >= n_rows) m = 0;
                    surv.add_block(m, 0u, q0, (int)row, count, cand_s, stride, kprev, cand_r, cap, flags);
                }
            }
            a += 2;
            if (a >= NACC) { a -= NACC; ++k; }
        }
        surv.flush(count, cand_s, stride, kprev, cand_r, cap, flags);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, TM_COLS);
}

}  // namespace

static bool filter_bf16() {
    static int on = -1;                      // RB200_FLAT_FILTER_TF32=1: the tf32 form of the filter (tuning / testing knob)
    if (on < 0) { const char* e = getenv("RB200_FLAT_FILTER_TF32"); on = (e && atoi(e)) ? 0 : 1; }
    return on != 0;
}

// does a round over nq queries use the one-pass filter (its survivors then need rb_flat_rescore)?  RB200_FLAT_FILTER=0: 3xTF32 scores
// straight from flat_scan_tc_kernel (testing knob).
bool rb_flat_filtered(int nq) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("RB200_FLAT_FILTER"); on = e ? atoi(e) : 1; }
    return on && nq > 128 && nq <= MAX_Q;
}

int rb_flat_qimage(const float* q, int nq, int nq_pad, int block_rows, unsigned char* qimg, float* qmarg, cudaStream_t st);

// the filter's query image (128-query blocks) and the queries' margins; nq_pad a multiple of 128
int rb_flat_filter_image(const float* q, int nq, int nq_pad, unsigned char* qimg, float* qmarg, cudaStream_t st) {
    if (!filter_bf16()) return rb_flat_qimage(q, nq, nq_pad, 128, qimg, qmarg, st);
    RB_REQUIRE(nq_pad % 128 == 0 && nq_pad >= nq, "flat_filter_image: bad padding");
    flat_filter_image_bf16_kernel<<<nq_pad * 8 / 256, 256, 0, st>>>(q, nq, qimg, qmarg);
    RB_LAUNCH_CHECK("flat_filter_image_bf16_kernel");
    return RB200_OK;
}

// start of a filtered round (after thr has been written): the threshold rows of the query image
int rb_flat_filter_thresholds(const float* thr, const float* qmarg, int nq_pad, unsigned char* qimg, cudaStream_t st) {
    if (filter_bf16()) flat_filter_ext_kernel<true><<<(nq_pad + 255) / 256, 256, 0, st>>>(thr, qmarg, nq_pad, qimg);
    else flat_filter_ext_kernel<false><<<(nq_pad + 255) / 256, 256, 0, st>>>(thr, qmarg, nq_pad, qimg);
    RB_LAUNCH_CHECK("flat_filter_ext_kernel");
    return RB200_OK;
}

template <bool BF16>
static int launch_filter(const float* x, long long n_rows, const unsigned char* qimg, int n_blocks, const float* thr, const float* qmarg, int* count,
                         float* cand_s, long long stride, int kprev, int* cand_r, int cap, int* flags, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(flat_filter_tc_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f<BF16>()));
        attr_set = true;
    }
    const long long n_cta = (n_rows + TILES * VT - 1) / (TILES * VT);
    flat_filter_tc_kernel<BF16><<<(unsigned)n_cta, NT_F, smem_f<BF16>(), st>>>(x, n_rows, qimg, n_blocks, thr, qmarg, count, cand_s, stride, kprev,
                                                                             cand_r, cap, flags);
    RB_LAUNCH_CHECK("flat_filter_tc_kernel");
    return RB200_OK;
}

// one round: rows [0, n_rows) of x (the caller offsets x) against n_blocks·128 queries (image: rb_flat_filter_image + this round's
// threshold rows: rb_flat_filter_thresholds), count [n_blocks·128]
int rb_flat_filter_tc(const float* x, long long n_rows, const unsigned char* qimg, int n_blocks, const float* thr, const float* qmarg, int* count,
                      float* cand_s, long long stride, int kprev, int* cand_r, int cap, int* flags, cudaStream_t st) {
    RB_REQUIRE(n_rows >= 1 && n_rows < (1ll << 31) && n_blocks >= 1 && n_blocks * QB <= MAX_Q, "flat_filter: 1..2^31 rows, at most 65536 queries");
    if (filter_bf16()) return launch_filter<true>(x, n_rows, qimg, n_blocks, thr, qmarg, count, cand_s, stride, kprev, cand_r, cap, flags, st);
    return launch_filter<false>(x, n_rows, qimg, n_blocks, thr, qmarg, count, cand_s, stride, kprev, cand_r, cap, flags, st);
}
