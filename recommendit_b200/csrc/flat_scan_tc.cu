// Exhaustive inner-product scan with threshold pruning on the tensor cores (faiss IndexFlatIP.search at BASELINE cfg 5 sizes:
// 12.5 M rows per GPU × thousands of queries).
//
// The chunked path in ivf.cu (score GEMM → [nq × chunk] scores in HBM → select) writes and re-reads every score: 205 GB of
// traffic per 4096-query batch over a 12.5 M-row shard.  Here a score leaves the SM only if it can still enter its query's
// top-k: the caller passes, per query, the k-th best score over the rows seen so far (`thr`), and the epilogue appends
// (score, row) to the query's survivor list only when score > thr (strictly: rows are visited in ascending order, so an equal
// score loses the tie to the earlier row, as in the reference's heap).  The host loop in ivf.cu grows the row range
// geometrically (×4 per round), so a round is expected to leave 3·k survivors per query whatever the database size.
//
// Database-stationary: one CTA owns TILES × 128 database rows, staged once as K-major A operands (hi/lo split: 3xTF32 = fp32-grade
// products, same arithmetic as the IVF list scan), and streams the whole query set past them in chunks of 64 queries.  The
// query chunks come from a pre-split image (`qimg`, built once per search by flat_qimage_kernel; 2 MB at nq = 4096, L2-resident)
// by ONE bulk asynchronous copy (TMA engine) per chunk into a ring of NSLOT shared-memory slots; S[128 rows × 64 queries] per
// tile is accumulated in TMEM, double-buffered.  Warp-specialised: warp 8 issues the copies and the MMAs (one thread), warps 0-7
// stage the rows once and then run the epilogues — the MMAs of chunk c+1 run under the epilogue of chunk c, and hand-offs go
// through mbarriers only (copy → MMA: expect_tx; MMA → epilogue and MMA → slot reuse: tcgen05.commit; epilogue → MMA: arrive).
// FILTER (the default for more than two query chunks): a ONE-pass TF32 filter instead of 3xTF32.  Only the hi images of rows and
// queries are multiplied (1 MMA per logical MMA instead of 3), so the accumulator holds S̃ = Σ hi(q_i)·hi(x_i) with
// |S − S̃| ≤ (2ε + ε²)·Σ|q_i||x_i| ≤ 2⁻¹⁰·(1 + 2⁻¹²)·‖q‖‖x‖ (ε = 2⁻¹¹: round-to-nearest tf32 split; fp32 accumulation error is four
// orders of magnitude smaller).  A score survives when S̃ > thr[q] − qmarg[q]·max‖x‖ with qmarg[q] = 1.5·2⁻¹⁰·‖q‖ (flat_qimage_kernel)
// and max‖x‖ taken over the CTA's rows: no row that could beat the threshold is lost, a few per cent more survive (the score density at
// the k-th best of N rows is ≈ k·z/(N·σ) per unit score), and the host loop re-scores every survivor in fp32 (flat_rescore_kernel)
// before the select — same ids, a third of the MMAs.  With a third of the tensor work per unit the epilogue gets one warp set PER TILE
// (16 epilogue warps at TILES = 2) so that a warp sees every other unit.
// D = 64.  <TILES 2, NSLOT 3>: 224 KB of shared memory, 256 TMEM columns, 1 CTA per SM (many query chunks: tensor-bound);
// <TILES 1, NSLOT 1>: 96 KB, 2 CTAs per SM (one or two query chunks: the scan is bound by reading the rows).
#include <stdlib.h>

#include "common.cuh"
#include "survivors.cuh"
#include "umma.cuh"

namespace {

constexpr int VT = 128, QT = 64, DD = 64;
constexpr int V_BYTES = VT * DD * 4;             // one half (hi or lo) of one tile: 32 KB
constexpr int Q_HALF = QT * DD * 4;              // 16 KB
constexpr int Q_IMG = 2 * Q_HALF;                // hi image then lo image of one 64-query chunk: 32 KB
constexpr int TH_SM_CHUNKS = 128;                // the filter stages the thresholds of up to 8192 queries in shared memory (32 KB)

// q [nq, 64] fp32 → per chunk of 64 queries [hi | lo], each in the K-major core-matrix layout of umma.cuh; rows past nq are zero
// qmarg[r] = FILTER_C·‖q_r‖ (0 for the padding rows): the filter's per-query safety margin per unit of row norm
constexpr float FILTER_C = 1.5f / 1024.f;
// blk = rows per image block: 64 (flat_scan_tc / flat_stream_tc) or 128 (flat_filter_tc): block b = [hi | lo] of rows [b·blk, (b+1)·blk)
__global__ void __launch_bounds__(256) flat_qimage_kernel(const float* __restrict__ q, int nq, int blk, unsigned char* __restrict__ qimg,
                                                          float* __restrict__ qmarg) {
    const int i = blockIdx.x * 256 + threadIdx.x;                // (row, 16-byte chunk); grid covers the padded rows × 16 exactly
    const int r = i >> 4, c4 = i & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nq) v = __ldg(reinterpret_cast<const float4*>(q + (long long)r * DD) + c4);
    float ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (c4 == 0) qmarg[r] = FILTER_C * sqrtf(ss) * 1.0001f;
    float4 hi, lo;
    umma::split4(v, hi, lo);
    const size_t half_bytes = (size_t)blk * DD * 4;
    unsigned char* base = qimg + (size_t)(r / blk) * 2 * half_bytes;
    const uint32_t off = umma::kmajor_offset(blk, r % blk, c4 * 4);
    *reinterpret_cast<float4*>(base + off) = hi;
    if (blk == 64) *reinterpret_cast<float4*>(base + half_bytes + off) = lo;      // 128-row blocks (the filter): hi only, threshold rows behind it
}

template <int TILES, int NSLOT, bool A_TMEM, bool FILTER>
__global__ void __launch_bounds__((FILTER ? 8 * TILES : 8) * 32 + 32, 1)
flat_scan_tc_kernel(const float* __restrict__ x, long long n_rows, const unsigned char* __restrict__ qimg, int n_chunks,
                    const float* __restrict__ thr, const float* __restrict__ qmarg, int* __restrict__ count,
                    float* __restrict__ cand_s, long long stride, int kprev, int* __restrict__ cand_r, int cap, int* __restrict__ flags) {
    static_assert(!FILTER || A_TMEM, "the one-pass filter keeps the rows in TMEM");
    constexpr int EW = FILTER ? 8 * TILES : 8;                   // epilogue warps (FILTER: one set of 8 per tile); warp EW = copy + MMA issuer
    constexpr int QSLOT = FILTER ? Q_HALF : Q_IMG;               // the filter multiplies the hi images only
    constexpr int A_COLS = FILTER ? DD : 2 * DD;                 // TMEM columns of one tile's row image(s)
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* v_hi = smem;                                  // [TILES][V_BYTES]
    unsigned char* v_lo = smem + TILES * V_BYTES;
    unsigned char* qbuf = smem + (A_TMEM ? 0 : 2 * TILES * V_BYTES);      // [NSLOT][QSLOT]
    constexpr int CAPW = TILES == 1 ? 128 : 256;                 // survivor buffer entries per epilogue warp (survivors.cuh)
    unsigned char* surv_mem = qbuf + NSLOT * QSLOT;              // [EW][WarpSurvivors<CAPW>::BYTES]
    float* th_sm = reinterpret_cast<float*>(surv_mem + EW * WarpSurvivors<CAPW>::BYTES);     // FILTER: [n_chunks·64] effective thresholds
    constexpr int NBUF = FILTER ? 3 : 2;                         // accumulator buffers per tile (the filter's epilogue is the critical path)
    __shared__ __align__(8) uint64_t bar_qfull[NSLOT], bar_qfree[NSLOT], bar_m[NBUF * TILES], bar_accfree[NBUF * TILES];   // per (buffer, tile) unit
    __shared__ uint32_t tmem_slot;
    __shared__ int dead;
    __shared__ float nrm_s[TILES * 4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long row0 = (long long)blockIdx.x * (TILES * VT);
    // TMEM columns: [NBUF buffers][TILES] accumulators of 64, then (A_TMEM) per tile the hi (and the lo) image of the rows, 64 columns each
    constexpr uint32_t ACC_COLS = NBUF * TILES * QT, USED_COLS = A_TMEM ? ACC_COLS + TILES * A_COLS : ACC_COLS;
    constexpr uint32_t TM_COLS = USED_COLS <= 128 ? 128 : USED_COLS <= 256 ? 256 : 512;
    static_assert(USED_COLS <= 512, "TMEM budget");

    if (warp == 0) umma::tmem_alloc(&tmem_slot, TM_COLS);
    if (tid == EW * 32) {
        for (int i = 0; i < NSLOT; ++i) { umma::mbar_init(&bar_qfull[i], 1); umma::mbar_init(&bar_qfree[i], 1); }
        for (int i = 0; i < NBUF * TILES; ++i) { umma::mbar_init(&bar_m[i], 1); umma::mbar_init(&bar_accfree[i], 256); }
        umma::fence_mbar_init();
        dead = 0;
        for (int i = 0; i < NSLOT && i < n_chunks; ++i) {        // the first query chunks are on their way while the rows are staged
            umma::mbar_expect_tx(&bar_qfull[i], QSLOT);
            umma::bulk_g2s(qbuf + i * QSLOT, qimg + (size_t)i * Q_IMG, QSLOT, &bar_qfull[i]);
        }
    }
    // ---- stage the database rows (warps 0-7): all loads of a thread are issued before the first is used.  A warp covers 8 rows ×
    //      4 16-byte chunks per step (64-byte global segments, conflict-free 128-byte shared-memory phases) ------------------- //
    if (A_TMEM) {
        // rows go to TMEM as the A operand (lane = row, column = k): the tensor core then reads only the query chunk from shared
        // memory (with both operands there an M128×N64×K8 MMA needs 6 KB = 48 cycles of shared-memory bandwidth for 32 cycles of
        // math).  Thread = one row: 16 independent 16-byte loads, hi/lo split, 8 tcgen05.st of 16 columns.
        const bool stager = warp < TILES * 4;
        const long long row = row0 + (warp >> 2) * VT + ((warp & 3) << 5) + lane;
        float4 vv[DD / 4];
        if (stager) {
#pragma unroll
            for (int i = 0; i < DD / 4; ++i)
                vv[i] = row < n_rows ? __ldcs(reinterpret_cast<const float4*>(x + row * DD) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (FILTER) {                                        // max ‖x‖ over the CTA's rows: the unit of the queries' margins
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < DD / 4; ++i) ss = fmaf(vv[i].x, vv[i].x, fmaf(vv[i].y, vv[i].y, fmaf(vv[i].z, vv[i].z, fmaf(vv[i].w, vv[i].w, ss))));
                const uint32_t mx = __reduce_max_sync(0xffffffffu, __float_as_uint(ss));     // ss ≥ 0: the bit patterns order like the values
                if (lane == 0) nrm_s[warp] = sqrtf(__uint_as_float(mx)) * 1.0001f;
            }
        }
        umma::fence_before_sync();
        __syncthreads();                                         // TMEM base address published (tmem_slot)
        umma::fence_after_sync();
        if (stager) {
            const uint32_t a_base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + ACC_COLS + (uint32_t)(warp >> 2) * A_COLS;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 h, l;
                    umma::split4(vv[g * 4 + i], h, l);
                    hi[4 * i] = h.x; hi[4 * i + 1] = h.y; hi[4 * i + 2] = h.z; hi[4 * i + 3] = h.w;
                    lo[4 * i] = l.x; lo[4 * i + 1] = l.y; lo[4 * i + 2] = l.z; lo[4 * i + 3] = l.w;
                }
                umma::tmem_st16(a_base + g * 16, hi);
                if (!FILTER) umma::tmem_st16(a_base + DD + g * 16, lo);
            }
            umma::tmem_st_wait();
        }
    } else if (warp < 8) {
        const int r8 = lane & 7, c4l = lane >> 3;
        float4 vv[8 * TILES];
#pragma unroll
        for (int i = 0; i < 8 * TILES; ++i) {
            const int u = warp * 8 * TILES + i;                  // 0 .. 64·TILES
            const int row = (u >> 2) * 8 + r8, c4 = (u & 3) * 4 + c4l;
            vv[i] = (row0 + row < n_rows) ? __ldcs(reinterpret_cast<const float4*>(x + (row0 + row) * DD) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8 * TILES; ++i) {
            const int u = warp * 8 * TILES + i;
            const int row = (u >> 2) * 8 + r8, c4 = (u & 3) * 4 + c4l;
            const int t = row / VT;
            const uint32_t off = (uint32_t)t * V_BYTES + umma::kmajor_offset(VT, row % VT, c4 * 4);
            float4 hi, lo;
            umma::split4(vv[i], hi, lo);
            *reinterpret_cast<float4*>(v_hi + off) = hi;
            *reinterpret_cast<float4*>(v_lo + off) = lo;
        }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp == EW) {
        // ================================ copy + MMA issuer (one thread) ================================ //
        // The whole warp walks the loop (warp-uniform control flow and operands); single instructions are issued by one elected lane.
        const uint32_t idesc = umma::idesc_tf32(VT, QT);
        constexpr uint32_t lbo_a = (VT / 8) * 128, lbo_b = (QT / 8) * 128;
        const uint32_t v_hi_s = umma::smem_u32(v_hi), v_lo_s = umma::smem_u32(v_lo), q_s = umma::smem_u32(qbuf);
        uint32_t ph_full = 0, ph_free = 0, ph_acc = 0;          // one phase bit per barrier (bit i = barrier i)
        bool ok = true;
        for (int c = 0; c < n_chunks && ok; ++c) {
            const int b = c % NBUF, sl = c % NSLOT;
            ok = umma::mbar_wait(&bar_qfull[sl], (ph_full >> sl) & 1);
            ph_full ^= 1u << sl;
            if (!ok) break;
            // descriptors differ only in the start address field (bits 0-13, in 16-byte units)
            const uint64_t dbh = umma::smem_desc(q_s + sl * QSLOT, lbo_b, 128);
            const uint64_t dbl = umma::smem_desc(q_s + sl * QSLOT + (FILTER ? 0 : Q_HALF), lbo_b, 128);
            // one unit = (chunk, tile): 24 MMAs (8 in the filter) into its own accumulator, its own completion barrier — the epilogue
            // of tile 0 starts while the MMAs of tile 1 run, and an accumulator is needed again only four units later
#pragma unroll
            for (int t = 0; t < TILES; ++t) {
                const int un = b * TILES + t;
                if (c >= NBUF) { ok = ok && umma::mbar_wait(&bar_accfree[un], (ph_acc >> un) & 1); ph_acc ^= 1u << un; }
                if (!ok) break;
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    const uint32_t acc = tmem + (uint32_t)un * QT;
                    const uint32_t a_hi = tmem + ACC_COLS + (uint32_t)t * A_COLS, a_lo = a_hi + DD;
                    const uint64_t dah = A_TMEM ? 0 : umma::smem_desc(v_hi_s + t * V_BYTES, lbo_a, 128);
                    const uint64_t dal = A_TMEM ? 0 : umma::smem_desc(v_lo_s + t * V_BYTES, lbo_a, 128);
#pragma unroll
                    for (int j = 0; j < DD / 8; ++j) {
                        const uint64_t oa = (uint64_t)((2 * j * lbo_a) >> 4), ob = (uint64_t)((2 * j * lbo_b) >> 4);
                        if (FILTER) {
                            umma::mma_tf32_ts(acc, a_hi + 8 * j, dbh + ob, idesc, j > 0);
                        } else if (A_TMEM) {
                            umma::mma_tf32_ts(acc, a_lo + 8 * j, dbh + ob, idesc, j > 0);
                            umma::mma_tf32_ts(acc, a_hi + 8 * j, dbl + ob, idesc, true);
                            umma::mma_tf32_ts(acc, a_hi + 8 * j, dbh + ob, idesc, true);
                        } else {
                            umma::mma_tf32(acc, dal + oa, dbh + ob, idesc, j > 0);
                            umma::mma_tf32(acc, dah + oa, dbl + ob, idesc, true);
                            umma::mma_tf32(acc, dah + oa, dbh + ob, idesc, true);
                        }
                    }
                    umma::commit(&bar_m[un]);                    // → epilogue of this unit
                    if (t == TILES - 1) umma::commit(&bar_qfree[sl]);      // → slot sl may be refilled
                }
                __syncwarp();
            }
            if (!ok) break;
            __syncwarp();
            // refill: with a ring of NSLOT > 1 the slot of the PREVIOUS chunk is recycled (its MMAs finish while this chunk's
            // are queued behind them); a single slot has to wait for this chunk itself
            const int cc = NSLOT > 1 ? c - 1 : c;
            if (cc >= 0 && cc + NSLOT < n_chunks) {
                const int s2 = cc % NSLOT;
                ok = umma::mbar_wait(&bar_qfree[s2], (ph_free >> s2) & 1);
                ph_free ^= 1u << s2;
                if (ok && umma::elect_one()) {
                    umma::mbar_expect_tx(&bar_qfull[s2], QSLOT);
                    umma::bulk_g2s(qbuf + s2 * QSLOT, qimg + (size_t)(cc + NSLOT) * Q_IMG, QSLOT, &bar_qfull[s2]);
                }
                __syncwarp();
            }
        }
        if (!ok && lane == 0) { dead = 1; atomicOr(flags + 1, 1); }
    } else {
        // ================================ epilogue warps ================================ //
        const int wq = warp & 7;                                 // FILTER: warps 8t … 8t+7 own tile t; otherwise every warp walks all tiles
        const int r_own = ((wq & 3) << 5) + lane, half = wq >> 2;
        const uint32_t lane_off = (uint32_t)((wq & 3) * 32) << 16;
        const int t_begin = FILTER ? (warp >> 3) : 0, t_end = FILTER ? t_begin + 1 : TILES;
        float nxmax = 0.f;
        if (FILTER) {
#pragma unroll
            for (int i = 0; i < TILES * 4; ++i) nxmax = fmaxf(nxmax, nrm_s[i]);
        }
        WarpSurvivors<CAPW> surv;
        surv.init(surv_mem + warp * WarpSurvivors<CAPW>::BYTES, lane);
        // FILTER: the thresholds of ALL queries, already lowered by each query's margin at this CTA's largest row norm, are staged in
        // shared memory once (a global load per chunk sat on the epilogue's per-chunk latency chain, which bounds this kernel)
        const bool th_staged = FILTER && n_chunks <= TH_SM_CHUNKS;
        if (th_staged) {
            for (int i = tid; i < n_chunks * QT; i += EW * 32) th_sm[i] = fmaf(-__ldg(qmarg + i), nxmax, __ldg(thr + i));
            asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        }
        uint32_t ph_m = 0;
        for (int c = 0; c < n_chunks; ++c) {
            const int b = c % NBUF;
            float4 th4[8];                                       // thresholds of this thread's 32 queries (uniform loads)
            if (th_staged) {
#pragma unroll
                for (int i = 0; i < 8; ++i) th4[i] = *reinterpret_cast<const float4*>(th_sm + c * QT + half * 32 + i * 4);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) th4[i] = __ldg(reinterpret_cast<const float4*>(thr + (long long)c * QT + half * 32) + i);
                if (FILTER) {                                    // lowered by the query's margin at this CTA's largest row norm
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 m = __ldg(reinterpret_cast<const float4*>(qmarg + (long long)c * QT + half * 32) + i);
                        th4[i].x = fmaf(-m.x, nxmax, th4[i].x); th4[i].y = fmaf(-m.y, nxmax, th4[i].y);
                        th4[i].z = fmaf(-m.z, nxmax, th4[i].z); th4[i].w = fmaf(-m.w, nxmax, th4[i].w);
                    }
                }
            }
            const float* th = reinterpret_cast<const float*>(th4);
            const int q0 = c * QT + half * 32;
            bool alive = true;
#pragma unroll
            for (int t = t_begin; t < t_end; ++t) {
                const int un = b * TILES + t;
                if (!umma::mbar_wait(&bar_m[un], (ph_m >> un) & 1)) { atomicOr(flags + 1, 2); alive = false; break; }
                ph_m ^= 1u << un;
                umma::fence_after_sync();
                const uint32_t acc = tmem + lane_off + (uint32_t)un * QT + half * 32;
                float s[32];
                umma::tmem_ld32(acc, s);
                const long long row = row0 + t * VT + r_own;
                // branch-free (a branch per score made the epilogue the bottleneck: 64 reconvergence regions per chunk,
                // instruction-fetch bound).  One predicate-accumulating compare per score decides whether the warp has any
                // survivor at all; only then is the per-score mask built and handed to the warp's survivor buffer (survivors.cuh).
                bool any = false;
#pragma unroll
                for (int j = 0; j < 32; ++j) any = any || (s[j] > th[j]);
                if (row >= n_rows) any = false;
                if (__any_sync(0xffffffffu, any)) {
                    uint32_t m = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) m |= (s[j] > th[j] ? 1u : 0u) << j;
                    if (row >= n_rows) m = 0;
                    surv.add_block(m, acc, q0, (int)row, count, cand_s, stride, kprev, cand_r, cap, flags);   // FILTER: S̃, replaced by flat_rescore_kernel
                }
                umma::fence_before_sync();
                umma::mbar_arrive(&bar_accfree[un]);             // this accumulator may be overwritten by chunk c + 2
            }
            if (!alive) break;
        }
        surv.flush(count, cand_s, stride, kprev, cand_r, cap, flags);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, TM_COLS);
}

// exact fp32 scores of the filter's survivors: 16 lanes per (query, survivor) — lane i holds the i-th 16-byte chunk of the query and of
// the row, four fused multiply-adds, then a fixed xor-shuffle tree (deterministic).  grid (ceil(cap / 128), nq), 256 threads.
__global__ void __launch_bounds__(256) flat_rescore_kernel(const float* __restrict__ q, const float* __restrict__ x, const int* __restrict__ count,
                                                           const int* __restrict__ cand_r, int cap, float* __restrict__ cand_s,
                                                           long long stride, int kprev) {
    const int qi = blockIdx.y;
    const int n = min(__ldg(count + qi), cap);
    const int p0 = blockIdx.x * 128;
    if (p0 >= n) return;
    const int g = threadIdx.x >> 4, l = threadIdx.x & 15;
    const float4 qv = __ldg(reinterpret_cast<const float4*>(q + (long long)qi * DD) + l);
    float4 xv[8];
    int rows[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int p = p0 + i * 16 + g;
        rows[i] = p < n ? __ldg(cand_r + (long long)qi * cap + p) : -1;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
        xv[i] = rows[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (long long)rows[i] * DD) + l) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float d = fmaf(qv.w, xv[i].w, fmaf(qv.z, xv[i].z, fmaf(qv.y, xv[i].y, qv.x * xv[i].x)));
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (l == 0 && rows[i] >= 0) cand_s[(long long)qi * stride + kprev + p0 + i * 16 + g] = d;
    }
}

template <int TILES, int NSLOT, bool A_TMEM, bool FILTER>
int launch_flat_scan(const float* x, long long n_rows, const unsigned char* qimg, int n_chunks, const float* thr, const float* qmarg, int* count,
                     float* cand_s, long long stride, int kprev, int* cand_r, int cap, int* flags, cudaStream_t st) {
    static bool attr_set = false;
    constexpr int EWARPS = FILTER ? 8 * TILES : 8;
    constexpr size_t smem = (A_TMEM ? 0 : (size_t)TILES * 2 * V_BYTES) + (size_t)NSLOT * (FILTER ? Q_HALF : Q_IMG)
                            + (size_t)EWARPS * WarpSurvivors<(TILES == 1 ? 128 : 256)>::BYTES + (FILTER ? (size_t)TH_SM_CHUNKS * QT * 4 : 0);
    constexpr int NTHR = EWARPS * 32 + 32;
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(flat_scan_tc_kernel<TILES, NSLOT, A_TMEM, FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const long long n_cta = (n_rows + TILES * VT - 1) / (TILES * VT);
    RB_REQUIRE(n_rows >= 1 && n_rows < (1ll << 31) && n_cta < (1ll << 31), "flat_scan: a round holds at most 2^31 rows");
    flat_scan_tc_kernel<TILES, NSLOT, A_TMEM, FILTER><<<(unsigned)n_cta, NTHR, smem, st>>>(x, n_rows, qimg, n_chunks, thr, qmarg, count, cand_s,
                                                                                         stride, kprev, cand_r, cap, flags);
    RB_LAUNCH_CHECK("flat_scan_tc_kernel");
    return RB200_OK;
}

}  // namespace

// one round: rows [0, n_rows) of x (the caller offsets x) against n_chunks·64 queries (image), thresholds thr[n_chunks·64]
int rb_flat_scan_tc(const float* x, long long n_rows, const unsigned char* qimg, int n_chunks, const float* thr, const float* qmarg, int* count,
                    float* cand_s, long long stride, int kprev, int* cand_r, int cap, int* flags, cudaStream_t st) {
    static int variant = -1;                // RB200_FLAT_VARIANT: 0 = rows in shared memory (SS MMAs), 1 = rows in TMEM (TS MMAs)
    if (variant < 0) { const char* e = getenv("RB200_FLAT_VARIANT"); variant = e ? atoi(e) : 1; }
    if (n_chunks <= 2)
        return launch_flat_scan<1, 1, false, false>(x, n_rows, qimg, n_chunks, thr, qmarg, count, cand_s, stride, kprev, cand_r, cap, flags, st);
    if (variant == 1)
        return launch_flat_scan<2, 4, true, false>(x, n_rows, qimg, n_chunks, thr, qmarg, count, cand_s, stride, kprev, cand_r, cap, flags, st);
    return launch_flat_scan<2, 2, false, false>(x, n_rows, qimg, n_chunks, thr, qmarg, count, cand_s, stride, kprev, cand_r, cap, flags, st);
}

// survivors of a filtered round → their fp32 scores (x = the round's first row, as passed to rb_flat_scan_tc)
int rb_flat_rescore(const float* q, int nq, const float* x, const int* count, const int* cand_r, int cap, float* cand_s, long long stride,
                    int kprev, cudaStream_t st) {
    flat_rescore_kernel<<<dim3((cap + 127) / 128, nq), 256, 0, st>>>(q, x, count, cand_r, cap, cand_s, stride, kprev);
    RB_LAUNCH_CHECK("flat_rescore_kernel");
    return RB200_OK;
}

// nq_pad: rows of the image (a multiple of 128; rows past nq are zero), block_rows: 64 or 128
int rb_flat_qimage(const float* q, int nq, int nq_pad, int block_rows, unsigned char* qimg, float* qmarg, cudaStream_t st) {
    RB_REQUIRE((block_rows == 64 || block_rows == 128) && nq_pad % 128 == 0 && nq_pad >= nq, "flat_qimage: bad padding");
    const int n = nq_pad * (DD / 4);
    flat_qimage_kernel<<<n / 256, 256, 0, st>>>(q, nq, block_rows, qimg, qmarg);
    RB_LAUNCH_CHECK("flat_qimage_kernel");
    return RB200_OK;
}
