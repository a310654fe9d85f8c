// IVF list scan (step 3 of the search; faiss IndexIVFFlat::search scan of the probed lists) on the tensor cores.
//
// List-major like the FFMA kernel in ivf.cu: one CTA per (list, 128-vector tile).  The tile is staged once as the A
// operand (K-major, hi/lo split: 3xTF32 = fp32-grade inner products); the queries that probe the list — found through
// the radix-sorted (list, query) pairs — are gathered in chunks of 64 as the B operand; S[128 vectors × 64 queries] is
// accumulated in TMEM by tcgen05.mma and written to the compact candidate buffer (for a fixed query, the 32 lanes of a
// warp hold 32 consecutive vectors → 128-byte stores).  D = 64.  96 KB of shared memory → 2 CTAs per SM.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int VT = 128, QT = 64, DD = 64, NT_SC = 256;
constexpr int V_BYTES = VT * DD * 4, Q_BYTES = QT * DD * 4;
constexpr size_t SC_SMEM = 2 * V_BYTES + 2 * Q_BYTES;     // 96 KB

__device__ __forceinline__ void put4s(unsigned char* hi_base, unsigned char* lo_base, int R, int r, int k, const float4& v) {
    const uint32_t off = umma::kmajor_offset(R, r, k);
    float4 hi, lo;
    umma::split4(v, hi, lo);
    *reinterpret_cast<float4*>(hi_base + off) = hi;
    *reinterpret_cast<float4*>(lo_base + off) = lo;
}

__global__ void __launch_bounds__(NT_SC, 2) list_scan_tc_kernel(const float* __restrict__ q, const float* __restrict__ list_vecs,
                                                                const int64_t* __restrict__ offsets, const int* __restrict__ pair_qp,
                                                                const int* __restrict__ list_qstart, int nprobe,
                                                                const int* __restrict__ cand_base, const long long* __restrict__ cand_off,
                                                                float* __restrict__ cand, const int* __restrict__ tile_list,
                                                                const int* __restrict__ tile_idx, int* __restrict__ err_flag) {
    // the index's tile table has 64-vector granularity: even entries own a 128-vector tile, odd entries have nothing to do
    const int ti = tile_idx[blockIdx.x];
    if (ti & 1) return;
    const int l = tile_list[blockIdx.x];
    const long long lbeg = offsets[l], lend = offsets[l + 1];
    const long long v0 = lbeg + (long long)ti * 64;
    const int qs = list_qstart[l], qe = list_qstart[l + 1];
    if (v0 >= lend || qs == qe) return;

    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* v_hi = smem;
    unsigned char* v_lo = v_hi + V_BYTES;
    unsigned char* q_hi = v_lo + V_BYTES;
    unsigned char* q_lo = q_hi + Q_BYTES;
    __shared__ __align__(8) uint64_t bar_mem;
    __shared__ uint32_t tmem_slot;
    __shared__ int dead;
    __shared__ long long dst[QT];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, QT);
    if (tid == 0) { umma::mbar_init(&bar_mem, 1); umma::fence_mbar_init(); dead = 0; }
    // Every global load of the first chunk is issued before any of them is used: the (list, query) pairs first, then the
    // vector tile, then the query rows the pairs point to (measured: the serial version spent 1.9 us staging the tile and
    // another 2.7 us gathering the queries per work item).  A warp covers 8 rows × 4 16-byte chunks per step (64-byte
    // global segments, conflict-free 128-byte shared-memory phases).
    const int r8 = lane & 7, c4l = lane >> 3;
    const int nv = (int)min((long long)VT, lend - v0);
    int qrow[4];
    auto load_pairs = [&](int p0, long long& d) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int pi = p0 + ((warp * 4 + i) >> 2) * 8 + r8;
            qrow[i] = pi < qe ? __ldg(pair_qp + pi) / nprobe : -1;
        }
        d = -1;
        if (tid < QT && p0 + tid < qe) {
            const int qp = __ldg(pair_qp + p0 + tid);
            d = cand_off[qp / nprobe] + cand_base[qp] + (v0 - lbeg);
        }
    };
    auto load_queries = [&](float4 (&qv)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c4 = ((warp * 4 + i) & 3) * 4 + c4l;
            qv[i] = qrow[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(q + (long long)qrow[i] * DD) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_queries = [&](const float4 (&qv)[4], long long d) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int u = warp * 4 + i;
            put4s(q_hi, q_lo, QT, (u >> 2) * 8 + r8, ((u & 3) * 4 + c4l) * 4, qv[i]);
        }
        if (tid < QT) dst[tid] = d;
    };
    long long d_first;
    float4 qv[4], vv[8];
    load_pairs(qs, d_first);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int u = warp * 8 + i, row = (u >> 2) * 8 + r8, c4 = (u & 3) * 4 + c4l;
        vv[i] = (v0 + row < lend) ? __ldg(reinterpret_cast<const float4*>(list_vecs + (v0 + row) * DD) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    load_queries(qv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int u = warp * 8 + i;
        put4s(v_hi, v_lo, VT, (u >> 2) * 8 + r8, ((u & 3) * 4 + c4l) * 4, vv[i]);
    }
    store_queries(qv, d_first);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int r_own = ((warp & 3) << 5) + lane, half = warp >> 2;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc = umma::idesc_tf32(VT, QT);
    const uint32_t lbo_a = (VT / 8) * 128, lbo_b = (QT / 8) * 128;
    uint32_t phase = 0;
    for (int p0 = qs; p0 < qe; p0 += QT) {
        if (p0 != qs) {                        // further chunks of a popular list (rare): gather + stage, then the same MMA path
            long long d;
            load_pairs(p0, d);
            load_queries(qv);
            store_queries(qv, d);
            umma::fence_proxy_async();
            umma::fence_before_sync();
            __syncthreads();
        }
        if (!dead && umma::elect_issuer(tid)) {
            umma::fence_after_sync();
            const uint32_t ah = umma::smem_u32(v_hi), al = umma::smem_u32(v_lo), bh = umma::smem_u32(q_hi), bl = umma::smem_u32(q_lo);
#pragma unroll
            for (int j = 0; j < DD / 8; ++j) {
                const uint32_t oa = 2 * j * lbo_a, ob = 2 * j * lbo_b;
                umma::mma_tf32(tmem, umma::smem_desc(al + oa, lbo_a, 128), umma::smem_desc(bh + ob, lbo_b, 128), idesc, j > 0);
                umma::mma_tf32(tmem, umma::smem_desc(ah + oa, lbo_a, 128), umma::smem_desc(bl + ob, lbo_b, 128), idesc, true);
                umma::mma_tf32(tmem, umma::smem_desc(ah + oa, lbo_a, 128), umma::smem_desc(bh + ob, lbo_b, 128), idesc, true);
            }
            umma::commit(&bar_mem);
        }
        if (!dead && !umma::mbar_wait(&bar_mem, phase)) { dead = 1; if (err_flag) atomicOr(err_flag, 2); }
        phase ^= 1;
        umma::fence_after_sync();
        float s[32];
        if (!dead) umma::tmem_ld32(tmem + lane_off + half * 32, s);
        if (r_own < nv) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const long long d = dst[half * 32 + c];
                if (d >= 0) __stcs(cand + d + r_own, s[c]);
            }
        }
        umma::fence_before_sync();
        __syncthreads();                      // dst / the query operand / the accumulator are reused by the next chunk
    }
    if (warp == 0) umma::tmem_free(tmem, QT);
}

// ------------------------------------------------------------------------------------------------------------------------ //
// The same scan as a PERSISTENT, WARP-SPECIALISED kernel.  ncu on list_scan_tc_kernel (profiles/r02_ivf.md): warps active 22 %, 27 % of
// the stall samples at the block barrier behind four dependent global round trips per work item (tile table → offsets / query range →
// (list, query) pairs → query rows and destinations), and every unit — a 128-vector tile against a chunk of ≤ 64 probing queries; the
// chunks beyond the first of a popular list are 40 % of the units on the bench's skewed index — ran load → stage → MMA → store as one
// serial chain with two CTAs per SM to overlap.  (A first persistent form that only took the dependent chains off the main warps
// measured the same 270 us per batch: the chain stage → MMA → store itself, ≈ 4.4 us per unit, was the bound.)
//
// One CTA per SM walks the tile table with stride gridDim.x and, per tile, the chunks of the queries that probe its list:
//   * descriptor warp — owns every dependent chain: tile entry two tiles ahead and list geometry one tile ahead (registers), then per
//     chunk one load of the pairs and of their destinations (pair_dst, precomputed by the plan) → a unit descriptor (query rows,
//     destinations, tile geometry) in a ring of four shared-memory slots;
//   * sixteen loader warps — global loads of unit u+2 go out (two register sets) before unit u+1 is staged: vector tile (when the tile
//     changes; two shared-memory buffers) and query chunk (two buffers), hi/lo split, K-major operand layout;
//   * one MMA warp — 24 tcgen05.mma per unit (3xTF32) with descriptors built once, two accumulators in TMEM;
//   * eight epilogue warps — TMEM → candidate buffer (128-byte stores per (warp, query)).
// Hand-offs through mbarriers only; every barrier has one waiting role that sees its phases in order.
// Measured (device timestamps per role, one CTA): ≈ 3 us per unit — the loaders wait ≈ 1.6 us for the vector tile they requested a unit
// earlier and the epilogue needs ≈ 2 us for its 32 stores per thread: the SM's load/store path, shared by the loaders' 16-byte gathers
// and the 256 misaligned 128-byte candidate stores of a unit, is what bounds the kernel now (C3 batch 0.67 → 0.63 ms; a third query
// buffer / accumulator and an L2 prefetch of the tiles changed nothing).  Next: the vector tile by tensor-map TMA (off the LSU).
struct UnitDesc {
    long long v0;              // first vector (row of list_vecs) of the tile
    int nv;                    // vectors in the tile (≤ 128)
    int new_tile;              // the vector tile has to be staged (first chunk of the tile)
    int last;                  // last chunk of the tile (its vector buffer is free once these MMAs are done)
    int valid;                 // 0: end of this CTA's work
    unsigned dst[QT];          // per query of the chunk: where its scores for this tile start in cand (0xFFFFFFFF: no query)
    int qrow[QT];              // per query of the chunk: its row in q (−1: none)
};
constexpr unsigned NO_DST = 0xFFFFFFFFu;
constexpr int NQB = 2, NACC = 3;                            // query-chunk buffers and accumulators
constexpr int RAW_BYTES = VT * DD * 4;                      // a vector tile as it arrives by TMA: two 128-byte-swizzled [128 × 32] atoms
constexpr int RING = 4, N_LOAD = 16, N_EPI = 8;            // (eight loader warps were busy 84 % of the time: staging bounded the kernel)
constexpr int VPT = 64 / N_LOAD, QPT = 32 / N_LOAD;        // 16-byte pieces of the vector tile / the query chunk per loader thread
constexpr int W_MMA = N_LOAD + N_EPI, W_DESC = W_MMA + 1, NT_PIPE = (W_DESC + 1) * 32;
constexpr int N_BARS = 2 * RING + 2 * NQB + 2 + 2 * NACC + 2;
// 224 KB of operands + the descriptor ring + the barriers: everything in the dynamic block (a static block next to a 1024-byte aligned
// dynamic one is padded to 1 KB, which would not fit the 227 KB any more)
constexpr size_t PIPE_SMEM = 2 * (2 * V_BYTES) + NQB * (2 * Q_BYTES) + RAW_BYTES + RING * sizeof(UnitDesc) + N_BARS * 8 + 16;

__device__ __forceinline__ void tma_rows_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(umma::smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(umma::smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(NT_PIPE, 1) list_scan_pipe_kernel(const __grid_constant__ CUtensorMap vmap, const float* __restrict__ q,
                                                                    const int64_t* __restrict__ offsets, const int* __restrict__ pair_qp,
                                                                    const long long* __restrict__ pair_dst,
                                                                    const int* __restrict__ list_qstart, int nprobe, float* __restrict__ cand,
                                                                    const int* __restrict__ tile_list, const int* __restrict__ tile_idx,
                                                                    int n_tiles, int* __restrict__ err_flag) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* vbuf = smem;                                  // [2][hi V_BYTES | lo V_BYTES]
    unsigned char* qbuf = smem + 4 * V_BYTES;                    // [NQB][hi Q_BYTES | lo Q_BYTES]
    unsigned char* raw = qbuf + NQB * 2 * Q_BYTES;               // [2 atoms][128 rows × 128 B], SWIZZLE_128B (1024-byte aligned)
    UnitDesc* ring = reinterpret_cast<UnitDesc*>(raw + RAW_BYTES);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(ring + RING);
    uint64_t* bar_empty = bar_full + RING;
    uint64_t* bar_sfull = bar_empty + RING;
    uint64_t* bar_qfree = bar_sfull + NQB;
    uint64_t* bar_vfree = bar_qfree + NQB;
    uint64_t* bar_done = bar_vfree + 2;
    uint64_t* bar_free = bar_done + NACC;
    uint64_t* bar_rawfull = bar_free + NACC;
    uint64_t* bar_rawfree = bar_rawfull + 1;
    uint32_t* tmem_slot_p = reinterpret_cast<uint32_t*>(bar_rawfree + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc(tmem_slot_p, 4 * QT);
    if (tid == 32) {
        for (int i = 0; i < RING; ++i) { umma::mbar_init(&bar_full[i], 1); umma::mbar_init(&bar_empty[i], N_LOAD + 1 + N_EPI); }
        for (int i = 0; i < NQB; ++i) { umma::mbar_init(&bar_sfull[i], N_LOAD); umma::mbar_init(&bar_qfree[i], 1); }
        for (int i = 0; i < NACC; ++i) { umma::mbar_init(&bar_done[i], 1); umma::mbar_init(&bar_free[i], N_EPI); }
        for (int i = 0; i < 2; ++i) umma::mbar_init(&bar_vfree[i], 1);
        umma::mbar_init(bar_rawfull, 1); umma::mbar_init(bar_rawfree, N_LOAD);
        umma::fence_mbar_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot_p;
    bool ok = true;

    if (warp == W_DESC) {
        // ================================ descriptor warp ================================ //
        const int G = (int)gridDim.x;
        int slot = 0, uses = 0, n_tma = 0;                       // uses: descriptors written so far (slot = uses % RING); tiles requested
        auto publish = [&](bool okk) -> bool {                   // wait until the slot is free (its previous unit has been consumed)
            if (uses >= RING && okk) okk = umma::mbar_wait(&bar_empty[slot], ((uses / RING) - 1) & 1);
            return okk;
        };
        // software pipeline over this CTA's tile entries: entry e+2G's (tile, list) and entry e+G's list geometry are in flight while
        // the chunks of entry e are described
        const int e0 = (int)blockIdx.x;
        int ti1 = 1, l1 = 0, ti2 = 1, l2 = 0;                    // entries e+G and e+2G (ti odd = nothing to do)
        long long lb1 = 0, le1 = 0; int qs1 = 0, qe1 = 0;
        int ti0 = 1, l0 = 0; long long lb0 = 0, le0 = 0; int qs0 = 0, qe0 = 0;
        if (e0 < n_tiles) { ti0 = __ldg(tile_idx + e0); l0 = __ldg(tile_list + e0); }
        if (e0 + G < n_tiles) { ti1 = __ldg(tile_idx + e0 + G); l1 = __ldg(tile_list + e0 + G); }
        if (!(ti0 & 1)) { lb0 = __ldg(offsets + l0); le0 = __ldg(offsets + l0 + 1); qs0 = __ldg(list_qstart + l0); qe0 = __ldg(list_qstart + l0 + 1); }
        for (int e = e0; e < n_tiles && ok; e += G) {
            if (e + 2 * G < n_tiles) { ti2 = __ldg(tile_idx + e + 2 * G); l2 = __ldg(tile_list + e + 2 * G); } else ti2 = 1;
            if (e + G < n_tiles && !(ti1 & 1)) {
                lb1 = __ldg(offsets + l1); le1 = __ldg(offsets + l1 + 1); qs1 = __ldg(list_qstart + l1); qe1 = __ldg(list_qstart + l1 + 1);
            }
            const long long v0 = lb0 + (long long)ti0 * 64;
            if (!(ti0 & 1) && v0 < le0 && qs0 < qe0) {
                const int nv = (int)min((long long)VT, le0 - v0);
                for (int p0 = qs0; p0 < qe0 && ok; p0 += QT) {
                    int qp[2]; long long pd[2];                  // this chunk's pairs first (the only dependent loads of a unit), then the slot
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int pi = p0 + h * 32 + lane;
                        qp[h] = pi < qe0 ? __ldg(pair_qp + pi) : -1;
                        pd[h] = pi < qe0 ? __ldg(pair_dst + pi) : -1;
                    }
                    ok = publish(ok);
                    if (!ok) break;
                    UnitDesc& u = ring[slot];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        u.qrow[h * 32 + lane] = qp[h] >= 0 ? qp[h] / nprobe : -1;
                        u.dst[h * 32 + lane] = qp[h] >= 0 ? (unsigned)(pd[h] + (v0 - lb0)) : NO_DST;
                    }
                    if (lane == 0) { u.v0 = v0; u.nv = nv; u.new_tile = p0 == qs0; u.last = p0 + QT >= qe0; u.valid = 1; }
                    __syncwarp();
                    if (lane == 0) { __threadfence_block(); umma::mbar_arrive(&bar_full[slot]); }
                    ++uses; slot = uses % RING;
                    if (p0 == qs0) {
                        // The tile's rows by tensor-map TMA into the raw buffer (off the SM's load/store path, where the loaders' 16-byte
                        // gathers queued behind the epilogue's stores), as soon as the loaders have taken the previous tile out of it —
                        // AFTER the descriptor is out: the loaders look one descriptor ahead before they stage (and free the raw
                        // buffer of) the current unit, so waiting here first would dead-lock.  Rows past the tile belong to the next
                        // list (or are zero-filled past the table): their scores are not stored.
                        if (n_tma >= 1 && ok) ok = umma::mbar_wait(bar_rawfree, (n_tma - 1) & 1);
                        if (!ok) break;
                        if (umma::elect_one()) {
                            umma::mbar_expect_tx(bar_rawfull, RAW_BYTES);
                            tma_rows_2d(raw, &vmap, 0, (int)v0, bar_rawfull);
                            tma_rows_2d(raw + RAW_BYTES / 2, &vmap, 32, (int)v0, bar_rawfull);
                        }
                        __syncwarp();
                        ++n_tma;
                    }
                }
            }
            ti0 = ti1; l0 = l1; lb0 = lb1; le0 = le1; qs0 = qs1; qe0 = qe1;
            ti1 = ti2; l1 = l2;
        }
        ok = publish(ok);                                        // end marker
        if (ok && lane == 0) { ring[slot].valid = 0; __threadfence_block(); umma::mbar_arrive(&bar_full[slot]); }
        if (!ok && lane == 0 && err_flag) atomicOr(err_flag, 4);
    } else if (warp < N_LOAD) {
        // ================================ loader warps ================================ //
        const int r8 = lane & 7, c4l = lane >> 3;
        float4 qvA[QPT], qvB[QPT];
        // this thread's pieces: (row, 16-byte column) of the tile, where they sit in the swizzled raw tile and their K-major
        // shared-memory offsets in the operand buffers — fixed for the whole kernel
        int qrow_i[QPT], qc4[QPT];
        uint32_t voff[VPT], roff[VPT], qoff[QPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int uu = warp * VPT + i;
            const int vrow = (uu >> 2) * 8 + r8, vc4 = (uu & 3) * 4 + c4l;
            voff[i] = umma::kmajor_offset(VT, vrow, vc4 * 4);
            roff[i] = (uint32_t)((vc4 >> 3) * (RAW_BYTES / 2) + vrow * 128 + (((vc4 & 7) ^ (vrow & 7)) << 4));
        }
#pragma unroll
        for (int i = 0; i < QPT; ++i) {
            const int uu = warp * QPT + i;
            qrow_i[i] = (uu >> 2) * 8 + r8; qc4[i] = (uu & 3) * 4 + c4l;
            qoff[i] = umma::kmajor_offset(QT, qrow_i[i], qc4[i] * 4);
        }
        auto load_unit = [&](const UnitDesc& u, float4 (&qv)[QPT]) {
#pragma unroll
            for (int i = 0; i < QPT; ++i) {
                const int qr = u.qrow[qrow_i[i]];
                qv[i] = qr >= 0 ? __ldg(reinterpret_cast<const float4*>(q + (long long)qr * DD) + qc4[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto put = [&](unsigned char* hi_base, unsigned char* lo_base, uint32_t off, const float4& v) {
            float4 hi, lo;
            umma::split4(v, hi, lo);
            *reinterpret_cast<float4*>(hi_base + off) = hi;
            *reinterpret_cast<float4*>(lo_base + off) = lo;
        };
        int n_tile = -1;                                         // tiles started so far − 1: the vector buffer of the current tile is n_tile & 1
        // unit u: descriptor in ring[u % RING]; stage it from the given register set
        auto stage_unit = [&](int u, const float4 (&qv)[QPT]) {
            const UnitDesc& d = ring[u % RING];
            const int qb = u % NQB;
            if (d.new_tile) {
                ++n_tile;
                const int vb = n_tile & 1;
                if (n_tile >= 2 && ok) ok = umma::mbar_wait(&bar_vfree[vb], ((n_tile >> 1) - 1) & 1);
                unsigned char* vh = vbuf + vb * 2 * V_BYTES;
                if (ok) ok = umma::mbar_wait(bar_rawfull, n_tile & 1);          // the tile has landed
#pragma unroll
                for (int i = 0; i < VPT; ++i) put(vh, vh + V_BYTES, voff[i], *reinterpret_cast<const float4*>(raw + roff[i]));
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(bar_rawfree);                  // this warp has taken its pieces out of the raw tile
            }
            if (u >= NQB && ok) ok = umma::mbar_wait(&bar_qfree[qb], ((u / NQB) - 1) & 1);
            unsigned char* qh = qbuf + qb * 2 * Q_BYTES;
#pragma unroll
            for (int i = 0; i < QPT; ++i) put(qh, qh + Q_BYTES, qoff[i], qv[i]);
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) { umma::mbar_arrive(&bar_sfull[qb]); umma::mbar_arrive(&bar_empty[u % RING]); }
        };
        auto desc_ready = [&](int u) -> bool {                   // descriptor of unit u present and valid
            if (ok) ok = umma::mbar_wait(&bar_full[u % RING], (u / RING) & 1);
            return ok && ring[u % RING].valid;
        };
        // two register sets: the loads of unit u+1 are in flight while unit u is staged
        bool vA = desc_ready(0), vB = false;
        if (vA) load_unit(ring[0], qvA);
        for (int u = 0; vA; u += 2) {
            vB = desc_ready(u + 1);
            if (vB) load_unit(ring[(u + 1) % RING], qvB);
            stage_unit(u, qvA);
            if (!vB) break;
            vA = desc_ready(u + 2);
            if (vA) load_unit(ring[(u + 2) % RING], qvA);
            stage_unit(u + 1, qvB);
        }
        if (!ok && lane == 0 && err_flag) atomicOr(err_flag, 8);
    } else if (warp == W_MMA) {
        // ================================ MMA warp ================================ //
        const uint32_t idesc = umma::idesc_tf32(VT, QT);
        const uint32_t lbo_a = (VT / 8) * 128, lbo_b = (QT / 8) * 128;
        const uint32_t v_s = umma::smem_u32(vbuf), q_s = umma::smem_u32(qbuf);
        int n_tile = -1;
        for (int u = 0; ok; ++u) {
            ok = umma::mbar_wait(&bar_full[u % RING], (u / RING) & 1);
            if (!ok) break;
            const UnitDesc& d = ring[u % RING];
            if (!d.valid) break;
            const int new_tile = d.new_tile, last = d.last;
            if (new_tile) ++n_tile;
            const int vb = n_tile & 1, qb = u % NQB, a = u % NACC;
            ok = umma::mbar_wait(&bar_sfull[qb], (u / NQB) & 1);
            if (ok && u >= NACC) ok = umma::mbar_wait(&bar_free[a], ((u / NACC) - 1) & 1);
            if (!ok) break;
            umma::fence_after_sync();
            if (umma::elect_one()) {
                const uint64_t dah = umma::smem_desc(v_s + vb * 2 * V_BYTES, lbo_a, 128), dal = umma::smem_desc(v_s + vb * 2 * V_BYTES + V_BYTES, lbo_a, 128);
                const uint64_t dbh = umma::smem_desc(q_s + qb * 2 * Q_BYTES, lbo_b, 128), dbl = umma::smem_desc(q_s + qb * 2 * Q_BYTES + Q_BYTES, lbo_b, 128);
                const uint32_t acc = tmem + (uint32_t)a * QT;
#pragma unroll
                for (int j = 0; j < DD / 8; ++j) {
                    const uint64_t oa = (uint64_t)((2 * j * lbo_a) >> 4), ob = (uint64_t)((2 * j * lbo_b) >> 4);
                    umma::mma_tf32(acc, dal + oa, dbh + ob, idesc, j > 0);
                    umma::mma_tf32(acc, dah + oa, dbl + ob, idesc, true);
                    umma::mma_tf32(acc, dah + oa, dbh + ob, idesc, true);
                }
                umma::commit(&bar_done[a]);                      // → epilogue of this unit
                umma::commit(&bar_qfree[qb]);                    // → the query buffer may be refilled
                if (last) umma::commit(&bar_vfree[vb]);          // → so may the vector buffer, when this was the tile's last chunk
                umma::mbar_arrive(&bar_empty[u % RING]);
            }
            __syncwarp();
        }
        if (!ok && lane == 0 && err_flag) atomicOr(err_flag, 1);
    } else {
        // ================================ epilogue warps ================================ //
        const int ew = warp - N_LOAD;
        const int r_own = ((ew & 3) << 5) + lane, half = ew >> 2;
        const uint32_t lane_off = (uint32_t)((ew & 3) * 32) << 16;
        for (int u = 0; ok; ++u) {
            ok = umma::mbar_wait(&bar_full[u % RING], (u / RING) & 1);
            if (!ok) break;
            const UnitDesc& d = ring[u % RING];
            if (!d.valid) break;
            const int a = u % NACC;
            ok = umma::mbar_wait(&bar_done[a], (u / NACC) & 1);
            if (!ok) break;
            umma::fence_after_sync();
            float s[32];
            umma::tmem_ld32(tmem + lane_off + (uint32_t)a * QT + half * 32, s);
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_free[a]);      // the scores are in registers: the accumulator goes back to the MMA warp
            if (r_own < d.nv) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const unsigned dd = d.dst[half * 32 + c];
                    if (dd != NO_DST) __stcs(cand + (size_t)dd + r_own, s[c]);
                }
            }
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bar_empty[u % RING]);
        }
        if (!ok && lane == 0 && err_flag) atomicOr(err_flag, 2);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, 4 * QT);
}

}  // namespace

// returns 1 when this kernel does not cover the shape (caller uses the FFMA kernel)
int rb_list_scan_tc(const float* q, int D, const float* list_vecs, const int64_t* offsets, const int* pair_qp, const int* list_qstart,
                    int nprobe, const int* cand_base, const long long* cand_off, float* cand, const int* tile_list,
                    const int* tile_idx, long long n_tiles, cudaStream_t st) {
    if (D != DD || !tile_list || !tile_idx || n_tiles <= 0) return 1;
    static bool attr_set = false;
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(list_scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM));
        attr_set = true;
    }
    list_scan_tc_kernel<<<(unsigned)n_tiles, NT_SC, SC_SMEM, st>>>(q, list_vecs, offsets, pair_qp, list_qstart, nprobe, cand_base,
                                                                   cand_off, cand, tile_list, tile_idx, nullptr);
    RB_LAUNCH_CHECK("list_scan_tc_kernel");
    return RB200_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// returns 1 when this kernel does not cover the shape or is switched off (RB200_IVF_PIPE=0: one CTA per tile, list_scan_tc_kernel)
int rb_list_scan_pipe(const float* q, int D, const float* list_vecs, long long n_vectors, const int64_t* offsets, const int* pair_qp, const long long* pair_dst,
                      const int* list_qstart, int nprobe, float* cand, long long total_candidates, const int* tile_list, const int* tile_idx,
                      long long n_tiles, int* err_flag, cudaStream_t st) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("RB200_IVF_PIPE"); on = e ? atoi(e) : 1; }
    // (the unit descriptors keep candidate offsets in 32 bits)
    if (!on || D != DD || !tile_list || !tile_idx || n_tiles <= 0 || n_tiles >= (1ll << 31) || total_candidates >= 0xFFFFFFFFll) return 1;
    // (the tensor map: int32 row coordinates, a 16-byte aligned base, at least one full box of rows — smaller indexes take the
    //  one-CTA-per-tile kernel, where speed does not matter)
    if (n_vectors < VT || n_vectors >= (1ll << 31) || (reinterpret_cast<uintptr_t>(list_vecs) & 15)) return 1;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return 1;
    CUtensorMap vmap;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)DD, (cuuint64_t)n_vectors};
        const cuuint64_t strides[1] = {(cuuint64_t)DD * 4};
        const cuuint32_t box[2] = {32, (cuuint32_t)VT};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&vmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)list_vecs, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 1;                        // (not fatal: the other list-scan kernel needs no tensor map)
    }
    static bool attr_set = false;
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(list_scan_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_SMEM));
        attr_set = true;
    }
    long long grid = rb_sm_count();
    if (grid > n_tiles) grid = n_tiles;
    list_scan_pipe_kernel<<<(unsigned)grid, NT_PIPE, PIPE_SMEM, st>>>(vmap, q, offsets, pair_qp, pair_dst, list_qstart, nprobe, cand,
                                                                      tile_list, tile_idx, (int)n_tiles, err_flag);
    RB_LAUNCH_CHECK("list_scan_pipe_kernel");
    return RB200_OK;
}
