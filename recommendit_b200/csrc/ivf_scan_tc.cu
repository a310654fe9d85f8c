// IVF list scan (step 3 of the search; faiss IndexIVFFlat::search scan of the probed lists) on the tensor cores.
//
// List-major like the FFMA kernel in ivf.cu: one CTA per (list, 128-vector tile).  The tile is staged once as the A
// operand (K-major, hi/lo split: 3xTF32 = fp32-grade inner products); the queries that probe the list — found through
// the radix-sorted (list, query) pairs — are gathered in chunks of 64 as the B operand; S[128 vectors × 64 queries] is
// accumulated in TMEM by tcgen05.mma and written to the compact candidate buffer (for a fixed query, the 32 lanes of a
// warp hold 32 consecutive vectors → 128-byte stores).  D = 64.  96 KB of shared memory → 2 CTAs per SM.
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
