// Exhaustive inner-product scan for FEW queries (≤ 128: one or two 64-query chunks) — the round kernel of rb200_flat_search where
// the scan is bound by READING THE ROWS (BASELINE cfg 5 at serving batch sizes: 12.5 M × 64 fp32 rows = 3.2 GB per query batch).
//
// flat_scan_tc.cu's database-stationary CTAs (one per 128 rows: TMEM allocation, barrier set-up, a 32 KB query image and thread-staged
// rows per CTA, two CTAs per SM) reach 23 % of the HBM peak here.  This kernel is the streaming form:
//
//   * persistent — one CTA per SM walks the 128-row tiles blockIdx.x, blockIdx.x + gridDim.x, …; the query image is loaded ONCE;
//   * the rows arrive by TENSOR-MAP TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B): each tile is two [128 rows × 32 floats] boxes that land
//     in shared memory already in the canonical K-major 128-byte-swizzle operand layout, straight from the row-major fp32 table — no
//     thread touches a row on its way to the tensor core.  A ring of NST = 5 stages (160 KB) keeps ≈ 24 MB in flight over the chip;
//   * ONE-pass TF32 filter (see flat_scan_tc.cu): kind::tf32 reads the raw fp32 rows and ignores their low 13 mantissa bits
//     (ε_x ≤ 2⁻¹⁰), the queries are the round-to-nearest hi image (ε_q ≤ 2⁻¹¹): |S − S̃| ≤ (ε_x + ε_q + ε_xε_q)·‖q‖‖x‖ < 1.51·2⁻¹⁰·‖q‖‖x‖.
//     A score survives when S̃ > thr[q] − (4/3)·qmarg[q]·max‖x‖ = thr[q] − 2⁻⁹·‖q‖·max‖x‖ (max over the tile's rows, which four epilogue
//     warps read back from the stage); survivors are re-scored in fp32 by flat_rescore_kernel before the select;
//   * warp-specialised: warps 0-7 epilogue (two groups of four warps taking alternate tiles; thread = row), warp 8 MMA issuer, warp 9 TMA
//     producer; a ring of four accumulators [128 rows × 128 queries] in TMEM (512 columns); mbarriers only (full / empty per stage, done /
//     free per accumulator); survivors go through per-warp shared-memory buffers (survivors.cuh).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "survivors.cuh"
#include "umma.cuh"

namespace {

constexpr int VT = 128, QT = 64, DD = 64;
constexpr int NST = 5, NACC = 4;
constexpr int STAGE_BYTES = VT * DD * 4;         // 32 KB: two 16 KB swizzle atoms columns [0, 32) and [32, 64)
constexpr int ATOM_BYTES = VT * 128;             // [128 rows × 128 B]
constexpr int Q_HALF = QT * DD * 4;              // hi image of one 64-query chunk (flat_qimage_kernel: [hi | lo] per chunk)
constexpr int Q_IMG = 2 * Q_HALF;
constexpr int NT_S = 10 * 32;
constexpr int CAPW = 256;                        // survivor buffer entries per epilogue warp

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart (SBO); LBO is not used by swizzled
// K-major layouts (1); version 1 (Blackwell); layout type 2 = SWIZZLE_128B (bits 61-63)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(umma::smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(umma::smem_u32(bar))
                 : "memory");
}

template <int NCH>
__global__ void __launch_bounds__(NT_S, 1)
flat_stream_tc_kernel(const __grid_constant__ CUtensorMap xmap, long long n_rows, int n_tiles, const unsigned char* __restrict__ qimg,
                      const float* __restrict__ thr, const float* __restrict__ qmarg, float marg_scale, int* __restrict__ count,
                      float* __restrict__ cand_s, long long stride, int kprev, int* __restrict__ cand_r, int cap, int* __restrict__ flags) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* stages = smem;                                // [NST][STAGE_BYTES], 1024-byte aligned atoms
    unsigned char* qbuf = smem + NST * STAGE_BYTES;              // [NCH][Q_HALF]
    unsigned char* surv_mem = qbuf + NCH * Q_HALF;               // [8][WarpSurvivors<CAPW, false>::BYTES] (survivors.cuh)
    __shared__ __align__(8) uint64_t bar_full[NST], bar_empty[NST], bar_done[NACC], bar_free[NACC], bar_q;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int dead;                                // a barrier wait timed out: every role stops waiting (flags[1] tells the host)
    __shared__ float nrm_s[NACC][4];
    __shared__ __align__(16) float th_s[NCH * QT], qm_s[NCH * QT];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t TM_COLS = 512;                            // NACC accumulators of 128 columns (NCH·64 used)

    if (warp == 0) umma::tmem_alloc(&tmem_slot, TM_COLS);
    if (tid == 8 * 32) {
        for (int i = 0; i < NST; ++i) { umma::mbar_init(&bar_full[i], 1); umma::mbar_init(&bar_empty[i], 1 + 4); }
        for (int i = 0; i < NACC; ++i) { umma::mbar_init(&bar_done[i], 1); umma::mbar_init(&bar_free[i], 128); }
        umma::mbar_init(&bar_q, 1);
        umma::fence_mbar_init();
        dead = 0;
        umma::mbar_expect_tx(&bar_q, NCH * Q_HALF);
        for (int c = 0; c < NCH; ++c) umma::bulk_g2s(qbuf + c * Q_HALF, qimg + (size_t)c * Q_IMG, Q_HALF, &bar_q);
    }
    if (tid < NCH * QT) { th_s[tid] = __ldg(thr + tid); qm_s[tid] = __ldg(qmarg + tid) * marg_scale; }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int n_mine = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // tiles of this CTA

    if (warp == 9) {
        // ================================ TMA producer ================================ //
        bool ok = true;
        for (int i = 0; i < n_mine && ok; ++i) {
            const int st = i % NST;
            if (i >= NST) ok = umma::mbar_wait(&bar_empty[st], ((i / NST) - 1) & 1);
            if (ok && umma::elect_one()) {
                const int r0 = ((int)blockIdx.x + i * (int)gridDim.x) * VT;
                umma::mbar_expect_tx(&bar_full[st], STAGE_BYTES);            // rows past the table are zero-filled and still counted
                tma_load_2d(stages + st * STAGE_BYTES, &xmap, 0, r0, &bar_full[st]);
                tma_load_2d(stages + st * STAGE_BYTES + ATOM_BYTES, &xmap, 32, r0, &bar_full[st]);
            }
            __syncwarp();
        }
        if (!ok && lane == 0) { dead = 1; atomicOr(flags + 1, 4); }
    } else if (warp == 8) {
        // ================================ MMA issuer ================================ //
        const uint32_t idesc = umma::idesc_tf32(VT, QT);
        constexpr uint32_t lbo_b = (QT / 8) * 128;
        const uint32_t st_s = umma::smem_u32(stages), q_s = umma::smem_u32(qbuf);
        bool ok = umma::mbar_wait(&bar_q, 0);
        for (int i = 0; i < n_mine && ok; ++i) {
            const int st = i % NST, a = i % NACC;
            ok = umma::mbar_wait(&bar_full[st], (i / NST) & 1);
            if (ok && i >= NACC) ok = umma::mbar_wait(&bar_free[a], ((i / NACC) - 1) & 1);
            if (!ok) break;
            umma::fence_after_sync();
            if (umma::elect_one()) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const uint32_t acc = tmem + (uint32_t)a * 128 + (uint32_t)c * QT;
                    const uint64_t db = umma::smem_desc(q_s + c * Q_HALF, lbo_b, 128);
#pragma unroll
                    for (int j = 0; j < DD / 8; ++j) {
                        // K step j: 32 bytes further inside the 128-byte swizzle row; the second atom holds k ≥ 32
                        const uint64_t da = smem_desc_sw128(st_s + st * STAGE_BYTES + (j >> 2) * ATOM_BYTES + (j & 3) * 32);
                        umma::mma_tf32(acc, da, db + (uint64_t)((2 * j * lbo_b) >> 4), idesc, j > 0);
                    }
                }
                umma::commit(&bar_done[a]);                      // → epilogue of this tile
                umma::commit(&bar_empty[st]);                    // → the stage may be refilled (once the norm readers are done too)
            }
            __syncwarp();
        }
        if (!ok && lane == 0) { dead = 1; atomicOr(flags + 1, 1); }
    } else {
        // ================================ epilogue warps 0-7 ================================ //
        // Two groups of four warps (thread = one row of the tile, all NCH·64 queries), group g takes the tiles i ≡ g (mod 2): the
        // epilogue of tile i+1 runs under the epilogue of tile i, so neither the norm pass nor a survivor append is on every tile's path.
        const int grp = warp >> 2, wq = warp & 3;
        const int r_own = (wq << 5) + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        WarpSurvivors<CAPW, false> surv;
        surv.init(surv_mem + warp * WarpSurvivors<CAPW, false>::BYTES, lane);
        for (int i = grp; i < n_mine; i += 2) {
            const int st = i % NST, a = i % NACC;
            const long long row = ((long long)blockIdx.x + (long long)i * gridDim.x) * VT + r_own;
            // max ‖x‖ over the tile: thread = row, its 16 swizzled 16-byte chunks (conflict-free: 8 consecutive rows hit 8 chunks),
            // four independent partial sums
            if (!dead && !umma::mbar_wait(&bar_full[st], (i / NST) & 1)) { dead = 1; atomicOr(flags + 1, 8); }
            {
                const unsigned char* base = stages + st * STAGE_BYTES + r_own * 128;
                float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = *reinterpret_cast<const float4*>(base + h * ATOM_BYTES + ((c ^ (r_own & 7)) << 4));
                        ss[c & 3] = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss[c & 3]))));
                    }
                const float st2 = (ss[0] + ss[1]) + (ss[2] + ss[3]);
                const uint32_t mx = __reduce_max_sync(0xffffffffu, __float_as_uint(st2));
                if (lane == 0) { nrm_s[a][wq] = sqrtf(__uint_as_float(mx)) * 1.0001f; umma::mbar_arrive(&bar_empty[st]); }
            }
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");     // the four warps of the group: norms of this tile visible
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            const float nxmax = fmaxf(fmaxf(nrm_s[a][0], nrm_s[a][1]), fmaxf(nrm_s[a][2], nrm_s[a][3]));
            if (!dead && !umma::mbar_wait(&bar_done[a], (i / NACC) & 1)) { dead = 1; atomicOr(flags + 1, 2); }
            if (dead) continue;                                  // (the loop still meets the other warps at the named barrier)
            umma::fence_after_sync();
#pragma unroll
            for (int c = 0; c < NCH * 2; ++c) {                  // blocks of 32 queries
                const uint32_t acc = tmem + lane_off + (uint32_t)a * 128 + (uint32_t)c * 32;
                const int q0 = c * 32;
                float s[32];
                umma::tmem_ld32(acc, s);
                bool any = false;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(&th_s[q0 + j4 * 4]);
                    const float4 m4 = *reinterpret_cast<const float4*>(&qm_s[q0 + j4 * 4]);
                    any = any || (fmaf(m4.x, nxmax, s[j4 * 4 + 0]) > t4.x) || (fmaf(m4.y, nxmax, s[j4 * 4 + 1]) > t4.y)
                          || (fmaf(m4.z, nxmax, s[j4 * 4 + 2]) > t4.z) || (fmaf(m4.w, nxmax, s[j4 * 4 + 3]) > t4.w);
                }
                if (row >= n_rows) any = false;
                if (__any_sync(0xffffffffu, any)) {
                    uint32_t m = 0;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(&th_s[q0 + j4 * 4]);
                        const float4 m4 = *reinterpret_cast<const float4*>(&qm_s[q0 + j4 * 4]);
                        m |= (fmaf(m4.x, nxmax, s[j4 * 4 + 0]) > t4.x ? 1u : 0u) << (j4 * 4 + 0);
                        m |= (fmaf(m4.y, nxmax, s[j4 * 4 + 1]) > t4.y ? 1u : 0u) << (j4 * 4 + 1);
                        m |= (fmaf(m4.z, nxmax, s[j4 * 4 + 2]) > t4.z ? 1u : 0u) << (j4 * 4 + 2);
                        m |= (fmaf(m4.w, nxmax, s[j4 * 4 + 3]) > t4.w ? 1u : 0u) << (j4 * 4 + 3);
                    }
                    if (row >= n_rows) m = 0;
                    surv.add_block(m, acc, q0, (int)row, count, cand_s, stride, kprev, cand_r, cap, flags);   // S̃, replaced by flat_rescore_kernel
                }
            }
            umma::fence_before_sync();
            umma::mbar_arrive(&bar_free[a]);                     // the accumulator may be overwritten by tile i + NACC
        }
        surv.flush(count, cand_s, stride, kprev, cand_r, cap, flags);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, TM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

template <int NCH>
int launch_flat_stream(const CUtensorMap& map, long long n_rows, int n_tiles, const unsigned char* qimg, const float* thr, const float* qmarg,
                       int* count, float* cand_s, long long stride, int kprev, int* cand_r, int cap, int* flags, cudaStream_t st) {
    static bool attr_set = false;
    constexpr size_t smem = (size_t)NST * STAGE_BYTES + (size_t)NCH * Q_HALF + 8 * (size_t)WarpSurvivors<CAPW, false>::BYTES;
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(flat_stream_tc_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int grid = n_tiles < rb_sm_count() ? n_tiles : rb_sm_count();
    flat_stream_tc_kernel<NCH><<<grid, NT_S, smem, st>>>(map, n_rows, n_tiles, qimg, thr, qmarg, 4.f / 3.f, count, cand_s, stride, kprev,
                                                          cand_r, cap, flags);
    RB_LAUNCH_CHECK("flat_stream_tc_kernel");
    return RB200_OK;
}

}  // namespace

// is the streaming round kernel used for this many 64-query chunks?  (RB200_FLAT_STREAM=0: flat_scan_tc.cu's per-tile CTAs, 3xTF32)
bool rb_flat_streamed(int n_chunks) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("RB200_FLAT_STREAM"); on = e ? atoi(e) : 1; }
    return on && n_chunks <= 2;
}

// one round over rows [0, n_rows) of x (row-major [n_rows, 64] fp32, 16-byte aligned); survivors carry S̃ and need rb_flat_rescore
int rb_flat_stream_tc(const float* x, long long n_rows, const unsigned char* qimg, int n_chunks, const float* thr, const float* qmarg, int* count,
                      float* cand_s, long long stride, int kprev, int* cand_r, int cap, int* flags, cudaStream_t st) {
    RB_REQUIRE(n_rows >= 1 && n_rows < (1ll << 31) && n_chunks >= 1 && n_chunks <= 2, "flat_stream: 1..2^31 rows, at most 128 queries");
    EncodeTiledFn enc = encode_tiled();
    RB_REQUIRE(enc != nullptr, "flat_stream: cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)DD, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)DD * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)VT};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RB_REQUIRE(r == CUDA_SUCCESS, "flat_stream: cuTensorMapEncodeTiled failed (%d)", (int)r);
    const int n_tiles = (int)((n_rows + VT - 1) / VT);
    if (n_chunks == 1)
        return launch_flat_stream<1>(map, n_rows, n_tiles, qimg, thr, qmarg, count, cand_s, stride, kprev, cand_r, cap, flags, st);
    return launch_flat_stream<2>(map, n_rows, n_tiles, qimg, thr, qmarg, count, cand_s, stride, kprev, cand_r, cap, flags, st);
}
