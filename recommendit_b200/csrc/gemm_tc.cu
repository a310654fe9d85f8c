// C[M,N] = A[M,K] · B[N,K]ᵀ on the 5th-generation tensor cores (tcgen05.mma.kind::tf32, fp32 accumulate in TMEM).
//
// Used where the hot path has a plain "scores" product: the IVF coarse quantizer (queries × centroids,
// src/models/faiss_index.py:113 → faiss IndexFlatIP), exhaustive search chunks (BASELINE cfg 5) and the in-batch score
// matrix.  mode 1 = single TF32 (≈1e-3 relative, stated fast mode), mode 2 = 3xTF32 error-compensated (fp32-grade).
//
// One CTA (128 threads) per 128×128 output tile; K is consumed in chunks of 32 floats staged in shared memory in the
// canonical no-swizzle K-major UMMA layout (umma.cuh); thread 0 issues the MMAs, completion comes back through an
// mbarrier (tcgen05.commit); the epilogue reads the accumulator with tcgen05.ld, one TMEM lane (= output row) per thread.
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int TM = 128, TN = 128, KC = 32;
constexpr int CHUNK_BYTES = TM * KC * 4;              // 16 KB per operand chunk
constexpr uint32_t SBO = 128, LBO = (TM / 8) * 128;   // see umma.cuh

template <int MODE>
__global__ void __launch_bounds__(128, 1) gemm_nt_tc_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ C, int M, int N, int K, long long ldc,
                                                            int* __restrict__ err_flag) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA_hi = smem;
    unsigned char* sA_lo = smem + CHUNK_BYTES;
    unsigned char* sB_hi = smem + 2 * CHUNK_BYTES;
    unsigned char* sB_lo = smem + 3 * CHUNK_BYTES;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    __shared__ int dead;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, TN);
    if (tid == 0) { umma::mbar_init(&mbar, 1); umma::fence_mbar_init(); dead = 0; }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const long long m0 = (long long)blockIdx.y * TM, n0 = (long long)blockIdx.x * TN;
    constexpr uint32_t idesc = umma::idesc_tf32(TM, TN);
    uint32_t phase = 0;
    const int n_chunks = (K + KC - 1) / KC;
    for (int c = 0; c < n_chunks; ++c) {
        const int k0 = c * KC;
        // stage A and B chunks: thread = row, loop over the 8 float4 of the chunk
        {
            const long long ra = m0 + tid, rb = n0 + tid;
#pragma unroll
            for (int c4 = 0; c4 < KC / 4; ++c4) {
                const int k = k0 + c4 * 4;
                float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
                if (k < K) {
                    if (ra < M) va = __ldg(reinterpret_cast<const float4*>(A + ra * K + k));
                    if (rb < N) vb = __ldg(reinterpret_cast<const float4*>(B + rb * K + k));
                }
                const uint32_t off = umma::kmajor_offset(TM, tid, c4 * 4);
                float4 hi, lo;
                umma::split4(va, hi, lo);
                *reinterpret_cast<float4*>(sA_hi + off) = hi;
                if (MODE == 2) *reinterpret_cast<float4*>(sA_lo + off) = lo;
                umma::split4(vb, hi, lo);
                *reinterpret_cast<float4*>(sB_hi + off) = hi;
                if (MODE == 2) *reinterpret_cast<float4*>(sB_lo + off) = lo;
            }
        }
        umma::fence_proxy_async();
        __syncthreads();
        if (!dead && umma::elect_issuer(tid)) {
            umma::fence_after_sync();
            const uint32_t a_hi = umma::smem_u32(sA_hi), a_lo = umma::smem_u32(sA_lo);
            const uint32_t b_hi = umma::smem_u32(sB_hi), b_lo = umma::smem_u32(sB_lo);
#pragma unroll
            for (int j = 0; j < KC / 8; ++j) {
                const uint32_t ko = 2 * j * LBO;
                const bool first = (c == 0 && j == 0);
                if (MODE == 2) {
                    umma::mma_tf32(tmem, umma::smem_desc(a_lo + ko, LBO, SBO), umma::smem_desc(b_hi + ko, LBO, SBO), idesc, !first);
                    umma::mma_tf32(tmem, umma::smem_desc(a_hi + ko, LBO, SBO), umma::smem_desc(b_lo + ko, LBO, SBO), idesc, true);
                    umma::mma_tf32(tmem, umma::smem_desc(a_hi + ko, LBO, SBO), umma::smem_desc(b_hi + ko, LBO, SBO), idesc, true);
                } else {
                    umma::mma_tf32(tmem, umma::smem_desc(a_hi + ko, LBO, SBO), umma::smem_desc(b_hi + ko, LBO, SBO), idesc, !first);
                }
            }
            umma::commit(&mbar);
        }
        // the chunk buffers are reused: wait until the tensor core has consumed them
        if (!dead) {
            if (!umma::mbar_wait(&mbar, phase)) { dead = 1; if (err_flag) atomicOr(err_flag, 2); }
        }
        phase ^= 1;
        umma::fence_after_sync();
        __syncthreads();
    }
    // epilogue: TMEM lane = output row.  The tile goes through shared memory (the operand buffers are free now) so that
    // global stores are 128-byte row segments instead of one element per row per instruction.
    float* tile = reinterpret_cast<float*>(smem);          // [128][33]
    const int lane = tid & 31;
#pragma unroll 1
    for (int cb = 0; cb < TN / 32; ++cb) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
        __syncthreads();                                    // previous block's readers are done
#pragma unroll
        for (int i = 0; i < 32; ++i) tile[tid * 33 + i] = v[i];
        __syncthreads();
        const long long col = n0 + cb * 32 + lane;
        for (int r = warp; r < TM; r += 4) {
            const long long row = m0 + r;
            if (row < M && col < N) C[row * ldc + col] = tile[r * 33 + lane];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tmem, TN);
}

}  // namespace

int rb_gemm_nt_tc(const float* A, int M, const float* B, int N, int K, int mode, float* C, long long ldc, int* err_flag,
                  cudaStream_t st) {
    RB_REQUIRE(A && B && C && M >= 1 && N >= 1 && K >= 4 && K % 4 == 0, "gemm_nt: bad arguments (K must be a multiple of 4)");
    RB_REQUIRE(mode == 1 || mode == 2, "gemm_nt: tensor-core modes are 1 (TF32) and 2 (3xTF32)");
    RB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0, "gemm_nt: operands must be 16-byte aligned");
    const dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
    RB_REQUIRE(grid.y <= 65535, "gemm_nt: M too large for one launch");
    const size_t smem = 4 * CHUNK_BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RB_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    if (mode == 1) gemm_nt_tc_kernel<1><<<grid, 128, smem, st>>>(A, B, C, M, N, K, ldc, err_flag);
    else gemm_nt_tc_kernel<2><<<grid, 128, smem, st>>>(A, B, C, M, N, K, ldc, err_flag);
    RB_LAUNCH_CHECK("gemm_nt_tc_kernel");
    return RB200_OK;
}

extern "C" int rb200_gemm_nt(const float* A, int M, const float* B, int N, int K, int mode, float* C, int64_t ldc, int* err_flag,
                             void* stream) {
    return rb_gemm_nt_tc(A, M, B, N, K, mode, C, ldc, err_flag, (cudaStream_t)stream);
}
