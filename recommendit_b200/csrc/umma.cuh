// tcgen05 (5th-gen tensor core) building blocks for sm_100a, inline PTX.
//
//   - TMEM allocation, tcgen05.mma.kind::tf32 with both operands in shared memory ("SS"), tcgen05.commit → mbarrier,
//     tcgen05.ld of the fp32 accumulator (32x32b: one TMEM lane = one accumulator row per thread).
//   - Shared-memory operand layout: the canonical *no-swizzle, K-major* UMMA layout.  In 16-byte units
//     ((8,n),2):((1,SBO),LBO) — a "core matrix" is 8 rows × 16 B stored contiguously (128 B); SBO is the byte distance
//     between consecutive 8-row groups, LBO between the two 16-byte K chunks one K=8 (tf32) instruction consumes.
//     We store an [R rows × KC floats] operand chunk core-matrix-major:  offset(r, k) =
//         ((k/4)·(R/8) + r/8)·128 + (r%8)·16 + (k%4)·4      ⇒  SBO = 128 B, LBO = (R/8)·128 B,
//     so any thread can drop a float4 (4 consecutive k of one row) with one 16-byte store.
//   - "3xTF32": fp32 operands are split hi = tf32(x), lo = x − hi; a·b ≈ a_hi·b_hi + a_hi·b_lo + a_lo·b_hi with fp32
//     accumulation in TMEM gives ~2^-21 relative error per product (fp32-grade), at 3 MMAs per logical MMA.
//
// Every mbarrier wait is BOUNDED: a wrong descriptor must never hang the GPU; on timeout the kernel sets an error flag
// and carries on (results are garbage, the host reports the flag).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a converged warp (the form the compiler recognises: code predicated on it is issued once, with warp-uniform operands
// kept in uniform registers — a `lane == 0` branch makes ptxas wrap every tcgen05.mma in an elect/R2UR/branch loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// the MMA-issuing lane of warp 0: every lane of warp 0 must reach this call converged (it follows a __syncthreads everywhere)
__device__ __forceinline__ bool elect_issuer(int tid) { return (tid >> 5) == 0 && elect_one(); }

// ---- TMEM ------------------------------------------------------------------------------------------------ //
// one full warp; writes the TMEM base address to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes → visible to the async proxy (tensor core reads operands through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier ---------------------------------------------------------------------------------------------- //
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// returns false on timeout
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, uint32_t max_polls = 4000000u) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t i = 0; i < max_polls; ++i) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- bulk async copy (TMA engine, 1-D: no tensor map) ------------------------------------------------------- //
// arrive(1) + expect `bytes` of asynchronous copy traffic on bar
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global → shared, `bytes` multiple of 16, both addresses 16-byte aligned; completion is signalled on bar (complete_tx)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// all previously issued tcgen05.mma of this thread → arrive(1) on bar when they complete
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- descriptors ------------------------------------------------------------------------------------------- //
// instruction descriptor, kind::tf32, fp32 accumulate, both operands K-major (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major = false, bool b_mn_major = false) {
    return (1u << 4) /*c = F32*/ | (2u << 7) /*a = TF32*/ | (2u << 10) /*b = TF32*/ | ((a_mn_major ? 1u : 0u) << 15) |
           ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// NOTE (measured on B200, driver 580): with the MN-major bits (15/16) set and a no-swizzle MN-major operand layout
// (8 K-rows × 16 B core matrices) kind::tf32 MMAs complete but leave the accumulator at zero.  All operands in this
// library are therefore K-major; operands whose global layout is MN-contiguous (batch-major activations in the weight
// gradients) are transposed while being staged (tower_tc.cu::put_block_t).
// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor bit layout, version 1 = Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// byte offset of element (r, k) in an [R × KC] operand chunk stored core-matrix-major (see header)
__device__ __forceinline__ uint32_t kmajor_offset(int R, int r, int k) {
    return (uint32_t)((((k >> 2) * (R >> 3) + (r >> 3)) << 7) + ((r & 7) << 4) + ((k & 3) << 2));
}

// D[tmem] (+)= A[smem] · B[smem]ᵀ, M×N×8 (tf32).  Issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u, z = 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}

// D[tmem] (+)= A[tmem] · B[smem]ᵀ, M×N×8 (tf32), A taken from TMEM (lane = row, one 32-bit column per k): the tensor core reads
// only B from shared memory.  Issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u, z = 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}

// ---- kind::f16 with bf16 operands (fp32 accumulate): M×N×16 per instruction.  A 16-byte chunk of a K-major core-matrix row holds 8
// consecutive k; a TMEM A operand packs two consecutive k per 32-bit column (low half = even k), 8 columns per instruction.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) /*c = F32*/ | (1u << 7) /*a = BF16*/ | (1u << 10) /*b = BF16*/ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u, z = 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u, z = 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z)
        : "memory");
}

// registers → 16 consecutive 32-bit columns of this thread's lane (no wait: follow the last store with tmem_st_wait)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
// registers → 8 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMEM → registers --------------------------------------------------------------------------------------- //
// 32 consecutive fp32 columns of this thread's lane (warp w may only touch lanes 32·(w%4) … +31)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// one fp32 column of this thread's lane (warp-collective: every lane passes the same column)
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return __uint_as_float(r);
}

// 16 consecutive fp32 columns of this thread's lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- 3xTF32 split ------------------------------------------------------------------------------------------- //
// round-to-nearest (ties away from zero) tf32 with the low 13 mantissa bits zero — what cvt.rna.tf32.f32 returns, done on
// the integer pipe: adding half a tf32 ulp to the sign-magnitude bit pattern and truncating IS that rounding (the carry into
// the exponent is the correct result).  The conversion instruction issues at a fraction of the ALU rate and showed up as the
// top stall in the staging code of every tensor-core kernel.
__device__ __forceinline__ float tf32_hi(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo) {
    hi.x = tf32_hi(x.x); hi.y = tf32_hi(x.y); hi.z = tf32_hi(x.z); hi.w = tf32_hi(x.w);
    lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;   // exact in fp32; the MMA truncates it to tf32
}

}  // namespace umma
