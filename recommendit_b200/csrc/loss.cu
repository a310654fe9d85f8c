// BPR losses, fp32 FFMA parity path (sm_100a).
//
//   rb200_bpr_pair    : TwoTowerModel.bpr_loss          (src/models/two_tower.py:117-130)
//   rb200_bpr_inbatch : TwoTowerModel.in_batch_bpr_loss (src/models/two_tower.py:132-160), mode 0
//                       (modes 1/2 — tcgen05 — live in inbatch_tc.cu)
//
// The in-batch loss is evaluated tile by tile: the B×B score matrix S = U·Iᵀ exists only as 64×64
// tiles in shared memory / registers, both in the forward and in the recompute backward.
#include "common.cuh"

int rb_inbatch_tc(const float* U, const float* I, int B, int D, int mode, float* loss, float* dU, float* dI,
                  float grad_scale, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t rb_inbatch_tc_workspace_bytes(int B, int D);

namespace {

constexpr int NT = 256;

// deterministic block-wide sum (fixed shuffle tree, warp partials added in warp order)
__device__ __forceinline__ double block_sum(double v, double* scratch /*[NT/32]*/) {
    v = rb_warp_sum_d(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w];
    return t;   // valid on thread 0
}

// one warp: lane-strided sums + fixed shuffle tree (deterministic)
__global__ void finalize_sum_kernel(const double* __restrict__ partials, int n, double scale, float* __restrict__ out,
                                    rb200_opt_state* __restrict__ opt) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) t += partials[i];
    t = rb_warp_sum_d(t);
    if (threadIdx.x == 0) {
        out[0] = (float)(t * scale);
        if (opt) opt->loss = out[0];
    }
}

// ---------------------------------------------------------------------------------------- //
// pairwise BPR: one warp per sample
// ---------------------------------------------------------------------------------------- //
__global__ void __launch_bounds__(NT) bpr_pair_kernel(const float* __restrict__ u, const float* __restrict__ p,
                                                      const float* __restrict__ n, int B, int D, float gscale,
                                                      float* __restrict__ du, float* __restrict__ dp,
                                                      float* __restrict__ dn, double* __restrict__ partials) {
    __shared__ double scratch[NT / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = NT / 32;
    double lsum = 0.0;
    for (int row = blockIdx.x * wpb + warp; row < B; row += gridDim.x * wpb) {
        const float* ur = u + (long long)row * D;
        const float* pr = p + (long long)row * D;
        const float* nr = n + (long long)row * D;
        float pos = 0.f, neg = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float uv = __ldg(ur + d);
            pos = fmaf(uv, __ldg(pr + d), pos);
            neg = fmaf(uv, __ldg(nr + d), neg);
        }
        pos = rb_warp_sum(pos);
        neg = rb_warp_sum(neg);
        const float diff = pos - neg;
        if (lane == 0) lsum += (double)rb_softplus(-diff);
        if (du) {
            const float g = -rb_sigmoid(-diff) * gscale;   // dL/d(diff), gscale = grad_scale / B
            for (int d = lane; d < D; d += 32) {
                const float uv = __ldg(ur + d), pv = __ldg(pr + d), nv = __ldg(nr + d);
                du[(long long)row * D + d] = g * (pv - nv);
                dp[(long long)row * D + d] = g * uv;
                dn[(long long)row * D + d] = -g * uv;
            }
        }
    }
    const double t = block_sum(lsum, scratch);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// ---------------------------------------------------------------------------------------- //
// in-batch BPR (mode 0).  X = "own" rows (a 64-row tile resident in the CTA), Y = streamed tiles.
//   pass A: X = U, Y = I : m = S[x][y] − diag[x];  loss, rowsum r_x = Σ_y G[x][y], dU
//   pass B: X = I, Y = U : m = S[y][x] − diag[y];  dI  (uses rowsum from pass A)
// with G = σ(m)/(B(B−1)) off the diagonal, and the diagonal term  −r·(other row)  added at the end.
// ---------------------------------------------------------------------------------------- //
constexpr int TT = 64;

template <int RPT, bool ZERO>
__device__ __forceinline__ void tile_nn(const float* __restrict__ As, int lda, int row0, const float* __restrict__ Bs,
                                        int ldb, int col0, int K, float (&acc)[RPT][4]) {
    if (ZERO) {
#pragma unroll
        for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    }
    const float* a_base = As + row0 * lda;
    const float* b_base = Bs + col0;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(b_base + (k + 0) * ldb);
        const float4 b1 = *reinterpret_cast<const float4*>(b_base + (k + 1) * ldb);
        const float4 b2 = *reinterpret_cast<const float4*>(b_base + (k + 2) * ldb);
        const float4 b3 = *reinterpret_cast<const float4*>(b_base + (k + 3) * ldb);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(a_base + r * lda + k);
            acc[r][0] = fmaf(a.w, b3.x, fmaf(a.z, b2.x, fmaf(a.y, b1.x, fmaf(a.x, b0.x, acc[r][0]))));
            acc[r][1] = fmaf(a.w, b3.y, fmaf(a.z, b2.y, fmaf(a.y, b1.y, fmaf(a.x, b0.y, acc[r][1]))));
            acc[r][2] = fmaf(a.w, b3.z, fmaf(a.z, b2.z, fmaf(a.y, b1.z, fmaf(a.x, b0.z, acc[r][2]))));
            acc[r][3] = fmaf(a.w, b3.w, fmaf(a.z, b2.w, fmaf(a.y, b1.w, fmaf(a.x, b0.w, acc[r][3]))));
        }
    }
}

__global__ void __launch_bounds__(NT) rowdot_kernel(const float* __restrict__ U, const float* __restrict__ I, int B,
                                                    int D, float* __restrict__ diag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (NT / 32) + warp;
    if (row >= B) return;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(__ldg(U + (long long)row * D + d), __ldg(I + (long long)row * D + d), s);
    s = rb_warp_sum(s);
    if (lane == 0) diag[row] = s;
}

template <int RPTD, bool PASS_B, bool WITH_GRAD>   // RPTD = D/16
__global__ void __launch_bounds__(NT, 1) inbatch_kernel(const float* __restrict__ X, const float* __restrict__ Y, int B,
                                                        const float* __restrict__ diag, float inv_denom,
                                                        float* __restrict__ rowsum, float* __restrict__ dX,
                                                        double* __restrict__ partials) {
    constexpr int D = RPTD * 16;
    constexpr int ldx = ((D / 4) & 1) ? D : D + 4;      // A operand of S, rows differ across lanes
    constexpr int ldy = ldx;                            // NT operand of S, NN operand of dX
    constexpr int ldg = TT + 4;                         // 68: 17 is odd
    constexpr int NTXS = TT / 4, NTXD = D / 4;
    extern __shared__ __align__(16) float smem[];
    __shared__ double scratch[NT / 32];
    float* Xs = smem;                // [TT][ldx]
    float* Ys = Xs + TT * ldx;       // [TT][ldy]
    float* Gs = Ys + TT * ldy;       // [TT][ldg]
    float* dgx = Gs + TT * ldg;      // [TT] diag of own rows
    float* dgy = dgx + TT;           // [TT] diag of streamed rows
    float* rs = dgy + TT;            // [TT] row sums
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TT;

    for (int idx = tid; idx < TT * (D / 4); idx += NT) {
        const int r = idx / (D / 4), c = idx - r * (D / 4);
        const int gx = x0 + r;
        *reinterpret_cast<float4*>(Xs + r * ldx + c * 4) =
            gx < B ? __ldg(reinterpret_cast<const float4*>(X + (long long)gx * D) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid < TT) dgx[tid] = (x0 + tid < B) ? __ldg(diag + x0 + tid) : 0.f;

    const int txs = tid % NTXS, tys = tid / NTXS;     // S tile mapping: rows tys*4+r, cols txs+16c
    const int txd = tid % NTXD, tyd = tid / NTXD;     // dX tile mapping: rows tyd*RPTD+r, cols txd*4..
    float dacc[RPTD][4];
#pragma unroll
    for (int r = 0; r < RPTD; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) dacc[r][c] = 0.f;
    float rsum[4] = {0.f, 0.f, 0.f, 0.f};
    float lacc = 0.f;
    double lsum = 0.0;

    const int n_tiles = (B + TT - 1) / TT;
    for (int t = 0; t < n_tiles; ++t) {
        const int y0 = t * TT;
        __syncthreads();   // previous iteration's readers of Ys / Gs are done
        for (int idx = tid; idx < TT * (D / 4); idx += NT) {
            const int r = idx / (D / 4), c = idx - r * (D / 4);
            const int gy = y0 + r;
            *reinterpret_cast<float4*>(Ys + r * ldy + c * 4) =
                gy < B ? __ldg(reinterpret_cast<const float4*>(Y + (long long)gy * D) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (PASS_B && tid < TT) dgy[tid] = (y0 + tid < B) ? __ldg(diag + y0 + tid) : 0.f;
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
        {   // S[x][y] = Σ_k Xs[x][k]·Ys[y][k]
            const float* a_base = Xs + (tys * 4) * ldx;
            const float* w_base = Ys + txs * ldy;
#pragma unroll 2
            for (int k = 0; k < D; k += 4) {
                float4 b[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4*>(w_base + c * NTXS * ldy + k);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 a = *reinterpret_cast<const float4*>(a_base + r * ldx + k);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float v = acc[r][c];
                        v = fmaf(a.x, b[c].x, v); v = fmaf(a.y, b[c].y, v);
                        v = fmaf(a.z, b[c].z, v); v = fmaf(a.w, b[c].w, v);
                        acc[r][c] = v;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int lx = tys * 4 + r, gx = x0 + lx;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int ly = txs + c * NTXS, gy = y0 + ly;
                const bool live = gx < B && gy < B && gx != gy;
                const float m = acc[r][c] - (PASS_B ? dgy[ly] : dgx[lx]);
                float g = 0.f;
                if (live) {
                    if (!PASS_B) lacc += rb_softplus(m);
                    if (WITH_GRAD) g = rb_sigmoid(m) * inv_denom;
                }
                if (WITH_GRAD) {
                    Gs[lx * ldg + ly] = g;
                    if (!PASS_B) rsum[r] += g;
                }
            }
        }
        if (!PASS_B) { lsum += (double)lacc; lacc = 0.f; }
        if (WITH_GRAD) {
            __syncthreads();
            // dX[x][:] += Σ_y G[x][y]·Ys[y][:]
            tile_nn<RPTD, false>(Gs, ldg, tyd * RPTD, Ys, ldy, txd * 4, TT, dacc);
        }
    }

    if (WITH_GRAD) {
        // row sums of G over all y (pass A computes and publishes them; pass B reads the *other*
        // side's: the diagonal entry of G is −rowsum of the user row with the same index)
        __syncthreads();
        if (!PASS_B) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float v = rsum[r];
#pragma unroll
                for (int o = NTXS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(RB_FULL_MASK, v, o);
                if (txs == 0) {
                    rs[tys * 4 + r] = v;
                    if (x0 + tys * 4 + r < B) rowsum[x0 + tys * 4 + r] = v;
                }
            }
        } else if (tid < TT) {
            rs[tid] = (x0 + tid < B) ? __ldg(rowsum + x0 + tid) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RPTD; ++r) {
            const int lx = tyd * RPTD + r, gx = x0 + lx;
            if (gx < B) {
                const float4 o = __ldg(reinterpret_cast<const float4*>(Y + (long long)gx * D) + txd);
                const float rr = rs[lx];
                reinterpret_cast<float4*>(dX + (long long)gx * D)[txd] =
                    make_float4(dacc[r][0] - rr * o.x, dacc[r][1] - rr * o.y, dacc[r][2] - rr * o.z, dacc[r][3] - rr * o.w);
            }
        }
    }
    if (!PASS_B) {
        const double t = block_sum(lsum, scratch);
        if (tid == 0) partials[blockIdx.x] = t;
    }
}

size_t inbatch_smem(int D) {
    const int ld = rb_ld_odd4(D);
    return sizeof(float) * ((size_t)2 * TT * ld + (size_t)TT * (TT + 4) + 3 * TT);
}

template <int RPTD>
int launch_inbatch(const float* U, const float* I, int B, const float* diag, float inv, float* rowsum, float* dU,
                   float* dI, double* partials, cudaStream_t st) {
    constexpr int D = RPTD * 16;
    const size_t smem = inbatch_smem(D);
    const int grid = (B + TT - 1) / TT;
    static bool attr_set = false;
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(inbatch_kernel<RPTD, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RB_CUDA(cudaFuncSetAttribute(inbatch_kernel<RPTD, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RB_CUDA(cudaFuncSetAttribute(inbatch_kernel<RPTD, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    if (dU) {
        inbatch_kernel<RPTD, false, true><<<grid, NT, smem, st>>>(U, I, B, diag, inv, rowsum, dU, partials);
        RB_LAUNCH_CHECK("inbatch_kernel<A>");
        inbatch_kernel<RPTD, true, true><<<grid, NT, smem, st>>>(I, U, B, diag, inv, rowsum, dI, nullptr);
        RB_LAUNCH_CHECK("inbatch_kernel<B>");
    } else {
        inbatch_kernel<RPTD, false, false><<<grid, NT, smem, st>>>(U, I, B, diag, inv, nullptr, nullptr, partials);
        RB_LAUNCH_CHECK("inbatch_kernel<fwd>");
    }
    return RB200_OK;
}

}  // namespace

extern "C" size_t rb200_bpr_pair_workspace_bytes(int B) {
    (void)B;
    return 256 + sizeof(double) * (size_t)(rb_sm_count() * 8);
}

// The gradient kernel runs on `st`; the loss reduction (block partials → loss, optionally also opt->loss) runs on
// `st_fin` after `fork` (both may be NULL/equal to st: plain sequential).  csrc/step.cu passes its side stream so the
// reduction leaves the critical path of the training step.
int rb_bpr_pair(const float* u, const float* p, const float* n, int B, int D, float* loss, float* du, float* dp, float* dn,
                float grad_scale, float loss_scale, void* workspace, size_t workspace_bytes, rb200_opt_state* opt, cudaStream_t st,
                cudaStream_t st_fin, cudaEvent_t fork);

extern "C" int rb200_bpr_pair(const float* u, const float* p, const float* n, int B, int D, float* loss, float* du,
                              float* dp, float* dn, float grad_scale, void* workspace, size_t workspace_bytes,
                              void* stream) {
    return rb_bpr_pair(u, p, n, B, D, loss, du, dp, dn, grad_scale, 1.f, workspace, workspace_bytes, nullptr, (cudaStream_t)stream,
                       (cudaStream_t)stream, nullptr);
}

int rb_bpr_pair(const float* u, const float* p, const float* n, int B, int D, float* loss, float* du, float* dp, float* dn,
                float grad_scale, float loss_scale, void* workspace, size_t workspace_bytes, rb200_opt_state* opt, cudaStream_t st,
                cudaStream_t st_fin, cudaEvent_t fork) {
    RB_REQUIRE(u && p && n && loss && B >= 1 && D >= 1, "bpr_pair: bad arguments");
    RB_REQUIRE((du == nullptr) == (dp == nullptr) && (du == nullptr) == (dn == nullptr), "bpr_pair: du/dp/dn must be all set or all NULL");
    int grid = (B + NT / 32 - 1) / (NT / 32);
    const int cap = rb_sm_count() * 8;          // 8192 samples → one sample per warp, every warp resident at once
    if (grid > cap) grid = cap;
    RbArena ar(workspace, workspace_bytes);
    double* partials = ar.take<double>(cap);
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "bpr_pair: workspace too small");
    bpr_pair_kernel<<<grid, NT, 0, st>>>(u, p, n, B, D, grad_scale / (float)B, du, dp, dn, partials);
    RB_LAUNCH_CHECK("bpr_pair_kernel");
    if (fork && st_fin != st) {
        RB_CUDA(cudaEventRecord(fork, st));
        RB_CUDA(cudaStreamWaitEvent(st_fin, fork, 0));
    }
    finalize_sum_kernel<<<1, 32, 0, st_fin>>>(partials, grid, (double)loss_scale / (double)B, loss, opt);
    RB_LAUNCH_CHECK("finalize_sum_kernel");
    return RB200_OK;
}

extern "C" size_t rb200_bpr_inbatch_workspace_bytes(int B, int D) {
    const size_t tiles = (size_t)(B + TT - 1) / TT;
    size_t simt = 256 * 4 + sizeof(float) * 2 * (size_t)B + sizeof(double) * tiles;
    size_t tc = rb_inbatch_tc_workspace_bytes(B, D);
    return simt > tc ? simt : tc;
}

extern "C" int rb200_bpr_inbatch(const float* U, const float* I, int B, int D, int mode, float* loss, float* dU,
                                 float* dI, float grad_scale, void* workspace, size_t workspace_bytes, void* stream) {
    RB_REQUIRE(U && I && loss && B >= 1, "bpr_inbatch: bad arguments");
    RB_REQUIRE((dU == nullptr) == (dI == nullptr), "bpr_inbatch: dU/dI must both be set or both NULL");
    RB_REQUIRE(D == 32 || D == 64 || D == 128, "bpr_inbatch: D must be 32, 64 or 128 (got %d)", D);
    cudaStream_t st = (cudaStream_t)stream;
    // (B = 1 has no negatives: the loss is NaN as in the reference loop — left to the SIMT kernel)
    if (mode != 0 && B >= 2) return rb_inbatch_tc(U, I, B, D, mode, loss, dU, dI, grad_scale, workspace, workspace_bytes, st);
    const int tiles = (B + TT - 1) / TT;
    RbArena ar(workspace, workspace_bytes);
    float* diag = ar.take<float>(B);
    float* rowsum = ar.take<float>(B);
    double* partials = ar.take<double>(tiles);
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "bpr_inbatch: workspace too small");
    rowdot_kernel<<<(B + NT / 32 - 1) / (NT / 32), NT, 0, st>>>(U, I, B, D, diag);
    RB_LAUNCH_CHECK("rowdot_kernel");
    const double denom = (double)B * (double)(B - 1);
    const float inv = (float)((double)grad_scale / denom);
    int rc;
    switch (D) {
        case 32: rc = launch_inbatch<2>(U, I, B, diag, inv, rowsum, dU, dI, partials, st); break;
        case 64: rc = launch_inbatch<4>(U, I, B, diag, inv, rowsum, dU, dI, partials, st); break;
        default: rc = launch_inbatch<8>(U, I, B, diag, inv, rowsum, dU, dI, partials, st); break;
    }
    if (rc) return rc;
    finalize_sum_kernel<<<1, 32, 0, st>>>(partials, tiles, 1.0 / denom, loss, nullptr);
    RB_LAUNCH_CHECK("finalize_sum_kernel");
    return RB200_OK;
}
