// Two-tower MLP kernels, fp32 FFMA parity path (sm_100a).
//
//   forward  : gather(table, ids) ++ extra → Linear → ReLU → dropout → Linear → L2-normalise
//              (src/models/two_tower.py:39-42, :68-72 of the reference), one launch for up to three
//              tower evaluations (user / positive items / negative items).
//   backward : normalize-bwd → dH = dpre·W2 → ReLU/dropout mask → dRows = dact·W1[:, :D]   (data)
//              dW2 = dpreᵀ·hid, dW1 = dactᵀ·X, db = column sums, split over the batch with a
//              fixed-order two-stage reduction                                             (weights)
//
// Layout: a CTA of 256 threads owns a tile of TM = 64 samples.  Both weight matrices stay resident
// in shared memory for the CTA's lifetime (persistent over tiles); activations live in shared
// memory between the two GEMMs and never touch HBM in inference.  Register-tiled 128-bit
// shared-memory operand loads; leading dimensions are chosen so that the loads are conflict-free.
#include "common.cuh"
#include "tower_common.cuh"

namespace {

constexpr int TM = 64;    // samples per tile
constexpr int NT = 256;   // threads per CTA

// C[r][c] (r < RPT rows starting at row0; 4 columns col0 + c*cstride) = Σ_k A[row][k] · W[col][k]
// A: [rows][lda] row-major, W: [cols][ldw] row-major (torch Linear layout) — "NT" product.
template <int RPT>
__device__ __forceinline__ void gemm_nt(const float* __restrict__ As, int lda, int row0,
                                        const float* __restrict__ Ws, int ldw, int col0, int cstride,
                                        int K, float (&acc)[RPT][4]) {
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    const float* a_base = As + row0 * lda;
    const float* w_base = Ws + col0 * ldw;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 b[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4*>(w_base + c * cstride * ldw + k);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(a_base + r * lda + k);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v = acc[r][c];
                v = fmaf(a.x, b[c].x, v);
                v = fmaf(a.y, b[c].y, v);
                v = fmaf(a.z, b[c].z, v);
                v = fmaf(a.w, b[c].w, v);
                acc[r][c] = v;
            }
        }
    }
}

// C[r][c] (4 contiguous columns col0..col0+3) = Σ_k A[row][k] · Bm[k][col] — "NN" product.
template <int RPT>
__device__ __forceinline__ void gemm_nn(const float* __restrict__ As, int lda, int row0,
                                        const float* __restrict__ Bs, int ldb, int col0, int K,
                                        float (&acc)[RPT][4]) {
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
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

// Gather TM rows of [table[ids] ; extra ; 0-pad] into Xs[TM][ldx].  One warp per row at a time.
template <int D>
__device__ __forceinline__ void gather_tile(float* Xs, int ldx, int Kp, const float* __restrict__ table,
                                            const int64_t* __restrict__ ids, const float* __restrict__ extra,
                                            int E, int extra_by_id, long long n_rows, int row0, int B, int rows,
                                            int* err_flag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < rows; r += NT / 32) {
        const int row = row0 + r;
        const bool valid = row < B;
        long long id = valid ? ids[row] : 0;
        if ((unsigned long long)id >= (unsigned long long)n_rows) {
            if (lane == 0 && err_flag) atomicOr(err_flag, 1);
            id = 0;
        }
        const float4* src = reinterpret_cast<const float4*>(table + id * D);
        float4* dst = reinterpret_cast<float4*>(Xs + r * ldx);
        for (int c = lane; c < D / 4; c += 32) dst[c] = valid ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = lane; e < Kp - D; e += 32)
            Xs[r * ldx + D + e] = (valid && e < E) ? __ldg(extra + (extra_by_id ? id : (long long)row) * E + e) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------ //
// forward
// ------------------------------------------------------------------------------------------ //
template <int RPT1, int RPT2>   // RPT1 = H/16, RPT2 = D/16
__global__ void __launch_bounds__(NT, 1) tower_fwd_kernel(const FwdParams p) {
    constexpr int H = RPT1 * 16, D = RPT2 * 16;
    constexpr int NTX1 = H / 4, NTX2 = D / 4;
    constexpr int ldw2 = ((H / 4) & 1) ? H : H + 4;
    constexpr int ldh = ldw2;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;

    int j = 0;
#pragma unroll
    for (int t = 1; t < MAX_JOBS; ++t)
        if (t < p.n_jobs && (int)blockIdx.x >= p.job[t].cta_begin) j = t;
    const FwdJob J = p.job[j];
    const int E = J.E, Din = D + E, Kp = (Din + 3) & ~3;
    const int ldw1 = rb_ld_odd4(Kp), ldx = ldw1;

    float* W1s = smem;                 // [H][ldw1]
    float* W2s = W1s + H * ldw1;       // [D][ldw2]
    float* b1s = W2s + D * ldw2;       // [H]
    float* b2s = b1s + H;              // [D]
    float* Xs = b2s + D;               // [TM][ldx]
    float* Hs = Xs + TM * ldx;         // [TM][ldh]

    for (int idx = tid; idx < H * Kp; idx += NT) {
        const int h = idx / Kp, k = idx - h * Kp;
        W1s[h * ldw1 + k] = (k < Din) ? __ldg(J.W1 + h * Din + k) : 0.f;
    }
    for (int idx = tid; idx < D * (H / 4); idx += NT) {
        const int d = idx / (H / 4), k4 = idx - d * (H / 4);
        *reinterpret_cast<float4*>(W2s + d * ldw2 + k4 * 4) = __ldg(reinterpret_cast<const float4*>(J.W2 + d * H) + k4);
    }
    for (int idx = tid; idx < H; idx += NT) b1s[idx] = __ldg(J.b1 + idx);
    for (int idx = tid; idx < D; idx += NT) b2s[idx] = __ldg(J.b2 + idx);

    const int B = J.B;
    const int n_tiles = (B + TM - 1) / TM;
    const bool do_drop = p.drop_p > 0.f;
    const float keep_scale = do_drop ? 1.f / (1.f - p.drop_p) : 1.f;
    const unsigned long long drop_off = p.offset + (unsigned long long)j +
                                        (p.offset_dev ? (unsigned long long)__ldg(p.offset_dev) * MAX_JOBS : 0ull);

    for (int tile = (int)blockIdx.x - J.cta_begin; tile < n_tiles; tile += J.cta_count) {
        const int row0 = tile * TM;
        gather_tile<D>(Xs, ldx, Kp, J.table, J.ids, J.extra, E, J.extra_by_id, J.n_rows, row0, B, TM, p.err_flag);
        __syncthreads();
        {   // hidden = dropout(relu(X·W1ᵀ + b1))
            const int tx = tid % NTX1, ty = tid / NTX1;
            float acc[RPT1][4];
            gemm_nt<RPT1>(Xs, ldx, ty * RPT1, W1s, ldw1, tx, NTX1, Kp, acc);
#pragma unroll
            for (int r = 0; r < RPT1; ++r) {
                const int lr = ty * RPT1 + r, row = row0 + lr;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int col = tx + c * NTX1;
                    float v = fmaxf(acc[r][c] + b1s[col], 0.f);
                    if (do_drop) {
                        bool keep;
                        if (J.keep_mask) keep = (row < B) ? (J.keep_mask[(long long)row * H + col] != 0) : true;
                        else keep = rb_dropout_keep(p.seed, drop_off, row, col, p.drop_p);
                        v = keep ? v * keep_scale : 0.f;
                    }
                    Hs[lr * ldh + col] = v;
                    if (J.hid && row < B) J.hid[(long long)row * H + col] = v;
                }
            }
        }
        __syncthreads();
        {   // y = normalize(hidden·W2ᵀ + b2)
            const int tx = tid % NTX2, ty = tid / NTX2;
            float acc[RPT2][4];
            gemm_nt<RPT2>(Hs, ldh, ty * RPT2, W2s, ldw2, tx, NTX2, H, acc);
#pragma unroll
            for (int r = 0; r < RPT2; ++r) {
                const int row = row0 + ty * RPT2 + r;
                float ss = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc[r][c] += b2s[tx + c * NTX2];
                    ss = fmaf(acc[r][c], acc[r][c], ss);
                }
#pragma unroll
                for (int o = NTX2 / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(RB_FULL_MASK, ss, o);
                const float den = fmaxf(sqrtf(ss), NORM_EPS);
                if (row < B) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) J.out[(long long)row * D + tx + c * NTX2] = acc[r][c] / den;
                    if (J.denom && tx == 0) J.denom[row] = den;
                }
            }
        }
        // next iteration's gather writes Xs (free since the barrier above) and its first barrier
        // orders this tile's reads of Hs before the next tile's writes.
    }
}

size_t fwd_smem_bytes(int D, int H, int E) {
    const int Kp = (D + E + 3) & ~3;
    const int ldw1 = rb_ld_odd4(Kp), ldw2 = rb_ld_odd4(H);
    return sizeof(float) * ((size_t)H * ldw1 + (size_t)D * ldw2 + H + D + (size_t)TM * ldw1 + (size_t)TM * ldw2);
}

template <int RPT1, int RPT2>
int launch_fwd(const FwdParams& p, int grid, size_t smem, cudaStream_t st) {
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(tower_fwd_kernel<RPT1, RPT2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     rb_max_smem_optin()));
        attr_set = true;
    }
    tower_fwd_kernel<RPT1, RPT2><<<grid, NT, smem, st>>>(p);
    RB_LAUNCH_CHECK("tower_fwd_kernel");
    return RB200_OK;
}

// ------------------------------------------------------------------------------------------ //
// backward, data part
// ------------------------------------------------------------------------------------------ //
template <int RPT1, int RPT2>
__global__ void __launch_bounds__(NT, 1) tower_bwd_data_kernel(const BwdParams p) {
    constexpr int H = RPT1 * 16, D = RPT2 * 16;
    constexpr int NTX1 = H / 4, NTX2 = D / 4;
    constexpr int ldg = ((D / 4) & 1) ? D : D + 4;
    constexpr int lda = ((H / 4) & 1) ? H : H + 4;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    int j = 0;
#pragma unroll
    for (int t = 1; t < MAX_JOBS; ++t)
        if (t < p.n_jobs && (int)blockIdx.x >= p.job[t].cta_begin) j = t;
    const BwdJob J = p.job[j];
    const int Din = D + J.E;

    float* W2s = smem;             // [D][H]   B operand of dH = G·W2     (k = d, n = h)
    float* W1s = W2s + D * H;      // [H][D]   B operand of dX = dA·W1[:, :D]
    float* Gs = W1s + H * D;       // [TM][ldg]
    float* As = Gs + TM * ldg;     // [TM][lda]

    for (int idx = tid; idx < D * H / 4; idx += NT)
        reinterpret_cast<float4*>(W2s)[idx] = __ldg(reinterpret_cast<const float4*>(J.W2) + idx);
    for (int idx = tid; idx < H * D; idx += NT) {
        const int h = idx / D, d = idx - h * D;
        W1s[idx] = __ldg(J.W1 + h * Din + d);
    }

    const int B = J.B;
    const int n_tiles = (B + TM - 1) / TM;
    for (int tile = (int)blockIdx.x - J.cta_begin; tile < n_tiles; tile += J.cta_count) {
        const int row0 = tile * TM;
        // normalize backward: dpre = (g − y (y·g)) / denom   (g / denom when the eps clamp was active)
        for (int r = warp; r < TM; r += NT / 32) {
            const int row = row0 + r;
            const bool valid = row < B;
            float g[D / 32], yv[D / 32];
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < D / 32; ++i) {
                g[i] = valid ? __ldg(J.dY + (long long)row * D + lane + 32 * i) : 0.f;
                yv[i] = valid ? __ldg(J.y + (long long)row * D + lane + 32 * i) : 0.f;
                dot = fmaf(g[i], yv[i], dot);
            }
            dot = rb_warp_sum(dot);
            const float den = valid ? __ldg(J.denom + row) : 1.f;
            const bool clamped = den <= NORM_EPS;
#pragma unroll
            for (int i = 0; i < D / 32; ++i) {
                const float v = clamped ? g[i] / den : (g[i] - yv[i] * dot) / den;
                Gs[r * ldg + lane + 32 * i] = v;
                if (valid) J.dpre[(long long)row * D + lane + 32 * i] = v;
            }
        }
        __syncthreads();
        {   // dact = (dpre·W2) ⊙ [hid > 0] · keep_scale
            const int tx = tid % NTX1, ty = tid / NTX1;
            float acc[RPT1][4];
            gemm_nn<RPT1>(Gs, ldg, ty * RPT1, W2s, H, tx * 4, D, acc);
#pragma unroll
            for (int r = 0; r < RPT1; ++r) {
                const int lr = ty * RPT1 + r, row = row0 + lr;
                float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < B) hv = __ldg(reinterpret_cast<const float4*>(J.hid + (long long)row * H) + tx);
                float4 o;
                o.x = hv.x > 0.f ? acc[r][0] * p.keep_scale : 0.f;
                o.y = hv.y > 0.f ? acc[r][1] * p.keep_scale : 0.f;
                o.z = hv.z > 0.f ? acc[r][2] * p.keep_scale : 0.f;
                o.w = hv.w > 0.f ? acc[r][3] * p.keep_scale : 0.f;
                *reinterpret_cast<float4*>(As + lr * lda + tx * 4) = o;
                if (row < B) reinterpret_cast<float4*>(J.dact + (long long)row * H)[tx] = o;
            }
        }
        __syncthreads();
        {   // dRows = dact·W1[:, :D]
            const int tx = tid % NTX2, ty = tid / NTX2;
            float acc[RPT2][4];
            gemm_nn<RPT2>(As, lda, ty * RPT2, W1s, D, tx * 4, H, acc);
#pragma unroll
            for (int r = 0; r < RPT2; ++r) {
                const int row = row0 + ty * RPT2 + r;
                if (row < B)
                    reinterpret_cast<float4*>(J.dRows + (long long)row * D)[tx] =
                        make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            }
        }
    }
}

size_t bwd_data_smem_bytes(int D, int H) {
    return sizeof(float) * ((size_t)2 * D * H + (size_t)TM * rb_ld_odd4(D) + (size_t)TM * rb_ld_odd4(H));
}

template <int RPT1, int RPT2>
int launch_bwd_data(const BwdParams& p, int grid, size_t smem, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(tower_bwd_data_kernel<RPT1, RPT2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     rb_max_smem_optin()));
        attr_set = true;
    }
    tower_bwd_data_kernel<RPT1, RPT2><<<grid, NT, smem, st>>>(p);
    RB_LAUNCH_CHECK("tower_bwd_data_kernel");
    return RB200_OK;
}

// ------------------------------------------------------------------------------------------ //
// backward, weight part: split over the batch, register accumulators, per-split partials
// ------------------------------------------------------------------------------------------ //
constexpr int RT = 32;   // batch rows per shared-memory stage

// blockIdx.y == 0: dW2[D][H] = Σ_r dpre[r][:]ᵀ hid[r][:],  db2 = Σ_r dpre[r]
// blockIdx.y == 1: dW1[H][Din] = Σ_r dact[r][:]ᵀ X[r][:],  db1 = Σ_r dact[r]
template <int RPT1, int RPT2, int MAXMT>
__global__ void __launch_bounds__(NT, 1) tower_bwd_weights_kernel(const BwdParams p) {
    constexpr int H = RPT1 * 16, D = RPT2 * 16;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int which = blockIdx.y, s = blockIdx.x;
    const int E = p.job[0].E, Din = D + E, Kp = (Din + 3) & ~3;
    const int M = which == 0 ? D : H;
    const int Np = which == 0 ? H : Kp;
    const int NG = Np / 4, nMT = (M / 8) * NG;
    float* As = smem;            // [RT][M]
    float* Bs = As + RT * M;     // [RT][Np]

    float acc[MAXMT][8][4];
#pragma unroll
    for (int i = 0; i < MAXMT; ++i)
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][a][c] = 0.f;
    float bias_acc = 0.f;

    for (int j = 0; j < p.n_jobs; ++j) {
        const BwdJob& J = p.job[j];
        const int chunk = (J.B + p.nsplit - 1) / p.nsplit;
        const int r_begin = min(J.B, s * chunk), r_end = min(J.B, r_begin + chunk);
        for (int r0 = r_begin; r0 < r_end; r0 += RT) {
            const int nr = min(RT, r_end - r0);
            if (which == 0) {
                for (int idx = tid; idx < RT * (D / 4); idx += NT) {
                    const int k = idx / (D / 4), c = idx - k * (D / 4);
                    reinterpret_cast<float4*>(As)[idx] =
                        k < nr ? __ldg(reinterpret_cast<const float4*>(J.dpre + (long long)(r0 + k) * D) + c)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                for (int idx = tid; idx < RT * (H / 4); idx += NT) {
                    const int k = idx / (H / 4), c = idx - k * (H / 4);
                    reinterpret_cast<float4*>(Bs)[idx] =
                        k < nr ? __ldg(reinterpret_cast<const float4*>(J.hid + (long long)(r0 + k) * H) + c)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                for (int idx = tid; idx < RT * (H / 4); idx += NT) {
                    const int k = idx / (H / 4), c = idx - k * (H / 4);
                    reinterpret_cast<float4*>(As)[idx] =
                        k < nr ? __ldg(reinterpret_cast<const float4*>(J.dact + (long long)(r0 + k) * H) + c)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                gather_tile<D>(Bs, Kp, Kp, J.table, J.ids, J.extra, E, J.extra_by_id, J.n_rows, r0, r_end, RT, nullptr);
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < MAXMT; ++i) {
                const int mt = tid + i * NT;
                if (mt < nMT) {
                    const int mg = mt / NG, ng = mt - mg * NG;
                    const float* ap = As + mg * 8;
                    const float* bp = Bs + ng * 4;
#pragma unroll 4
                    for (int k = 0; k < RT; ++k) {
                        const float4 a0 = *reinterpret_cast<const float4*>(ap + k * M);
                        const float4 a1 = *reinterpret_cast<const float4*>(ap + k * M + 4);
                        const float4 b = *reinterpret_cast<const float4*>(bp + k * Np);
                        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                        for (int a = 0; a < 8; ++a) {
                            acc[i][a][0] = fmaf(av[a], b.x, acc[i][a][0]);
                            acc[i][a][1] = fmaf(av[a], b.y, acc[i][a][1]);
                            acc[i][a][2] = fmaf(av[a], b.z, acc[i][a][2]);
                            acc[i][a][3] = fmaf(av[a], b.w, acc[i][a][3]);
                        }
                    }
                }
            }
            if (tid < M) {
#pragma unroll 8
                for (int k = 0; k < RT; ++k) bias_acc += As[k * M + tid];
            }
            __syncthreads();
        }
    }

    // partial block layout: [W1 (H*Din) | b1 (H) | W2 (D*H) | b2 (D)]
    float* part = p.part + (long long)s * p.P;
    const int N = which == 0 ? H : Din;
    float* wout = which == 0 ? part + H * Din + H : part;
    float* bout = which == 0 ? part + H * Din + H + D * H : part + H * Din;
#pragma unroll
    for (int i = 0; i < MAXMT; ++i) {
        const int mt = tid + i * NT;
        if (mt < nMT) {
            const int mg = mt / NG, ng = mt - mg * NG;
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int n = ng * 4 + c;
                    if (n < N) wout[(mg * 8 + a) * N + n] = acc[i][a][c];
                }
        }
    }
    if (tid < M) bout[tid] = bias_acc;
}

__global__ void reduce_partials_kernel(const float* __restrict__ part, int nsplit, int P, float* __restrict__ out,
                                       int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < nsplit; ++k) s += __ldg(part + (long long)k * P + i);   // fixed order; loads issued 8 deep
    out[i] = accumulate ? out[i] + s : s;
}

template <int RPT1, int RPT2, int MAXMT>
int launch_bwd_weights(const BwdParams& p, size_t smem, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        RB_CUDA(cudaFuncSetAttribute(tower_bwd_weights_kernel<RPT1, RPT2, MAXMT>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, rb_max_smem_optin()));
        attr_set = true;
    }
    tower_bwd_weights_kernel<RPT1, RPT2, MAXMT><<<dim3(p.nsplit, 2), NT, smem, st>>>(p);
    RB_LAUNCH_CHECK("tower_bwd_weights_kernel");
    return RB200_OK;
}

template <int RPT1, int RPT2>
int dispatch_bwd_weights(const BwdParams& p, int maxmt, size_t smem, cudaStream_t st) {
    switch (maxmt) {
        case 1: return launch_bwd_weights<RPT1, RPT2, 1>(p, smem, st);
        case 2: return launch_bwd_weights<RPT1, RPT2, 2>(p, smem, st);
        case 3: return launch_bwd_weights<RPT1, RPT2, 3>(p, smem, st);
        default: return rb_set_error(RB200_ERR_INVALID, "tower_bwd: weight tile too large (needs %d register tiles)", maxmt);
    }
}

bool dims_supported(int D, int H) {
    return (D == 32 || D == 64 || D == 128) && (H == 64 || H == 128 || H == 256);
}

// CTA partition between jobs, proportional to flops; returns grid size.
template <typename JobT>
int partition_ctas(JobT* jobs, int n_jobs, int D, int H) {
    const int n_sm = rb_sm_count();
    double w[MAX_JOBS], total = 0;
    int tiles[MAX_JOBS];
    for (int j = 0; j < n_jobs; ++j) {
        tiles[j] = (jobs[j].B + TM - 1) / TM;
        w[j] = (double)jobs[j].B * ((double)(D + jobs[j].E) * H + (double)H * D);
        total += w[j];
    }
    int begin = 0, left = n_sm;
    for (int j = 0; j < n_jobs; ++j) {
        int c;
        if (j == n_jobs - 1) c = left;
        else c = (int)(n_sm * w[j] / total + 0.5);
        const int jobs_after = n_jobs - 1 - j;
        if (c > left - jobs_after) c = left - jobs_after;
        if (c < 1) c = 1;
        if (c > tiles[j]) c = tiles[j] > 0 ? tiles[j] : 1;
        jobs[j].cta_begin = begin;
        jobs[j].cta_count = c;
        begin += c;
        left -= c;
        if (left < jobs_after) left = jobs_after;
    }
    return begin;
}

#define RB_DISPATCH_DH(D, H, CALL)                                               \
    switch ((H) * 1000 + (D)) {                                                  \
        case 64 * 1000 + 32: { constexpr int R1 = 4, R2 = 2; CALL; } break;      \
        case 64 * 1000 + 64: { constexpr int R1 = 4, R2 = 4; CALL; } break;      \
        case 64 * 1000 + 128: { constexpr int R1 = 4, R2 = 8; CALL; } break;     \
        case 128 * 1000 + 32: { constexpr int R1 = 8, R2 = 2; CALL; } break;     \
        case 128 * 1000 + 64: { constexpr int R1 = 8, R2 = 4; CALL; } break;     \
        case 128 * 1000 + 128: { constexpr int R1 = 8, R2 = 8; CALL; } break;    \
        case 256 * 1000 + 32: { constexpr int R1 = 16, R2 = 2; CALL; } break;    \
        case 256 * 1000 + 64: { constexpr int R1 = 16, R2 = 4; CALL; } break;    \
        case 256 * 1000 + 128: { constexpr int R1 = 16, R2 = 8; CALL; } break;   \
        default: rc = rb_set_error(RB200_ERR_INVALID, "unsupported (D=%d, H=%d)", D, H); \
    }

}  // namespace

int rb_tower_fwd_tc(FwdParams& p, int D, int H, int mode, void* workspace, size_t workspace_bytes, cudaStream_t st);
int rb_tower_bwd_tc(BwdParams& p, int D, int H, int mode, float* grads_out, int accumulate, unsigned char* img_ws, cudaStream_t st);
bool rb_tower_tc_supported(int D, int H, int E);
bool rb_tower_use_ts(int D);
size_t rb_tower_img_bytes(int D, int H, int E);

extern "C" size_t rb200_tower_fwd_workspace_bytes(int n_jobs, int D, int H, int extra_dim, int mode) {
    if (mode == 0 || !rb_tower_tc_supported(D, H, extra_dim)) return 0;
    return 256 + (size_t)(n_jobs < 2 ? n_jobs : 2) * (rb_tower_img_bytes(D, H, extra_dim) + 256);
}

extern "C" int rb200_tower_fwd(const rb200_tower_job* jobs, int n_jobs, int D, int H, float dropout_p,
                               uint64_t seed, uint64_t offset, const int64_t* offset_dev, int mode, int* err_flag, void* workspace,
                               size_t workspace_bytes, void* stream) {
    RB_REQUIRE(jobs && n_jobs >= 1 && n_jobs <= MAX_JOBS, "tower_fwd: n_jobs must be 1..3");
    RB_REQUIRE(mode >= 0 && mode <= 2, "tower_fwd: mode must be 0 (fp32 FFMA), 1 (tcgen05 TF32) or 2 (tcgen05 3xTF32)");
    RB_REQUIRE(dims_supported(D, H), "tower_fwd: unsupported widths D=%d H=%d (D in {32,64,128}, H in {64,128,256})", D, H);
    RB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "tower_fwd: dropout_p must be in [0,1)");
    FwdParams p{};
    p.n_jobs = 0; p.drop_p = dropout_p; p.seed = seed; p.offset = offset; p.offset_dev = (const long long*)offset_dev; p.err_flag = err_flag;
    size_t smem = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const rb200_tower_job& s = jobs[j];
        RB_REQUIRE(s.B >= 0 && s.extra_dim >= 0 && s.extra_dim <= 64, "tower_fwd: bad B/extra_dim");
        RB_REQUIRE(s.extra_dim == 0 || s.extra != nullptr, "tower_fwd: extra is NULL but extra_dim > 0");
        if (s.B == 0) continue;
        RB_REQUIRE(s.table && s.ids && s.W1 && s.b1 && s.W2 && s.b2 && s.out, "tower_fwd: NULL pointer in job %d", j);
        FwdJob& d = p.job[p.n_jobs++];
        d.table = s.table; d.ids = s.ids; d.extra = s.extra; d.W1 = s.W1; d.b1 = s.b1; d.W2 = s.W2; d.b2 = s.b2;
        d.out = s.out; d.hid = s.hid; d.denom = s.denom; d.keep_mask = s.keep_mask; d.img = (const unsigned char*)s.img;
        d.n_rows = s.n_rows; d.B = s.B; d.E = s.extra_dim; d.extra_by_id = s.extra_by_id;
        const size_t b = fwd_smem_bytes(D, H, s.extra_dim);
        if (b > smem) smem = b;
    }
    if (p.n_jobs == 0) return RB200_OK;
    if (mode != 0) {
        for (int j = 0; j < p.n_jobs; ++j)
            RB_REQUIRE(rb_tower_tc_supported(D, H, p.job[j].E), "tower_fwd: tcgen05 mode supports D in {64,128}, H=128, extra_dim<=24 (got D=%d H=%d E=%d)", D, H, p.job[j].E);
        return rb_tower_fwd_tc(p, D, H, mode, workspace, workspace_bytes, (cudaStream_t)stream);
    }
    RB_REQUIRE((int)smem <= rb_max_smem_optin(), "tower_fwd: D=%d H=%d needs %zu B of shared memory (> %d)", D, H, smem,
               rb_max_smem_optin());
    const int grid = partition_ctas(p.job, p.n_jobs, D, H);
    int rc = RB200_OK;
    RB_DISPATCH_DH(D, H, (rc = launch_fwd<R1, R2>(p, grid, smem, (cudaStream_t)stream)));
    return rc;
}

extern "C" size_t rb200_tower_bwd_workspace_bytes(int D, int H, int extra_dim) {
    const size_t P = (size_t)H * (D + extra_dim) + H + (size_t)D * H + D;
    // split-K partials (at most one per SM) + room for one weight image of the tcgen05 path
    return 512 + sizeof(float) * P * (size_t)rb_sm_count() + (rb_tower_tc_supported(D, H, extra_dim) ? rb_tower_img_bytes(D, H, extra_dim) + 256 : 0);
}

// `defer` != NULL (tensor-core modes only): the weight-gradient partials are NOT reduced; *defer describes them and the
// caller reduces them (csrc/step.cu → rb_grad_finish).  Returns 1 if the mode cannot defer (caller passes grads_out then).
int rb_tower_bwd(const rb200_tower_bwd_job* jobs, int n_jobs, int D, int H, float dropout_p, int mode, float* grads_out,
                 int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream, RbPartials* defer);

extern "C" int rb200_tower_bwd(const rb200_tower_bwd_job* jobs, int n_jobs, int D, int H, float dropout_p, int mode,
                               float* grads_out, int accumulate, void* workspace, size_t workspace_bytes,
                               void* stream) {
    RB_REQUIRE(grads_out != nullptr, "tower_bwd: grads_out is NULL");
    return rb_tower_bwd(jobs, n_jobs, D, H, dropout_p, mode, grads_out, accumulate, workspace, workspace_bytes, (cudaStream_t)stream,
                        nullptr);
}

int rb_tower_bwd(const rb200_tower_bwd_job* jobs, int n_jobs, int D, int H, float dropout_p, int mode, float* grads_out,
                 int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream, RbPartials* defer) {
    RB_REQUIRE(jobs && n_jobs >= 1 && n_jobs <= MAX_JOBS, "tower_bwd: n_jobs must be 1..3");
    RB_REQUIRE(mode >= 0 && mode <= 2, "tower_bwd: mode must be 0, 1 or 2");
    RB_REQUIRE(dims_supported(D, H), "tower_bwd: unsupported widths D=%d H=%d", D, H);
    RB_REQUIRE(grads_out != nullptr || (defer != nullptr && mode != 0), "tower_bwd: grads_out is NULL");
    BwdParams p{};
    p.n_jobs = 0;
    p.keep_scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
    const int E = jobs[0].extra_dim;
    long long total_rows = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const rb200_tower_bwd_job& s = jobs[j];
        RB_REQUIRE(s.extra_dim == E, "tower_bwd: jobs sharing weights must share extra_dim");
        RB_REQUIRE(s.B >= 0, "tower_bwd: bad B");
        if (s.B == 0) continue;
        RB_REQUIRE(s.table && s.ids && s.W1 && s.W2 && s.dY && s.y && s.denom && s.hid && s.dpre && s.dact && s.dRows,
                   "tower_bwd: NULL pointer in job %d", j);
        RB_REQUIRE(E == 0 || s.extra, "tower_bwd: extra is NULL but extra_dim > 0");
        BwdJob& d = p.job[p.n_jobs++];
        d.table = s.table; d.ids = s.ids; d.extra = s.extra; d.n_rows = s.n_rows; d.B = s.B; d.E = E; d.extra_by_id = s.extra_by_id;
        d.W1 = s.W1; d.W2 = s.W2; d.dY = s.dY; d.y = s.y; d.denom = s.denom; d.hid = s.hid;
        d.dpre = s.dpre; d.dact = s.dact; d.dRows = s.dRows; d.img = (const unsigned char*)s.img;
        total_rows += s.B;
    }
    const int Din = D + E, Kp = (Din + 3) & ~3;
    const int P = H * Din + H + D * H + D;
    cudaStream_t st = stream;
    if (defer) *defer = RbPartials{nullptr, 0, P, H, Din};
    if (p.n_jobs == 0) {
        if (grads_out && !accumulate) RB_CUDA(cudaMemsetAsync(grads_out, 0, sizeof(float) * P, st));
        return RB200_OK;
    }
    if (mode != 0) {
        RB_REQUIRE(rb_tower_tc_supported(D, H, E), "tower_bwd: tcgen05 mode supports D in {64,128}, H=128, extra_dim<=24");
        RbArena tar(workspace, workspace_bytes);
        // split-K over the batch: short accumulation chains in TMEM (the tensor core's fp32 accumulation error grows with
        // the chain length; ≤ ~128 rows per CTA keeps the weight gradients inside the 1e-5 bound), all SMs busy
        int ns = (int)((total_rows + 63) / 64);
        // (the TMEM-operand weight-gradient kernel runs two CTAs per split — one per product — so half the splits fill the SMs)
        const int ns_max = rb_tower_use_ts(D) ? (rb_sm_count() + 1) / 2 : rb_sm_count();
        if (ns > ns_max) ns = ns_max;
        if (ns < 1) ns = 1;
        p.nsplit = ns;
        p.P = P;
        p.part = tar.take<float>((size_t)p.nsplit * P);
        unsigned char* img_ws = tar.take<unsigned char>(rb_tower_img_bytes(D, H, E));
        if (!workspace || !tar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "tower_bwd: workspace too small (%zu given)", workspace_bytes);
        if (defer) { defer->part = p.part; defer->nsplit = p.nsplit; }
        return rb_tower_bwd_tc(p, D, H, mode, defer ? nullptr : grads_out, accumulate, img_ws, st);
    }
    int nsplit = rb_sm_count() / 2;
    const long long stages = (total_rows + RT - 1) / RT;
    if (nsplit > stages) nsplit = (int)stages;
    if (nsplit < 1) nsplit = 1;
    RbArena ar(workspace, workspace_bytes);
    p.part = ar.take<float>((size_t)nsplit * P);
    p.nsplit = nsplit; p.P = P;
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "tower_bwd: workspace too small (%zu given)", workspace_bytes);

    const size_t smem_d = bwd_data_smem_bytes(D, H);
    RB_REQUIRE((int)smem_d <= rb_max_smem_optin(), "tower_bwd: D=%d H=%d exceeds shared memory", D, H);
    const int grid = partition_ctas(p.job, p.n_jobs, D, H);
    int rc = RB200_OK;
    RB_DISPATCH_DH(D, H, (rc = launch_bwd_data<R1, R2>(p, grid, smem_d, st)));
    if (rc) return rc;

    const int mt_w2 = ((D / 8) * (H / 4) + NT - 1) / NT;
    const int mt_w1 = ((H / 8) * (Kp / 4) + NT - 1) / NT;
    const int maxmt = mt_w1 > mt_w2 ? mt_w1 : mt_w2;
    const size_t smem_w = sizeof(float) * (size_t)RT * ((H > D ? H : D) + (H > Kp ? H : Kp));
    RB_DISPATCH_DH(D, H, (rc = dispatch_bwd_weights<R1, R2>(p, maxmt, smem_w, st)));
    if (rc) return rc;
    reduce_partials_kernel<<<(P + 255) / 256, 256, 0, st>>>(p.part, nsplit, P, grads_out, accumulate);
    RB_LAUNCH_CHECK("reduce_partials_kernel");
    return RB200_OK;
}
