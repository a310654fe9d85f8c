// IVFFlat / flat inner-product retrieval (sm_100a), fp32 FFMA path.
//
// Restates on the GPU what the reference gets from faiss.IndexIVFFlat (src/models/faiss_index.py:
// 68-74 build, :113/:145 search; algorithm per SURVEY.md Appendix B):
//
//   build   normalise rows → assign to argmax-IP centroid → stable counting sort into CSR lists
//           (offsets / list_ids / list_vecs contiguous per list)
//   train   spherical k-means: assign + deterministic segment mean + renormalise
//   search  (1) coarse scores q·Cᵀ  (2) top-nprobe lists per query, descending
//           (3) LIST-MAJOR scan: (list, query) pairs are radix-sorted by list so that one CTA
//               stages a 64-vector tile of a list in shared memory ONCE and scores it against every
//               query that probes the list — the database is read from HBM once per batch instead
//               of once per probing query
//           (4) per-query radix-select of the k best candidates + bitonic sort, ties by scan order
//   flat    exhaustive search = the same score/select kernels over row chunks with a running top-k
#include <cub/cub.cuh>
#include <float.h>
#include <stdlib.h>

#include "common.cuh"

// tcgen05 score product (gemm_tc.cu): C[M,N] = A[M,K]·B[N,K]ᵀ, mode 2 = 3xTF32 (fp32-grade)
int rb_list_scan_pipe(const float* q, int D, const float* list_vecs, long long n_vectors, const int64_t* offsets, const int* pair_qp, const long long* pair_dst,
                      const int* list_qstart, int nprobe, float* cand, long long total_candidates, const int* tile_list, const int* tile_idx,
                      long long n_tiles, int* err_flag, cudaStream_t st);
int rb_list_scan_tc(const float* q, int D, const float* list_vecs, const int64_t* offsets, const int* pair_qp, const int* list_qstart,
                    int nprobe, const int* cand_base, const long long* cand_off, float* cand, const int* tile_list,
                    const int* tile_idx, long long n_tiles, cudaStream_t st);
int rb_gemm_nt_tc(const float* A, int M, const float* B, int N, int K, int mode, float* C, long long ldc, int* err_flag,
                  cudaStream_t st);
// threshold-pruned exhaustive scan (flat_scan_tc.cu)
int rb_flat_scan_tc(const float* x, long long n_rows, const unsigned char* qimg, int n_chunks, const float* thr, const float* qmarg, int* count,
                    float* cand_s, long long stride, int kprev, int* cand_r, int cap, int* flags, cudaStream_t st);
int rb_flat_qimage(const float* q, int nq, int nq_pad, int block_rows, unsigned char* qimg, float* qmarg, cudaStream_t st);
// one-pass TF32 filter for 129 … 8192 queries (flat_filter_tc.cu: N = 128 MMAs, one issuer warp per tile)
bool rb_flat_filtered(int nq);
int rb_flat_filter_thresholds(const float* thr, const float* qmarg, int nq_pad, unsigned char* qimg, cudaStream_t st);
int rb_flat_filter_image(const float* q, int nq, int nq_pad, unsigned char* qimg, float* qmarg, cudaStream_t st);
int rb_flat_filter_tc(const float* x, long long n_rows, const unsigned char* qimg, int n_blocks, const float* thr, const float* qmarg, int* count,
                      float* cand_s, long long stride, int kprev, int* cand_r, int cap, int* flags, cudaStream_t st);
// streaming round kernel for at most 128 queries (flat_stream_tc.cu: persistent, tensor-map TMA, one-pass filter)
bool rb_flat_streamed(int n_chunks);
int rb_flat_stream_tc(const float* x, long long n_rows, const unsigned char* qimg, int n_chunks, const float* thr, const float* qmarg, int* count,
                      float* cand_s, long long stride, int kprev, int* cand_r, int cap, int* flags, cudaStream_t st);
int rb_flat_rescore(const float* q, int nq, const float* x, const int* count, const int* cand_r, int cap, float* cand_s, long long stride,
                    int kprev, cudaStream_t st);

namespace {

constexpr int NT = 256;
constexpr int TT = 64;          // score tile: 64 queries × 64 vectors

// ------------------------------------------------------------------------------------------ //
// small utilities
// ------------------------------------------------------------------------------------------ //
__global__ void __launch_bounds__(NT) normalize_rows_kernel(const float* __restrict__ x, long long n, int D, float eps,
                                                            float* __restrict__ out) {
    const long long row = ((long long)blockIdx.x * NT + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = x[row * D + d]; ss = fmaf(v, v, ss); }
    ss = rb_warp_sum(ss);
    const float den = fmaxf(sqrtf(ss), eps);
    for (int d = lane; d < D; d += 32) out[row * D + d] = x[row * D + d] / den;
}

// load a 64-row tile of row-major [*, D] data (rows given by an optional index list) into smem
template <int D, typename IdxT>
__device__ __forceinline__ void load_tile(float* __restrict__ S, int ld, const float* __restrict__ src, long long row0,
                                          long long row_end, const IdxT* __restrict__ index, int index_div) {
    for (int idx = threadIdx.x; idx < TT * (D / 4); idx += NT) {
        const int r = idx / (D / 4), c = idx - r * (D / 4);
        const long long g = row0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < row_end) {
            const long long srow = index ? (long long)(index[g] / index_div) : g;
            v = __ldg(reinterpret_cast<const float4*>(src + srow * D) + c);
        }
        *reinterpret_cast<float4*>(S + r * ld + c * 4) = v;
    }
}

// acc[r][c] = Σ_k A[ty*4+r][k] · W[tx+16c][k]
template <int D, int LD>
__device__ __forceinline__ void score_tile(const float* __restrict__ As, const float* __restrict__ Ws, int tx, int ty,
                                           float (&acc)[4][4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    const float* a_base = As + (ty * 4) * LD;
    const float* w_base = Ws + tx * LD;
#pragma unroll 2
    for (int k = 0; k < D; k += 4) {
        float4 b[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4*>(w_base + c * 16 * LD + k);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(a_base + r * LD + k);
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

template <int D> struct Ld { static constexpr int v = ((D / 4) & 1) ? D : D + 4; };
template <int D> constexpr size_t tile_pair_bytes() { return sizeof(float) * 2 * TT * Ld<D>::v; }
// opt in to > 48 KB of dynamic shared memory (D = 128) once per kernel instantiation
template <typename K>
int allow_smem(K kernel, size_t bytes, bool& done) {
    if (!done && bytes > 48 * 1024) RB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    done = true;
    return RB200_OK;
}
#define RB_TILE_LAUNCH(KERNEL, DD, GRID, ST, ...)                                   \
    do {                                                                            \
        static bool done__ = false;                                                 \
        int rc__ = allow_smem(KERNEL<DD>, tile_pair_bytes<DD>(), done__);           \
        if (rc__) return rc__;                                                      \
        KERNEL<DD><<<GRID, NT, tile_pair_bytes<DD>(), ST>>>(__VA_ARGS__);           \
    } while (0)

// ------------------------------------------------------------------------------------------ //
// assignment: argmax_c x·centroid_c  (lowest index on ties)
// ------------------------------------------------------------------------------------------ //
template <int D>
__global__ void __launch_bounds__(NT) assign_kernel(const float* __restrict__ x, long long n, const float* __restrict__ C,
                                                    int nlist, int* __restrict__ assign, float* __restrict__ best_score) {
    constexpr int LD = Ld<D>::v;
    extern __shared__ __align__(16) float tile_smem[];
    float* Xs = tile_smem;
    float* Cs = tile_smem + TT * LD;
    const long long r0 = (long long)blockIdx.x * TT;
    load_tile<D, int>(Xs, LD, x, r0, n, nullptr, 1);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float bv[4]; int bi[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { bv[r] = -FLT_MAX; bi[r] = 0x7fffffff; }
    for (int c0 = 0; c0 < nlist; c0 += TT) {
        __syncthreads();
        load_tile<D, int>(Cs, LD, C, c0, nlist, nullptr, 1);
        __syncthreads();
        float acc[4][4];
        score_tile<D, LD>(Xs, Cs, tx, ty, acc);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int ci = c0 + tx + 16 * c;
                if (ci < nlist && (acc[r][c] > bv[r] || (acc[r][c] == bv[r] && ci < bi[r]))) { bv[r] = acc[r][c]; bi[r] = ci; }
            }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(RB_FULL_MASK, bv[r], o);
            const int oi = __shfl_xor_sync(RB_FULL_MASK, bi[r], o);
            if (ov > bv[r] || (ov == bv[r] && oi < bi[r])) { bv[r] = ov; bi[r] = oi; }
        }
        const long long row = r0 + ty * 4 + r;
        if (tx == 0 && row < n) {
            assign[row] = bi[r];
            if (best_score) best_score[row] = bv[r];
        }
    }
}

// ------------------------------------------------------------------------------------------ //
// CSR construction / k-means update (shared: stable sort of rows by list)
// ------------------------------------------------------------------------------------------ //
__global__ void iota_kernel(int* __restrict__ v, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int)i;
}
__global__ void histogram_kernel(const int* __restrict__ keys, long long n, int nbins, int* __restrict__ counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const int k = keys[i]; if ((unsigned)k < (unsigned)nbins) atomicAdd(counts + k, 1); }
}
__global__ void offsets_to_i64_kernel(const int* __restrict__ excl, const int* __restrict__ counts, int nlist,
                                      int64_t* __restrict__ offsets) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nlist) offsets[i] = excl[i];
    if (i == nlist - 1) offsets[nlist] = (int64_t)excl[i] + counts[i];
}
__global__ void __launch_bounds__(NT) gather_rows_kernel(const float* __restrict__ x, const int* __restrict__ order, long long n,
                                                         int D, int64_t* __restrict__ list_ids, float* __restrict__ list_vecs) {
    const long long i = ((long long)blockIdx.x * NT + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const long long src = order[i];
    if (lane == 0) list_ids[i] = src;
    for (int c = lane; c < D / 4; c += 32)
        reinterpret_cast<float4*>(list_vecs + i * D)[c] = __ldg(reinterpret_cast<const float4*>(x + src * D) + c);
}
// one warp per list: mean of member rows in ascending row order, then L2 renormalise
__global__ void __launch_bounds__(NT) centroid_update_kernel(const float* __restrict__ x, const int* __restrict__ order,
                                                             const int* __restrict__ excl, const int* __restrict__ counts,
                                                             int nlist, int D, float* __restrict__ centroids) {
    const int l = (blockIdx.x * NT + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (l >= nlist) return;
    const int cnt = counts[l], beg = excl[l];
    if (cnt == 0) return;   // empty list keeps its centroid (host-side split policy decides what to do)
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
        float s = 0.f;
        for (int j = 0; j < cnt; ++j) s += __ldg(x + (long long)order[beg + j] * D + d);
        s /= (float)cnt;
        centroids[(long long)l * D + d] = s;
        ss = fmaf(s, s, ss);
    }
    ss = rb_warp_sum(ss);
    const float den = fmaxf(sqrtf(ss), 1e-30f);
    for (int d = lane; d < D; d += 32) centroids[(long long)l * D + d] /= den;
}

struct SortWs {
    int* iota; int* order; int* keys_sorted; int* counts; int* excl; char* temp; size_t temp_bytes;
};
size_t sort_ws_temp(long long n, int nlist) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, nlist);
    return a > b ? a : b;
}
size_t sort_ws_bytes(long long n, int nlist) {
    return 256 * 8 + sizeof(int) * (3 * (size_t)n + 2 * (size_t)nlist) + sort_ws_temp(n, nlist);
}
bool carve_sort_ws(RbArena& ar, long long n, int nlist, SortWs& w) {
    w.iota = ar.take<int>(n); w.order = ar.take<int>(n); w.keys_sorted = ar.take<int>(n);
    w.counts = ar.take<int>(nlist); w.excl = ar.take<int>(nlist);
    w.temp_bytes = sort_ws_temp(n, nlist);
    w.temp = ar.take<char>(w.temp_bytes);
    return ar.ok();
}
int key_bits_i32(int nlist) { int b = 1; while (b < 31 && (1 << b) < nlist) ++b; return b; }

int sort_rows_by_list(const int* assign, long long n, int nlist, SortWs& w, cudaStream_t st) {
    iota_kernel<<<(unsigned)((n + NT - 1) / NT), NT, 0, st>>>(w.iota, n);
    RB_LAUNCH_CHECK("iota_kernel");
    size_t tb = w.temp_bytes;
    RB_CUDA(cub::DeviceRadixSort::SortPairs(w.temp, tb, assign, w.keys_sorted, (const int*)w.iota, w.order, (int)n, 0,
                                            key_bits_i32(nlist), st));
    RB_CUDA(cudaMemsetAsync(w.counts, 0, sizeof(int) * nlist, st));
    histogram_kernel<<<(unsigned)((n + NT - 1) / NT), NT, 0, st>>>(assign, n, nlist, w.counts);
    RB_LAUNCH_CHECK("histogram_kernel");
    tb = w.temp_bytes;
    RB_CUDA(cub::DeviceScan::ExclusiveSum(w.temp, tb, (const int*)w.counts, w.excl, nlist, st));
    return RB200_OK;
}

// ------------------------------------------------------------------------------------------ //
// search, step 2: top-nprobe lists per query (one CTA per query) + candidate bookkeeping
// ------------------------------------------------------------------------------------------ //
// After the generic top-k kernel picked the nprobe best lists per query (score desc, list id asc): candidate
// bookkeeping, one warp per query (lane = probe rank, exclusive scan of the list lengths by shuffles).
__global__ void __launch_bounds__(NT) probe_finish_kernel(const int64_t* __restrict__ probe_ids, int nq, int nprobe,
                                                          const int64_t* __restrict__ offsets, int* __restrict__ probes,
                                                          int* __restrict__ cand_base, long long* __restrict__ totals,
                                                          int* __restrict__ list_qcount) {
    const int q = (blockIdx.x * NT + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    int running = 0;
    for (int p0 = 0; p0 < nprobe; p0 += 32) {
        const int p = p0 + lane;
        long long l = -1;
        int len = 0;
        if (p < nprobe) {
            l = probe_ids[(long long)q * nprobe + p];
            if (l >= 0) { len = (int)(offsets[l + 1] - offsets[l]); atomicAdd(list_qcount + l, 1); }
        }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(RB_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (p < nprobe) {
            probes[(long long)q * nprobe + p] = (int)l;
            cand_base[(long long)q * nprobe + p] = running + incl - len;
        }
        running += __shfl_sync(RB_FULL_MASK, incl, 31);
    }
    if (lane == 0) totals[q] = running;
}

// One block: list_qstart = exclusive scan of the per-list probe counts (n_lists + 1 entries), cand_off = exclusive scan of the per-query
// candidate totals (nq + 1 entries), out2 = {Σ, max} of the totals — the plan's two cub scans, a memset and a reduction in one launch.
constexpr int PS_NT = 1024;
// block-wide exclusive scan of n values (+ the total at index n): thread t owns a contiguous run of ceil((n+1)/PS_NT) elements — its
// loads are independent of each other and of the block's one shuffle / shared-memory round (a chunked scan paid a dependent global
// round trip and three barriers per 1024 elements: 15.6 us for the plan's two 4097-element scans)
template <typename TIn, typename TOut>
__device__ __forceinline__ long long block_excl_scan(const TIn* __restrict__ in, int n, TOut* __restrict__ out, long long* wsum /*[PS_NT/32 + 1]*/,
                                                     long long* vmax) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + 1 + PS_NT - 1) / PS_NT;
    const int i0 = tid * per, i1 = min(i0 + per, n);              // inputs [i0, i1) (index n has no input: it receives the total)
    long long local = 0, mx = 0;
    for (int i = i0; i < i1; ++i) { const long long v = (long long)in[i]; local += v; mx = v > mx ? v : mx; }
    long long incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(RB_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();                                             // (wsum may still be read by a previous call)
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    long long before = incl - local, total = 0;
    for (int w = 0; w < PS_NT / 32; ++w) { const long long c = wsum[w]; if (w < warp) before += c; total += c; }
    long long run = before;
    for (int i = i0; i < min(i0 + per, n + 1); ++i) {
        out[i] = (TOut)run;
        if (i < n) run += (long long)in[i];
    }
    if (vmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const long long om = __shfl_xor_sync(RB_FULL_MASK, mx, o); mx = om > mx ? om : mx; }
        __syncthreads();
        if (lane == 0) wsum[warp] = mx;
        __syncthreads();
        long long m = 0;
        for (int w = 0; w < PS_NT / 32; ++w) m = wsum[w] > m ? wsum[w] : m;
        *vmax = m;
    }
    return total;
}
__global__ void __launch_bounds__(PS_NT) plan_scans_kernel(const int* __restrict__ list_qcount, int n_lists, int* __restrict__ list_qstart,
                                                           const long long* __restrict__ totals, int nq, long long* __restrict__ cand_off,
                                                           long long* __restrict__ out2) {
    __shared__ long long wsum[PS_NT / 32];
    long long vmax = 0;
    block_excl_scan<int, int>(list_qcount, n_lists, list_qstart, wsum, nullptr);
    const long long total = block_excl_scan<long long, long long>(totals, nq, cand_off, wsum, &vmax);
    if (threadIdx.x == 0) { out2[0] = total; out2[1] = vmax; }
}

// (list, query) pairs grouped by list without a sort: pair i of list l goes to one of the list's slots [list_qstart[l], list_qstart[l+1]),
// handed out by counting the list's counter down.  The order of a list's pairs is arbitrary (it only decides which query sits in which
// column of the scan's MMAs: every candidate's destination comes from cand_base, so the results do not depend on it).
// pair_dst: where the pair's candidate scores start in the candidate buffer (cand_off of its query + cand_base of the probe) — the
// scan then finds a chunk's destinations with one load instead of a pair → query → offset chain.
__global__ void pair_scatter_kernel(const int* __restrict__ probes, long long n_pairs, int nprobe, const int* __restrict__ list_qstart,
                                    int* __restrict__ list_qcount, const long long* __restrict__ cand_off, const int* __restrict__ cand_base,
                                    int* __restrict__ pair_qp, long long* __restrict__ pair_dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int l = probes[i];
    if (l < 0) return;
    const int slot = atomicSub(list_qcount + l, 1) - 1;
    const int pos = list_qstart[l] + slot;
    pair_qp[pos] = (int)i;
    pair_dst[pos] = cand_off[i / nprobe] + cand_base[i];
}

// ------------------------------------------------------------------------------------------ //
// search, step 3: list-major scan.  grid = (vector tiles, lists)
// ------------------------------------------------------------------------------------------ //
template <int D>
__global__ void __launch_bounds__(NT) list_scan_kernel(const float* __restrict__ q, const float* __restrict__ list_vecs,
                                                       const int64_t* __restrict__ offsets, const int* __restrict__ pair_qp,
                                                       const int* __restrict__ list_qstart, int nprobe,
                                                       const int* __restrict__ cand_base, const long long* __restrict__ cand_off,
                                                       float* __restrict__ cand, const int* __restrict__ tile_list,
                                                       const int* __restrict__ tile_idx) {
    constexpr int LD = Ld<D>::v;
    extern __shared__ __align__(16) float tile_smem[];
    float* Vs = tile_smem;
    float* Qs = tile_smem + TT * LD;
    __shared__ long long dst[TT];
    // work item = (list, 64-vector tile): from the index's precomputed tile table when given (no empty CTAs), else 2-D grid
    const int l = tile_list ? tile_list[blockIdx.x] : (int)blockIdx.y;
    const int ti = tile_list ? tile_idx[blockIdx.x] : (int)blockIdx.x;
    const long long lbeg = offsets[l], lend = offsets[l + 1];
    const long long v0 = lbeg + (long long)ti * TT;
    if (v0 >= lend) return;
    const int qs = list_qstart[l], qe = list_qstart[l + 1];
    if (qs == qe) return;
    load_tile<D, int>(Vs, LD, list_vecs, v0, lend, nullptr, 1);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int nv = (int)min((long long)TT, lend - v0);
    for (int p0 = qs; p0 < qe; p0 += TT) {
        __syncthreads();
        load_tile<D, int>(Qs, LD, q, p0, qe, pair_qp, nprobe);
        if (threadIdx.x < TT) {
            const int pi = p0 + threadIdx.x;
            long long d = -1;
            if (pi < qe) {
                const int qp = pair_qp[pi];
                d = cand_off[qp / nprobe] + cand_base[qp] + (v0 - lbeg);
            }
            dst[threadIdx.x] = d;
        }
        __syncthreads();
        float acc[4][4];
        score_tile<D, LD>(Qs, Vs, tx, ty, acc);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const long long d = dst[ty * 4 + r];
            if (d < 0) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int vj = tx + 16 * c;
                if (vj < nv) cand[d + vj] = acc[r][c];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ //
// step 4: per-query top-k.  Radix select (4 × 8 bit, MSB first) of the k-th largest key, ordered
// compaction (ties by candidate position), bitonic sort of the winners.
// ------------------------------------------------------------------------------------------ //
__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// id resolution modes for the winners
struct ResolveIvf {      // candidate position → (probe, offset in list) → list_ids
    const int* probes; const int* cand_base; const int64_t* offsets; const int64_t* list_ids; int nprobe;
    const int* s_base; const int* s_probe;      // shared-memory copies for the current query (set by stage)
    __host__ __device__ int aux_ints() const { return 2 * nprobe; }
    __device__ void stage(int q, int* aux, int tid, int nt) {
        for (int i = tid; i < nprobe; i += nt) { aux[i] = cand_base[(long long)q * nprobe + i]; aux[nprobe + i] = probes[(long long)q * nprobe + i]; }
        s_base = aux; s_probe = aux + nprobe;
    }
};
struct ResolveIdentity {   // candidate position is the id (coarse quantizer: list number)
    int unused;
    __host__ __device__ int aux_ints() const { return 0; }
    __device__ void stage(int, int*, int, int) {}
};
struct ResolveFlat {     // [running top-k (kprev entries) | chunk rows]
    const int64_t* prev_ids; int kprev; long long row_base;
    __host__ __device__ int aux_ints() const { return 0; }
    __device__ void stage(int, int*, int, int) {}
};
struct ResolveSurv {     // [running top-k (kprev entries) | survivors of a pruned round: explicit rows relative to row_base]
    const int64_t* prev_ids; int kprev; const int* rows; int cap; long long row_base;
    __host__ __device__ int aux_ints() const { return 0; }
    __device__ void stage(int, int*, int, int) {}
};

template <typename R> __device__ long long resolve_id(const R& r, int q, int c);
template <> __device__ long long resolve_id<ResolveIvf>(const ResolveIvf& r, int q, int c) {
    (void)q;
    const int* base = r.s_base;
    int lo = 0, hi = r.nprobe - 1;           // last probe p with base[p] <= c  (empty lists share a base: take the last)
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (base[mid] <= c) lo = mid; else hi = mid - 1; }
    int p = lo;                              // step back over empty / invalid probes that share the base
    while (p > 0) {
        const int l = r.s_probe[p];
        if (l >= 0 && c - base[p] < r.offsets[l + 1] - r.offsets[l]) break;
        --p;
    }
    const int l = r.s_probe[p];
    return r.list_ids[r.offsets[l] + (c - base[p])];
}
template <> __device__ long long resolve_id<ResolveIdentity>(const ResolveIdentity&, int, int c) { return c; }
template <> __device__ long long resolve_id<ResolveFlat>(const ResolveFlat& r, int q, int c) {
    if (c < r.kprev) return r.prev_ids[(long long)q * r.kprev + c];
    return r.row_base + (c - r.kprev);
}
template <> __device__ long long resolve_id<ResolveSurv>(const ResolveSurv& r, int q, int c) {
    if (c < r.kprev) return r.prev_ids[(long long)q * r.kprev + c];
    return r.row_base + r.rows[(long long)q * r.cap + (c - r.kprev)];
}

constexpr int SORT_CAP = 2048;   // winners + boundary bucket must fit the in-CTA buffer (falls back to radix select otherwise); the buffer
                                 // of a launch is sized by its k (sort_cap: a power of two ≥ 4·k in 256 … SORT_CAP) — the coarse
                                 // quantizer's top-nprobe select keeps 8 CTAs per SM that way instead of 6
constexpr int NBK = 1024;        // value-range buckets of the fast path (fine enough that top-500 + boundary bucket ≤ 512 usually)
constexpr int RANK_MAX = 512;    // fullest bucket for which the winners are ordered by in-bucket ranks instead of the bitonic sort

// Per-query top-k (one CTA per query).
//   fast path : value-range bucket select — min/max of the scores, NBK linear buckets over [min, max] (well spread even
//               though cosine scores share their leading float bits), locate the bucket b* that holds the k-th score.
//               The buckets are already an order: the histogram becomes, in place, each bucket's first output position
//               (suffix sums from the top), every candidate in buckets ≥ b* takes a slot inside its bucket's range, and its
//               exact place is its rank by (score desc, candidate position asc) among the few entries of the same bucket —
//               written straight to the output when < k.  No sort: the 45-step bitonic network over 512 (key, position)
//               pairs was ≈ 40 % of this kernel's instructions and the separate counting pass another ≈ 8 %.
//               (A bucket with more than RANK_MAX entries — heavily duplicated scores — takes the older form: compact
//               winners + boundary bucket, bitonic-sort that set, emit the first k.)
//   fallback  : exact MSB-first radix select (many equal scores / boundary bucket too large).
// Bitonic sort of NT·EPT entries, descending by (key, then ascending position): EPT consecutive entries per thread as 64-bit
// composites (key, inverted position) in registers; strides < EPT are register swaps, strides < 32·EPT warp shuffles, only the
// larger ones go through shared memory (6 of the 45 / 55 steps at 512 / 1024 entries) — the all-shared-memory network costs a
// __syncthreads per step.  skey|sidx (16 KB) double as the exchange buffer.
template <int EPT>
__device__ __forceinline__ void reg_bitonic_sort(uint32_t* skey, int* sidx, int tid) {
    constexpr int N = NT * EPT;
    unsigned long long v[EPT];
    unsigned long long* xbuf = reinterpret_cast<unsigned long long*>(skey);
    __syncthreads();                                                             // entries (and padding) written by other threads
#pragma unroll
    for (int e = 0; e < EPT; ++e) v[e] = ((unsigned long long)skey[EPT * tid + e] << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)sidx[EPT * tid + e]);
    __syncthreads();
    for (int size = 2; size <= N; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32 * EPT) {
#pragma unroll
                for (int e = 0; e < EPT; ++e) xbuf[EPT * tid + e] = v[e];
                __syncthreads();
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    const int i = EPT * tid + e;
                    const unsigned long long pv = xbuf[i ^ stride];
                    const bool keep_max = ((i & stride) == 0) == ((i & size) == 0);
                    v[e] = keep_max ? (v[e] > pv ? v[e] : pv) : (v[e] < pv ? v[e] : pv);
                }
                __syncthreads();
            } else if (stride >= EPT) {
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    const int i = EPT * tid + e;
                    const unsigned long long pv = __shfl_xor_sync(RB_FULL_MASK, v[e], stride / EPT);
                    const bool keep_max = ((i & stride) == 0) == ((i & size) == 0);
                    v[e] = keep_max ? (v[e] > pv ? v[e] : pv) : (v[e] < pv ? v[e] : pv);
                }
            } else {
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    if ((e & stride) == 0) {
                        const bool desc = ((EPT * tid + e) & size) == 0;
                        const unsigned long long a = v[e], b = v[e | stride];
                        const unsigned long long hi = a > b ? a : b, lo = a > b ? b : a;
                        v[e] = desc ? hi : lo; v[e | stride] = desc ? lo : hi;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) { skey[EPT * tid + e] = (uint32_t)(v[e] >> 32); sidx[EPT * tid + e] = (int)(0xFFFFFFFFu - (uint32_t)v[e]); }
}

template <typename R>
__global__ void __launch_bounds__(NT) select_topk_kernel(const float* __restrict__ cand, const long long* __restrict__ cand_off,
                                                         long long fixed_stride, const long long* __restrict__ counts,
                                                         int fixed_count, int k, int cache_cap, int sort_cap, R res,
                                                         float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint32_t* skey = reinterpret_cast<uint32_t*>(sm_raw);           // [sort_cap]
    int* sidx = reinterpret_cast<int*>(skey + sort_cap);            // [sort_cap]
    float* cache = reinterpret_cast<float*>(sidx + sort_cap);       // [cache_cap]
    int* aux = reinterpret_cast<int*>(cache + cache_cap);           // [res.aux_ints()] per-query lookup tables of the resolver
    __shared__ int hist[256];
    __shared__ int hist_f[NBK];           // fast path: NBK linear buckets over [min, max]
    __shared__ int wsum[2][NT / 32];
    __shared__ uint32_t s_prefix;
    __shared__ int s_a, s_b, s_c, s_count;
    const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* src = cand + (cand_off ? cand_off[q] : (long long)q * fixed_stride);
    const int n = counts ? (int)counts[q] : fixed_count;

    float lo = FLT_MAX, hi = -FLT_MAX;          // range of the scores, gathered while they are being cached
    bool have_range = false;
    if (n <= cache_cap) {
        // 8 independent loads in flight per thread (a plain copy loop serialises on every load's latency)
        for (int i0 = tid; i0 < n; i0 += 8 * NT) {
            float v[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) { const int i = i0 + b * NT; v[b] = i < n ? __ldg(src + i) : 0.f; }
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int i = i0 + b * NT;
                if (i < n) { cache[i] = v[b]; lo = fminf(lo, v[b]); hi = fmaxf(hi, v[b]); }
            }
        }
        __syncthreads();
        src = cache;
        have_range = true;
    }
    res.stage(q, aux, tid, NT);      // resolver copies its per-query tables into shared memory (no dependent global chains); every
                                     // path below passes a block barrier before it resolves ids
    int m = 0;                       // number of entries placed in skey/sidx
    bool done = false, emitted = false;
    if (n <= sort_cap && n <= 2 * k) {
        // small input: sort everything
        for (int i = tid; i < n; i += NT) { skey[i] = f2key(src[i]); sidx[i] = i; }
        m = n; done = true;
    } else {
        // ---- fast path ------------------------------------------------------------------------------------ //
        if (!have_range)
            for (int i = tid; i < n; i += NT) { const float v = src[i]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(RB_FULL_MASK, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(RB_FULL_MASK, hi, o));
        }
        if (lane == 0) { wsum[0][warp] = __float_as_int(lo); wsum[1][warp] = __float_as_int(hi); }
#pragma unroll
        for (int c = 0; c < NBK / NT; ++c) hist_f[tid * (NBK / NT) + c] = 0;
        if (tid == 0) s_count = 0;
        __syncthreads();
        for (int w = 0; w < NT / 32; ++w) { lo = fminf(lo, __int_as_float(wsum[0][w])); hi = fmaxf(hi, __int_as_float(wsum[1][w])); }
        const float scale = (hi > lo) ? (float)NBK / (hi - lo) : 0.f;
        for (int i = tid; i < n; i += NT) {
            const int b = min(NBK - 1, (int)((src[i] - lo) * scale));
            atomicAdd(&hist_f[b], 1);
        }
        __syncthreads();
        constexpr int PB = NBK / NT;
        int part = 0, suf = 0, any_big = 0;
        {
            // boundary bucket b* = largest b with count(bucket ≥ b) ≥ k, found with a block-wide suffix sum
            int hmax = 0;
#pragma unroll
            for (int c = 0; c < PB; ++c) { const int h = hist_f[tid * PB + c]; part += h; hmax = max(hmax, h); }
            suf = part;                                  // Σ over lanes ≥ lane (within the warp)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_down_sync(RB_FULL_MASK, suf, o);
                if (lane + o < 32) suf += t;
            }
            if (lane == 0) wsum[0][warp] = suf;
            any_big = __syncthreads_or(hmax > RANK_MAX);
            for (int w = warp + 1; w < NT / 32; ++w) suf += wsum[0][w];
            if (suf >= k && suf - part < k) {            // the boundary lies inside this thread's PB buckets
                int cum = suf - part, b = tid * PB + PB - 1;
                for (; b > tid * PB; --b) { if (cum + hist_f[b] >= k) break; cum += hist_f[b]; }
                s_a = b; s_b = cum + hist_f[b]; s_c = hist_f[b];
            }
        }
        __syncthreads();
        const int bstar = s_a, m_fast = s_b;
        auto bucket_of = [&](float v) { return min(NBK - 1, (int)((v - lo) * scale)); };
        if (m_fast <= sort_cap && scale > 0.f && !any_big) {
            // hist_f[b] := entries in buckets > b = first output position of bucket b (each thread converts the PB buckets it summed;
            // nobody reads the counts any more).  The scatter's atomicAdd then hands out the slots of the bucket's range and leaves
            // hist_f[b] at the range's end, which is the beginning of bucket b − 1's.
            int excl = suf - part;
#pragma unroll
            for (int c = PB - 1; c >= 0; --c) { const int h = hist_f[tid * PB + c]; hist_f[tid * PB + c] = excl; excl += h; }
            __syncthreads();
            for (int i = tid; i < n; i += NT) {
                const float v = src[i];
                const int b = bucket_of(v);
                if (b >= bstar) { const int slot = atomicAdd(&hist_f[b], 1); skey[slot] = f2key(v); sidx[slot] = i; }
            }
            __syncthreads();
            for (int j = tid; j < m_fast; j += NT) {
                const uint32_t kj = skey[j]; const int ij = sidx[j];
                const int b = bucket_of(key2f(kj));
                const int beg = b == NBK - 1 ? 0 : hist_f[b + 1], end = hist_f[b];
                int pos = beg;
                for (int t = beg; t < end; ++t) {
                    const uint32_t kt = skey[t]; const int it = sidx[t];
                    pos += (kt > kj || (kt == kj && it < ij)) ? 1 : 0;
                }
                if (pos < k) {                           // (the boundary bucket's tail falls off the end)
                    out_scores[(long long)q * k + pos] = key2f(kj);
                    out_ids[(long long)q * k + pos] = resolve_id<R>(res, q, ij);
                }
            }
            done = true; emitted = true;
        } else if (m_fast <= sort_cap && scale > 0.f) {
            // compaction without a per-iteration atomic (measured: the ballot + atomicAdd loop was 29 % of the kernel): every
            // thread counts its own survivors, one block-wide exclusive scan places them, a second pass writes them.  The
            // order of the survivors is irrelevant (they are sorted by (score, position) below).
            // Entries above the boundary bucket are certainly among the k best; of the c_b entries IN the boundary bucket only the
            // best r = k − (count above) are.  When that makes exactly k <= 512 entries the final sort runs on 512 instead of
            // 1024 slots, so the boundary bucket (a handful of entries) is cut by exact rank first.
            const int c_b = s_c, n_above = m_fast - c_b, r_keep = k - n_above;
            const bool cut = m_fast > 512 && k <= 512 && c_b <= 512;
            int cnt_a = 0, cnt_b = 0;
            for (int i = tid; i < n; i += NT) {
                const int b = min(NBK - 1, (int)((src[i] - lo) * scale));
                cnt_a += b > bstar ? 1 : 0; cnt_b += b == bstar ? 1 : 0;
            }
            int inc_a = cnt_a, inc_b = cnt_b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ta = __shfl_up_sync(RB_FULL_MASK, inc_a, o), tb = __shfl_up_sync(RB_FULL_MASK, inc_b, o);
                if (lane >= o) { inc_a += ta; inc_b += tb; }
            }
            if (lane == 31) { wsum[0][warp] = inc_a; wsum[1][warp] = inc_b; }
            __syncthreads();
            int pa = inc_a - cnt_a, pb = inc_b - cnt_b;
            for (int w = 0; w < warp; ++w) { pa += wsum[0][w]; pb += wsum[1][w]; }
            // boundary entries: behind the others (plain compaction) or parked at the end of the sort buffer (exact cut)
            const int b_base = cut ? sort_cap - c_b : n_above;
            for (int i = tid; i < n; i += NT) {
                const float v = src[i];
                const int b = min(NBK - 1, (int)((v - lo) * scale));
                if (b > bstar) { skey[pa] = f2key(v); sidx[pa] = i; ++pa; }
                else if (b == bstar) { skey[b_base + pb] = f2key(v); sidx[b_base + pb] = i; ++pb; }
            }
            m = m_fast;
            if (cut) {
                __syncthreads();
                for (int j = tid; j < c_b; j += NT) {
                    const uint32_t kj = skey[b_base + j]; const int ij = sidx[b_base + j];
                    int rank = 0;
                    for (int t = 0; t < c_b; ++t) {
                        const uint32_t kt = skey[b_base + t]; const int it = sidx[b_base + t];
                        rank += (kt > kj || (kt == kj && it < ij)) ? 1 : 0;
                    }
                    if (rank < r_keep) { skey[n_above + rank] = kj; sidx[n_above + rank] = ij; }     // ranks are distinct
                }
                m = k;
            }
            done = true;
        }
    }
    if (!done) {
        // ---- fallback: exact radix select (4 × 8 bit, MSB first) + position-ordered compaction ------------------ //
        __syncthreads();
        uint32_t prefix = 0u, mask = 0u;
        int krem = k;
        for (int shift = 24; shift >= 0; shift -= 8) {
            hist[tid] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += NT) {
                const uint32_t key = f2key(src[i]);
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int cum = 0, d = 255;
                for (; d > 0; --d) { if (cum + hist[d] >= krem) break; cum += hist[d]; }
                s_prefix = prefix | ((uint32_t)d << shift);
                s_a = krem - cum;
            }
            __syncthreads();
            prefix = s_prefix; krem = s_a; mask |= 255u << shift;
            __syncthreads();
        }
        const uint32_t T = prefix;
        const int k_eq = krem, n_gt = k - krem;
        int run_gt = 0, run_eq = 0;
        for (int base = 0; base < n; base += NT) {
            const int i = base + tid;
            uint32_t key = 0u; bool gt = false, eq = false;
            if (i < n) { key = f2key(src[i]); gt = key > T; eq = key == T; }
            const uint32_t bg = __ballot_sync(RB_FULL_MASK, gt), be = __ballot_sync(RB_FULL_MASK, eq);
            if (lane == 0) { wsum[0][warp] = __popc(bg); wsum[1][warp] = __popc(be); }
            __syncthreads();
            int og = run_gt, oe = run_eq, tg = 0, te = 0;
            for (int w = 0; w < NT / 32; ++w) {
                if (w < warp) { og += wsum[0][w]; oe += wsum[1][w]; }
                tg += wsum[0][w]; te += wsum[1][w];
            }
            const uint32_t lt = (1u << lane) - 1u;
            if (gt) { const int p = og + __popc(bg & lt); skey[p] = key; sidx[p] = i; }
            if (eq) { const int e = oe + __popc(be & lt); if (e < k_eq) { skey[n_gt + e] = key; sidx[n_gt + e] = i; } }
            run_gt += tg; run_eq += te;
            __syncthreads();
        }
        m = k;
    }
    if (emitted) return;             // (block-uniform)
    // ---- bitonic sort of the m collected entries, descending by (key, then ascending position) ---------------- //
    int sort_n = 2;
    while (sort_n < m) sort_n <<= 1;
    __syncthreads();
    for (int i = m + tid; i < sort_n; i += NT) { skey[i] = 0u; sidx[i] = 0x7fffffff; }
    if (sort_n == 1024) reg_bitonic_sort<4>(skey, sidx, tid);
    else if (sort_n == 512) reg_bitonic_sort<2>(skey, sidx, tid);
    else
    for (int size = 2; size <= sort_n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < sort_n / 2; t += NT) {
                const int lo_i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1)), hi_i = lo_i + stride;   // stride is a power of two
                const bool desc = ((lo_i & size) == 0);
                const uint32_t ka = skey[lo_i], kb = skey[hi_i];
                const int ia = sidx[lo_i], ib = sidx[hi_i];
                const bool a_first = ka > kb || (ka == kb && ia < ib);    // a ranks before b
                if (a_first != desc) { skey[lo_i] = kb; skey[hi_i] = ka; sidx[lo_i] = ib; sidx[hi_i] = ia; }
            }
        }
    }
    __syncthreads();
    const int n_out = m < k ? m : k;
    for (int j = tid; j < k; j += NT) {
        if (j < n_out) {
            out_scores[(long long)q * k + j] = key2f(skey[j]);
            out_ids[(long long)q * k + j] = resolve_id<R>(res, q, sidx[j]);
        } else {
            out_scores[(long long)q * k + j] = -FLT_MAX;
            out_ids[(long long)q * k + j] = -1;
        }
    }
}

int next_pow2(int v) { int p = 2; while (p < v) p <<= 1; return p; }

template <typename R>
int launch_select(const float* cand, const long long* cand_off, long long fixed_stride, const long long* counts,
                  int fixed_count, long long max_count, int nq, int k, const R& res, float* out_scores, int64_t* out_ids,
                  cudaStream_t st) {
    // candidates of a query are cached in shared memory when they fit 8192 floats (32 KB + the 16 KB sort buffer keeps 4 CTAs
    // per SM: measured 0.746 ms per C3 batch against 0.776 ms at 12288 floats / 3 CTAs); larger queries stream their
    // candidates from L2 in every pass.  RB200_SELECT_CACHE overrides the size (tuning knob).
    static int cap_max = 0;
    if (!cap_max) { const char* e = getenv("RB200_SELECT_CACHE"); cap_max = e ? atoi(e) : 8192; if (cap_max < 1024) cap_max = 1024; }
    int cache_cap = (int)(max_count < cap_max ? max_count : cap_max);
    if (cache_cap < 0) cache_cap = 0;
    int sort_cap = 256;
    while (sort_cap < 4 * k && sort_cap < SORT_CAP) sort_cap <<= 1;
    const size_t smem = (size_t)sort_cap * 8 + (size_t)cache_cap * 4 + (size_t)res.aux_ints() * 4 + 16;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        RB_CUDA(cudaFuncSetAttribute(select_topk_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        attr_smem = smem;
    }
    select_topk_kernel<R><<<nq, NT, smem, st>>>(cand, cand_off, fixed_stride, counts, fixed_count, k, cache_cap, sort_cap, res, out_scores,
                                                out_ids);
    RB_LAUNCH_CHECK("select_topk_kernel");
    return RB200_OK;
}

// k-way merge of `parts` sorted lists per query: treat the concatenation as candidates
struct ResolveMerge {
    const int64_t* ids; int parts; int nq; int k;
    __host__ __device__ int aux_ints() const { return 0; }
    __device__ void stage(int, int*, int, int) {}
};
template <> __device__ long long resolve_id<ResolveMerge>(const ResolveMerge& r, int q, int c) {
    const int part = c / r.k, j = c - part * r.k;
    return r.ids[((long long)part * r.nq + q) * r.k + j];
}
__global__ void __launch_bounds__(NT) merge_gather_kernel(const float* __restrict__ scores, int parts, int nq, int k,
                                                          float* __restrict__ cand) {
    const long long i = (long long)blockIdx.x * NT + threadIdx.x;
    const long long total = (long long)nq * parts * k;
    if (i >= total) return;
    const int q = (int)(i / ((long long)parts * k));
    const int rem = (int)(i - (long long)q * parts * k);
    const int part = rem / k, j = rem - part * k;
    cand[i] = scores[((long long)part * nq + q) * k + j];
}

struct PlanLayout {
    float* coarse; int* probes; int* cand_base; long long* totals; long long* cand_off; long long* tot2;
    int* list_qcount; int* list_qstart; int* pair_qp; long long* pair_dst;
    float* probe_scores; int64_t* probe_ids;
    int* tc_err;               // set by the tcgen05 coarse GEMM when one of its bounded barrier waits times out
};
bool carve_plan(RbArena& ar, int nq, int nlist, int nprobe, PlanLayout& L) {
    const size_t np = (size_t)nq * nprobe;
    L.coarse = ar.take<float>((size_t)nq * nlist);
    L.probes = ar.take<int>(np); L.cand_base = ar.take<int>(np);
    L.totals = ar.take<long long>((size_t)nq + 1); L.cand_off = ar.take<long long>((size_t)nq + 1);
    L.tot2 = ar.take<long long>(2);
    L.list_qcount = ar.take<int>((size_t)nlist + 2); L.list_qstart = ar.take<int>((size_t)nlist + 2);
    L.pair_qp = ar.take<int>(np); L.pair_dst = ar.take<long long>(np);
    L.probe_scores = ar.take<float>(np); L.probe_ids = ar.take<int64_t>(np);
    L.tc_err = ar.take<int>(4);
    return ar.ok();
}

#define RB_DISPATCH_D(D, CALL)                                            \
    switch (D) {                                                          \
        case 32: { constexpr int DD = 32; CALL; } break;                  \
        case 64: { constexpr int DD = 64; CALL; } break;                  \
        case 128: { constexpr int DD = 128; CALL; } break;                \
        default: return rb_set_error(RB200_ERR_INVALID, "unsupported D=%d (32, 64 or 128)", D); \
    }

}  // namespace

extern "C" int rb200_normalize_rows(const float* x, int64_t n, int D, float eps, float* out, void* stream) {
    RB_REQUIRE(x && out && n >= 0 && D >= 1, "normalize_rows: bad arguments");
    if (n == 0) return RB200_OK;
    normalize_rows_kernel<<<(unsigned)((n * 32 + NT - 1) / NT), NT, 0, (cudaStream_t)stream>>>(x, n, D, eps, out);
    RB_LAUNCH_CHECK("normalize_rows_kernel");
    return RB200_OK;
}

extern "C" int rb200_ivf_assign(const float* x, int64_t n, int D, const float* centroids, int nlist, int32_t* assign_out,
                                float* best_score, void* stream) {
    RB_REQUIRE(x && centroids && assign_out && n >= 0 && nlist >= 1, "ivf_assign: bad arguments");
    if (n == 0) return RB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    RB_DISPATCH_D(D, RB_TILE_LAUNCH(assign_kernel, DD, (unsigned)((n + TT - 1) / TT), st, x, n, centroids, nlist, assign_out, best_score));
    RB_LAUNCH_CHECK("assign_kernel");
    return RB200_OK;
}

extern "C" size_t rb200_kmeans_update_workspace_bytes(int64_t n, int nlist) { return sort_ws_bytes(n, nlist); }

extern "C" int rb200_kmeans_update(const float* x, int64_t n, int D, const int32_t* assign, int nlist, float* centroids,
                                   int32_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
    RB_REQUIRE(x && assign && centroids && n >= 1 && nlist >= 1 && n < (1ll << 31), "kmeans_update: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    RbArena ar(workspace, workspace_bytes);
    SortWs w;
    if (!workspace || !carve_sort_ws(ar, n, nlist, w)) return rb_set_error(RB200_ERR_WORKSPACE, "kmeans_update: workspace too small");
    int rc = sort_rows_by_list(assign, n, nlist, w, st);
    if (rc) return rc;
    centroid_update_kernel<<<(nlist * 32 + NT - 1) / NT, NT, 0, st>>>(x, w.order, w.excl, w.counts, nlist, D, centroids);
    RB_LAUNCH_CHECK("centroid_update_kernel");
    if (counts) RB_CUDA(cudaMemcpyAsync(counts, w.counts, sizeof(int) * nlist, cudaMemcpyDeviceToDevice, st));
    return RB200_OK;
}

extern "C" size_t rb200_ivf_build_workspace_bytes(int64_t n, int nlist) { return sort_ws_bytes(n, nlist); }

extern "C" int rb200_ivf_build(const float* x, int64_t n, int D, const int32_t* assign, int nlist, int64_t* offsets,
                               int64_t* list_ids, float* list_vecs, void* workspace, size_t workspace_bytes, void* stream) {
    RB_REQUIRE(x && assign && offsets && list_ids && list_vecs && n >= 1 && nlist >= 1 && n < (1ll << 31) && D % 4 == 0,
               "ivf_build: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    RbArena ar(workspace, workspace_bytes);
    SortWs w;
    if (!workspace || !carve_sort_ws(ar, n, nlist, w)) return rb_set_error(RB200_ERR_WORKSPACE, "ivf_build: workspace too small");
    int rc = sort_rows_by_list(assign, n, nlist, w, st);
    if (rc) return rc;
    offsets_to_i64_kernel<<<(nlist + NT - 1) / NT, NT, 0, st>>>(w.excl, w.counts, nlist, offsets);
    RB_LAUNCH_CHECK("offsets_to_i64_kernel");
    gather_rows_kernel<<<(unsigned)((n * 32 + NT - 1) / NT), NT, 0, st>>>(x, w.order, n, D, list_ids, list_vecs);
    RB_LAUNCH_CHECK("gather_rows_kernel");
    return RB200_OK;
}

extern "C" size_t rb200_ivf_plan_workspace_bytes(int nq, int nlist, int nprobe) {
    const size_t np = (size_t)nq * nprobe;
    return 256 * 20 + sizeof(float) * ((size_t)nq * nlist + np) + sizeof(int64_t) * 2 * np + sizeof(int) * (3 * np + 2 * ((size_t)nlist + 2)) +
           sizeof(long long) * (2 * ((size_t)nq + 1) + 2);
}

extern "C" int rb200_ivf_search_plan(const float* q, int nq, int D, const float* centroids, int nlist, int nprobe,
                                     const int64_t* offsets, void* plan_ws, size_t plan_ws_bytes,
                                     int64_t* total_candidates_host, int64_t* max_candidates_host, void* stream) {
    RB_REQUIRE(q && centroids && offsets && nq >= 1 && nlist >= 1 && nprobe >= 1 && nprobe <= nlist, "ivf_search_plan: bad arguments");
    RB_REQUIRE((total_candidates_host == nullptr) == (max_candidates_host == nullptr), "ivf_search_plan: pass both host outputs or neither");
    RB_REQUIRE(nprobe <= 2048, "ivf_search_plan: nprobe must be <= 2048");
    cudaStream_t st = (cudaStream_t)stream;
    RbArena ar(plan_ws, plan_ws_bytes);
    PlanLayout L;
    if (!plan_ws || !carve_plan(ar, nq, nlist, nprobe, L)) return rb_set_error(RB200_ERR_WORKSPACE, "ivf_search_plan: workspace too small");
    {   // coarse quantizer q·Cᵀ on the tensor cores (3xTF32, fp32 accumulate in TMEM)
        RB_CUDA(cudaMemsetAsync(L.tc_err, 0, sizeof(int), st));
        int rc = rb_gemm_nt_tc(q, nq, centroids, nlist, D, 2, L.coarse, nlist, L.tc_err, st);
        if (rc) return rc;
    }
    RB_CUDA(cudaMemsetAsync(L.list_qcount, 0, sizeof(int) * ((size_t)nlist + 2), st));
    {
        ResolveIdentity res{0};
        int rc = launch_select<ResolveIdentity>(L.coarse, nullptr, nlist, nullptr, nlist, nlist, nq, nprobe, res, L.probe_scores,
                                                L.probe_ids, st);
        if (rc) return rc;
    }
    probe_finish_kernel<<<(unsigned)(((long long)nq * 32 + NT - 1) / NT), NT, 0, st>>>(L.probe_ids, nq, nprobe, offsets, L.probes, L.cand_base, L.totals,
                                                         L.list_qcount);
    RB_LAUNCH_CHECK("probe_finish_kernel");
    // the (list, query) pairs grouped by list: a counting sort on the per-list probe counts of probe_finish_kernel — one scan launch
    // (lists and per-query candidate offsets together) and one scatter, instead of cub's radix sort + two scans + a reduction
    // (54 of the plan's 193 us in profiles/r01_launches_ivf_v7.csv)
    const long long np = (long long)nq * nprobe;
    plan_scans_kernel<<<1, PS_NT, 0, st>>>(L.list_qcount, nlist, L.list_qstart, L.totals, nq, L.cand_off, L.tot2);
    RB_LAUNCH_CHECK("plan_scans_kernel");
    pair_scatter_kernel<<<(unsigned)((np + NT - 1) / NT), NT, 0, st>>>(L.probes, np, nprobe, L.list_qstart, L.list_qcount, L.cand_off,
                                                                       L.cand_base, L.pair_qp, L.pair_dst);
    RB_LAUNCH_CHECK("pair_scatter_kernel");
    if (!total_candidates_host) return RB200_OK;      // asynchronous form (CUDA-graph capturable): the caller sizes by upper bounds
    long long h[2] = {0, 0};
    int tc_err = 0;
    RB_CUDA(cudaMemcpyAsync(h, L.tot2, sizeof(h), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaMemcpyAsync(&tc_err, L.tc_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    // (the synchronous form is where a timed-out tensor-core pipeline can be reported: garbage probe lists must not pass as a result)
    RB_REQUIRE(tc_err == 0, "ivf_search_plan: the coarse-quantizer GEMM's tensor-core pipeline timed out (flag %d)", tc_err);
    *total_candidates_host = h[0];
    *max_candidates_host = h[1];
    return RB200_OK;
}

extern "C" int rb200_ivf_search_status(const void* plan_ws, size_t plan_ws_bytes, int nq, int nlist, int nprobe, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    RbArena pa(const_cast<void*>(plan_ws), plan_ws_bytes);
    PlanLayout L;
    if (!plan_ws || !carve_plan(pa, nq, nlist, nprobe, L)) return rb_set_error(RB200_ERR_WORKSPACE, "ivf_search_status: plan workspace too small");
    int flag = 0;
    RB_CUDA(cudaMemcpyAsync(&flag, L.tc_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    RB_REQUIRE(flag == 0, "ivf_search: the list scan's tensor-core pipeline timed out (flag %d): the results are not valid", flag);
    return RB200_OK;
}

extern "C" size_t rb200_ivf_search_workspace_bytes(int64_t total_candidates) {
    return 256 + sizeof(float) * (size_t)(total_candidates > 0 ? total_candidates : 1);
}

extern "C" int rb200_ivf_search_run(const float* q, int nq, int D, int nlist, int nprobe, const int64_t* offsets,
                                    const int64_t* list_ids, const float* list_vecs, int64_t n_vectors, int64_t max_list_len,
                                    const int32_t* tile_list, const int32_t* tile_idx, int n_tiles, int k,
                                    void* plan_ws, size_t plan_ws_bytes, int64_t total_candidates, int64_t max_candidates,
                                    float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                                    void* stream) {
    RB_REQUIRE(q && offsets && list_ids && list_vecs && out_scores && out_ids, "ivf_search_run: NULL pointer");
    RB_REQUIRE(nq >= 1 && k >= 1 && k <= 2048 && nprobe >= 1 && max_list_len >= 0 && n_vectors >= 0, "ivf_search_run: bad sizes (k must be 1..2048)");
    cudaStream_t st = (cudaStream_t)stream;
    RbArena pa(plan_ws, plan_ws_bytes);
    PlanLayout L;
    if (!plan_ws || !carve_plan(pa, nq, nlist, nprobe, L)) return rb_set_error(RB200_ERR_WORKSPACE, "ivf_search_run: plan workspace too small");
    RbArena ar(workspace, workspace_bytes);
    float* cand = ar.take<float>((size_t)(total_candidates > 0 ? total_candidates : 1));
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "ivf_search_run: workspace too small");
    if (total_candidates > 0 && max_list_len > 0) {
        // D = 64 with a tile table: the persistent, prefetching tcgen05 kernel (ivf_scan_tc.cu); RB200_IVF_PIPE=0: one CTA per tile
        int tc = rb_list_scan_pipe(q, D, list_vecs, n_vectors, offsets, L.pair_qp, L.pair_dst, L.list_qstart, nprobe, cand, total_candidates, tile_list,
                                   tile_idx, n_tiles, L.tc_err, st);
        if (tc > 0)                 // (otherwise the FFMA tile kernel below)
            tc = rb_list_scan_tc(q, D, list_vecs, offsets, L.pair_qp, L.list_qstart, nprobe, L.cand_base, L.cand_off, cand,
                                 tile_list, tile_idx, n_tiles, st);
        if (tc < 0) return tc;
        if (tc != 0) {
            dim3 g((unsigned)((max_list_len + TT - 1) / TT), nlist);
            if (tile_list && tile_idx && n_tiles > 0) g = dim3((unsigned)n_tiles, 1);
            else { tile_list = nullptr; tile_idx = nullptr; RB_REQUIRE(nlist <= 65535, "ivf_search_run: nlist must be <= 65535 without a tile table"); }
            RB_DISPATCH_D(D, RB_TILE_LAUNCH(list_scan_kernel, DD, g, st, q, list_vecs, offsets, L.pair_qp, L.list_qstart, nprobe,
                                            L.cand_base, L.cand_off, cand, tile_list, tile_idx));
            RB_LAUNCH_CHECK("list_scan_kernel");
        }
    }
    ResolveIvf res{L.probes, L.cand_base, offsets, list_ids, nprobe, nullptr, nullptr};
    return launch_select<ResolveIvf>(cand, L.cand_off, 0, L.totals, 0, max_candidates, nq, k, res, out_scores, out_ids, st);
}

// ------------------------------------------------------------------------------------------ //
// flat (exhaustive) search: chunks of rows, running top-k carried as the first k candidates
// ------------------------------------------------------------------------------------------ //
static long long flat_chunk_rows(int nq, int64_t n) {
    long long c = (256ll << 20) / ((long long)nq * 4);        // ~256 MB of chunk scores
    c = c / TT * TT;
    if (c < 4096) c = 4096;
    if (c > 1 << 20) c = 1 << 20;
    if (c > n) c = (n + TT - 1) / TT * TT;
    return c;
}

// Threshold-pruned rounds (flat_scan_tc.cu) take over after the first exact chunk when D = 64: round j scans rows
// [seen, 4·seen) against thr[q] = the k-th best score over [0, seen), so it is expected to leave 3·k survivors per query;
// the survivor lists hold FLAT_CAP_K·k (at least 3584) entries — if one ever fills up (a database sorted by score towards a
// query), the search is redone on the chunked path, which has no such limit.
// At most 128 queries (the scan is bound by reading the rows and a round's fixed cost — prep, re-score, select, launch gaps ≈ 30 us —
// is what is left to save): rounds grow ×8 (7·k expected survivors, lists of 16·k), four rounds instead of six over 12.5 M rows.
constexpr int FLAT_CAP_MIN = 3584;
static int flat_growth(int nq) { return nq <= 128 ? 8 : 4; }
static int flat_cap(int k, int nq) {
    const int c = (nq <= 128 ? 16 : 7) * k;
    return ((c < FLAT_CAP_MIN ? FLAT_CAP_MIN : c) + 511) / 512 * 512;
}
static bool flat_use_rounds(int D, int64_t n, long long chunk) {
    static int off = -1;
    if (off < 0) { const char* e = getenv("RB200_FLAT_CHUNKED"); off = (e && atoi(e)) ? 1 : 0; }     // testing knob: chunked path only
    (void)chunk;
    return !off && D == 64 && n > 4 * 8192;
}

extern "C" size_t rb200_flat_search_workspace_bytes(int nq, int64_t n, int k) {
    const long long chunk = flat_chunk_rows(nq, n);
    const long long cap = flat_cap(k, nq), wide = chunk > cap ? chunk : cap;
    const size_t nq_pad = ((size_t)nq + 127) / 128 * 128;
    return 256 * 12 + sizeof(float) * (size_t)nq * (size_t)(wide + k) + (sizeof(float) + sizeof(int64_t)) * (size_t)nq * k
           + nq_pad * 64 * 8 + nq_pad * 12 + sizeof(long long) * (size_t)nq + sizeof(int) * (size_t)nq * cap + 64;
}

__global__ void copy_prev_kernel(const float* __restrict__ prev_scores, int nq, int k, float* __restrict__ cand, long long stride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nq * k) return;
    const int q = (int)(i / k), j = (int)(i - (long long)q * k);
    cand[q * stride + j] = prev_scores[i];
}

// start of a pruned round: running top-k in front of the survivor list, thr = its k-th score (+inf for the padding queries of the
// last 64-query chunk), survivor counters cleared
__global__ void flat_round_prep_kernel(const float* __restrict__ prev_scores, int nq, int nq_pad, int k, float* __restrict__ cand,
                                       long long stride, float* __restrict__ thr, int* __restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nq_pad * k) return;
    const int q = (int)(i / k), j = (int)(i - (long long)q * k);
    if (q < nq) cand[q * stride + j] = prev_scores[i];
    if (j == k - 1) { thr[q] = q < nq ? prev_scores[i] : FLT_MAX; count[q] = 0; }
}
__global__ void flat_round_lens_kernel(const int* __restrict__ count, int nq, int k, int cap, long long* __restrict__ lens) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) lens[q] = k + (count[q] < cap ? count[q] : cap);
}

// chunked exact path over rows [r_begin, r_end): score GEMM → select, running top-k carried in front of the chunk's scores.
// bufS/bufI: two (scores, ids) buffers used alternately; cur = the one holding the running top-k (-1: none yet).
static int flat_chunked_range(const float* q, int nq, const float* x, int D, long long r_begin, long long r_end, long long chunk,
                              long long stride, int k, int64_t id_base, float* cand, float* const bufS[2], int64_t* const bufI[2],
                              int& cur, cudaStream_t st) {
    for (long long r0 = r_begin; r0 < r_end; r0 += chunk) {
        const long long rows = (r_end - r0 < chunk) ? r_end - r0 : chunk;
        const int kprev = cur < 0 ? 0 : k, nxt = cur < 0 ? 0 : cur ^ 1;
        if (kprev) {
            copy_prev_kernel<<<(unsigned)(((long long)nq * k + NT - 1) / NT), NT, 0, st>>>(bufS[cur], nq, k, cand, stride);
            RB_LAUNCH_CHECK("copy_prev_kernel");
        }
        {   // chunk scores on the tensor cores (3xTF32)
            int rc = rb_gemm_nt_tc(q, nq, x + r0 * D, (int)rows, D, 2, cand + kprev, stride, nullptr, st);
            if (rc) return rc;
        }
        ResolveFlat res{kprev ? bufI[cur] : nullptr, kprev, id_base + r0};
        // previous winners may contain -FLT_MAX padding (when fewer than k rows so far): they carry id -1
        int rc = launch_select<ResolveFlat>(cand, nullptr, stride, nullptr, (int)(rows + kprev), rows + kprev, nq, k, res, bufS[nxt],
                                            bufI[nxt], st);
        if (rc) return rc;
        cur = nxt;
    }
    return RB200_OK;
}

extern "C" int rb200_flat_search(const float* q, int nq, const float* x, int64_t n, int D, int k, int64_t id_base,
                                 float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream) {
    RB_REQUIRE(q && x && out_scores && out_ids && nq >= 1 && n >= 1 && k >= 1 && k <= 2048, "flat_search: bad arguments (k must be 1..2048)");
    cudaStream_t st = (cudaStream_t)stream;
    const long long chunk = flat_chunk_rows(nq, n);
    const int cap = flat_cap(k, nq), growth = flat_growth(nq);
    const long long stride = (chunk > cap ? chunk : cap) + k;
    const int n_qchunks = (nq + 63) / 64, nq_pad = (nq + 127) / 128 * 128;     // 64-query chunks of the scan / stream kernels; padded rows
    RbArena ar(workspace, workspace_bytes);
    float* cand = ar.take<float>((size_t)nq * stride);
    float* tmp_scores = ar.take<float>((size_t)nq * k);
    int64_t* tmp_ids = ar.take<int64_t>((size_t)nq * k);
    unsigned char* qimg = ar.take<unsigned char>((size_t)nq_pad * 64 * 8);
    float* thr = ar.take<float>(nq_pad);
    int* count = ar.take<int>(nq_pad);
    float* qmarg = ar.take<float>(nq_pad);
    long long* lens = ar.take<long long>(nq);
    int* cand_r = ar.take<int>((size_t)nq * cap);
    int* flags = ar.take<int>(4);
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "flat_search: workspace too small");
    float* const bufS[2] = {out_scores, tmp_scores};
    int64_t* const bufI[2] = {out_ids, tmp_ids};
    int cur = -1, rc;
    const bool rounds = flat_use_rounds(D, n, chunk);
    // With pruned rounds behind it the exact prefix only has to seed the thresholds: 8192 rows fit the select kernel's shared-memory
    // candidate cache (a 16384-row first chunk spent 2.6 ms in select_topk_kernel per 4096-query batch, 7 % of the whole search)
    const long long n0 = rounds ? (chunk < 8192 ? chunk : 8192) : n;
    if ((rc = flat_chunked_range(q, nq, x, D, 0, n0, chunk, stride, k, id_base, cand, bufS, bufI, cur, st))) return rc;
    if (rounds) {
        RB_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
        const bool streamed = rb_flat_streamed(n_qchunks);     // ≤ 128 queries: the scan is bound by reading the rows
        const bool filter_tc = !streamed && rb_flat_filtered(nq);     // 129 … 8192 queries: bound by the tensor pipe
        const bool filtered = streamed || filter_tc;           // one-pass TF32 filter: survivors are re-scored in fp32 before the select
        if ((rc = filter_tc ? rb_flat_filter_image(q, nq, nq_pad, qimg, qmarg, st) : rb_flat_qimage(q, nq, nq_pad, 64, qimg, qmarg, st))) return rc;
        const long long rstride = (long long)k + cap;
        for (long long seen = n0; seen < n;) {
            const long long upto = (n / growth >= seen) ? seen * growth : n;
            const long long rows = (upto < n ? upto : n) - seen;
            flat_round_prep_kernel<<<(unsigned)(((long long)nq_pad * k + NT - 1) / NT), NT, 0, st>>>(bufS[cur], nq, nq_pad, k, cand, rstride,
                                                                                                  thr, count);
            RB_LAUNCH_CHECK("flat_round_prep_kernel");
            if (filter_tc && (rc = rb_flat_filter_thresholds(thr, qmarg, nq_pad, qimg, st))) return rc;
            rc = streamed    ? rb_flat_stream_tc(x + seen * D, rows, qimg, n_qchunks, thr, qmarg, count, cand, rstride, k, cand_r, cap, flags, st)
                 : filter_tc ? rb_flat_filter_tc(x + seen * D, rows, qimg, nq_pad / 128, thr, qmarg, count, cand, rstride, k, cand_r, cap, flags, st)
                             : rb_flat_scan_tc(x + seen * D, rows, qimg, n_qchunks, thr, qmarg, count, cand, rstride, k, cand_r, cap, flags, st);
            if (rc) return rc;
            if (filtered && (rc = rb_flat_rescore(q, nq, x + seen * D, count, cand_r, cap, cand, rstride, k, st))) return rc;
            flat_round_lens_kernel<<<(nq + NT - 1) / NT, NT, 0, st>>>(count, nq, k, cap, lens);
            RB_LAUNCH_CHECK("flat_round_lens_kernel");
            ResolveSurv res{bufI[cur], k, cand_r, cap, id_base + seen};
            if ((rc = launch_select<ResolveSurv>(cand, nullptr, rstride, lens, 0, rstride, nq, k, res, bufS[cur ^ 1], bufI[cur ^ 1], st)))
                return rc;
            cur ^= 1;
            seen += rows;
        }
        int h[4] = {0, 0, 0, 0};
        RB_CUDA(cudaMemcpyAsync(h, flags, sizeof(h), cudaMemcpyDeviceToHost, st));
        RB_CUDA(cudaStreamSynchronize(st));
        RB_REQUIRE(h[1] == 0, "flat_search: tensor-core pipeline timed out (flag %d, first wait 0x%x)", h[1], h[2]);
        if (h[0]) {            // a survivor list overflowed: redo everything on the chunked path (exact for any row order)
            cur = -1;
            if ((rc = flat_chunked_range(q, nq, x, D, 0, n, chunk, stride, k, id_base, cand, bufS, bufI, cur, st))) return rc;
        }
    }
    if (cur == 1) {
        RB_CUDA(cudaMemcpyAsync(out_scores, tmp_scores, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToDevice, st));
        RB_CUDA(cudaMemcpyAsync(out_ids, tmp_ids, sizeof(int64_t) * (size_t)nq * k, cudaMemcpyDeviceToDevice, st));
    }
    return RB200_OK;
}

extern "C" size_t rb200_topk_merge_workspace_bytes(int parts, int nq, int k) {
    return 256 + sizeof(float) * (size_t)parts * nq * k;
}

extern "C" int rb200_topk_merge(const float* scores, const int64_t* ids, int parts, int nq, int k, float* out_scores,
                                int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream) {
    RB_REQUIRE(scores && ids && out_scores && out_ids && parts >= 1 && nq >= 1 && k >= 1 && k <= 2048, "topk_merge: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    RbArena ar(workspace, workspace_bytes);
    float* cand = ar.take<float>((size_t)parts * nq * k);
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "topk_merge: workspace too small");
    const long long total = (long long)parts * nq * k;
    merge_gather_kernel<<<(unsigned)((total + NT - 1) / NT), NT, 0, st>>>(scores, parts, nq, k, cand);
    RB_LAUNCH_CHECK("merge_gather_kernel");
    ResolveMerge res{ids, parts, nq, k};
    return launch_select<ResolveMerge>(cand, nullptr, (long long)parts * k, nullptr, parts * k, (long long)parts * k, nq, k, res,
                                       out_scores, out_ids, st);
}
