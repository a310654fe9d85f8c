// Sparse embedding-gradient reduction, gradient clipping and Adam (sm_100a).
//
//   rb200_scatter_rows      autograd of nn.Embedding(padding_idx=0) (two_tower.py:27,54): deterministic
//                           sorted-segment sum — radix sort of (id, sample) pairs, then one warp per
//                           distinct id adds that id's rows in ascending sample order.
//   rb200_sumsq_* / clip    torch.nn.utils.clip_grad_norm_ (train_embeddings.py:191)
//   rb200_adam_*            torch.optim.Adam(weight_decay=1e-5) (train_embeddings.py:160,192)
//
// All of it is HBM-bound streaming: 128-bit loads/stores, one pass per tensor.
#include <cub/cub.cuh>

#include "common.cuh"

static_assert(sizeof(rb200_opt_state) == 128, "rb200_opt_state must be 128 bytes");

namespace {

constexpr int NT = 256;

__global__ void iota_kernel(int* __restrict__ v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// flag[i] = 1 when sorted position i starts a new id (padding id never starts a segment)
__global__ void head_flags_kernel(const int64_t* __restrict__ keys, int n, long long padding_idx, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { flags[n] = 0; return; }
    const long long k = keys[i];
    flags[i] = (k != padding_idx && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
}

// one warp per segment head
__global__ void __launch_bounds__(NT) segment_sum_kernel(const int64_t* __restrict__ keys, const int* __restrict__ pos,
                                                         const int* __restrict__ flags, const int* __restrict__ slots,
                                                         int n, int D, long long n_rows, const float* __restrict__ rows,
                                                         float* __restrict__ dense, int64_t* __restrict__ uniq_ids,
                                                         float* __restrict__ uniq_grads, int* __restrict__ n_uniq,
                                                         int* __restrict__ row_slot) {
    const int warp = (blockIdx.x * NT + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0 && n_uniq) n_uniq[0] = slots[n];
    if (warp >= n || !flags[warp]) return;
    const long long key = keys[warp];
    if ((unsigned long long)key >= (unsigned long long)n_rows) return;   // out-of-range ids are dropped (fwd flagged them)
    const int slot = slots[warp];
    int end = warp + 1;
    while (end < n && keys[end] == key) ++end;
    const int D4 = D >> 2;
    for (int c = lane; c < D4; c += 32) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = warp; j < end; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(rows + (long long)pos[j] * D) + c);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        if (uniq_grads) reinterpret_cast<float4*>(uniq_grads + (long long)slot * D)[c] = s;
        if (dense) {
            float4* dp = reinterpret_cast<float4*>(dense + key * D) + c;
            float4 o = *dp;
            o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
            *dp = o;
        }
    }
    if (lane == 0) {
        if (uniq_ids) uniq_ids[slot] = key;
        if (row_slot) row_slot[key] = slot;
    }
}

__global__ void set_slots_kernel(const int64_t* __restrict__ uniq_ids, const int* __restrict__ n_uniq, int* __restrict__ row_slot) {
    const int n = n_uniq[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) row_slot[uniq_ids[i]] = i;
}

// out[i] = table[rows[i]] — plain coalesced 128-bit row gather (row-sharded tables: owners serve the requested rows)
__global__ void __launch_bounds__(NT) gather_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ rows,
                                                         long long n, int D4, long long n_table_rows, float* __restrict__ out) {
    const long long total = n * D4, stride = (long long)gridDim.x * NT;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += stride) {
        const long long r = i / D4;
        const int c = (int)(i - r * D4);
        long long src = rows[r];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((unsigned long long)src < (unsigned long long)n_table_rows) v = __ldg(reinterpret_cast<const float4*>(table) + src * D4 + c);
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

__global__ void reset_slots_kernel(const int64_t* __restrict__ uniq_ids, const int* __restrict__ n_uniq, int* __restrict__ row_slot) {
    const int n = n_uniq[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) row_slot[uniq_ids[i]] = -1;
}
__global__ void reset_slots2_kernel(const int64_t* __restrict__ ids0, const int* __restrict__ n0, int* __restrict__ slot0,
                                    const int64_t* __restrict__ ids1, const int* __restrict__ n1, int* __restrict__ slot1) {
    const int a = n0[0], b = n1[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a + b; i += gridDim.x * blockDim.x) {
        if (i < a) slot0[ids0[i]] = -1;
        else slot1[ids1[i - a]] = -1;
    }
}


// ---------------------------------------------------------------------------------------- //
// Fast path (≤ 16384 sample-rows per table, i.e. every batch of BASELINE configs C1/C2/C4): ONE CTA per table
// radix-sorts the (id, sample) pairs in shared memory (cub::BlockRadixSort), finds the segment heads, scans them and
// emits [sorted sample positions | segment starts | unique ids | row slots] — 1 launch instead of 9.
// ---------------------------------------------------------------------------------------- //
constexpr int BS_THREADS = 1024;
constexpr unsigned KEY_SENTINEL = 0xFFFFFFFFu;

struct SortJob {
    const int64_t* ids_a; const int64_t* ids_b; int n_a, n_b;   // the id list is the concatenation a ++ b
    long long n_rows, padding_idx; int key_bits;
    int* pos; int* seg_start; int64_t* uniq_ids; int* n_uniq; int* row_slot;
};
struct SortParams { SortJob job[2]; };

template <int ITEMS>
__global__ void __launch_bounds__(BS_THREADS, 1) block_sort_segments_kernel(const SortParams p) {
    using Sort = cub::BlockRadixSort<unsigned, BS_THREADS, ITEMS, int>;
    using Scan = cub::BlockScan<int, BS_THREADS>;
    extern __shared__ __align__(16) unsigned char bs_smem[];
    typename Sort::TempStorage& sort_tmp = *reinterpret_cast<typename Sort::TempStorage*>(bs_smem);
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ unsigned last_key[BS_THREADS];
    const SortJob& J = p.job[blockIdx.x];
    const int n = J.n_a + J.n_b, tid = threadIdx.x;
    unsigned keys[ITEMS];
    int vals[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int idx = tid * ITEMS + i;
        unsigned key = KEY_SENTINEL;
        if (idx < n) {
            const long long id = idx < J.n_a ? J.ids_a[idx] : J.ids_b[idx - J.n_a];
            if (id != J.padding_idx && (unsigned long long)id < (unsigned long long)J.n_rows) key = (unsigned)id;
        }
        keys[i] = key;
        vals[i] = idx;
    }
    // 2^key_bits > n_rows, so the sentinel's low bits (all ones) sort strictly after every valid id;
    // the sort is stable ⇒ equal ids keep ascending sample order
    Sort(sort_tmp).Sort(keys, vals, 0, J.key_bits);
    last_key[tid] = keys[ITEMS - 1];
    __syncthreads();
    int heads = 0, valid = 0;
    bool is_head[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const unsigned prev = i > 0 ? keys[i - 1] : (tid > 0 ? last_key[tid - 1] : KEY_SENTINEL);
        const bool v = keys[i] != KEY_SENTINEL;
        is_head[i] = v && ((tid == 0 && i == 0) || keys[i] != prev);
        heads += is_head[i];
        valid += v;
    }
    int base, total, vbase, vtotal;
    Scan(scan_tmp).ExclusiveSum(heads, base, total);
    __syncthreads();
    Scan(scan_tmp).ExclusiveSum(valid, vbase, vtotal);
    (void)vbase;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int gi = tid * ITEMS + i;
        if (gi < n) J.pos[gi] = vals[i];
        if (is_head[i]) {
            J.seg_start[base] = gi;
            J.uniq_ids[base] = (int64_t)keys[i];
            if (J.row_slot) J.row_slot[keys[i]] = base;
            ++base;
        }
    }
    if (tid == 0) {
        J.seg_start[total] = vtotal;
        J.n_uniq[0] = total;
    }
}

struct SegJob {
    const int* pos; const int* seg_start; const int64_t* uniq_ids; const int* n_uniq;
    const float* rows; float* uniq_grads; float* dense; int cap;   // cap = upper bound of n_uniq (launch sizing)
    const int* first_pos;                                          // optional: pos[seg_start[s]] per segment (shorter load chain)
};
struct SegParams { SegJob job[2]; int n_jobs; int D4; int skip_long; int* long_count; };   // skip_long: segments over LONG_SEG rows are left to long_segments_kernel
constexpr int LONG_SEG = 8;           // (a chain of 48 rows still cost 55 us of the C4 step: two dependent loads per link)

// s += rows[pos[k]] (column c) for k = k0 … k1-1 in ascending k.  One pos → row dependent-latency chain per row: fine for the short
// segments this is used for; segments of a popular id (Zipf-skewed users: one id can own 9 % of a batch, 700 links, 333 us of the
// C4 step while the rest of the grid had long finished) are summed by many warps instead — long_segments_kernel after
// segment_sum2_kernel, the whole block inside grad_finish_kernel.
__device__ __forceinline__ void seg_accumulate(float4& s, const float4* __restrict__ rows, const int* __restrict__ pos, int k0, int k1,
                                               int D4, int c) {
    for (int k = k0; k < k1; ++k) {
        const float4 v = __ldg(rows + (long long)__ldg(pos + k) * D4 + c);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
}

// one warp per SEG_PER_WARP consecutive unique ids; the first row of each segment is fetched for all of them at once
// (most segments of a large batch have length 1, so this keeps SEG_PER_WARP independent 16·D4-byte row reads in
// flight per warp instead of one at the end of a seg_start → pos → row dependency chain).  Rows are added in ascending
// sample order (deterministic).
template <int SEG_PER_WARP>
__host__ __device__ inline int seg_warps(int cap) { return (cap + SEG_PER_WARP - 1) / SEG_PER_WARP; }

template <int SEG_PER_WARP>
__global__ void __launch_bounds__(NT) segment_sum2_kernel(const SegParams p) {
    int w = (blockIdx.x * NT + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    int j = 0;
    if (p.n_jobs > 1 && w >= seg_warps<SEG_PER_WARP>(p.job[0].cap)) { w -= seg_warps<SEG_PER_WARP>(p.job[0].cap); j = 1; }
    const SegJob& J = p.job[j];
    const int n = J.n_uniq[0];
    const int w0 = w * SEG_PER_WARP;
    if (w0 >= n) return;
    const int mine = (lane <= SEG_PER_WARP && w0 + lane <= n) ? J.seg_start[w0 + lane] : 0;
    const int first = (lane < SEG_PER_WARP && w0 + lane < n) ? (J.first_pos ? J.first_pos[w0 + lane] : J.pos[mine]) : 0;
    long long dst = (J.dense && lane < SEG_PER_WARP && w0 + lane < n) ? (long long)J.uniq_ids[w0 + lane] : 0;
    int beg[SEG_PER_WARP], end[SEG_PER_WARP], p0[SEG_PER_WARP];
    long long drow[SEG_PER_WARP];
#pragma unroll
    for (int i = 0; i < SEG_PER_WARP; ++i) {
        beg[i] = __shfl_sync(0xffffffffu, mine, i);
        end[i] = __shfl_sync(0xffffffffu, mine, i + 1);
        p0[i] = __shfl_sync(0xffffffffu, first, i);
        drow[i] = __shfl_sync(0xffffffffu, dst, i);
    }
    const int D4 = p.D4;
    const float4* __restrict__ rows = reinterpret_cast<const float4*>(J.rows);
    if (p.skip_long) {
        // a long segment (popular id) keeps only its FIRST row here; long_segments_kernel adds the rest with 32 warps
        bool any = false;
#pragma unroll
        for (int i = 0; i < SEG_PER_WARP; ++i)
            if (w0 + i < n && end[i] - beg[i] > LONG_SEG) { end[i] = beg[i] + 1; any = true; }
        if (any && p.long_count && lane == 0) atomicAdd(p.long_count, 1);
    }
    for (int c = lane; c < D4; c += 32) {
        float4 s[SEG_PER_WARP];
#pragma unroll
        for (int i = 0; i < SEG_PER_WARP; ++i)
            s[i] = (w0 + i < n) ? __ldg(rows + (long long)p0[i] * D4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < SEG_PER_WARP; ++i) {
            if (w0 + i >= n) break;
            for (int k = beg[i] + 1; k < end[i]; ++k) {
                const float4 v = __ldg(rows + (long long)J.pos[k] * D4 + c);
                s[i].x += v.x; s[i].y += v.y; s[i].z += v.z; s[i].w += v.w;
            }
            if (J.uniq_grads) __stcs(reinterpret_cast<float4*>(J.uniq_grads) + (long long)(w0 + i) * D4 + c, s[i]);
            if (J.dense) {
                float4* dp = reinterpret_cast<float4*>(J.dense) + drow[i] * D4 + c;
                float4 o = *dp;
                o.x += s[i].x; o.y += s[i].y; o.z += s[i].z; o.w += s[i].w;
                *dp = o;
            }
        }
    }
}

// Segments of more than LONG_SEG rows (a popular id under Zipf-skewed traffic: hundreds of rows) — one CTA of 32 warps per segment
// instead of one warp: segment_sum2_kernel has written the segment's first row; warp w adds rows beg+1+w, beg+1+w+32, … (ascending),
// the 32 partial sums are added in warp order and the total is added to the first row.
// Deterministic (the split depends only on the segment's length).  Block b looks at segments b, b+G, b+2G, ….
constexpr int LS_THREADS = 1024, LS_MAX_D = 256;
__global__ void __launch_bounds__(LS_THREADS) long_segments_kernel(const SegParams p) {
    __shared__ float part[32][LS_MAX_D + 4];
    __shared__ int list[LS_THREADS], wcount[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D4 = p.D4, D = D4 * 4;
    if (p.long_count && *p.long_count == 0) return;            // the segment kernel met no long segment: nothing to do
    for (int j = 0; j < p.n_jobs; ++j) {
        const SegJob& J = p.job[j];
        const int n = J.n_uniq[0];
        const float4* __restrict__ rows = reinterpret_cast<const float4*>(J.rows);
        for (int base = blockIdx.x; base < n; base += gridDim.x * LS_THREADS) {
            // this block's long segments among base, base+G, …, listed in index order (ballot + prefix over the warps)
            const int sidx = base + tid * gridDim.x;
            const bool is_long = sidx < n && J.seg_start[sidx + 1] - J.seg_start[sidx] > LONG_SEG;
            const unsigned bal = __ballot_sync(0xffffffffu, is_long);
            __syncthreads();                                   // previous round's list / partials are no longer read
            if (lane == 0) wcount[warp] = __popc(bal);
            __syncthreads();
            int off = __popc(bal & ((1u << lane) - 1u)), nl = 0;
            for (int w = 0; w < 32; ++w) { const int cw = wcount[w]; if (w < warp) off += cw; nl += cw; }
            if (is_long) list[off] = sidx;
            __syncthreads();
            for (int li = 0; li < nl; ++li) {
                const int sg = list[li];
                const int beg = J.seg_start[sg], end = J.seg_start[sg + 1];
                // warp w owns rows beg+1+w, beg+1+w+32, … (row `beg` was taken by segment_sum2_kernel).  Their positions are fetched 32 at a
                // time (one per lane) and handed round by shuffles, and four row reads are in flight per lane with one partial sum each
                // (i mod 4; added as (a0 + a1) + (a2 + a3)): a 5 900-row segment — the hottest user of a 65 536-sample Zipf batch, all of it
                // on one rank of the sharded step — took 66 us with one dependent pos → row chain per warp.
                const int n_mine = end > beg + 1 + warp ? (end - (beg + 1 + warp) + 31) / 32 : 0;      // rows of this warp
                for (int c0 = 0; c0 < D4; c0 += 32) {                   // (every lane walks the loop: the shuffles below are warp-wide)
                    const int c = c0 + lane;
                    const bool col = c < D4;
                    float4 a4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) a4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int i0 = 0; i0 < n_mine; i0 += 32) {
                        const int ki = i0 + lane;
                        const int pk = ki < n_mine ? __ldg(J.pos + beg + 1 + warp + 32 * ki) : 0;
                        const int lim = min(32, n_mine - i0);
                        for (int j0 = 0; j0 < lim; j0 += 4) {
                            float4 v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int pj = __shfl_sync(0xffffffffu, pk, (j0 + u) & 31);
                                v[u] = (col && j0 + u < lim) ? __ldg(rows + (long long)pj * D4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) { a4[u].x += v[u].x; a4[u].y += v[u].y; a4[u].z += v[u].z; a4[u].w += v[u].w; }
                        }
                    }
                    float4 a;
                    a.x = (a4[0].x + a4[1].x) + (a4[2].x + a4[3].x); a.y = (a4[0].y + a4[1].y) + (a4[2].y + a4[3].y);
                    a.z = (a4[0].z + a4[1].z) + (a4[2].z + a4[3].z); a.w = (a4[0].w + a4[1].w) + (a4[2].w + a4[3].w);
                    if (col) *reinterpret_cast<float4*>(&part[warp][c * 4]) = a;
                }
                __syncthreads();
                for (int d = tid; d < D; d += LS_THREADS) {
                    float t = part[0][d];
#pragma unroll 8
                    for (int w = 1; w < 32; ++w) t += part[w][d];
                    if (J.uniq_grads) J.uniq_grads[(long long)sg * D + d] += t;
                    if (J.dense) J.dense[J.uniq_ids[sg] * (long long)D + d] += t;
                }
                __syncthreads();
            }
        }
    }
}
static int launch_long_segments(const SegParams& gp, cudaStream_t st) {
    if (gp.D4 * 4 > LS_MAX_D) return RB200_OK;      // (callers clear skip_long for such widths)
    long_segments_kernel<<<64, LS_THREADS, 0, st>>>(gp);
    RB_LAUNCH_CHECK("long_segments_kernel");
    return RB200_OK;
}

int key_bits_strict(long long n_rows) {     // smallest b with 2^b > n_rows
    int b = 1;
    while (b < 32 && (1ll << b) <= n_rows) ++b;
    return b;
}

template <int ITEMS>
int launch_block_sort(const SortParams& sp, int n_jobs, cudaStream_t st) {
    using Sort = cub::BlockRadixSort<unsigned, BS_THREADS, ITEMS, int>;
    const size_t smem = sizeof(typename Sort::TempStorage);
    static bool attr_set = false;
    if (!attr_set) {
        if (smem > 48 * 1024)
            RB_CUDA(cudaFuncSetAttribute(block_sort_segments_kernel<ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    block_sort_segments_kernel<ITEMS><<<n_jobs, BS_THREADS, smem, st>>>(sp);
    RB_LAUNCH_CHECK("block_sort_segments_kernel");
    return RB200_OK;
}

// ---- large batches (> 16384 sample-rows): device-wide radix sort on 32-bit keys, then the same segment-sum kernel ---- //
__global__ void prep_keys32_kernel(const int64_t* __restrict__ ids, int n, long long n_rows, long long padding_idx,
                                   unsigned* __restrict__ keys, int* __restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long id = ids[i];
    keys[i] = (id != padding_idx && (unsigned long long)id < (unsigned long long)n_rows) ? (unsigned)id : KEY_SENTINEL;
    pos[i] = i;
}
// head flag of sorted position i (1 where a new id starts; sentinels and i == n are 0), computed on the fly for the scan
struct HeadFlag32 {
    const unsigned* keys; int n;
    __host__ __device__ int operator()(int i) const {
        if (i >= n) return 0;
        const unsigned k = keys[i];
        return (k != KEY_SENTINEL && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
    }
};
using HeadFlagIter = cub::TransformInputIterator<int, HeadFlag32, cub::CountingInputIterator<int>>;
// slots = exclusive scan of the head flags (n + 1 entries): position i is a head iff slots[i + 1] != slots[i]
__global__ void emit_heads32_kernel(const unsigned* __restrict__ keys, const int* __restrict__ pos, const int* __restrict__ slots, int n,
                                    int* __restrict__ seg_start, int* __restrict__ first_pos, int64_t* __restrict__ uniq_ids, int* __restrict__ n_uniq,
                                    int* __restrict__ row_slot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = keys[i];
    const int slot = slots[i];
    if (i == 0) n_uniq[0] = slots[n];
    if (k == KEY_SENTINEL) { if (i == 0) seg_start[0] = 0; return; }                // nothing valid at all
    if (i == n - 1 || keys[i + 1] == KEY_SENTINEL) seg_start[slots[i + 1]] = i + 1;   // end of the last segment (sentinels sort last)
    if (slots[i + 1] == slot) return;
    seg_start[slot] = i;
    first_pos[slot] = pos[i];
    uniq_ids[slot] = (int64_t)k;
    if (row_slot) row_slot[k] = slot;
}
size_t sort32_temp_bytes(int B, int bits) {
    size_t t = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr, (int*)nullptr, B, 0, bits);
    HeadFlagIter it(cub::CountingInputIterator<int>(0), HeadFlag32{nullptr, B});
    cub::DeviceScan::ExclusiveSum(nullptr, t2, it, (int*)nullptr, B + 1);
    return t > t2 ? t : t2;
}

int key_bits(long long n_rows) {
    int b = 1;
    while (b < 63 && (1ll << b) < n_rows) ++b;
    return b;
}

size_t sort_temp_bytes(int B, int bits) {
    size_t t = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t, (const int64_t*)nullptr, (int64_t*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, B, 0, bits);
    size_t t2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t2, (const int*)nullptr, (int*)nullptr, B + 1);
    return t > t2 ? t : t2;
}

// ---------------------------------------------------------------------------------------- //
// optimiser state
// ---------------------------------------------------------------------------------------- //
__global__ void opt_begin_step_kernel(rb200_opt_state* st) {
    if (threadIdx.x || blockIdx.x) return;
    rb_opt_begin_step_dev(st);
}

struct SumsqSegs { rb200_sumsq_seg s[4]; int n; };

// Block partials in fp64; the LAST block to finish (ticket in the optimizer state, self-resetting) adds them up in a
// fixed lane-strided order — deterministic whichever block that is — and, with do_clip, derives the clip coefficient.
__global__ void __launch_bounds__(NT) sumsq_kernel(const SumsqSegs segs, double* __restrict__ partials, rb200_opt_state* st,
                                                   int do_clip) {
    __shared__ double scratch[NT / 32];
    __shared__ bool is_last;
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * NT;
    for (int seg = 0; seg < segs.n; ++seg) {
        const float* x = segs.s[seg].x;
        long long n = segs.s[seg].n;
        if (segs.s[seg].count) n = (long long)segs.s[seg].count[0] * segs.s[seg].row_len;
        if (!x || n <= 0) continue;
        const long long n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? (n >> 2) : 0;
        for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += stride) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
            acc += (double)fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
        }
        for (long long i = n4 * 4 + (long long)blockIdx.x * NT + threadIdx.x; i < n; i += stride) {
            const float v = __ldg(x + i);
            acc += (double)(v * v);
        }
    }
    acc = rb_warp_sum_d(acc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) scratch[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < NT / 32; ++w) t += scratch[w];
        partials[blockIdx.x] = t;
        __threadfence();
        is_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last || threadIdx.x >= 32) return;
    __threadfence();
    double t = 0.0;
    for (int i = lane; i < (int)gridDim.x; i += 32) t += __ldcg(partials + i);
    t = rb_warp_sum_d(t);
    if (lane == 0) {
        st->ticket = 0u;
        t += st->sumsq;
        st->sumsq = t;
        if (do_clip) {
            const float total = (float)sqrt(t);
            st->total_norm = total;
            const float coef = st->max_norm / (total + 1e-6f);
            st->clip_coef = coef < 1.f ? coef : 1.f;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------ //
// Gradient finish of the fused training step, ONE launch instead of four:
//   blocks [0, nb_red[0])         reduction of the user tower's split-K weight-gradient partials   (reduce_partials_tc)
//   blocks [.., +nb_red[1])       … of the item towers'
//   blocks [.., +nb_seg)          deterministic segment sums of the embedding-row gradients (segment_sum2, 1 segment / warp)
// and, with do_sumsq, Σ g² of everything those blocks produce (fp64 block partials; the last block to finish adds them in
// index order and derives total_norm / clip_coef): clip_grad_norm_ needs no pass of its own.
// ------------------------------------------------------------------------------------------------------------ //
struct RedSet { const float* part; int nsplit; int P; int H; int Din; float* out; };
constexpr int FIN_LONG_SEG = 16, FIN_MAX_D = 256;      // grad_finish_kernel: segments above 16 rows are summed by the whole block
struct FinishParams { SegParams seg; RedSet red[2]; int nb_seg; int nb_red[2]; int do_sumsq; int do_clip; };

__global__ void __launch_bounds__(NT) grad_finish_kernel(const FinishParams p, double* __restrict__ partials) {
    __shared__ double scratch[NT / 32];
    __shared__ float sm[8][33];
    __shared__ __align__(16) float lg_part[NT / 32][FIN_MAX_D];
    __shared__ int lg_seg[NT / 32], lg_job[NT / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double sq = 0.0;
    // the partial-reduction blocks come FIRST in the grid: each is one long chain of dependent-latency loads, so they should
    // be resident from the start and overlap the many short segment blocks behind them
    const int nb_red_all = p.nb_red[0] + p.nb_red[1];
    if ((int)blockIdx.x >= nb_red_all) {
        int w = ((int)blockIdx.x - nb_red_all) * (NT / 32) + warp;
        int j = 0;
        if (p.seg.n_jobs > 1 && w >= p.seg.job[0].cap) { w -= p.seg.job[0].cap; j = 1; }
        const SegJob& J = p.seg.job[j];
        const int D4 = p.seg.D4;
        const bool valid = w < J.cap && w < J.n_uniq[0];
        const int beg = valid ? J.seg_start[w] : 0, end = valid ? J.seg_start[w + 1] : 0;
        // A popular id (Zipf-skewed users: one id can own > 100 rows of a batch) would make this warp walk a chain of that many
        // dependent loads while the rest of the grid has long finished.  Such segments are summed by the block's 8 warps together:
        // warp v adds rows beg+v, beg+v+8, … (ascending), the 8 partial rows are then added in warp order — a fixed split, so the
        // result is as deterministic as the single-warp sum.
        const bool is_long = valid && end - beg > FIN_LONG_SEG && D4 * 4 <= FIN_MAX_D;
        if (valid && !is_long) {
            const float4* __restrict__ rows = reinterpret_cast<const float4*>(J.rows);
            for (int c = lane; c < D4; c += 32) {
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                seg_accumulate(s, rows, J.pos, beg, end, D4, c);
                if (J.uniq_grads) reinterpret_cast<float4*>(J.uniq_grads)[(long long)w * D4 + c] = s;
                if (J.dense) {
                    float4* dp = reinterpret_cast<float4*>(J.dense) + J.uniq_ids[w] * D4 + c;
                    float4 o = *dp;
                    o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
                    *dp = o;
                }
                sq += (double)fmaf(s.x, s.x, fmaf(s.y, s.y, fmaf(s.z, s.z, s.w * s.w)));
            }
        }
        if (__syncthreads_or(is_long)) {                       // (one barrier is all a block without long segments pays)
        if (lane == 0) { lg_seg[warp] = is_long ? w : -1; lg_job[warp] = j; }
        __syncthreads();
        for (int i = 0; i < NT / 32; ++i) {
            const int sg = lg_seg[i];
            if (sg < 0) continue;                              // block-uniform
            const SegJob& L = p.seg.job[lg_job[i]];
            const int lb = L.seg_start[sg], le = L.seg_start[sg + 1];
            const float4* __restrict__ rows = reinterpret_cast<const float4*>(L.rows);
            for (int c = lane; c < D4; c += 32) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k = lb + warp; k < le; k += NT / 32) {
                    const float4 v = __ldg(rows + (long long)__ldg(L.pos + k) * D4 + c);
                    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
                }
                *reinterpret_cast<float4*>(&lg_part[warp][c * 4]) = a;
            }
            __syncthreads();
            for (int d = threadIdx.x; d < D4 * 4; d += NT) {
                float t = lg_part[0][d];
#pragma unroll
                for (int v = 1; v < NT / 32; ++v) t += lg_part[v][d];
                if (L.uniq_grads) L.uniq_grads[(long long)sg * (D4 * 4) + d] = t;
                if (L.dense) L.dense[L.uniq_ids[sg] * (long long)(D4 * 4) + d] += t;
                sq += (double)(t * t);
            }
            __syncthreads();
        }
        }
    } else {
        int b = (int)blockIdx.x, r = 0;
        if (b >= p.nb_red[0]) { b -= p.nb_red[0]; r = 1; }
        const RedSet& R = p.red[r];
        const int x = lane, y = warp;
        const int i = b * 32 + x;
        float a = 0.f;
        if (i < R.P) {
            // six independent loads in flight per thread; the additions keep the order k = y, y+8, y+16, …
            for (int k0 = y; k0 < R.nsplit; k0 += 48) {
                float v[6];
#pragma unroll
                for (int u = 0; u < 6; ++u) { const int k = k0 + 8 * u; v[u] = k < R.nsplit ? __ldg(R.part + (long long)k * R.P + i) : 0.f; }
#pragma unroll
                for (int u = 0; u < 6; ++u) a += v[u];
            }
        }
        sm[y][x] = a;
        __syncthreads();
        if (y == 0 && i < R.P) {
            float t = 0.f;
#pragma unroll
            for (int g = 0; g < 8; ++g) t += sm[g][x];
            int o = i;
            if (i < R.H * R.Din) { const int k = i / R.H, h = i - k * R.H; o = h * R.Din + k; }
            R.out[o] = t;
            sq = (double)(t * t);
        }
    }
    if (!p.do_sumsq) return;
    sq = rb_warp_sum_d(sq);
    if (lane == 0) scratch[warp] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < NT / 32; ++w) t += scratch[w];
        partials[blockIdx.x] = t;      // summed in index order by the consumer (the Adam launch, or norm_from_partials_kernel)
    }
}

// Σ of n fp64 block partials in a fixed order (thread t takes t, t+NT, …; then warps, then the block); valid on thread 0.
__device__ __forceinline__ double block_sum_partials(const double* __restrict__ partials, int n, double* scratch /*[NT/32]*/) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += NT) t += __ldcg(partials + i);
    t = rb_warp_sum_d(t);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = t;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < NT / 32; ++w) tot += scratch[w];
    return tot;
}
__device__ __forceinline__ float clip_from_sumsq(double sumsq, float max_norm, float* total_norm) {
    const float total = (float)sqrt(sumsq);
    *total_norm = total;
    const float coef = max_norm / (total + 1e-6f);
    return coef < 1.f ? coef : 1.f;
}
__global__ void __launch_bounds__(NT) norm_from_partials_kernel(const double* __restrict__ partials, int n, rb200_opt_state* st) {
    __shared__ double scratch[NT / 32];
    const double tot = block_sum_partials(partials, n, scratch);
    if (threadIdx.x == 0) {
        const double t = tot + st->sumsq;
        st->sumsq = t;
        float tn;
        st->clip_coef = clip_from_sumsq(t, st->max_norm, &tn);
        st->total_norm = tn;
    }
}

__global__ void grad_norm_clip_kernel(rb200_opt_state* st) {
    if (threadIdx.x || blockIdx.x) return;
    const float total = (float)sqrt(st->sumsq);
    st->total_norm = total;
    const float coef = st->max_norm / (total + 1e-6f);
    st->clip_coef = coef < 1.f ? coef : 1.f;
}

struct AdamK {
    float clip, wd, omb1, omb2, beta2, step_size, bc2s, eps;
};
__device__ __forceinline__ AdamK load_adam(const rb200_opt_state* st) {
    AdamK k;
    k.clip = st->clip_coef; k.wd = st->weight_decay; k.omb1 = st->one_minus_beta1; k.omb2 = st->one_minus_beta2;
    k.beta2 = st->beta2_f; k.step_size = st->step_size; k.bc2s = st->bias_corr2_sqrt; k.eps = st->eps;
    return k;
}
// torch.optim.Adam single-tensor update, same operation order
__device__ __forceinline__ void adam1(float& w, float g, float& m, float& v, const AdamK& k) {
    g = fmaf(k.wd, w, g * k.clip);                  // clip_grad_norm_ scaling, then grad.add(param, alpha=wd)
    m = fmaf(k.omb1, g - m, m);                     // exp_avg.lerp_(grad, 1-beta1)
    v = fmaf(k.omb2 * g, g, v * k.beta2);           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    const float denom = sqrtf(v) / k.bc2s + k.eps;
    w = w - k.step_size * (m / denom);              // param.addcdiv_(exp_avg, denom, value=-step_size)
}
__device__ __forceinline__ void adam4(float4& w, const float4& g, float4& m, float4& v, const AdamK& k) {
    adam1(w.x, g.x, m.x, v.x, k); adam1(w.y, g.y, m.y, v.y, k);
    adam1(w.z, g.z, m.z, v.z, k); adam1(w.w, g.w, m.w, v.w, k);
}

__global__ void __launch_bounds__(NT) adam_dense_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        const rb200_opt_state* __restrict__ st) {
    const AdamK k = load_adam(st);
    const long long stride = (long long)gridDim.x * NT;
    const bool al = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                      reinterpret_cast<uintptr_t>(g)) & 15) == 0;
    const long long n4 = al ? (n >> 2) : 0;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += stride) {
        float4 wv = reinterpret_cast<float4*>(w)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gv = g ? __ldg(reinterpret_cast<const float4*>(g) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        adam4(wv, gv, mv, vv, k);
        reinterpret_cast<float4*>(w)[i] = wv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * NT + threadIdx.x; i < n; i += stride) {
        float wv = w[i], mv = m[i], vv = v[i];
        adam1(wv, g ? __ldg(g + i) : 0.f, mv, vv, k);
        w[i] = wv; m[i] = mv; v[i] = vv;
    }
}

// up to two tensors per launch (the two MLP blocks, or the two embedding tables)
struct AdamDenseJob { float* w; const float* g; float* m; float* v; long long n; };
struct AdamDenseParams { AdamDenseJob job[2]; int n_jobs; };
__global__ void __launch_bounds__(NT) adam_dense2_kernel(const AdamDenseParams p, const rb200_opt_state* __restrict__ st) {
    const AdamK k = load_adam(st);
    const long long stride = (long long)gridDim.x * NT;
    for (int j = 0; j < p.n_jobs; ++j) {
        const AdamDenseJob& J = p.job[j];
        const long long n4 = J.n >> 2;         // callers guarantee 16-byte alignment and n % 4 == 0 handled by the tail loop
        for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += stride) {
            float4 wv = reinterpret_cast<float4*>(J.w)[i], mv = reinterpret_cast<float4*>(J.m)[i], vv = reinterpret_cast<float4*>(J.v)[i];
            const float4 gv = __ldg(reinterpret_cast<const float4*>(J.g) + i);
            adam4(wv, gv, mv, vv, k);
            reinterpret_cast<float4*>(J.w)[i] = wv; reinterpret_cast<float4*>(J.m)[i] = mv; reinterpret_cast<float4*>(J.v)[i] = vv;
        }
        for (long long i = n4 * 4 + (long long)blockIdx.x * NT + threadIdx.x; i < J.n; i += stride) {
            float wv = J.w[i], mv = J.m[i], vv = J.v[i];
            adam1(wv, __ldg(J.g + i), mv, vv, k);
            J.w[i] = wv; J.m[i] = mv; J.v[i] = vv;
        }
    }
}

struct AdamTableJob { float* w; float* m; float* v; long long n_rows; int* row_slot; const float* uniq_grads; };
struct AdamTableParams { AdamTableJob job[2]; int n_jobs; int D4; };

// The optimizer step of the fused training step in ONE launch: both embedding tables (dense, reference-exact: every
// row moves) and both MLP parameter blocks.  Requires D4 a power of two <= 32, so that the D4 lanes of a table row sit
// in one warp: the first of them reads the row's gradient slot, resets it to -1 for the next step and broadcasts it.
// `partials` != NULL: the gradient norm has not been finalised yet — every block adds the n_partials fp64 block partials of
// grad_finish_kernel in the same fixed order (no atomics, no extra launch) and derives the clip coefficient itself; block 0
// also records sumsq / total_norm / clip_coef in the optimizer state.
__global__ void __launch_bounds__(NT) adam_step_all_kernel(const AdamTableParams tp, const AdamDenseParams dp, rb200_opt_state* st,
                                                           const double* __restrict__ partials, int n_partials) {
    AdamK k = load_adam(st);
    if (partials) {
        __shared__ double scratch[NT / 32];
        __shared__ float s_clip;
        const double tot = block_sum_partials(partials, n_partials, scratch);
        if (threadIdx.x == 0) {
            float tn;
            s_clip = clip_from_sumsq(tot, st->max_norm, &tn);
            if (blockIdx.x == 0) { st->sumsq = tot; st->total_norm = tn; st->clip_coef = s_clip; }
        }
        __syncthreads();
        k.clip = s_clip;
    }
    const long long stride = (long long)gridDim.x * NT;
    const int D4 = tp.D4, lane = threadIdx.x & 31;
    for (int j = 0; j < tp.n_jobs; ++j) {
        const AdamTableJob& J = tp.job[j];
        const long long n4 = J.n_rows * D4;
        for (long long base = (long long)blockIdx.x * NT + (threadIdx.x & ~31); base < n4; base += stride) {   // warp-uniform
            const long long i = base + lane;
            const bool live = i < n4;
            const long long row = i / D4;
            const int c = (int)(i - row * D4);
            int slot = -1;
            if (live && c == 0) {
                slot = J.row_slot[row];
                if (slot >= 0) J.row_slot[row] = -1;
            }
            slot = __shfl_sync(0xffffffffu, slot, lane & ~(D4 - 1));
            if (!live) continue;
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (slot >= 0) gv = __ldg(reinterpret_cast<const float4*>(J.uniq_grads) + (long long)slot * D4 + c);
            float4 wv = reinterpret_cast<float4*>(J.w)[i], mv = reinterpret_cast<float4*>(J.m)[i], vv = reinterpret_cast<float4*>(J.v)[i];
            adam4(wv, gv, mv, vv, k);
            reinterpret_cast<float4*>(J.w)[i] = wv; reinterpret_cast<float4*>(J.m)[i] = mv; reinterpret_cast<float4*>(J.v)[i] = vv;
        }
    }
    for (int j = 0; j < dp.n_jobs; ++j) {
        const AdamDenseJob& J = dp.job[j];
        const long long n4 = J.n >> 2;
        for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += stride) {
            float4 wv = reinterpret_cast<float4*>(J.w)[i], mv = reinterpret_cast<float4*>(J.m)[i], vv = reinterpret_cast<float4*>(J.v)[i];
            const float4 gv = __ldg(reinterpret_cast<const float4*>(J.g) + i);
            adam4(wv, gv, mv, vv, k);
            reinterpret_cast<float4*>(J.w)[i] = wv; reinterpret_cast<float4*>(J.m)[i] = mv; reinterpret_cast<float4*>(J.v)[i] = vv;
        }
        for (long long i = n4 * 4 + (long long)blockIdx.x * NT + threadIdx.x; i < J.n; i += stride) {
            float wv = J.w[i], mv = J.m[i], vv = J.v[i];
            adam1(wv, __ldg(J.g + i), mv, vv, k);
            J.w[i] = wv; J.m[i] = mv; J.v[i] = vv;
        }
    }
}
// dense (reference-exact) table update for up to two tables in one launch
__global__ void __launch_bounds__(NT) adam_table_dense2_kernel(const AdamTableParams p, const rb200_opt_state* __restrict__ st) {
    const AdamK k = load_adam(st);
    const long long stride = (long long)gridDim.x * NT;
    const int D4 = p.D4;
    for (int j = 0; j < p.n_jobs; ++j) {
        const AdamTableJob& J = p.job[j];
        const long long n4 = J.n_rows * D4;
        for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += stride) {
            const long long row = i / D4;
            const int c = (int)(i - row * D4);
            const int slot = J.row_slot[row];
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (slot >= 0) gv = __ldg(reinterpret_cast<const float4*>(J.uniq_grads) + (long long)slot * D4 + c);
            float4 wv = reinterpret_cast<float4*>(J.w)[i], mv = reinterpret_cast<float4*>(J.m)[i], vv = reinterpret_cast<float4*>(J.v)[i];
            adam4(wv, gv, mv, vv, k);
            reinterpret_cast<float4*>(J.w)[i] = wv; reinterpret_cast<float4*>(J.m)[i] = mv; reinterpret_cast<float4*>(J.v)[i] = vv;
        }
    }
}

__global__ void __launch_bounds__(NT) adam_table_dense_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                                                              long long n_rows, int D4, const int* __restrict__ row_slot,
                                                              const float* __restrict__ uniq_grads,
                                                              const rb200_opt_state* __restrict__ st) {
    const AdamK k = load_adam(st);
    const long long n4 = n_rows * D4, stride = (long long)gridDim.x * NT;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += stride) {
        const long long row = i / D4;
        const int c = (int)(i - row * D4);
        const int slot = row_slot ? __ldg(row_slot + row) : -1;
        float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (slot >= 0) gv = __ldg(reinterpret_cast<const float4*>(uniq_grads) + (long long)slot * D4 + c);
        float4 wv = reinterpret_cast<float4*>(w)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        adam4(wv, gv, mv, vv, k);
        reinterpret_cast<float4*>(w)[i] = wv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
    }
}

__global__ void __launch_bounds__(NT) adam_rows_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, int D4,
                                                       const int64_t* __restrict__ uniq_ids, const float* __restrict__ uniq_grads,
                                                       const int* __restrict__ n_uniq, const rb200_opt_state* __restrict__ st) {
    const AdamK k = load_adam(st);
    const long long n4 = (long long)n_uniq[0] * D4, stride = (long long)gridDim.x * NT;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += stride) {
        const long long u = i / D4;
        const int c = (int)(i - u * D4);
        const long long o = __ldg(uniq_ids + u) * D4 + c;
        const float4 gv = __ldg(reinterpret_cast<const float4*>(uniq_grads) + i);
        float4 wv = reinterpret_cast<float4*>(w)[o], mv = reinterpret_cast<float4*>(m)[o], vv = reinterpret_cast<float4*>(v)[o];
        adam4(wv, gv, mv, vv, k);
        reinterpret_cast<float4*>(w)[o] = wv; reinterpret_cast<float4*>(m)[o] = mv; reinterpret_cast<float4*>(v)[o] = vv;
    }
}

int stream_grid(long long work_items) {
    long long g = (work_items + NT - 1) / NT;
    const long long cap = (long long)rb_sm_count() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace


// Scratch needed by the fast path for one table of n sample-rows.
static size_t fast_scratch_bytes(int n) { return 256 * 3 + sizeof(int) * (2 * (size_t)n + 2); }

// Internal: up to two tables.  Phase 1 (rb_scatter_sort) depends only on the ids, so the fused step runs it on a side
// stream concurrently with the towers; phase 2 (rb_scatter_sum) needs the row gradients.  ids of a table may be the
// concatenation of two arrays (positive ++ negative item ids).  Requires n ≤ 16384 and n_rows < 2^31 per table.
struct ScatterPlan { SortParams sp; SegParams gp; int max_n; };

static int scatter_plan(ScatterPlan& pl, int n_tables, const int64_t* const ids_a[2], const int64_t* const ids_b[2],
                        const int n_a[2], const int n_b[2], const float* const rows[2], int D, const long long n_rows[2],
                        long long padding_idx, float* const dense[2], int64_t* const uniq_ids[2], float* const uniq_grads[2],
                        int* const n_uniq[2], int* const row_slot[2], void* workspace, size_t workspace_bytes) {
    RbArena ar(workspace, workspace_bytes);
    pl.sp = SortParams{};
    pl.gp = SegParams{};
    pl.gp.n_jobs = n_tables; pl.gp.D4 = D / 4;
    pl.max_n = 0;
    for (int t = 0; t < n_tables; ++t) {
        const int n = n_a[t] + n_b[t];
        if (n > pl.max_n) pl.max_n = n;
        SortJob& J = pl.sp.job[t];
        J.ids_a = ids_a[t]; J.ids_b = ids_b[t]; J.n_a = n_a[t]; J.n_b = n_b[t];
        J.n_rows = n_rows[t]; J.padding_idx = padding_idx; J.key_bits = key_bits_strict(n_rows[t]);
        J.pos = ar.take<int>(n); J.seg_start = ar.take<int>((size_t)n + 1);
        J.uniq_ids = uniq_ids[t]; J.n_uniq = n_uniq[t]; J.row_slot = row_slot[t];
        SegJob& G = pl.gp.job[t];
        G.pos = J.pos; G.seg_start = J.seg_start; G.uniq_ids = J.uniq_ids; G.n_uniq = J.n_uniq;
        G.rows = rows[t]; G.uniq_grads = uniq_grads[t]; G.dense = dense[t]; G.cap = n; G.first_pos = nullptr;
    }
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "scatter: workspace too small (%zu given)", workspace_bytes);
    return RB200_OK;
}

static int scatter_sort(const ScatterPlan& pl, cudaStream_t st) {
    const int nt = pl.gp.n_jobs;
    if (pl.max_n <= BS_THREADS * 4) return launch_block_sort<4>(pl.sp, nt, st);
    if (pl.max_n <= BS_THREADS * 8) return launch_block_sort<8>(pl.sp, nt, st);
    return launch_block_sort<16>(pl.sp, nt, st);
}

static int scatter_sum(const ScatterPlan& pl, cudaStream_t st) {
    // small batches (the fused step): one segment per warp keeps every SM busy; see the large-batch path for 8 per warp
    long long warps = 0;
    for (int t = 0; t < pl.gp.n_jobs; ++t) warps += seg_warps<1>(pl.gp.job[t].cap);
    SegParams gp = pl.gp;
    gp.skip_long = gp.D4 * 4 <= LS_MAX_D;
    segment_sum2_kernel<1><<<(unsigned)((warps * 32 + NT - 1) / NT), NT, 0, st>>>(gp);
    RB_LAUNCH_CHECK("segment_sum2_kernel");
    if (gp.skip_long) return launch_long_segments(gp, st);
    return RB200_OK;
}

// sort_stream == sum_stream ⇒ plain sequential execution.  Otherwise the caller has already made sort_stream wait for
// the ids and must make sum_stream wait for the sort (see csrc/step.cu).
int rb_scatter_tables(int phase, int n_tables, const int64_t* const ids_a[2], const int64_t* const ids_b[2], const int n_a[2],
                      const int n_b[2], const float* const rows[2], int D, const long long n_rows[2], long long padding_idx,
                      float* const dense[2], int64_t* const uniq_ids[2], float* const uniq_grads[2], int* const n_uniq[2],
                      int* const row_slot[2], void* workspace, size_t workspace_bytes, cudaStream_t st) {
    ScatterPlan pl;
    int rc = scatter_plan(pl, n_tables, ids_a, ids_b, n_a, n_b, rows, D, n_rows, padding_idx, dense, uniq_ids, uniq_grads, n_uniq,
                          row_slot, workspace, workspace_bytes);
    if (rc) return rc;
    if (phase == 0 || phase == 1) { if ((rc = scatter_sort(pl, st))) return rc; }
    if (phase == 0 || phase == 2) { if ((rc = scatter_sum(pl, st))) return rc; }
    return RB200_OK;
}

// Phase 2 of rb_scatter_tables fused with the reduction of up to two towers' weight-gradient partials and (do_sumsq) the
// global gradient norm.  `red[t].nsplit == 0` → that gradient block is zero-filled.  `sumsq_ws` holds the fp64 block partials.
size_t rb_grad_finish_workspace_bytes(int n_seg_rows, const RbPartials red[2]) {
    size_t blocks = (size_t)(n_seg_rows + NT / 32 - 1) / (NT / 32);
    for (int t = 0; t < 2; ++t) blocks += (size_t)(red[t].P + 31) / 32;
    return 256 + sizeof(double) * blocks;
}

int rb_grad_finish(int n_tables, const int64_t* const ids_a[2], const int64_t* const ids_b[2], const int n_a[2], const int n_b[2],
                   const float* const rows[2], int D, const long long n_rows[2], long long padding_idx, float* const dense[2],
                   int64_t* const uniq_ids[2], float* const uniq_grads[2], int* const n_uniq[2], int* const row_slot[2],
                   void* scatter_ws, size_t scatter_ws_bytes, const RbPartials red[2], float* const red_out[2], int do_sumsq,
                   void* sumsq_ws, size_t sumsq_ws_bytes, const double** norm_partials, int* n_norm_partials, cudaStream_t s) {
    ScatterPlan pl;
    int rc = scatter_plan(pl, n_tables, ids_a, ids_b, n_a, n_b, rows, D, n_rows, padding_idx, dense, uniq_ids, uniq_grads, n_uniq,
                          row_slot, scatter_ws, scatter_ws_bytes);
    if (rc) return rc;
    FinishParams fp{};
    fp.seg = pl.gp;
    long long warps = 0;
    for (int t = 0; t < pl.gp.n_jobs; ++t) warps += pl.gp.job[t].cap;
    fp.nb_seg = (int)((warps + NT / 32 - 1) / (NT / 32));
    for (int t = 0; t < 2; ++t) {
        if (red[t].nsplit == 0 && red_out[t]) RB_CUDA(cudaMemsetAsync(red_out[t], 0, sizeof(float) * red[t].P, s));
        fp.red[t] = RedSet{red[t].part, red[t].nsplit, red[t].P, red[t].H, red[t].Din, red_out[t]};
        fp.nb_red[t] = (red_out[t] && red[t].nsplit > 0) ? (red[t].P + 31) / 32 : 0;
    }
    fp.do_sumsq = do_sumsq; fp.do_clip = do_sumsq;
    const int grid = fp.nb_seg + fp.nb_red[0] + fp.nb_red[1];
    RbArena ar(sumsq_ws, sumsq_ws_bytes);
    double* partials = ar.take<double>(grid);
    if (do_sumsq && (!sumsq_ws || !ar.ok())) return rb_set_error(RB200_ERR_WORKSPACE, "grad_finish: workspace too small");
    RB_REQUIRE(grid >= 1, "grad_finish: nothing to do");
    grad_finish_kernel<<<grid, NT, 0, s>>>(fp, partials);
    RB_LAUNCH_CHECK("grad_finish_kernel");
    if (norm_partials) { *norm_partials = do_sumsq ? partials : nullptr; *n_norm_partials = do_sumsq ? grid : 0; }
    return RB200_OK;
}

// Σg² block partials of rb_grad_finish → sumsq / total_norm / clip_coef (for consumers other than the fused Adam launch)
int rb_norm_from_partials(const double* partials, int n, rb200_opt_state* st, cudaStream_t s) {
    norm_from_partials_kernel<<<1, NT, 0, s>>>(partials, n, st);
    RB_LAUNCH_CHECK("norm_from_partials_kernel");
    return RB200_OK;
}

extern "C" size_t rb200_scatter_workspace_bytes(int B, int64_t n_rows) {
    if (B < 1) B = 1;
    // (fast path additionally needs an int64 scratch for unique ids and an int for the count when the caller
    //  asks only for the dense output)
    return 256 * 16 + sizeof(int64_t) * (size_t)2 * B + sizeof(int) * ((size_t)8 * B + 3 * ((size_t)B + 1) + 4) +
           fast_scratch_bytes(B) + sort_temp_bytes(B, key_bits(n_rows)) + sort32_temp_bytes(B, 32);
}

extern "C" int rb200_scatter_rows(const int64_t* ids, const float* rows, int B, int D, int64_t n_rows, int64_t padding_idx,
                                  float* dense_grad, int64_t* uniq_ids, float* uniq_grads, int* n_uniq, int* row_slot,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    RB_REQUIRE(ids && rows && B >= 0 && D >= 4 && (D % 4) == 0 && n_rows >= 1, "scatter_rows: bad arguments");
    RB_REQUIRE((uniq_grads == nullptr) == (uniq_ids == nullptr), "scatter_rows: uniq_ids and uniq_grads go together");
    RB_REQUIRE(row_slot == nullptr || uniq_grads != nullptr, "scatter_rows: row_slot needs the compact outputs");
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) {
        if (n_uniq) RB_CUDA(cudaMemsetAsync(n_uniq, 0, sizeof(int), st));
        return RB200_OK;
    }
    if (B <= BS_THREADS * 16 && n_rows < (1ll << 31)) {
        // single-CTA sort fast path
        RbArena fa(workspace, workspace_bytes);
        int64_t* u_ids = uniq_ids ? uniq_ids : fa.take<int64_t>(B);
        int* n_u = n_uniq ? n_uniq : fa.take<int>(1);
        char* rest = fa.take<char>(fast_scratch_bytes(B));
        if (!workspace || !fa.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "scatter_rows: workspace too small (%zu given)", workspace_bytes);
        const int64_t* ia[2] = {ids, nullptr}; const int64_t* ib[2] = {nullptr, nullptr};
        const int na[2] = {B, 0}, nb[2] = {0, 0};
        const float* rw[2] = {rows, nullptr}; const long long nr[2] = {n_rows, 0};
        float* dn[2] = {dense_grad, nullptr}; int64_t* ui[2] = {u_ids, nullptr}; float* ug[2] = {uniq_grads, nullptr};
        int* nu[2] = {n_u, nullptr}; int* rs[2] = {row_slot, nullptr};
        return rb_scatter_tables(0, 1, ia, ib, na, nb, rw, D, nr, padding_idx, dn, ui, ug, nu, rs, rest, fast_scratch_bytes(B), st);
    }
    if (n_rows < (1ll << 31)) {
        const int bits32 = key_bits_strict(n_rows);
        RbArena a2(workspace, workspace_bytes);
        int64_t* u_ids = uniq_ids ? uniq_ids : a2.take<int64_t>(B);
        int* n_u = n_uniq ? n_uniq : a2.take<int>(1);
        unsigned* k_in = a2.take<unsigned>(B); unsigned* k_out = a2.take<unsigned>(B);
        int* p_in = a2.take<int>(B); int* p_out = a2.take<int>(B);
        int* slots = a2.take<int>((size_t)B + 1);
        int* seg_start = a2.take<int>((size_t)B + 1);
        int* first_pos = a2.take<int>(B);
        const size_t tbytes = sort32_temp_bytes(B, bits32);
        char* temp = a2.take<char>(tbytes);
        if (!workspace || !a2.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "scatter_rows: workspace too small (%zu given)", workspace_bytes);
        prep_keys32_kernel<<<(B + NT - 1) / NT, NT, 0, st>>>(ids, B, n_rows, padding_idx, k_in, p_in);
        RB_LAUNCH_CHECK("prep_keys32_kernel");
        size_t tb = tbytes;
        RB_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, (const unsigned*)k_in, k_out, (const int*)p_in, p_out, B, 0, bits32, st));
        tb = tbytes;
        HeadFlagIter heads(cub::CountingInputIterator<int>(0), HeadFlag32{k_out, B});
        RB_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, heads, slots, B + 1, st));
        emit_heads32_kernel<<<(B + NT - 1) / NT, NT, 0, st>>>(k_out, p_out, slots, B, seg_start, first_pos, u_ids, n_u, row_slot);
        RB_LAUNCH_CHECK("emit_heads32_kernel");
        SegParams gp{};
        gp.n_jobs = 1; gp.D4 = D / 4;
        gp.job[0] = SegJob{p_out, seg_start, u_ids, n_u, rows, uniq_grads, dense_grad, B, first_pos};
        gp.skip_long = D <= LS_MAX_D;
        gp.long_count = reinterpret_cast<int*>(k_in);            // the unsorted keys are dead: their first word counts the long segments
        if (gp.skip_long) RB_CUDA(cudaMemsetAsync(gp.long_count, 0, sizeof(int), st));
        segment_sum2_kernel<8><<<(unsigned)(((long long)seg_warps<8>(B) * 32 + NT - 1) / NT), NT, 0, st>>>(gp);
        RB_LAUNCH_CHECK("segment_sum2_kernel");
        if (gp.skip_long) return launch_long_segments(gp, st);
        return RB200_OK;
    }
    const int bits = key_bits(n_rows);
    RbArena ar(workspace, workspace_bytes);
    int64_t* keys = ar.take<int64_t>(B);
    int* pos_in = ar.take<int>(B);
    int* pos = ar.take<int>(B);
    int* flags = ar.take<int>((size_t)B + 1);
    int* slots = ar.take<int>((size_t)B + 1);
    size_t temp_bytes = sort_temp_bytes(B, bits);
    char* temp = ar.take<char>(temp_bytes);
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "scatter_rows: workspace too small (%zu given)", workspace_bytes);
    iota_kernel<<<(B + NT - 1) / NT, NT, 0, st>>>(pos_in, B);
    RB_LAUNCH_CHECK("iota_kernel");
    size_t tb = temp_bytes;
    RB_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, ids, keys, (const int*)pos_in, pos, B, 0, bits, st));
    head_flags_kernel<<<(B + 1 + NT - 1) / NT, NT, 0, st>>>(keys, B, padding_idx, flags);
    RB_LAUNCH_CHECK("head_flags_kernel");
    tb = temp_bytes;
    RB_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, (const int*)flags, slots, B + 1, st));
    segment_sum_kernel<<<(int)(((long long)B * 32 + NT - 1) / NT), NT, 0, st>>>(keys, pos, flags, slots, B, D, n_rows, rows,
                                                                              dense_grad, uniq_ids, uniq_grads, n_uniq, row_slot);
    RB_LAUNCH_CHECK("segment_sum_kernel");
    return RB200_OK;
}

// Two-phase form of rb200_scatter_rows (same result): the PLAN depends only on the ids — sort, segment heads, compact id list — and
// can run on a side stream before the gradient rows exist (the row-sharded step sorts its received row list under the towers); APPLY
// adds the rows up.  Both carve the same workspace (rb200_scatter_workspace_bytes); it must stay untouched in between.
namespace {
struct ScatterPlanWs { unsigned* k_in; unsigned* k_out; int* p_in; int* p_out; int* slots; int* seg_start; int* first_pos; char* temp; size_t tbytes; };
bool carve_scatter_plan(void* workspace, size_t workspace_bytes, int B, int64_t n_rows, ScatterPlanWs& w) {
    RbArena a(workspace, workspace_bytes);
    w.k_in = a.take<unsigned>(B); w.k_out = a.take<unsigned>(B);
    w.p_in = a.take<int>(B); w.p_out = a.take<int>(B);
    w.slots = a.take<int>((size_t)B + 1);
    w.seg_start = a.take<int>((size_t)B + 1);
    w.first_pos = a.take<int>(B);
    w.tbytes = sort32_temp_bytes(B, key_bits_strict(n_rows));
    w.temp = a.take<char>(w.tbytes);
    return workspace && a.ok();
}
}  // namespace

extern "C" int rb200_scatter_plan(const int64_t* ids, int B, int64_t n_rows, int64_t padding_idx, int64_t* uniq_ids, int* n_uniq,
                                  int* row_slot, void* workspace, size_t workspace_bytes, void* stream) {
    RB_REQUIRE(ids && uniq_ids && n_uniq && B >= 1 && n_rows >= 1 && n_rows < (1ll << 31), "scatter_plan: bad arguments (1 <= n_rows < 2^31)");
    cudaStream_t st = (cudaStream_t)stream;
    ScatterPlanWs w;
    if (!carve_scatter_plan(workspace, workspace_bytes, B, n_rows, w))
        return rb_set_error(RB200_ERR_WORKSPACE, "scatter_plan: workspace too small (%zu given)", workspace_bytes);
    prep_keys32_kernel<<<(B + NT - 1) / NT, NT, 0, st>>>(ids, B, n_rows, padding_idx, w.k_in, w.p_in);
    RB_LAUNCH_CHECK("prep_keys32_kernel");
    size_t tb = w.tbytes;
    RB_CUDA(cub::DeviceRadixSort::SortPairs(w.temp, tb, (const unsigned*)w.k_in, w.k_out, (const int*)w.p_in, w.p_out, B, 0,
                                            key_bits_strict(n_rows), st));
    tb = w.tbytes;
    HeadFlagIter heads(cub::CountingInputIterator<int>(0), HeadFlag32{w.k_out, B});
    RB_CUDA(cub::DeviceScan::ExclusiveSum(w.temp, tb, heads, w.slots, B + 1, st));
    emit_heads32_kernel<<<(B + NT - 1) / NT, NT, 0, st>>>(w.k_out, w.p_out, w.slots, B, w.seg_start, w.first_pos, uniq_ids, n_uniq, row_slot);
    RB_LAUNCH_CHECK("emit_heads32_kernel");
    RB_CUDA(cudaMemsetAsync(w.k_in, 0, sizeof(int), st));      // the unsorted keys are dead: their first word counts the long segments
    return RB200_OK;
}

extern "C" int rb200_scatter_apply(const float* rows, int B, int D, int64_t n_rows, float* dense_grad, const int64_t* uniq_ids,
                                   float* uniq_grads, const int* n_uniq, void* workspace, size_t workspace_bytes, void* stream) {
    RB_REQUIRE(rows && uniq_ids && n_uniq && B >= 1 && D >= 4 && (D % 4) == 0 && n_rows >= 1 && n_rows < (1ll << 31) && (dense_grad || uniq_grads),
               "scatter_apply: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    ScatterPlanWs w;
    if (!carve_scatter_plan(workspace, workspace_bytes, B, n_rows, w))
        return rb_set_error(RB200_ERR_WORKSPACE, "scatter_apply: workspace too small (%zu given)", workspace_bytes);
    SegParams gp{};
    gp.n_jobs = 1; gp.D4 = D / 4;
    gp.job[0] = SegJob{w.p_out, w.seg_start, const_cast<int64_t*>(uniq_ids), const_cast<int*>(n_uniq), rows, uniq_grads, dense_grad, B, w.first_pos};
    gp.skip_long = D <= LS_MAX_D;
    gp.long_count = reinterpret_cast<int*>(w.k_in);
    segment_sum2_kernel<8><<<(unsigned)(((long long)seg_warps<8>(B) * 32 + NT - 1) / NT), NT, 0, st>>>(gp);
    RB_LAUNCH_CHECK("segment_sum2_kernel");
    if (gp.skip_long) return launch_long_segments(gp, st);
    return RB200_OK;
}

extern "C" int rb200_scatter_reset_slots(const int64_t* uniq_ids, const int* n_uniq, int max_uniq, int* row_slot, void* stream) {
    RB_REQUIRE(uniq_ids && n_uniq && row_slot, "scatter_reset_slots: NULL pointer");
    if (max_uniq <= 0) return RB200_OK;
    reset_slots_kernel<<<stream_grid(max_uniq), NT, 0, (cudaStream_t)stream>>>(uniq_ids, n_uniq, row_slot);
    RB_LAUNCH_CHECK("reset_slots_kernel");
    return RB200_OK;
}

extern "C" int rb200_scatter_set_slots(const int64_t* uniq_ids, const int* n_uniq, int max_uniq, int* row_slot, void* stream) {
    RB_REQUIRE(uniq_ids && n_uniq && row_slot, "scatter_set_slots: NULL pointer");
    if (max_uniq <= 0) return RB200_OK;
    set_slots_kernel<<<stream_grid(max_uniq), NT, 0, (cudaStream_t)stream>>>(uniq_ids, n_uniq, row_slot);
    RB_LAUNCH_CHECK("set_slots_kernel");
    return RB200_OK;
}

extern "C" int rb200_gather_rows(const float* table, const int64_t* rows, int64_t n, int D, int64_t n_table_rows, float* out,
                                 void* stream) {
    RB_REQUIRE(table && rows && out && n >= 0 && D >= 4 && D % 4 == 0, "gather_rows: bad arguments");
    if (n == 0) return RB200_OK;
    gather_rows_kernel<<<stream_grid(n * (D / 4)), NT, 0, (cudaStream_t)stream>>>(table, rows, n, D / 4, n_table_rows, out);
    RB_LAUNCH_CHECK("gather_rows_kernel");
    return RB200_OK;
}

// ---------------------------------------------------------------------------------------- //
// Peer-memory exchange of the row-sharded step (one NVLink / NVSwitch box, tables mapped into every process with
// torch symmetric memory): the requester READS the rows it needs straight out of the owners' shards and WRITES its row
// gradients straight into the owners' receive buckets — the gather and its all-to-all, and the all-to-all of the gradients,
// are one kernel each; no NCCL on the row path.
// ---------------------------------------------------------------------------------------- //
struct PeerTables { const float* table[RB200_MAX_PEERS]; long long user_rows[RB200_MAX_PEERS]; };
struct PeerBuckets { float* grads[RB200_MAX_PEERS]; int64_t* rows[RB200_MAX_PEERS]; };

// out[r] = shard[owner(id_r)][local_row(id_r)], requests in sample order [user ids | item ids].  One warp per request, four requests
// in flight per warp (a peer read over NVLink takes ≈ 2 us); the per-request arithmetic (owner = id mod world, local row = id / world)
// is done once per request, not once per 16 bytes.
__global__ void __launch_bounds__(NT) gather_rows_sharded_kernel(const PeerTables P, int world, const int64_t* __restrict__ user_ids,
                                                                 long long n_u, const int64_t* __restrict__ item_ids, long long n_i,
                                                                 long long n_user_rows, long long n_item_rows, int D4,
                                                                 float* __restrict__ out, int* __restrict__ err_flag) {
    const int lane = threadIdx.x & 31;
    const long long n = n_u + n_i, n_warps = (long long)gridDim.x * (NT / 32);
    for (long long r0 = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5); r0 < n; r0 += 4 * n_warps) {
        const float4* src[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long r = r0 + u * n_warps;
            src[u] = nullptr;
            if (r < n) {
                const bool is_user = r < n_u;
                long long id = is_user ? user_ids[r] : item_ids[r - n_u];
                if ((unsigned long long)id >= (unsigned long long)(is_user ? n_user_rows : n_item_rows)) {
                    if (err_flag && lane == 0) atomicOr(err_flag, 1);
                    id = 0;
                }
                const int owner = (int)(id % world);
                const long long local = id / world + (is_user ? 0 : P.user_rows[owner]);
                src[u] = reinterpret_cast<const float4*>(P.table[owner]) + local * D4;
            }
        }
        for (int c = lane; c < D4; c += 32) {
            float4 v[4];
            // (plain loads, not the read-only path: the source may be peer memory that other GPUs rewrite between steps)
#pragma unroll
            for (int u = 0; u < 4; ++u) if (src[u]) v[u] = src[u][c];
#pragma unroll
            for (int u = 0; u < 4; ++u) if (src[u]) reinterpret_cast<float4*>(out)[(r0 + u * n_warps) * D4 + c] = v[u];
        }
    }
}

// gradient rows (sample order) → bucket (this rank) of their owner's receive buffer, at the slot the exchange plan assigned;
// the plan's row list (owner-local row per slot, -1 = empty) goes with them unless it went ahead (send_rows NULL).  Requests that
// overflowed their bucket (slot == world·C) are dropped (counted by the plan).  One warp per row, four rows in flight per warp.
__global__ void __launch_bounds__(NT) push_rows_sharded_kernel(const PeerBuckets P, int world, int rank, long long C,
                                                               const float* __restrict__ drows, const int64_t* __restrict__ slot_of_sample,
                                                               long long n, int D4, const int64_t* __restrict__ send_rows) {
    const int lane = threadIdx.x & 31;
    const long long n_warps = (long long)gridDim.x * (NT / 32), stride = (long long)gridDim.x * NT;
    for (long long r0 = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5); r0 < n; r0 += 4 * n_warps) {
        float4* dst[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long r = r0 + u * n_warps;
            dst[u] = nullptr;
            if (r < n) {
                const long long slot = slot_of_sample[r];
                if (slot < (long long)world * C) {
                    const int owner = (int)(slot / C);
                    const long long off = slot - (long long)owner * C;
                    dst[u] = reinterpret_cast<float4*>(P.grads[owner]) + ((long long)rank * C + off) * D4;
                }
            }
        }
        for (int c = lane; c < D4; c += 32) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) if (dst[u]) v[u] = __ldg(reinterpret_cast<const float4*>(drows) + (r0 + u * n_warps) * D4 + c);
#pragma unroll
            for (int u = 0; u < 4; ++u) if (dst[u]) dst[u][c] = v[u];
        }
    }
    if (send_rows)
        for (long long j = (long long)blockIdx.x * NT + threadIdx.x; j < (long long)world * C; j += stride) {
            const int owner = (int)(j / C);
            const long long off = j - (long long)owner * C;
            P.rows[owner][(long long)rank * C + off] = send_rows[j];
        }
    // No fence here: the cross-GPU barrier that follows on the stream is a later kernel (its release at system scope is ordered after
    // this kernel's writes by the kernel boundary, and the peers' acquire makes them visible).  A __threadfence_system() per thread
    // cost this kernel most of its time (12.6 MB in 20 us at world 1).
}
// the plan's row list alone (it depends only on the ids: sent at the start of the step, so that the owners sort their received rows
// under the towers)
__global__ void __launch_bounds__(NT) push_row_lists_kernel(const PeerBuckets P, int world, int rank, long long C,
                                                            const int64_t* __restrict__ send_rows) {
    const long long stride = (long long)gridDim.x * NT;
    for (long long j = (long long)blockIdx.x * NT + threadIdx.x; j < (long long)world * C; j += stride) {
        const int owner = (int)(j / C);
        const long long off = j - (long long)owner * C;
        P.rows[owner][(long long)rank * C + off] = send_rows[j];
    }
    __threadfence_system();
}

// out[i] = Σ_k src[k][i], k ascending — every rank reads every rank's buffer (peer memory) and adds in the SAME order, so all
// ranks hold bit-identical sums without a broadcast ("one-shot" all-reduce; the MLP gradients are 273 KB at C4 widths)
struct PeerSrc { const float* p[RB200_MAX_PEERS]; };
__global__ void __launch_bounds__(NT) allreduce_oneshot_kernel(const PeerSrc P, int world, long long n4, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * NT;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += stride) {
        float4 a = *(reinterpret_cast<const float4*>(P.p[0]) + i);
        for (int k = 1; k < world; ++k) {
            const float4 v = *(reinterpret_cast<const float4*>(P.p[k]) + i);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        reinterpret_cast<float4*>(out)[i] = a;
    }
}
// Two-shot all-reduce IN PLACE over peer memory (the data-parallel step's dense gradients: 2.7 MB at the reference's sizes): rank r owns
// slice r of the buffer — it reads that slice from every rank (ascending rank order: every element is summed by exactly one rank, in
// the same order whatever the rank count, so all ranks end up with bit-identical sums), and writes the sum back into slice r of EVERY
// rank's buffer.  No other rank reads or writes slice r, so the reduction is in place; the caller puts a cross-GPU barrier before
// (all buffers complete) and after (all slices delivered).  Per rank (W − 1)/W · n floats cross NVLink in each direction.
struct PeerRW { float* p[RB200_MAX_PEERS]; };
__global__ void __launch_bounds__(NT) allreduce_twoshot_kernel(const PeerRW P, int world, int rank, long long n4) {
    const long long per = (n4 + world - 1) / world, beg = per * rank, end = beg + per < n4 ? beg + per : n4;
    const long long stride = (long long)gridDim.x * NT;
    for (long long i0 = beg + (long long)blockIdx.x * NT + threadIdx.x; i0 < end; i0 += 2 * stride) {
        const long long i1 = i0 + stride;
        const bool two = i1 < end;
        float4 a = *(reinterpret_cast<const float4*>(P.p[0]) + i0), b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (two) b = *(reinterpret_cast<const float4*>(P.p[0]) + i1);
        for (int k = 1; k < world; ++k) {
            const float4 v = *(reinterpret_cast<const float4*>(P.p[k]) + i0);
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (two) w = *(reinterpret_cast<const float4*>(P.p[k]) + i1);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            b.x += w.x; b.y += w.y; b.z += w.z; b.w += w.w;
        }
        for (int k = 0; k < world; ++k) {
            reinterpret_cast<float4*>(P.p[k])[i0] = a;
            if (two) reinterpret_cast<float4*>(P.p[k])[i1] = b;
        }
    }
}

// this rank's {Σg² of its table-shard gradients (fp64 as a hi/lo float pair), loss·scale} → its 4-float slot (peer-readable)
__global__ void scalars_publish_kernel(const rb200_opt_state* st, const float* loss, float scale, float* slot) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double s = st->sumsq;
        const float hi = (float)s;
        slot[0] = hi; slot[1] = (float)(s - (double)hi); slot[2] = loss[0] * scale; slot[3] = 0.f;
        __threadfence_system();
    }
}
__global__ void scalars_reduce_kernel(const PeerSrc P, int world, rb200_opt_state* st) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        float l = 0.f;
        for (int k = 0; k < world; ++k) {
            const volatile float* q = P.p[k];
            s += (double)q[0] + (double)q[1];
            l += q[2];
        }
        st->sumsq = s;
        st->loss = l;
    }
}

// The tail of the sharded step's norm in ONE launch (one block): Σ over the ranks' published {Σg², loss} slots (rank order), plus the
// sum of squares of the replicated MLP gradient (n floats; fp64, fixed order: per-thread strided, shuffle tree, warp order), then
// total_norm and clip_coef — scalars_reduce + sumsq_accumulate + grad_norm_clip before.
constexpr int SF_NT = 1024;
__global__ void __launch_bounds__(SF_NT) scalars_finish_kernel(const PeerSrc P, int world, const float* __restrict__ g, long long n,
                                                               rb200_opt_state* st) {
    __shared__ double wpart[SF_NT / 32];
    double a = 0.0;
    for (long long i = threadIdx.x; i < n; i += SF_NT) { const double x = (double)g[i]; a += x * x; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) wpart[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        float l = 0.f;
        for (int k = 0; k < world; ++k) {
            const volatile float* q = P.p[k];
            s += (double)q[0] + (double)q[1];
            l += q[2];
        }
        double mlp = 0.0;
        for (int w = 0; w < SF_NT / 32; ++w) mlp += wpart[w];
        s += mlp;
        st->sumsq = s;
        st->loss = l;
        float tn;
        st->clip_coef = clip_from_sumsq(s, st->max_norm, &tn);
        st->total_norm = tn;
    }
}

extern "C" int rb200_sharded_scalars_finish(const void* const* slot_ptrs, int world, const float* dense_grad, int64_t n, rb200_opt_state* st,
                                            void* stream) {
    RB_REQUIRE(slot_ptrs && st && world >= 1 && world <= RB200_MAX_PEERS && n >= 0 && (n == 0 || dense_grad), "sharded_scalars_finish: 1..%d ranks",
               RB200_MAX_PEERS);
    PeerSrc P{};
    for (int k = 0; k < world; ++k) { RB_REQUIRE(slot_ptrs[k], "sharded_scalars_finish: NULL pointer of rank %d", k); P.p[k] = (const float*)slot_ptrs[k]; }
    scalars_finish_kernel<<<1, SF_NT, 0, (cudaStream_t)stream>>>(P, world, dense_grad, (long long)n, st);
    RB_LAUNCH_CHECK("scalars_finish_kernel");
    return RB200_OK;
}

extern "C" int rb200_allreduce_oneshot(const void* const* src_ptrs, int world, int64_t n, float* out, void* stream) {
    RB_REQUIRE(src_ptrs && out && world >= 1 && world <= RB200_MAX_PEERS && n >= 4 && n % 4 == 0, "allreduce_oneshot: n must be a multiple of 4, 1..%d ranks", RB200_MAX_PEERS);
    PeerSrc P{};
    for (int k = 0; k < world; ++k) { RB_REQUIRE(src_ptrs[k], "allreduce_oneshot: NULL pointer of rank %d", k); P.p[k] = (const float*)src_ptrs[k]; }
    allreduce_oneshot_kernel<<<stream_grid(n / 4), NT, 0, (cudaStream_t)stream>>>(P, world, n / 4, out);
    RB_LAUNCH_CHECK("allreduce_oneshot_kernel");
    return RB200_OK;
}

// The same all-reduce through the NVSwitch's multicast object (NVLS): rank r reads slice r ONCE through the multicast address with the
// reduction done inside the switch (multimem.ld_reduce) and stores the sum once through the same address, which the switch delivers to
// every rank (multimem.st) — per rank n/W floats each way instead of (W − 1)/W · n.  One reduction per element, broadcast to all: every
// rank holds the same bits.  Same barriers around it as the two-shot form.
__global__ void __launch_bounds__(NT) allreduce_multimem_kernel(float* __restrict__ mc, int world, int rank, long long n4) {
    const long long per = (n4 + world - 1) / world, beg = per * rank, end = beg + per < n4 ? beg + per : n4;
    const long long stride = (long long)gridDim.x * NT;
    for (long long i = beg + (long long)blockIdx.x * NT + threadIdx.x; i < end; i += stride) {
        float4 v;
        float* p = mc + i * 4;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}

extern "C" int rb200_allreduce_multimem(void* multicast_ptr, int world, int rank, int64_t n, void* stream) {
    RB_REQUIRE(multicast_ptr && world >= 1 && world <= RB200_MAX_PEERS && rank >= 0 && rank < world && n >= 4 && n % 4 == 0,
               "allreduce_multimem: n must be a multiple of 4, 1..%d ranks, a multicast address", RB200_MAX_PEERS);
    const long long per = (n / 4 + world - 1) / world;
    allreduce_multimem_kernel<<<stream_grid(per), NT, 0, (cudaStream_t)stream>>>((float*)multicast_ptr, world, rank, n / 4);
    RB_LAUNCH_CHECK("allreduce_multimem_kernel");
    return RB200_OK;
}

extern "C" int rb200_allreduce_twoshot(void* const* buf_ptrs, int world, int rank, int64_t n, void* stream) {
    RB_REQUIRE(buf_ptrs && world >= 1 && world <= RB200_MAX_PEERS && rank >= 0 && rank < world && n >= 4 && n % 4 == 0,
               "allreduce_twoshot: n must be a multiple of 4, 1..%d ranks", RB200_MAX_PEERS);
    PeerRW P{};
    for (int k = 0; k < world; ++k) { RB_REQUIRE(buf_ptrs[k], "allreduce_twoshot: NULL pointer of rank %d", k); P.p[k] = (float*)buf_ptrs[k]; }
    const long long per = (n / 4 + world - 1) / world;
    allreduce_twoshot_kernel<<<stream_grid((per + 1) / 2), NT, 0, (cudaStream_t)stream>>>(P, world, rank, n / 4);
    RB_LAUNCH_CHECK("allreduce_twoshot_kernel");
    return RB200_OK;
}

extern "C" int rb200_sharded_scalars_publish(const rb200_opt_state* st, const float* loss, float scale, float* slot, void* stream) {
    RB_REQUIRE(st && loss && slot, "sharded_scalars_publish: NULL pointer");
    scalars_publish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(st, loss, scale, slot);
    RB_LAUNCH_CHECK("scalars_publish_kernel");
    return RB200_OK;
}

extern "C" int rb200_sharded_scalars_reduce(const void* const* slot_ptrs, int world, rb200_opt_state* st, void* stream) {
    RB_REQUIRE(slot_ptrs && st && world >= 1 && world <= RB200_MAX_PEERS, "sharded_scalars_reduce: 1..%d ranks", RB200_MAX_PEERS);
    PeerSrc P{};
    for (int k = 0; k < world; ++k) { RB_REQUIRE(slot_ptrs[k], "sharded_scalars_reduce: NULL pointer of rank %d", k); P.p[k] = (const float*)slot_ptrs[k]; }
    scalars_reduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, world, st);
    RB_LAUNCH_CHECK("scalars_reduce_kernel");
    return RB200_OK;
}

extern "C" int rb200_gather_rows_sharded(const void* const* shard_ptrs, const int64_t* user_rows_by_rank, int world,
                                         const int64_t* user_ids, int64_t n_user, const int64_t* item_ids, int64_t n_item,
                                         int64_t n_user_rows, int64_t n_item_rows, int D, float* out, int* err_flag, void* stream) {
    RB_REQUIRE(shard_ptrs && user_rows_by_rank && out && world >= 1 && world <= RB200_MAX_PEERS, "gather_rows_sharded: 1..%d ranks", RB200_MAX_PEERS);
    RB_REQUIRE(n_user >= 0 && n_item >= 0 && (n_user == 0 || user_ids) && (n_item == 0 || item_ids) && D >= 4 && D % 4 == 0,
               "gather_rows_sharded: bad arguments");
    if (n_user + n_item == 0) return RB200_OK;
    PeerTables P{};
    for (int k = 0; k < world; ++k) {
        RB_REQUIRE(shard_ptrs[k], "gather_rows_sharded: NULL shard pointer of rank %d", k);
        P.table[k] = (const float*)shard_ptrs[k];
        P.user_rows[k] = user_rows_by_rank[k];
    }
    gather_rows_sharded_kernel<<<stream_grid((n_user + n_item) * (D / 4)), NT, 0, (cudaStream_t)stream>>>(
        P, world, user_ids, n_user, item_ids, n_item, n_user_rows, n_item_rows, D / 4, out, err_flag);
    RB_LAUNCH_CHECK("gather_rows_sharded_kernel");
    return RB200_OK;
}

extern "C" int rb200_push_rows_sharded(void* const* grad_bucket_ptrs, void* const* row_bucket_ptrs, int world, int rank, int64_t capacity,
                                       const float* drows, const int64_t* slot_of_sample, int64_t n, int D, const int64_t* send_rows,
                                       void* stream) {
    RB_REQUIRE(grad_bucket_ptrs && row_bucket_ptrs && world >= 1 && world <= RB200_MAX_PEERS && rank >= 0 && rank < world,
               "push_rows_sharded: 1..%d ranks", RB200_MAX_PEERS);
    RB_REQUIRE(drows && slot_of_sample && n >= 1 && capacity >= 1 && D >= 4 && D % 4 == 0, "push_rows_sharded: bad arguments");
    PeerBuckets P{};
    for (int k = 0; k < world; ++k) {
        RB_REQUIRE(grad_bucket_ptrs[k] && (row_bucket_ptrs[k] || !send_rows), "push_rows_sharded: NULL bucket pointer of rank %d", k);
        P.grads[k] = (float*)grad_bucket_ptrs[k];
        P.rows[k] = (int64_t*)row_bucket_ptrs[k];
    }
    const long long work = n * (D / 4) > (long long)world * capacity ? n * (D / 4) : (long long)world * capacity;
    push_rows_sharded_kernel<<<stream_grid(work), NT, 0, (cudaStream_t)stream>>>(P, world, rank, capacity, drows, slot_of_sample, n, D / 4,
                                                                                send_rows);
    RB_LAUNCH_CHECK("push_rows_sharded_kernel");
    return RB200_OK;
}

extern "C" int rb200_push_row_lists_sharded(void* const* row_bucket_ptrs, int world, int rank, int64_t capacity, const int64_t* send_rows,
                                            void* stream) {
    RB_REQUIRE(row_bucket_ptrs && send_rows && world >= 1 && world <= RB200_MAX_PEERS && rank >= 0 && rank < world && capacity >= 1,
               "push_row_lists_sharded: 1..%d ranks", RB200_MAX_PEERS);
    PeerBuckets P{};
    for (int k = 0; k < world; ++k) {
        RB_REQUIRE(row_bucket_ptrs[k], "push_row_lists_sharded: NULL bucket pointer of rank %d", k);
        P.rows[k] = (int64_t*)row_bucket_ptrs[k];
    }
    push_row_lists_kernel<<<stream_grid((long long)world * capacity), NT, 0, (cudaStream_t)stream>>>(P, world, rank, capacity, send_rows);
    RB_LAUNCH_CHECK("push_row_lists_kernel");
    return RB200_OK;
}

extern "C" int rb200_opt_begin_step(rb200_opt_state* st, void* stream) {
    RB_REQUIRE(st, "opt_begin_step: NULL state");
    opt_begin_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(st);
    RB_LAUNCH_CHECK("opt_begin_step_kernel");
    return RB200_OK;
}

extern "C" size_t rb200_sumsq_workspace_bytes(void) { return 256 + sizeof(double) * (size_t)rb_sm_count() * 4; }

int rb_sumsq_accumulate(rb200_opt_state* st, const rb200_sumsq_seg* segs, int n_segs, int do_clip, void* workspace,
                        size_t workspace_bytes, cudaStream_t s);

extern "C" int rb200_sumsq_accumulate(rb200_opt_state* st, const rb200_sumsq_seg* segs, int n_segs, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    return rb_sumsq_accumulate(st, segs, n_segs, 0, workspace, workspace_bytes, (cudaStream_t)stream);
}

// do_clip != 0 also finalises total_norm / clip_coef in the same launch (this must then be the last accumulate call)
int rb_sumsq_accumulate(rb200_opt_state* st, const rb200_sumsq_seg* segs, int n_segs, int do_clip, void* workspace,
                        size_t workspace_bytes, cudaStream_t s) {
    RB_REQUIRE(st && segs && n_segs >= 1 && n_segs <= 4, "sumsq_accumulate: 1..4 segments");
    RbArena ar(workspace, workspace_bytes);
    const int cap = rb_sm_count() * 4;
    double* partials = ar.take<double>(cap);
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "sumsq_accumulate: workspace too small");
    SumsqSegs k{};
    k.n = n_segs;
    long long n = 1;
    for (int i = 0; i < n_segs; ++i) { k.s[i] = segs[i]; if (segs[i].n > n) n = segs[i].n; }
    int grid = (int)((n / 4 + NT - 1) / NT);
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    sumsq_kernel<<<grid, NT, 0, s>>>(k, partials, st, do_clip);
    RB_LAUNCH_CHECK("sumsq_kernel");
    return RB200_OK;
}

extern "C" int rb200_grad_norm_clip(rb200_opt_state* st, void* stream) {
    RB_REQUIRE(st, "grad_norm_clip: NULL state");
    grad_norm_clip_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(st);
    RB_LAUNCH_CHECK("grad_norm_clip_kernel");
    return RB200_OK;
}

extern "C" int rb200_adam_dense(float* w, const float* g, float* m, float* v, int64_t n, const rb200_opt_state* st, void* stream) {
    RB_REQUIRE(w && m && v && st && n >= 0, "adam_dense: bad arguments");
    if (n == 0) return RB200_OK;
    adam_dense_kernel<<<stream_grid(n / 4 + 1), NT, 0, (cudaStream_t)stream>>>(w, g, m, v, n, st);
    RB_LAUNCH_CHECK("adam_dense_kernel");
    return RB200_OK;
}

extern "C" int rb200_adam_table_dense(float* w, float* m, float* v, int64_t n_rows, int D, const int* row_slot,
                                      const float* uniq_grads, const rb200_opt_state* st, void* stream) {
    RB_REQUIRE(w && m && v && st && n_rows >= 1 && D >= 4 && D % 4 == 0, "adam_table_dense: bad arguments");
    RB_REQUIRE(row_slot == nullptr || uniq_grads != nullptr, "adam_table_dense: row_slot without uniq_grads");
    adam_table_dense_kernel<<<stream_grid(n_rows * (D / 4)), NT, 0, (cudaStream_t)stream>>>(w, m, v, n_rows, D / 4, row_slot,
                                                                                            uniq_grads, st);
    RB_LAUNCH_CHECK("adam_table_dense_kernel");
    return RB200_OK;
}

extern "C" int rb200_adam_rows(float* w, float* m, float* v, int D, const int64_t* uniq_ids, const float* uniq_grads,
                               const int* n_uniq, int max_uniq, const rb200_opt_state* st, void* stream) {
    RB_REQUIRE(w && m && v && st && uniq_ids && uniq_grads && n_uniq && D >= 4 && D % 4 == 0, "adam_rows: bad arguments");
    if (max_uniq <= 0) return RB200_OK;
    adam_rows_kernel<<<stream_grid((long long)max_uniq * (D / 4)), NT, 0, (cudaStream_t)stream>>>(w, m, v, D / 4, uniq_ids,
                                                                                                  uniq_grads, n_uniq, st);
    RB_LAUNCH_CHECK("adam_rows_kernel");
    return RB200_OK;
}

// The sharded step's optimizer in ONE launch: Adam on the touched rows of the table shard and on up to two dense parameter blocks (the
// replicated MLPs) — three launches before, at ≈ 3 us of launch boundary each inside the step's graph.
__global__ void __launch_bounds__(NT) adam_rows_dense2_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, int D4,
                                                              const int64_t* __restrict__ uniq_ids, const float* __restrict__ uniq_grads,
                                                              const int* __restrict__ n_uniq, const AdamDenseParams dp,
                                                              const rb200_opt_state* __restrict__ st) {
    const AdamK k = load_adam(st);
    const long long stride = (long long)gridDim.x * NT;
    const long long n4 = (long long)n_uniq[0] * D4;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += stride) {
        const long long u = i / D4;
        const int c = (int)(i - u * D4);
        const long long o = __ldg(uniq_ids + u) * D4 + c;
        const float4 gv = __ldg(reinterpret_cast<const float4*>(uniq_grads) + i);
        float4 wv = reinterpret_cast<float4*>(w)[o], mv = reinterpret_cast<float4*>(m)[o], vv = reinterpret_cast<float4*>(v)[o];
        adam4(wv, gv, mv, vv, k);
        reinterpret_cast<float4*>(w)[o] = wv; reinterpret_cast<float4*>(m)[o] = mv; reinterpret_cast<float4*>(v)[o] = vv;
    }
    for (int j = 0; j < dp.n_jobs; ++j) {
        const AdamDenseJob& J = dp.job[j];
        const long long d4 = J.n >> 2;
        for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < d4; i += stride) {
            float4 wv = reinterpret_cast<float4*>(J.w)[i], mv = reinterpret_cast<float4*>(J.m)[i], vv = reinterpret_cast<float4*>(J.v)[i];
            const float4 gv = __ldg(reinterpret_cast<const float4*>(J.g) + i);
            adam4(wv, gv, mv, vv, k);
            reinterpret_cast<float4*>(J.w)[i] = wv; reinterpret_cast<float4*>(J.m)[i] = mv; reinterpret_cast<float4*>(J.v)[i] = vv;
        }
        for (long long i = d4 * 4 + (long long)blockIdx.x * NT + threadIdx.x; i < J.n; i += stride) {
            float wv = J.w[i], mv = J.m[i], vv = J.v[i];
            adam1(wv, __ldg(J.g + i), mv, vv, k);
            J.w[i] = wv; J.m[i] = mv; J.v[i] = vv;
        }
    }
}

extern "C" int rb200_adam_rows_dense2(float* w, float* m, float* v, int D, const int64_t* uniq_ids, const float* uniq_grads, const int* n_uniq,
                                      int max_uniq, float* w0, const float* g0, float* m0, float* v0, int64_t n0, float* w1, const float* g1,
                                      float* m1, float* v1, int64_t n1, const rb200_opt_state* st, void* stream) {
    RB_REQUIRE(w && m && v && st && uniq_ids && uniq_grads && n_uniq && D >= 4 && D % 4 == 0 && max_uniq >= 0, "adam_rows_dense2: bad arguments");
    RB_REQUIRE(n0 >= 0 && n1 >= 0 && (n0 == 0 || (w0 && g0 && m0 && v0)) && (n1 == 0 || (w1 && g1 && m1 && v1)), "adam_rows_dense2: bad dense blocks");
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    RB_REQUIRE((n0 == 0 || (al(w0) && al(g0) && al(m0) && al(v0))) && (n1 == 0 || (al(w1) && al(g1) && al(m1) && al(v1))),
               "adam_rows_dense2: the dense blocks must be 16-byte aligned");
    AdamDenseParams dp{};
    if (n0 > 0) dp.job[dp.n_jobs++] = AdamDenseJob{w0, g0, m0, v0, (long long)n0};
    if (n1 > 0) dp.job[dp.n_jobs++] = AdamDenseJob{w1, g1, m1, v1, (long long)n1};
    long long work = (long long)max_uniq * (D / 4);
    if (n0 / 4 > work) work = n0 / 4;
    if (n1 / 4 > work) work = n1 / 4;
    if (work <= 0) return RB200_OK;
    adam_rows_dense2_kernel<<<stream_grid(work), NT, 0, (cudaStream_t)stream>>>(w, m, v, D / 4, uniq_ids, uniq_grads, n_uniq, dp, st);
    RB_LAUNCH_CHECK("adam_rows_dense2_kernel");
    return RB200_OK;
}

// ---- fused-step helpers: two tensors per launch ------------------------------------------------------------------ //
int rb_adam_dense2(float* w0, const float* g0, float* m0, float* v0, long long n0, float* w1, const float* g1, float* m1, float* v1,
                   long long n1, const rb200_opt_state* st, cudaStream_t s) {
    AdamDenseParams p{};
    p.n_jobs = 2;
    p.job[0] = {w0, g0, m0, v0, n0};
    p.job[1] = {w1, g1, m1, v1, n1};
    adam_dense2_kernel<<<stream_grid((n0 > n1 ? n0 : n1) / 4 + 1), NT, 0, s>>>(p, st);
    RB_LAUNCH_CHECK("adam_dense2_kernel");
    return RB200_OK;
}

int rb_adam_tables_dense2(float* w0, float* m0, float* v0, long long rows0, int* slot0, const float* ug0, float* w1, float* m1, float* v1,
                          long long rows1, int* slot1, const float* ug1, int D, const rb200_opt_state* st, cudaStream_t s) {
    AdamTableParams p{};
    p.n_jobs = 2; p.D4 = D / 4;
    p.job[0] = {w0, m0, v0, rows0, slot0, ug0};
    p.job[1] = {w1, m1, v1, rows1, slot1, ug1};
    adam_table_dense2_kernel<<<stream_grid((rows0 > rows1 ? rows0 : rows1) * (D / 4)), NT, 0, s>>>(p, st);
    RB_LAUNCH_CHECK("adam_table_dense2_kernel");
    return RB200_OK;
}

// returns 1 when D does not allow the fused launch (caller then uses the separate kernels)
int rb_adam_step_all(float* const mlp_w[2], const float* const mlp_g[2], float* const mlp_m[2], float* const mlp_v[2],
                     const long long mlp_n[2], float* const tab_w[2], float* const tab_m[2], float* const tab_v[2],
                     const long long tab_rows[2], int* const slot[2], const float* const ug[2], int D, rb200_opt_state* st,
                     const double* norm_partials, int n_norm_partials, cudaStream_t s) {
    const int D4 = D / 4;
    if (D % 4 != 0 || D4 > 32 || (D4 & (D4 - 1)) != 0) return 1;
    AdamTableParams tp{};
    AdamDenseParams dp{};
    tp.n_jobs = 2; tp.D4 = D4; dp.n_jobs = 2;
    for (int t = 0; t < 2; ++t) {
        tp.job[t] = {tab_w[t], tab_m[t], tab_v[t], tab_rows[t], slot[t], ug[t]};
        dp.job[t] = {mlp_w[t], mlp_g[t], mlp_m[t], mlp_v[t], mlp_n[t]};
    }
    const long long big = (tab_rows[0] > tab_rows[1] ? tab_rows[0] : tab_rows[1]) * D4;
    adam_step_all_kernel<<<stream_grid(big), NT, 0, s>>>(tp, dp, st, norm_partials, n_norm_partials);
    RB_LAUNCH_CHECK("adam_step_all_kernel");
    return RB200_OK;
}

int rb_reset_slots2(const int64_t* ids0, const int* n0, int cap0, int* slot0, const int64_t* ids1, const int* n1, int cap1, int* slot1,
                    cudaStream_t s) {
    reset_slots2_kernel<<<stream_grid(cap0 + cap1), NT, 0, s>>>(ids0, n0, slot0, ids1, n1, slot1);
    RB_LAUNCH_CHECK("reset_slots2_kernel");
    return RB200_OK;
}

// ------------------------------------------------------------------------------------------------------------ //
// exchange plan of the row-sharded step: stable partition of the requests by owner (one 1-pass radix sort on log2(world) bits)
// ------------------------------------------------------------------------------------------------------------ //
namespace {
__global__ void __launch_bounds__(NT) route_prep_kernel(const int64_t* __restrict__ user_ids, long long n_user,
                                                        const int64_t* __restrict__ item_ids, long long n, int world,
                                                        const int64_t* __restrict__ user_rows_by_rank, unsigned* __restrict__ keys,
                                                        int* __restrict__ vals, int64_t* __restrict__ local, int* __restrict__ counts) {
    // per-block histogram in shared memory, one global atomic per (block, owner): 24 576 threads hammering `world` global
    // counters took 19.5 us of the C4 step (integer counts: the result does not depend on the order)
    __shared__ int hist[64];
    const bool small = world <= 64;
    if (small && threadIdx.x < 64) hist[threadIdx.x] = 0;
    if (small) __syncthreads();
    const long long i = (long long)blockIdx.x * NT + threadIdx.x;
    if (i < n) {
        const bool is_item = i >= n_user;
        const long long id = is_item ? item_ids[i - n_user] : user_ids[i];
        const int owner = (int)(id % world);
        keys[i] = (unsigned)owner;
        vals[i] = (int)i;
        local[i] = id / world + (is_item ? user_rows_by_rank[owner] : 0);
        if (small) atomicAdd(&hist[owner], 1); else atomicAdd(&counts[owner], 1);
    }
    if (small) {
        __syncthreads();
        if (threadIdx.x < world && hist[threadIdx.x]) atomicAdd(&counts[threadIdx.x], hist[threadIdx.x]);
    }
}
// padded form: bucket w occupies slots [w·C, (w+1)·C); grid covers max(n, world·C)
__global__ void __launch_bounds__(NT) route_finish_padded_kernel(const unsigned* __restrict__ keys_sorted, const int* __restrict__ vals_sorted,
                                                                 const int64_t* __restrict__ local, long long n, int world,
                                                                 const int* __restrict__ counts, long long C,
                                                                 int64_t* __restrict__ slot_of_sample, int64_t* __restrict__ send_rows,
                                                                 unsigned long long* __restrict__ overflow) {
    extern __shared__ int start_s[];                           // [world + 1] exclusive prefix of the bucket sizes
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < world; ++b) { start_s[b] = acc; acc += counts[b]; }
        start_s[world] = acc;
    }
    __syncthreads();
    const long long j = (long long)blockIdx.x * NT + threadIdx.x;
    if (j < n) {
        const int smp = vals_sorted[j];
        const int owner = (int)keys_sorted[j];
        const long long off = j - start_s[owner];
        if (off < C) {
            const long long slot = (long long)owner * C + off;
            send_rows[slot] = local[smp];
            slot_of_sample[smp] = slot;
        } else {
            slot_of_sample[smp] = (long long)world * C;        // does not fit: the DUMMY slot past the buffer (a zero row on the way
                                                               // in, a dropped gradient on the way out) — counted, the caller reports it
            atomicAdd(overflow, 1ull);
        }
    }
    if (j < (long long)world * C) {
        const int b = (int)(j / C);
        const long long off = j - (long long)b * C;
        const long long cnt = start_s[b + 1] - start_s[b];
        if (off >= (cnt < C ? cnt : C)) send_rows[j] = -1;     // empty slot
    }
}
__global__ void __launch_bounds__(NT) route_finish_kernel(const int* __restrict__ vals_sorted, const int64_t* __restrict__ local,
                                                          long long n, int world, const int* __restrict__ counts,
                                                          int64_t* __restrict__ perm, int64_t* __restrict__ inv,
                                                          int64_t* __restrict__ local_rows, int64_t* __restrict__ send_counts) {
    const long long j = (long long)blockIdx.x * NT + threadIdx.x;
    if (j < world) send_counts[j] = counts[j];
    if (j >= n) return;
    const int smp = vals_sorted[j];
    perm[j] = smp;
    inv[smp] = j;
    local_rows[j] = local[smp];
}
size_t route_temp_bytes(long long n, int bits) {
    size_t t = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n, 0, bits);
    return t;
}
}  // namespace

extern "C" size_t rb200_route_plan_workspace_bytes(int64_t n, int world) {
    if (n < 1) n = 1;
    return 256 * 8 + sizeof(unsigned) * 2 * (size_t)n + sizeof(int) * (2 * (size_t)n + (size_t)world) + sizeof(int64_t) * (size_t)n +
           route_temp_bytes(n, 32);
}

extern "C" int rb200_route_plan(const int64_t* user_ids, int64_t n_user, const int64_t* item_ids, int64_t n_item, int world,
                                const int64_t* user_rows_by_rank, int64_t* perm, int64_t* inv, int64_t* local_rows,
                                int64_t* send_counts, void* workspace, size_t workspace_bytes, void* stream) {
    const long long n = n_user + n_item;
    RB_REQUIRE(n_user >= 0 && n_item >= 0 && world >= 1 && n < (1ll << 31), "route_plan: bad sizes");
    RB_REQUIRE((n_user == 0 || user_ids) && (n_item == 0 || item_ids) && user_rows_by_rank && perm && inv && local_rows && send_counts,
               "route_plan: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    RbArena ar(workspace, workspace_bytes);
    unsigned* keys = ar.take<unsigned>(n > 0 ? n : 1); unsigned* keys_out = ar.take<unsigned>(n > 0 ? n : 1);
    int* vals = ar.take<int>(n > 0 ? n : 1); int* vals_out = ar.take<int>(n > 0 ? n : 1);
    int* counts = ar.take<int>(world);
    int64_t* local = ar.take<int64_t>(n > 0 ? n : 1);
    int bits = 1;
    while ((1 << bits) < world) ++bits;
    size_t tb = route_temp_bytes(n > 0 ? n : 1, bits);
    char* temp = ar.take<char>(tb);
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "route_plan: workspace too small (%zu given)", workspace_bytes);
    RB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * world, st));
    if (n > 0) {
        const unsigned grid = (unsigned)((n + NT - 1) / NT);
        route_prep_kernel<<<grid, NT, 0, st>>>(user_ids, n_user, item_ids, n, world, user_rows_by_rank, keys, vals, local, counts);
        RB_LAUNCH_CHECK("route_prep_kernel");
        RB_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, (const unsigned*)keys, keys_out, (const int*)vals, vals_out, (int)n, 0, bits, st));
    }
    const long long m = n > world ? n : world;
    route_finish_kernel<<<(unsigned)((m + NT - 1) / NT), NT, 0, st>>>(vals_out, local, n, world, counts, perm, inv, local_rows, send_counts);
    RB_LAUNCH_CHECK("route_finish_kernel");
    return RB200_OK;
}

extern "C" size_t rb200_route_plan_padded_workspace_bytes(int64_t n, int world) { return rb200_route_plan_workspace_bytes(n, world); }

extern "C" int rb200_route_plan_padded(const int64_t* user_ids, int64_t n_user, const int64_t* item_ids, int64_t n_item, int world,
                                       const int64_t* user_rows_by_rank, int64_t capacity, int64_t* slot_of_sample,
                                       int64_t* send_rows, int64_t* overflow, void* workspace, size_t workspace_bytes, void* stream) {
    const long long n = n_user + n_item;
    RB_REQUIRE(n_user >= 0 && n_item >= 0 && n >= 1 && world >= 1 && world <= 4096 && n < (1ll << 31) && capacity >= 1 &&
               (long long)world * capacity < (1ll << 40), "route_plan_padded: bad sizes");
    RB_REQUIRE((n_user == 0 || user_ids) && (n_item == 0 || item_ids) && user_rows_by_rank && slot_of_sample && send_rows && overflow,
               "route_plan_padded: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    RbArena ar(workspace, workspace_bytes);
    unsigned* keys = ar.take<unsigned>(n); unsigned* keys_out = ar.take<unsigned>(n);
    int* vals = ar.take<int>(n); int* vals_out = ar.take<int>(n);
    int* counts = ar.take<int>(world);
    int64_t* local = ar.take<int64_t>(n);
    int bits = 1;
    while ((1 << bits) < world) ++bits;
    size_t tb = route_temp_bytes(n, bits);
    char* temp = ar.take<char>(tb);
    if (!workspace || !ar.ok()) return rb_set_error(RB200_ERR_WORKSPACE, "route_plan_padded: workspace too small (%zu given)", workspace_bytes);
    RB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * world, st));
    route_prep_kernel<<<(unsigned)((n + NT - 1) / NT), NT, 0, st>>>(user_ids, n_user, item_ids, n, world, user_rows_by_rank, keys, vals, local, counts);
    RB_LAUNCH_CHECK("route_prep_kernel");
    RB_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, (const unsigned*)keys, keys_out, (const int*)vals, vals_out, (int)n, 0, bits, st));
    const long long m = n > (long long)world * capacity ? n : (long long)world * capacity;
    route_finish_padded_kernel<<<(unsigned)((m + NT - 1) / NT), NT, sizeof(int) * (world + 1), st>>>(
        keys_out, vals_out, local, n, world, counts, capacity, slot_of_sample, send_rows, reinterpret_cast<unsigned long long*>(overflow));
    RB_LAUNCH_CHECK("route_finish_padded_kernel");
    return RB200_OK;
}
