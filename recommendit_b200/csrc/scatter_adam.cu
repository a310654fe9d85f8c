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

__global__ void reset_slots_kernel(const int64_t* __restrict__ uniq_ids, const int* __restrict__ n_uniq, int* __restrict__ row_slot) {
    const int n = n_uniq[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) row_slot[uniq_ids[i]] = -1;
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
    const long long step = st->step + 1;
    st->step = step;
    const double bc1 = 1.0 - pow(st->beta1, (double)step);
    const double bc2 = 1.0 - pow(st->beta2, (double)step);
    st->step_size = (float)(st->lr / bc1);
    st->bias_corr2_sqrt = (float)sqrt(bc2);
    st->sumsq = 0.0;
}

struct SumsqSegs { rb200_sumsq_seg s[4]; int n; };

__global__ void __launch_bounds__(NT) sumsq_kernel(const SumsqSegs segs, double* __restrict__ partials) {
    __shared__ double scratch[NT / 32];
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
    }
}

__global__ void sumsq_finalize_kernel(const double* __restrict__ partials, int n, rb200_opt_state* st) {
    if (threadIdx.x || blockIdx.x) return;
    double t = st->sumsq;
    for (int i = 0; i < n; ++i) t += partials[i];
    st->sumsq = t;
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

extern "C" size_t rb200_scatter_workspace_bytes(int B, int64_t n_rows) {
    if (B < 1) B = 1;
    return 256 * 6 + sizeof(int64_t) * (size_t)B + sizeof(int) * ((size_t)3 * B + 2 * ((size_t)B + 1)) +
           sort_temp_bytes(B, key_bits(n_rows));
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

extern "C" int rb200_scatter_reset_slots(const int64_t* uniq_ids, const int* n_uniq, int max_uniq, int* row_slot, void* stream) {
    RB_REQUIRE(uniq_ids && n_uniq && row_slot, "scatter_reset_slots: NULL pointer");
    if (max_uniq <= 0) return RB200_OK;
    reset_slots_kernel<<<stream_grid(max_uniq), NT, 0, (cudaStream_t)stream>>>(uniq_ids, n_uniq, row_slot);
    RB_LAUNCH_CHECK("reset_slots_kernel");
    return RB200_OK;
}

extern "C" int rb200_opt_begin_step(rb200_opt_state* st, void* stream) {
    RB_REQUIRE(st, "opt_begin_step: NULL state");
    opt_begin_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(st);
    RB_LAUNCH_CHECK("opt_begin_step_kernel");
    return RB200_OK;
}

extern "C" size_t rb200_sumsq_workspace_bytes(void) { return 256 + sizeof(double) * (size_t)rb_sm_count() * 4; }

extern "C" int rb200_sumsq_accumulate(rb200_opt_state* st, const rb200_sumsq_seg* segs, int n_segs, void* workspace,
                                      size_t workspace_bytes, void* stream) {
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
    cudaStream_t s = (cudaStream_t)stream;
    sumsq_kernel<<<grid, NT, 0, s>>>(k, partials);
    RB_LAUNCH_CHECK("sumsq_kernel");
    sumsq_finalize_kernel<<<1, 32, 0, s>>>(partials, grid, st);
    RB_LAUNCH_CHECK("sumsq_finalize_kernel");
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
