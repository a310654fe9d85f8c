// Device-side batch producer for the BPR step — the sample stream of the reference's UserItemDataset + DataLoader
// (src/training/train_embeddings.py:23-79, 144-151) produced where it is consumed:
//   * positives: (user, item) pairs with rating >= threshold; every epoch visits them in a fresh pseudo-random order
//     (DataLoader(shuffle=True, drop_last=True)).  The order is a keyed bijection of [0, n_pos) computed on the fly (4-round
//     Feistel network on the next power-of-four domain with cycle walking): no permutation array, no per-epoch sort.
//   * negatives: uniform over the catalog, rejected while the user has rated the item (:58-63) — membership by binary search in
//     the user's sorted rated list (CSR), draws from Philox4x32-10 keyed by (seed; epoch, sample, attempt).
// Counter-based throughout: batch (epoch, step) is a pure function of the seed — reproducible, order-independent, and exactly
// restated on the CPU by the test suite.
#include "common.cuh"

namespace {

constexpr int NT_S = 64;       // small blocks: a batch of 8192 samples spreads over 128 SMs (each thread is one latency chain)
constexpr int MAX_ATTEMPTS = 64;     // a user who rated (almost) the whole catalog: the last draw is taken

__host__ __device__ inline uint32_t mix32(uint32_t x) {        // murmur3 finaliser
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}

// bijection of [0, n): Feistel on 2·h bits (4^h >= n), cycle-walked back into range
__host__ __device__ inline uint32_t feistel_perm(uint32_t i, uint32_t n, int h, uint32_t k0, uint32_t k1) {
    const uint32_t mask = (1u << h) - 1u;
    uint32_t x = i;
    do {
        uint32_t L = x >> h, R = x & mask;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t F = mix32(R ^ (r & 1 ? k1 : k0) ^ (0x9E3779B9u * (uint32_t)(r + 1))) & mask;
            const uint32_t nl = R;
            R = L ^ F;
            L = nl;
        }
        x = (L << h) | R;
    } while (x >= n);
    return x;
}

// counter != NULL: epoch / step come from the device-resident global batch index (the optimizer's step counter)
__global__ void __launch_bounds__(NT_S) sample_batch_kernel(const rb200_sampler S, int B, long long epoch, long long step,
                                                            const long long* __restrict__ counter, int h,
                                                            int64_t* __restrict__ out_users, int64_t* __restrict__ out_pos,
                                                            int64_t* __restrict__ out_neg) {
    const int s = blockIdx.x * NT_S + threadIdx.x;
    if (s >= B) return;
    if (counter) { const long long g = counter[0]; epoch = g / S.batches_per_epoch; step = g - epoch * S.batches_per_epoch; }
    const uint64_t seed = S.seed;
    const uint32_t k0 = mix32((uint32_t)seed ^ 0xA511E9B3u) + (uint32_t)epoch * 0x632BE5ABu;
    const uint32_t k1 = mix32((uint32_t)(seed >> 32) ^ 0x94D049BBu) ^ mix32((uint32_t)epoch + 0x7F4A7C15u);
    const unsigned long long W = S.world > 1 ? (unsigned long long)S.world : 1ull, R = S.world > 1 ? (unsigned long long)S.rank : 0ull;
    const unsigned long long slot = ((unsigned long long)step * W + R) * (unsigned long long)B + (unsigned long long)s;   // < n_pos (drop_last)
    const uint32_t p = feistel_perm((uint32_t)slot, (uint32_t)S.n_pos, h, k0, k1);
    const long long u = S.pos_users[p];
    out_users[s] = u;
    out_pos[s] = S.pos_items[p];
    const uint32_t* bm = S.rated_bitmap ? S.rated_bitmap + u * S.bitmap_words : nullptr;
    long long rb = 0, re = 0;
    if (!bm) { rb = S.rated_offsets[u]; re = S.rated_offsets[u + 1]; }
    long long neg = 0;
    for (int a = 0; a < MAX_ATTEMPTS; a += 4) {
        const uint4 rnd = rb_philox4x32(make_uint4((uint32_t)slot, (uint32_t)(slot >> 32), (uint32_t)epoch, (uint32_t)(a >> 2)),
                                        make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
        long long cand[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) cand[e] = S.catalog[((unsigned long long)rw[e] * (unsigned long long)S.n_cat) >> 32];   // uniform
        bool found = false;
#pragma unroll
        for (int e = 0; e < 4 && !found; ++e) {
            neg = cand[e];
            bool rated;
            if (bm) {
                rated = (neg >> 5) < S.bitmap_words && ((bm[neg >> 5] >> (neg & 31)) & 1u);
            } else {
                long long lo = rb, hi = re;                                                      // is neg in rated[rb, re)?
                while (lo < hi) { const long long mid = (lo + hi) >> 1; if (S.rated_items[mid] < neg) lo = mid + 1; else hi = mid; }
                rated = lo < re && S.rated_items[lo] == neg;
            }
            found = !rated;
        }
        if (found) break;
    }
    out_neg[s] = neg;
}

int feistel_half_bits(long long n_pos) {
    int h = 1;
    while ((1ll << (2 * h)) < n_pos) ++h;
    return h;
}

}  // namespace

// used by csrc/step.cu (next_batch) and the two C entry points
int rb_sample_batch(const rb200_sampler& S, int B, long long epoch, long long step, const int64_t* counter_dev, int64_t* out_users,
                    int64_t* out_pos, int64_t* out_neg, cudaStream_t st) {
    RB_REQUIRE(S.pos_users && S.pos_items && S.catalog && (S.rated_bitmap || (S.rated_offsets && S.rated_items)) && out_users &&
               out_pos && out_neg, "sample_batch: NULL pointer");
    RB_REQUIRE(B >= 1 && S.n_cat >= 1 && S.n_pos >= B && S.n_pos < (1ll << 31) && S.n_cat < (1ll << 31), "sample_batch: bad sizes");
    const long long W = S.world > 1 ? S.world : 1;
    RB_REQUIRE(S.world <= 1 || (S.rank >= 0 && S.rank < S.world), "sample_batch: rank outside [0, world)");
    if (counter_dev) {
        RB_REQUIRE(S.batches_per_epoch >= 1 && S.batches_per_epoch * (long long)B * W <= S.n_pos,
                   "sample_batch: batches_per_epoch * world * B exceeds the number of positives (drop_last)");
    } else {
        RB_REQUIRE(epoch >= 0 && step >= 0 && (step + 1) * (long long)B * W <= S.n_pos, "sample_batch: step %lld is past the epoch (drop_last)",
                   step);
    }
    sample_batch_kernel<<<(B + NT_S - 1) / NT_S, NT_S, 0, st>>>(S, B, epoch, step, reinterpret_cast<const long long*>(counter_dev),
                                                              feistel_half_bits(S.n_pos), out_users, out_pos, out_neg);
    RB_LAUNCH_CHECK("sample_batch_kernel");
    return RB200_OK;
}

extern "C" int rb200_sample_batch(const int64_t* pos_users, const int64_t* pos_items, int64_t n_pos, const int64_t* rated_offsets,
                                  const int64_t* rated_items, const int64_t* catalog, int64_t n_cat, int B, uint64_t seed,
                                  int64_t epoch, int64_t step, int64_t rank, int64_t world, int64_t* out_users, int64_t* out_pos,
                                  int64_t* out_neg, void* stream) {
    rb200_sampler S{};
    S.pos_users = pos_users; S.pos_items = pos_items; S.n_pos = n_pos; S.rated_offsets = rated_offsets; S.rated_items = rated_items;
    S.catalog = catalog; S.n_cat = n_cat; S.seed = seed; S.batches_per_epoch = 1; S.rank = rank; S.world = world;
    return rb_sample_batch(S, B, epoch, step, nullptr, out_users, out_pos, out_neg, (cudaStream_t)stream);
}

extern "C" int rb200_sample_batch_dev(const rb200_sampler* s, int B, const int64_t* counter_dev, int64_t* out_users,
                                      int64_t* out_pos, int64_t* out_neg, void* stream) {
    RB_REQUIRE(s && counter_dev, "sample_batch_dev: NULL sampler / counter");
    return rb_sample_batch(*s, B, 0, 0, counter_dev, out_users, out_pos, out_neg, (cudaStream_t)stream);
}
