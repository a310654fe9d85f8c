"""CPU restatement of the device-side batch producer — TEST INFRASTRUCTURE ONLY (never imported by the product).

Semantics follow the reference's sample stream, src/training/train_embeddings.py:
  * :39-42   positives = (user, item) pairs with rating >= min_rating
  * :44-48   user_rated = every item a user has rated (any rating)
  * :58-63   negative = uniform draw from all_item_ids, redrawn while it is in user_rated[user]
  * :144-151 DataLoader(shuffle=True, drop_last=True): a fresh permutation of the positives per epoch, full batches only
The reference draws from NumPy's global RNG and torch's sampler, so its exact sample sequence is not reproducible elsewhere;
what IS pinned here is the counter-based procedure of csrc/sampler.cu (4-round Feistel permutation + Philox4x32-10 draws),
restated operation by operation so that the kernel can be checked bit-exactly, and the reference's semantic properties
(tests/test_oracle_sampler.py): permutation per epoch, negatives never rated, uniform over the unrated catalog.
"""
import numpy as np

MAX_ATTEMPTS = 64
M32 = np.uint64(0xFFFFFFFF)


def mix32(x):
    x = np.asarray(x, dtype=np.uint64) & M32
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x85EBCA6B)) & M32
    x ^= x >> np.uint64(13); x = (x * np.uint64(0xC2B2AE35)) & M32
    x ^= x >> np.uint64(16)
    return x


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011), vectorised over the counters; the round function of csrc/common.cuh:82-93."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & M32 for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0) & M32, np.uint64(k1) & M32
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & M32, p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0, k1 = (k0 + W0) & M32, (k1 + W1) & M32
    return c0, c1, c2, c3


def feistel_perm(i, n, k0, k1):
    """keyed bijection of [0, n): 4-round Feistel on 2·h bits (4^h >= n) with cycle walking (csrc/sampler.cu)"""
    h = 1
    while (1 << (2 * h)) < n:
        h += 1
    mask = np.uint64((1 << h) - 1)
    x = np.asarray(i, dtype=np.uint64).copy()
    todo = np.ones(x.shape, dtype=bool)
    while todo.any():
        L, R = x[todo] >> np.uint64(h), x[todo] & mask
        for r in range(4):
            key = np.uint64(k1 if r & 1 else k0)
            F = mix32(R ^ key ^ np.uint64((0x9E3779B9 * (r + 1)) & 0xFFFFFFFF)) & mask
            L, R = R, L ^ F
        x[todo] = (L << np.uint64(h)) | R
        todo = x >= np.uint64(n)
    return x.astype(np.int64)


def epoch_keys(seed, epoch):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    k0 = (int(mix32((seed & 0xFFFFFFFF) ^ 0xA511E9B3)) + ((int(epoch) * 0x632BE5AB) & 0xFFFFFFFF)) & 0xFFFFFFFF
    k1 = int(mix32((seed >> 32) ^ 0x94D049BB)) ^ int(mix32((int(epoch) + 0x7F4A7C15) & 0xFFFFFFFF))
    return k0, k1


def sample_batch(pos_users, pos_items, rated_offsets, rated_items, catalog, B, seed, epoch, step, rank=0, world=1):
    """→ (user_ids, pos_ids, neg_ids) int64[B] of batch `step` of `epoch`"""
    n_pos, n_cat = len(pos_users), len(catalog)
    assert (step + 1) * B * world <= n_pos, "drop_last: only full batches"
    k0, k1 = epoch_keys(seed, epoch)
    slot = np.arange((step * world + rank) * B, (step * world + rank + 1) * B, dtype=np.uint64)   # rank's slice of the step
    p = feistel_perm(slot, n_pos, k0, k1)
    users, pos = np.asarray(pos_users)[p].astype(np.int64), np.asarray(pos_items)[p].astype(np.int64)
    neg = np.zeros(B, dtype=np.int64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    for s in range(B):
        rb, re = int(rated_offsets[users[s]]), int(rated_offsets[users[s] + 1])
        rated = rated_items[rb:re]
        found = False
        for a in range(0, MAX_ATTEMPTS, 4):
            rnd = philox4x32(int(slot[s]) & 0xFFFFFFFF, int(slot[s]) >> 32, epoch & 0xFFFFFFFF, a >> 2, seed & 0xFFFFFFFF, seed >> 32)
            for e in range(4):
                cand = int(catalog[(int(rnd[e]) * n_cat) >> 32])
                neg[s] = cand
                j = int(np.searchsorted(rated, cand))
                if not (j < len(rated) and rated[j] == cand):
                    found = True
                    break
            if found:
                break
    return users, pos, neg


def build_rated_csr(user_ids, item_ids, n_users):
    """CSR of each user's rated items, ascending and de-duplicated (train_embeddings.py:44-48) → (offsets[n_users+2], items)"""
    pairs = np.unique(np.stack([np.asarray(user_ids, np.int64), np.asarray(item_ids, np.int64)], 1), axis=0)
    counts = np.bincount(pairs[:, 0], minlength=n_users + 1)
    offsets = np.zeros(n_users + 2, dtype=np.int64)
    offsets[1:] = np.cumsum(counts)
    return offsets, pairs[:, 1].copy()
