"""The batch-producer oracle against the reference's sampling semantics (train_embeddings.py:23-79, 144-151) — CPU."""
import numpy as np

from oracle import sampler_oracle as S


def _toy(seed=0, n_users=40, n_items=60, n_ratings=900):
    rng = np.random.default_rng(seed)
    u = rng.integers(1, n_users + 1, n_ratings)
    i = rng.integers(1, n_items + 1, n_ratings)
    r = rng.integers(1, 6, n_ratings)
    pairs = np.unique(np.stack([u, i], 1), axis=0, return_index=True)[1]
    u, i, r = u[pairs], i[pairs], r[pairs]
    pos = r >= 4
    offsets, rated = S.build_rated_csr(u, i, n_users)
    catalog = np.arange(1, n_items + 1, dtype=np.int64)
    return u[pos].astype(np.int64), i[pos].astype(np.int64), offsets, rated, catalog, (u, i)


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert [int(x) for x in S.philox4x32(0, 0, 0, 0, 0, 0)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert [int(x) for x in S.philox4x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)] == \
        [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


def test_feistel_is_a_permutation_and_differs_per_epoch():
    for n in (1, 2, 7, 64, 1000, 4097):
        k = S.epoch_keys(5, 0)
        p = S.feistel_perm(np.arange(n), n, *k)
        assert np.array_equal(np.sort(p), np.arange(n))
    a = S.feistel_perm(np.arange(1000), 1000, *S.epoch_keys(5, 0))
    b = S.feistel_perm(np.arange(1000), 1000, *S.epoch_keys(5, 1))
    assert (a != b).mean() > 0.95


def test_epoch_visits_every_positive_once_and_drops_the_last_partial_batch():
    pu, pi, off, rated, cat, _ = _toy()
    B = 32
    nb = len(pu) // B
    seen = []
    for step in range(nb):
        u, p, n = S.sample_batch(pu, pi, off, rated, cat, B, seed=9, epoch=2, step=step)
        seen.append(np.stack([u, p], 1))
    seen = np.concatenate(seen)
    assert len(np.unique(seen, axis=0)) == nb * B                       # no positive twice in an epoch
    allpos = set(map(tuple, np.stack([pu, pi], 1)))
    assert all(tuple(x) in allpos for x in seen)


def test_negatives_are_never_rated_and_cover_the_unrated_catalog_uniformly():
    pu, pi, off, rated, cat, (u_all, i_all) = _toy(seed=3, n_users=5, n_items=40, n_ratings=100)
    rated_set = {}
    for u, i in zip(u_all, i_all):
        rated_set.setdefault(int(u), set()).add(int(i))
    B = len(pu)
    counts = {}
    for epoch in range(300):
        u, p, n = S.sample_batch(pu, pi, off, rated, cat, B, seed=1, epoch=epoch, step=0)
        for uu, nn in zip(u, n):
            assert int(nn) not in rated_set.get(int(uu), set())
            counts.setdefault(int(uu), []).append(int(nn))
    for uu, draws in counts.items():
        free = sorted(set(range(1, 41)) - rated_set[uu])
        h = np.array([draws.count(f) for f in free], dtype=np.float64)
        exp = len(draws) / len(free)
        chi2 = ((h - exp) ** 2 / exp).sum()
        assert chi2 < 2.0 * len(free) + 30, (uu, chi2)                 # loose: uniform over the unrated items


def test_data_parallel_ranks_take_disjoint_slices_of_one_epoch():
    """world ranks with the same seed: a step consumes world·B positives, rank r the r-th slice — together the ranks visit
    exactly what a single process with batch world·B visits in the same steps."""
    pu, pi, off, rated, cat, _ = _toy()
    B, world = 8, 3
    nb = len(pu) // (B * world)
    assert nb >= 2
    for step in range(nb):
        parts = [S.sample_batch(pu, pi, off, rated, cat, B, seed=4, epoch=1, step=step, rank=r, world=world) for r in range(world)]
        whole = S.sample_batch(pu, pi, off, rated, cat, B * world, seed=4, epoch=1, step=step)
        for k in range(3):
            assert np.array_equal(np.concatenate([p[k] for p in parts]), whole[k])


def test_index_build_equals_the_reference_dataset():
    """Positives (order and repeats) and the rated relation against the reference's own UserItemDataset
    (tests/golden/sampler.npz, made by tests/golden/make_sampler_golden.py from train_embeddings.py:23-63); the reference's
    negatives for three users must lie in the support the oracle's rejection test allows."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "sampler.npz")
    u, i, r, cat = g["user_id"], g["item_id"], g["rating"], g["all_item_ids"]
    pos = r >= 4.0
    assert np.array_equal(u[pos], g["pos_users"]) and np.array_equal(i[pos], g["pos_items"]) and int(g["n_samples"]) == int(pos.sum())
    off, rated = S.build_rated_csr(u, i, int(u.max()))
    pairs = np.array([(usr, it) for usr in range(len(off) - 1) for it in rated[off[usr]:off[usr + 1]]], dtype=np.int64)
    assert np.array_equal(pairs, g["rated_pairs"])
    for usr, draws in zip(g["neg_user"], g["neg_draws"]):
        allowed = np.setdiff1d(cat, rated[off[usr]:off[usr + 1]])
        assert np.isin(draws, allowed).all()
        # and the oracle's own negatives for that user cover the same support
        idx = np.nonzero(u[pos] == usr)[0]
        if len(idx):
            seen = set()
            for epoch in range(60):
                _, _, n = S.sample_batch(u[pos], i[pos], off, rated, cat, int(pos.sum()), seed=2, epoch=epoch, step=0)
                perm_users = S.sample_batch(u[pos], i[pos], off, rated, cat, int(pos.sum()), seed=2, epoch=epoch, step=0)[0]
                seen.update(n[perm_users == usr].tolist())
            assert seen <= set(allowed.tolist())


def test_product_host_index_build_equals_the_reference_dataset_and_the_oracle():
    """recommendit_b200.producer.build_host_index (the host half of DeviceBatchProducer; pure NumPy) against the golden of the
    reference's UserItemDataset and the oracle's CSR; the bitmap must encode exactly the CSR's relation."""
    from pathlib import Path
    from recommendit_b200.producer import build_host_index
    g = np.load(Path(__file__).parent / "golden" / "sampler.npz")
    u, i, r, cat = g["user_id"], g["item_id"], g["rating"], g["all_item_ids"]
    n_users = int(u.max()) + 3                                   # ids need not fill the table
    ix = build_host_index(u, i, r, cat, n_users)
    assert np.array_equal(ix["pos_users"], g["pos_users"]) and np.array_equal(ix["pos_items"], g["pos_items"])
    off, rated = S.build_rated_csr(u, i, n_users)
    assert np.array_equal(ix["rated_offsets"], off) and np.array_equal(ix["rated_items"], rated)
    pairs = np.array([(usr, it) for usr in range(n_users + 1) for it in rated[off[usr]:off[usr + 1]]], dtype=np.int64)
    assert np.array_equal(pairs, g["rated_pairs"])
    W = ix["bitmap_words"]
    bm = ix["bitmap"].reshape(n_users + 1, W)
    dense = np.zeros((n_users + 1, W * 32), dtype=bool)
    dense[pairs[:, 0], pairs[:, 1]] = True
    bits = ((bm[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(n_users + 1, W * 32)
    assert np.array_equal(bits, dense)
    assert build_host_index(u, i, r, cat, n_users, bitmap_max_bytes=0)["bitmap"] is None
    import pytest
    with pytest.raises(ValueError):
        build_host_index(u, i, r, cat, 5)                        # a user id outside the table
    with pytest.raises(ValueError):
        build_host_index(u, i, np.zeros_like(r), cat, n_users)   # no positive pair
