"""Device-side batch producer (csrc/sampler.cu) against its CPU restatement — integer work, bit-exact — and against the
reference's sampling semantics (train_embeddings.py:23-79, 144-151)."""
import numpy as np
import pytest
import torch

from oracle import sampler_oracle as S

pytestmark = pytest.mark.gpu


def _ratings(seed, n_users, n_items, n_ratings):
    rng = np.random.default_rng(seed)
    u = rng.integers(1, n_users + 1, n_ratings)
    i = rng.integers(1, n_items + 1, n_ratings)
    r = rng.integers(1, 6, n_ratings).astype(np.float64)
    keep = np.unique(np.stack([u, i], 1), axis=0, return_index=True)[1]
    return u[keep], i[keep], r[keep]


@pytest.mark.parametrize("n_users,n_items,n_ratings,B", [(40, 60, 900, 32), (500, 300, 20000, 1000), (30, 12, 300, 7)])
def test_sampled_batches_equal_the_oracle_bit_for_bit(n_users, n_items, n_ratings, B):
    import recommendit_b200 as R
    u, i, r = _ratings(n_users, n_users, n_items, n_ratings)
    catalog = np.arange(1, n_items + 1, dtype=np.int64)
    prod = R.DeviceBatchProducer(u, i, r, catalog, n_users, seed=1234)
    pos = r >= 4
    off, rated = S.build_rated_csr(u, i, n_users)
    ou, op, on = (torch.empty(B, dtype=torch.int64, device="cuda") for _ in range(3))
    for epoch, step in ((0, 0), (0, prod.batches_per_epoch(B) - 1), (3, 1 % prod.batches_per_epoch(B))):
        prod.fill(ou, op, on, epoch, step)
        eu, ep, en = S.sample_batch(u[pos], i[pos], off, rated, catalog, B, 1234, epoch, step)
        assert np.array_equal(ou.cpu().numpy(), eu) and np.array_equal(op.cpu().numpy(), ep) and np.array_equal(on.cpu().numpy(), en)


def test_epoch_is_a_permutation_and_negatives_are_unrated_at_ml1m_shape():
    """BASELINE C1 shape (6040 users, 3883 catalog items, ~575 k positives): an epoch visits every positive exactly once
    (full batches), no negative was rated by its user, and past the end of the epoch the call fails loudly."""
    import recommendit_b200 as R
    from recommendit_b200 import RB200Error
    rng = np.random.default_rng(20240601)
    n_users, n_items, n = 6040, 3952, 1_000_209
    u = rng.integers(1, n_users + 1, n); i = rng.integers(1, n_items + 1, n)
    r = rng.choice([1, 2, 3, 4, 5], n, p=[0.056, 0.107, 0.261, 0.349, 0.227]).astype(np.float64)
    keep = np.unique(u * 4096 + i, return_index=True)[1]
    u, i, r = u[keep], i[keep], r[keep]
    catalog = np.sort(rng.choice(np.arange(1, n_items + 1), 3883, replace=False)).astype(np.int64)
    prod = R.DeviceBatchProducer(u, i, r, catalog, n_users, seed=7)
    B = 8192
    nb = prod.batches_per_epoch(B)
    ou, op, on = (torch.empty(B, dtype=torch.int64, device="cuda") for _ in range(3))
    rated_key = torch.as_tensor(np.sort(u * 4096 + i), device="cuda")
    seen = []
    for step in range(nb):
        prod.fill(ou, op, on, 0, step)
        seen.append(ou * 4096 + op)
        nk = ou * 4096 + on
        j = torch.searchsorted(rated_key, nk).clamp_(max=rated_key.numel() - 1)
        assert not bool((rated_key[j] == nk).any())                                  # negatives never rated by their user
        assert bool(torch.isin(on, torch.as_tensor(catalog, device="cuda")).all())
    seen = torch.cat(seen)
    assert seen.unique().numel() == nb * B                                           # each positive at most once per epoch
    with pytest.raises(RB200Error, match="past the epoch"):
        prod.fill(ou, op, on, 0, nb)


def test_train_epoch_from_the_device_producer_learns():
    import recommendit_b200 as R
    u, i, r = _ratings(5, 300, 200, 30000)
    catalog = np.arange(1, 201, dtype=np.int64)
    prod = R.DeviceBatchProducer(u, i, r, catalog, 300, seed=3)
    torch.manual_seed(0)
    model = R.TwoTowerModel(300, 200, 64, 128, dropout=0.0).cuda().train()
    genres = torch.zeros(201, 18)
    genres[torch.arange(201), torch.arange(201) % 18] = 1.0
    tr = R.FusedBPRTrainer(model, lr=1e-2, item_extra_table=genres)
    losses = [prod.train_epoch(tr, 512, e) for e in range(4)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0] - 0.01 and losses[0] < 0.75
    tr.check_ids()


def test_in_graph_producer_writes_the_same_batches_as_the_oracle_and_trains_identically():
    """rb200_step_params.next_batch: the step samples its own next batch on the device (graph replays only).  After step t
    the id buffers must hold batch t+1 of the oracle's stream, across the epoch boundary, and the trajectory must equal the one
    with an explicit sampling launch per step bit for bit."""
    import recommendit_b200 as R
    u, i, r = _ratings(11, 120, 90, 6000)
    catalog = np.arange(1, 91, dtype=np.int64)
    pos = r >= 4
    off, rated = S.build_rated_csr(u, i, 120)
    genres = torch.zeros(91, 18)
    genres[torch.arange(91), torch.arange(91) % 18] = 1.0
    B = 256
    finals = []
    for in_graph in (True, False):
        prod = R.DeviceBatchProducer(u, i, r, catalog, 120, seed=99)
        nb = prod.batches_per_epoch(B)
        torch.manual_seed(0)
        model = R.TwoTowerModel(120, 90, 64, 128, dropout=0.1).cuda().train()
        tr = R.FusedBPRTrainer(model, lr=1e-2, item_extra_table=genres, seed=5)
        if in_graph:
            tr._alloc(B)
            tr.attach_producer(prod.sampler(B), prod)
            counter = tr.opt_dev[32:40].view(torch.int64)
            prod.fill_from_counter(B, counter, tr.user_ids, tr.pos_ids, tr.neg_ids)
            for g in range(2 * nb + 1):                      # eager step, graph capture, replays; crosses two epoch boundaries
                e, s = divmod(g, nb)
                eu, ep, en = S.sample_batch(u[pos], i[pos], off, rated, catalog, B, 99, e, s)
                assert np.array_equal(tr.user_ids.cpu().numpy(), eu), g
                assert np.array_equal(tr.pos_ids.cpu().numpy(), ep), g
                assert np.array_equal(tr.neg_ids.cpu().numpy(), en), g
                tr.step()
            assert tr._graph is not None
        losses = []
        torch.manual_seed(0)
        model = R.TwoTowerModel(120, 90, 64, 128, dropout=0.1).cuda().train()
        tr = R.FusedBPRTrainer(model, lr=1e-2, item_extra_table=genres, seed=5)
        for e in range(3):
            losses.append(prod.train_epoch(tr, B, e, in_graph=in_graph))
        assert (tr._sampler is not None) == in_graph
        tr.check_ids()
        finals.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
    assert finals[0][0] == finals[1][0]
    for k in finals[0][1]:
        assert torch.equal(finals[0][1][k], finals[1][1][k]), k


def test_bitmap_and_csr_rejection_tests_agree():
    import recommendit_b200 as R
    from recommendit_b200 import producer as P
    u, i, r = _ratings(21, 200, 150, 12000)
    catalog = np.arange(1, 151, dtype=np.int64)
    a = R.DeviceBatchProducer(u, i, r, catalog, 200, seed=4)
    assert a.rated_bitmap is not None
    old, P._BITMAP_MAX_BYTES = P._BITMAP_MAX_BYTES, 0
    try:
        b = R.DeviceBatchProducer(u, i, r, catalog, 200, seed=4)
    finally:
        P._BITMAP_MAX_BYTES = old
    assert b.rated_bitmap is None
    B = 512
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    outs = []
    for prod in (a, b):
        o = [torch.empty(B, dtype=torch.int64, device="cuda") for _ in range(3)]
        res = []
        for g in (0, 3, prod.batches_per_epoch(B) + 1):
            cnt.fill_(g)
            prod.fill_from_counter(B, cnt, *o)
            res.append(torch.stack(o).clone())
        outs.append(torch.stack(res))
    assert torch.equal(outs[0], outs[1])


def test_rank_slices_equal_the_oracle_and_dp_trainer_runs_from_the_producer():
    """Data-parallel slicing of the sample stream (rank r of world takes the r-th B of every world·B step): bit-exact against
    the oracle through both kernel entry points; DataParallelBPRTrainer (world 1 here) trains from the in-graph producer
    exactly like FusedBPRTrainer."""
    import recommendit_b200 as R
    u, i, r = _ratings(31, 150, 120, 9000)
    catalog = np.arange(1, 121, dtype=np.int64)
    pos = r >= 4
    off, rated = S.build_rated_csr(u, i, 150)
    B, world = 64, 4
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for rank in range(world):
        prod = R.DeviceBatchProducer(u, i, r, catalog, 150, seed=8, rank=rank, world=world)
        nb = prod.batches_per_epoch(B)
        assert nb == int(pos.sum()) // (B * world)
        o = [torch.empty(B, dtype=torch.int64, device="cuda") for _ in range(3)]
        for epoch, step in ((0, 0), (2, nb - 1)):
            exp = S.sample_batch(u[pos], i[pos], off, rated, catalog, B, 8, epoch, step, rank=rank, world=world)
            prod.fill(*o, epoch, step)
            assert all(np.array_equal(a.cpu().numpy(), e) for a, e in zip(o, exp))
            cnt.fill_(epoch * nb + step)
            prod.fill_from_counter(B, cnt, *o)
            assert all(np.array_equal(a.cpu().numpy(), e) for a, e in zip(o, exp))
    genres = torch.zeros(121, 18)
    genres[torch.arange(121), torch.arange(121) % 18] = 1.0
    finals = []
    for cls in (R.FusedBPRTrainer, R.DataParallelBPRTrainer):
        prod = R.DeviceBatchProducer(u, i, r, catalog, 150, seed=8)
        torch.manual_seed(0)
        model = R.TwoTowerModel(150, 120, 64, 128, dropout=0.0).cuda().train()
        tr = cls(model, lr=1e-2, item_extra_table=genres, seed=5)
        losses = [prod.train_epoch(tr, 256, e) for e in range(2)]
        assert tr._sampler is not None
        finals.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
    assert np.allclose(finals[0][0], finals[1][0], atol=1e-6)
    for k in finals[0][1]:
        assert float((finals[0][1][k] - finals[1][1][k]).abs().max()) <= 1e-5, k
