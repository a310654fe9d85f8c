"""End-to-end mini pipeline on the GPU, written the way the reference's callers use the two classes
(train_embeddings.py:153-220 → build_index.py:86-138 → serving/recommender.py:203,311-313), on a synthetic
MovieLens-shaped problem.  Checks the trajectory against the CPU arm (oracle/torch_step.py) and retrieval against exact
search."""
import numpy as np
import pytest
import torch

from oracle import ivf_oracle as V
from oracle import torch_step as TS

pytestmark = pytest.mark.gpu
N_USERS, N_ITEMS, D, H, B = 900, 700, 64, 128, 256


def _dataset(seed=0):
    rng = np.random.default_rng(seed)
    genres = (rng.random((N_ITEMS + 1, 18)) < 0.12).astype(np.float32)
    # users like items whose id is congruent to theirs mod 7 (a learnable signal)
    users = rng.integers(1, N_USERS + 1, 6000)
    pos = ((rng.integers(0, N_ITEMS // 7, 6000) * 7 + users % 7) % N_ITEMS) + 1
    neg = rng.integers(1, N_ITEMS + 1, 6000)
    return users, pos, neg, genres


def test_training_loop_build_index_and_serve(tmp_path):
    import recommendit_b200 as R
    dev = torch.device("cuda")
    users, pos, neg, genres = _dataset()
    torch.manual_seed(0)
    model = R.TwoTowerModel(n_users=N_USERS, n_items=N_ITEMS, embed_dim=D, hidden_dim=H, dropout=0.0).to(dev)
    init = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
    # ---- train_embeddings.py:160-197 --------------------------------------------------------------------------- #
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=1e-5)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=2)
    T = TS.make_params(init)
    opt_ref = TS.make_optimizer(T, lr=1e-2)
    sch_ref = torch.optim.lr_scheduler.CosineAnnealingLR(opt_ref, T_max=2)
    losses, ref_losses = [], []
    for epoch in range(2):
        model.train()
        for s in range(0, 6000 - B + 1, B):
            sl = slice(s, s + B)
            batch = (torch.from_numpy(users[sl]), torch.from_numpy(pos[sl]), torch.from_numpy(genres[pos[sl]]),
                     torch.from_numpy(neg[sl]), torch.from_numpy(genres[neg[sl]]))
            user_ids, pos_ids, pos_genres, neg_ids, neg_genres = [b.to(dev) for b in batch]
            user_emb = model.user_tower(user_ids)
            pos_item_emb = model.item_tower(pos_ids, pos_genres)
            neg_item_emb = model.item_tower(neg_ids, neg_genres)
            loss = model.bpr_loss(user_emb, pos_item_emb, neg_item_emb)
            optimizer.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
            optimizer.step()
            losses.append(loss.item())
            ref_losses.append(TS.step(T, opt_ref, batch, dropout=0.0))
        scheduler.step(); sch_ref.step()
    # same trajectory as the reference's PyTorch CPU path, step by step (46 optimiser steps)
    assert np.allclose(losses, ref_losses, atol=5e-4), np.abs(np.array(losses) - np.array(ref_losses)).max()
    assert np.allclose(losses[:10], ref_losses[:10], atol=5e-5)
    assert np.mean(losses[-5:]) < np.mean(losses[:5])                      # and it learns
    for k, v in model.state_dict().items():                                # parameters track the CPU arm too
        assert np.abs(v.detach().cpu().numpy() - T[k].detach().numpy()).max() < 0.05, k

    # ---- train_embeddings.py:213-220, build_index.py:86-138 ---------------------------------------------------- #
    item_ids = list(range(1, N_ITEMS + 1))
    model.precompute_item_embeddings(item_ids, genres[1:], dev)
    path = tmp_path / "two_tower.pt"
    model.save(str(path))
    loaded = R.TwoTowerModel.load(str(path), device=dev)
    loaded.eval()
    item_embeddings = loaded.get_item_embeddings(item_ids, genres[1:], device=dev, batch_size=512)
    assert item_embeddings.shape == (N_ITEMS, D) and item_embeddings.dtype == np.float32
    assert np.allclose(np.linalg.norm(item_embeddings, axis=1), 1.0, atol=1e-5)
    n_lists = 100
    effective = n_lists if N_ITEMS >= 39 * n_lists else max(1, N_ITEMS // 39)          # build_index.py:119-121
    index = R.FAISSIndex(embed_dim=D, n_lists=effective, n_probe=10)
    index.build_ivf_index(item_embeddings, item_ids)
    index.save(str(tmp_path / "faiss.index"))
    index = R.FAISSIndex.load(str(tmp_path / "faiss.index"))
    assert index.stats()["n_vectors"] == N_ITEMS and index.stats()["n_lists"] == effective

    # ---- recommender.py:203, 311-313 ------------------------------------------------------------------------------ #
    xn = V.normalize_rows(item_embeddings)
    hits = 0
    for uid in (1, 17, 250, 899):
        user_emb = loaded.get_user_embedding(uid, dev)
        distances, candidate_item_ids = index.search(user_emb, k=50)
        assert len(candidate_item_ids) == 50 and (np.diff(distances) <= 1e-6).all()
        exact_s, exact_i = V.flat_search(V.normalize_rows(user_emb[None]), xn, 50)
        assert distances[0] <= exact_s[0, 0] + 2e-6                          # IVF (10 of 17 lists) never beats exact search
        hits += len(set(candidate_item_ids.tolist()) & set((exact_i[0] + 1).tolist()))
    assert hits / 200 > 0.8                                                # recall@50 of IVF (10 of 17 lists) vs exact
    index.set_n_probe(effective)                                           # probing every list = exact search
    for uid in (3, 444):
        user_emb = loaded.get_user_embedding(uid, dev)
        d, ids = index.search(user_emb, k=50)
        s_ref, i_ref = V.flat_search(V.normalize_rows(user_emb[None]), xn, 50)
        V.assert_topk_equivalent(d[None], ids[None], s_ref, i_ref + 1)


@pytest.mark.parametrize("d,nlist,nprobe", [(64, 16, 4), (32, 8, 8)])
def test_graph_captured_request_equals_the_two_drop_in_calls(d, nlist, nprobe):
    """serving micro-path: UserRecommender.recommend(u) — one CUDA-graph replay — must return exactly what
    `index.search(model.get_user_embedding(u), k)` returns (recommender.py:148-156, 203), for k above and below the candidates."""
    import recommendit_b200 as R
    from recommendit_b200 import RB200Error
    dev = torch.device("cuda")
    rng = np.random.default_rng(3)
    torch.manual_seed(1)
    model = R.TwoTowerModel(n_users=300, n_items=500, embed_dim=d, hidden_dim=128, dropout=0.1).to(dev)
    genres = (rng.random((500, 18)) < 0.15).astype(np.float32)
    item_ids = [int(i) for i in rng.permutation(np.arange(1, 501))]
    emb = model.get_item_embeddings(item_ids, genres, dev)
    index = R.FAISSIndex(embed_dim=d, n_lists=nlist, n_probe=nprobe)
    index.build_ivf_index(emb, item_ids)
    for k in (50, 500):
        rec = R.UserRecommender(model, index, k=k)
        for u in (1, 17, 300, 17, 0):
            s_ref, i_ref = index.search(model.get_user_embedding(u, dev), k)
            s, i = rec.recommend(u)
            assert np.array_equal(i, i_ref) and np.array_equal(s, s_ref), (k, u)
            assert i.dtype == np.int64 and s.dtype == np.float32 and np.all(np.diff(s) <= 0)
        assert rec._graph is not None
    # the model's weights are read in place: an update is seen by the captured request
    with torch.no_grad():
        model.user_tower.embedding.weight[17].add_(0.5)
    s_ref, i_ref = index.search(model.get_user_embedding(17, dev), 500)
    s, i = rec.recommend(17)
    assert np.array_equal(i, i_ref) and np.array_equal(s, s_ref)
    # a rebuilt index (new device arrays behind the same FAISSIndex) and a moved parameter block are noticed: the request is captured
    # again instead of replaying addresses that no longer hold the data
    old_ptr = index.index.list_vecs.data_ptr()
    keep_alive = index.index                                   # (keeps the old arrays allocated so that the new ones get new addresses)
    index.build_ivf_index(emb[::-1].copy(), item_ids[::-1])
    assert index.index.list_vecs.data_ptr() != old_ptr
    model.user_tower.embedding.weight.data = model.user_tower.embedding.weight.data.clone()
    for u in (5, 17):
        s_ref, i_ref = index.search(model.get_user_embedding(u, dev), 500)
        s, i = rec.recommend(u)
        assert np.array_equal(i, i_ref) and np.array_equal(s, s_ref), u
    del keep_alive
    index.set_n_probe(max(1, nprobe // 2))
    if nprobe // 2 >= 1 and nprobe // 2 != nprobe:
        with pytest.raises(RB200Error, match="n_probe changed"):
            rec.recommend(1)
