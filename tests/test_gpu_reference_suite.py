"""The reference's own acceptance tests for the hot path (tests/test_models.py::TestTwoTowerModel and
::TestFAISSIndex, lines 29-246 of the reference), restated against the drop-in classes on a CUDA device.
Same fixtures, sizes, seeds and assertions; the only change is `.cuda()` (the drop-in has no CPU path)."""
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


class TestTwoTowerModel:
    N_USERS, N_ITEMS, EMBED_DIM, BATCH = 100, 200, 32, 16

    @pytest.fixture
    def model(self):
        from recommendit_b200 import TwoTowerModel
        return TwoTowerModel(n_users=self.N_USERS, n_items=self.N_ITEMS, embed_dim=self.EMBED_DIM, hidden_dim=64).to(DEV)

    def _ids(self, n):
        return torch.randint(1, n + 1, (self.BATCH,), device=DEV)

    def test_user_tower_output_shape(self, model):
        assert model.user_tower(self._ids(self.N_USERS)).shape == (self.BATCH, self.EMBED_DIM)

    def test_item_tower_output_shape(self, model):
        from recommendit_b200 import N_GENRES
        emb = model.item_tower(self._ids(self.N_ITEMS), torch.rand(self.BATCH, N_GENRES, device=DEV))
        assert emb.shape == (self.BATCH, self.EMBED_DIM)

    def test_user_embeddings_l2_normalized(self, model):
        norms = torch.norm(model.user_tower(self._ids(self.N_USERS)), p=2, dim=-1)
        assert torch.allclose(norms, torch.ones(self.BATCH, device=DEV), atol=1e-5)

    def test_item_embeddings_l2_normalized(self, model):
        emb = model.item_tower(self._ids(self.N_ITEMS), torch.rand(self.BATCH, 18, device=DEV))
        assert torch.allclose(torch.norm(emb, p=2, dim=-1), torch.ones(self.BATCH, device=DEV), atol=1e-5)

    def test_forward_returns_tuple(self, model):
        result = model(self._ids(self.N_USERS), self._ids(self.N_ITEMS), torch.rand(self.BATCH, 18, device=DEV))
        assert isinstance(result, tuple) and len(result) == 2
        assert result[0].shape == (self.BATCH, self.EMBED_DIM) and result[1].shape == (self.BATCH, self.EMBED_DIM)

    def test_bpr_loss_positive(self, model):
        F = torch.nn.functional
        user_emb = F.normalize(torch.randn(self.BATCH, self.EMBED_DIM, device=DEV), p=2, dim=-1)
        pos_emb = F.normalize(user_emb + torch.randn_like(user_emb) * 0.1, p=2, dim=-1)
        neg_emb = F.normalize(-user_emb + torch.randn_like(user_emb) * 0.1, p=2, dim=-1)
        loss = model.bpr_loss(user_emb, pos_emb, neg_emb)
        assert isinstance(loss.item(), float) and loss.item() >= 0.0

    def test_bpr_loss_decreases_with_training(self, model):
        optimizer = torch.optim.Adam(model.parameters(), lr=0.01)
        losses = []
        for _ in range(20):
            genres = torch.rand(self.BATCH, 18, device=DEV)
            user_emb = model.user_tower(self._ids(self.N_USERS))
            pos_emb = model.item_tower(self._ids(self.N_ITEMS), genres)
            neg_emb = model.item_tower(self._ids(self.N_ITEMS), genres)
            loss = model.bpr_loss(user_emb, pos_emb, neg_emb)
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            losses.append(loss.item())
        assert losses[-1] < losses[0] * 2.0

    def test_get_user_embedding(self, model):
        emb = model.get_user_embedding(user_id=1, device=torch.device(DEV))
        assert isinstance(emb, np.ndarray) and emb.shape == (self.EMBED_DIM,)
        assert abs(np.linalg.norm(emb) - 1.0) < 1e-4

    def test_get_item_embeddings(self, model):
        genres = np.random.rand(20, 18).astype(np.float32)
        embs = model.get_item_embeddings(list(range(1, 21)), genres, device=torch.device(DEV))
        assert embs.shape == (20, self.EMBED_DIM)
        assert np.allclose(np.linalg.norm(embs, axis=1), 1.0, atol=1e-4)

    def test_save_and_load(self, model):
        """Passes here although it FAILS on the unmodified reference (hidden_dim is not checkpointed there)."""
        from recommendit_b200 import TwoTowerModel
        with tempfile.TemporaryDirectory() as tmpdir:
            path = str(Path(tmpdir) / "two_tower.pt")
            model.save(path)
            loaded = TwoTowerModel.load(path, device=torch.device(DEV))
            assert (loaded.n_users, loaded.n_items, loaded.embed_dim) == (model.n_users, model.n_items, model.embed_dim)
            user_ids = torch.tensor([1, 2, 3], device=DEV)
            model.eval()
            with torch.no_grad():
                assert np.allclose(model.user_tower(user_ids).cpu().numpy(), loaded.user_tower(user_ids).cpu().numpy(), atol=1e-5)

    def test_inference_helpers_match_reference_golden(self, golden):
        from recommendit_b200 import TwoTowerModel
        from oracle.two_tower_oracle import PARAM_KEYS
        g = golden("inference")
        m = TwoTowerModel(100, 200, 32, 64)
        m.load_state_dict({k: torch.from_numpy(g["init/" + k]) for k in PARAM_KEYS})
        m.to(DEV)
        embs = m.get_item_embeddings(g["item_ids"].tolist(), g["genres"], device=torch.device(DEV), batch_size=8)
        np.testing.assert_allclose(embs, g["item_embs"], atol=2e-6, rtol=0)
        for uid in (1, 100):
            np.testing.assert_allclose(m.get_user_embedding(uid, torch.device(DEV)), g[f"user_emb_{uid}"], atol=2e-6, rtol=0)
        assert not m.training


class TestFAISSIndex:
    EMBED_DIM, N_ITEMS = 32, 500

    @pytest.fixture
    def built_index(self):
        from recommendit_b200 import FAISSIndex
        np.random.seed(123)
        embeddings = np.random.randn(self.N_ITEMS, self.EMBED_DIM).astype(np.float32)
        embeddings = embeddings / np.linalg.norm(embeddings, axis=1, keepdims=True)
        item_ids = list(range(1, self.N_ITEMS + 1))
        index = FAISSIndex(embed_dim=self.EMBED_DIM, n_lists=10, n_probe=5)
        index.build_ivf_index(embeddings, item_ids)
        return index, embeddings, item_ids

    def test_index_built(self, built_index):
        index, _, _ = built_index
        assert index.index is not None and index.index.ntotal == self.N_ITEMS

    def test_search_returns_k_results(self, built_index):
        index, _, _ = built_index
        query = np.random.randn(self.EMBED_DIM).astype(np.float32)
        query /= np.linalg.norm(query)
        distances, retrieved_ids = index.search(query, k=20)
        assert len(distances) == 20 and len(retrieved_ids) == 20

    def test_search_result_type(self, built_index):
        index, _, _ = built_index
        distances, retrieved_ids = index.search(np.random.randn(self.EMBED_DIM).astype(np.float32), k=10)
        assert distances.dtype in [np.float32, np.float64]
        assert all(isinstance(i, (int, np.integer)) for i in retrieved_ids)

    def test_nearest_neighbor_is_self(self, built_index):
        index, embeddings, item_ids = built_index
        distances, retrieved_ids = index.search(embeddings[42].copy(), k=5)
        assert item_ids[42] in retrieved_ids

    def test_distances_descending(self, built_index):
        index, _, _ = built_index
        query = np.random.randn(self.EMBED_DIM).astype(np.float32)
        distances, _ = index.search(query / np.linalg.norm(query), k=20)
        assert (np.diff(distances) <= 0.01).all()

    def test_search_k_capped_at_n_items(self, built_index):
        index, _, _ = built_index
        distances, retrieved_ids = index.search(np.random.randn(self.EMBED_DIM).astype(np.float32), k=10000)
        assert len(retrieved_ids) <= self.N_ITEMS and len(retrieved_ids) == len(distances)

    def test_save_and_load(self, built_index):
        from recommendit_b200 import FAISSIndex
        index, embeddings, _ = built_index
        with tempfile.TemporaryDirectory() as tmpdir:
            path = str(Path(tmpdir) / "faiss.index")
            index.save(path)
            assert Path(path).exists() and Path(path).with_suffix(".meta.pkl").exists()
            loaded = FAISSIndex.load(path)
            assert loaded.index.ntotal == self.N_ITEMS and loaded.embed_dim == self.EMBED_DIM
            d1, ids1 = index.search(embeddings[0].copy(), k=10)
            d2, ids2 = loaded.search(embeddings[0].copy(), k=10)
            assert list(ids1) == list(ids2)

    def test_stats(self, built_index):
        stats = built_index[0].stats()
        assert stats["n_vectors"] == self.N_ITEMS and stats["embed_dim"] == self.EMBED_DIM
        assert stats["metric"] == "inner_product" and stats["n_item_ids"] == self.N_ITEMS

    def test_unnormalized_query_handled(self, built_index):
        index, _, _ = built_index
        unnorm = np.random.randn(self.EMBED_DIM).astype(np.float32) * 100
        d1, ids1 = index.search(unnorm, k=5)
        d2, ids2 = index.search(unnorm / np.linalg.norm(unnorm), k=5)
        assert list(ids1) == list(ids2)

    def test_set_n_probe_changes_recall(self, built_index):
        index, embeddings, _ = built_index
        index.set_n_probe(10)
        assert index.index.nprobe == 10
        d_full, ids_full = index.batch_search(embeddings[:8], k=50)
        index.set_n_probe(1)
        d_one, ids_one = index.batch_search(embeddings[:8], k=50)
        assert (d_full[:, 0] >= d_one[:, 0] - 1e-6).all() and d_full.shape == (8, 50) and ids_one.dtype == np.int64
