"""The reference's callers run UNCHANGED on the drop-in (BASELINE north_star: "train_embeddings.py, build_index.py and
serving/recommender.py drop it in unchanged").

oracle/_ref holds a byte-for-byte copy of the reference's `src/` package (oracle/make_ref.py; git-ignored, shipped to the GPU
box).  On a synthetic MovieLens-1M-shaped data directory (ratings.dat / movies.dat / users.dat in the `::` format) the test runs

    EmbeddingTrainer(...).train() -> IndexBuilder(...).build() -> RecommendationPipeline retrieval

twice: on the drop-in (cuda) and — the training part, faiss being absent — on the stock PyTorch CPU modules, with the same seeds,
and compares the per-step loss trajectories, the trained user embeddings and (drop-in) the served candidates with exact search."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GENRES = ["Action", "Adventure", "Animation", "Children's", "Comedy", "Crime", "Documentary", "Drama", "Fantasy", "Film-Noir",
          "Horror", "Musical", "Mystery", "Romance", "Sci-Fi", "Thriller", "War", "Western"]


def write_movielens(d: Path, n_users=400, n_items=300, n_ratings=9000, seed=0):
    rng = np.random.default_rng(seed)
    catalog = np.sort(rng.choice(np.arange(1, n_items + 1), n_items - 20, replace=False))
    with open(d / "movies.dat", "w", encoding="latin-1") as f:
        for i in catalog:
            g = "|".join(sorted(set(rng.choice(GENRES, rng.integers(1, 4)))))
            f.write(f"{i}::Movie {i} (19{i % 100:02d})::{g}\n")
    with open(d / "users.dat", "w", encoding="latin-1") as f:
        for u in range(1, n_users + 1):
            f.write(f"{u}::{'MF'[u % 2]}::{[1, 18, 25, 35, 45, 50, 56][u % 7]}::{u % 21}::{10000 + u}\n")
    users = rng.integers(1, n_users + 1, n_ratings)
    items = catalog[(rng.integers(0, len(catalog) // 7, n_ratings) * 7 + users % 7) % len(catalog)]     # a learnable signal
    pairs = np.unique(np.stack([users, items], 1), axis=0)
    pairs = pairs[rng.permutation(len(pairs))]
    ratings = rng.choice([1, 2, 3, 4, 5], len(pairs), p=[0.056, 0.107, 0.261, 0.349, 0.227])
    pairs[0] = (n_users, catalog[-1]); ratings[0] = 5            # the maxima the trainer sizes its tables from
    with open(d / "ratings.dat", "w") as f:
        for (u, i), r in zip(pairs, ratings):
            f.write(f"{u}::{i}::{r}::{978300000 + int(u) * 7 + int(i)}\n")


def run_arm(arm, data_dir, out_dir):
    out = subprocess.run([sys.executable, str(ROOT / "tests" / "unchanged_callers_worker.py"), arm, str(data_dir), str(out_dir)],
                         capture_output=True, text=True, timeout=900, cwd="/tmp")
    assert out.returncode == 0, f"{arm}: " + out.stdout[-1500:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_reference_callers_run_unchanged_on_the_drop_in(tmp_path):
    if not (ROOT / "oracle" / "_ref" / "src" / "training" / "train_embeddings.py").exists():
        pytest.skip("oracle/_ref not built (python oracle/make_ref.py needs /root/reference)")
    data = tmp_path / "ml-1m"
    data.mkdir()
    write_movielens(data)
    mine = run_arm("dropin", data, tmp_path)
    stock = run_arm("stock", data, tmp_path)
    # ---- training: the same per-step trajectory as the stock PyTorch CPU modules under the same seeds (same init, same
    #      shuffles, same sampled negatives: both arms consume torch's and numpy's generators identically) ---------------- #
    a, b = np.array(mine["losses"]), np.array(stock["losses"])
    assert len(a) == len(b) and len(a) >= 20
    assert np.abs(a[:10] - b[:10]).max() <= 5e-5, np.abs(a[:10] - b[:10]).max()
    assert np.abs(a - b).max() <= 5e-4, np.abs(a - b).max()
    assert mine["n_users"] == stock["n_users"] == 400
    ue_a, ue_b = np.array(mine["user_emb"]), np.array(stock["user_emb"])
    assert np.abs(ue_a - ue_b).max() <= 5e-3                       # 2 epochs of Adam at lr 1e-2 on both sides
    # ---- the drop-in's checkpoint is a reference checkpoint: the stock class loads it and reproduces the embeddings ------- #
    sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
    try:
        import importlib
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        ref_tt = importlib.import_module("src.models.two_tower")
        import torch
        m = ref_tt.TwoTowerModel.load(str(tmp_path / "two_tower_dropin.pt"))
        for u, e in zip([1, 2, 3, 57, 400], ue_a):
            np.testing.assert_allclose(m.get_user_embedding(u, torch.device("cpu")), e, atol=2e-6)
    finally:
        sys.path.pop(0)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
    # … and the other way round: a checkpoint written by the stock module (NumPy integer scalars inside) loads into the drop-in
    import recommendit_b200 as R
    import torch
    m2 = R.TwoTowerModel.load(str(tmp_path / "two_tower_stock.pt"))
    for u, e in zip([1, 2, 3, 57, 400], ue_b):
        np.testing.assert_allclose(m2.get_user_embedding(u, torch.device("cpu")), e, atol=2e-6)
    # ---- build_index + serving: candidates equal exact search over the item embeddings (nprobe >= nlist here) ------------- #
    st = mine["index_stats"]
    assert st["n_vectors"] == 280 and st["embed_dim"] == 64 and st["n_lists"] == 280 // 39 and st["metric"] == "inner_product"
    assert mine["unknown_user_embedding_is_none"] is True           # IndexError inside -> None -> popularity fallback upstream
    emb, ids = np.load(tmp_path / "item_emb.npy"), np.load(tmp_path / "item_ids.npy")
    emb = emb / np.maximum(np.linalg.norm(emb, axis=1, keepdims=True), 1e-8)
    for u, q in zip([1, 2, 3, 57, 400], ue_a):
        got = mine["served"][str(u)]
        q = np.asarray(q, np.float32); q = q / max(np.linalg.norm(q), 1e-8)
        s = emb @ q
        order = np.argsort(-s, kind="stable")[:50]
        np.testing.assert_allclose(got["scores"], s[order], atol=2e-6)
        exact = set(ids[order].tolist())
        assert len(exact ^ set(got["ids"])) <= 2                    # ties / 1-ulp swaps at the cut only
        assert all(x >= y - 1e-7 for x, y in zip(got["scores"], got["scores"][1:]))
