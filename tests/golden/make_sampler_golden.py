"""Golden fixture for the batch producer's index build from the reference's own ``UserItemDataset``
(src/training/train_embeddings.py:23-63): which pairs are positives (in which order) and which (user, item) pairs count as "rated".
Run in the build container (needs /root/reference):  python tests/golden/make_sampler_golden.py"""
import sys
from pathlib import Path

import numpy as np
import pandas as pd

sys.path.insert(0, "/root/reference")
from src.training.train_embeddings import UserItemDataset  # noqa: E402

rng = np.random.default_rng(17)
n = 600
df = pd.DataFrame({"user_id": rng.integers(1, 31, n), "item_id": rng.integers(1, 51, n), "rating": rng.integers(1, 6, n).astype(float)})
df = pd.concat([df, df.iloc[:25]], ignore_index=True)            # repeated (user, item) rows: positives keep them, the rated SET does not
all_item_ids = sorted(rng.choice(np.arange(1, 61), 45, replace=False).tolist())
ds = UserItemDataset(df, {}, all_item_ids, n_negatives=4, min_rating=4.0)
rated = np.array(sorted((int(u), int(i)) for u, s in ds.user_rated.items() for i in s), dtype=np.int64)
np.random.seed(0)
neg_user = np.array([3, 7, 19], dtype=np.int64)
neg_draws = np.array([[ds._sample_negative(int(u)) for _ in range(400)] for u in neg_user], dtype=np.int64)
np.savez(Path(__file__).parent / "sampler.npz", user_id=df["user_id"].values, item_id=df["item_id"].values,
         rating=df["rating"].values, all_item_ids=np.array(all_item_ids, dtype=np.int64), pos_users=np.asarray(ds.user_ids, dtype=np.int64),
         pos_items=np.asarray(ds.item_ids, dtype=np.int64), rated_pairs=rated, neg_user=neg_user, neg_draws=neg_draws, n_samples=len(ds))
print("wrote sampler.npz:", len(ds), "positives,", len(rated), "rated pairs")
