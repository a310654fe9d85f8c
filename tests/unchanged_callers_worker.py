"""Worker of tests/test_gpu_unchanged_callers.py: runs the reference's OWN callers, byte-identical copies under oracle/_ref
(see oracle/make_ref.py), either on the drop-in (`dropin`: recommendit_b200 registered as src.models.*, device cuda) or on the
stock CPU modules (`stock`).  usage: unchanged_callers_worker.py {dropin|stock} DATA_DIR OUT_DIR

    EmbeddingTrainer(...).train()                       src/training/train_embeddings.py:131-223
    IndexBuilder(...).build()                           src/training/build_index.py:67-140          (dropin only: no faiss here)
    RecommendationPipeline._load_model/_load_index/_get_user_embedding + faiss_index.search
                                                        src/serving/recommender.py:148-156, 200-207, 311-313
Nothing in the callers is edited.  Two things are set from OUTSIDE them: the name `TwoTowerModel` in train_embeddings' namespace
is bound to the same class with dropout=0.0 (the caller does not expose the dropout rate, and the two arms draw their masks from
different generators), and `bpr_loss` is wrapped to record the per-step loss (the caller only logs epoch means to 4 digits).
"""
import functools
import json
import logging
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
arm, data_dir, out_dir = sys.argv[1], sys.argv[2], Path(sys.argv[3])
sys.path.insert(0, str(ROOT))
sys.path.insert(1, str(ROOT / "oracle" / "_ref"))
logging.disable(logging.CRITICAL)
if arm == "dropin":
    import recommendit_b200.dropin as dropin
    dropin.install()
import src.training.train_embeddings as TE            # noqa: E402  (the reference's module, unmodified)
from src.models.two_tower import TwoTowerModel          # noqa: E402

if arm == "dropin":
    import recommendit_b200 as R
    assert TwoTowerModel is R.TwoTowerModel and TE.TwoTowerModel is R.TwoTowerModel
else:
    assert TwoTowerModel.__module__ == "src.models.two_tower" and "oracle/_ref" in sys.modules["src.models.two_tower"].__file__

LOSSES = []
_orig_loss = TwoTowerModel.bpr_loss


def _recording_loss(self, *a):
    loss = _orig_loss(self, *a)
    LOSSES.append(loss)
    return loss


TwoTowerModel.bpr_loss = _recording_loss
TE.TwoTowerModel = functools.partial(TwoTowerModel, dropout=0.0)

torch.manual_seed(1234)
np.random.seed(1234)
device = "cuda" if arm == "dropin" else "cpu"
model_path = out_dir / f"two_tower_{arm}.pt"
trainer = TE.EmbeddingTrainer(data_dir=data_dir, model_output_path=str(model_path), embed_dim=64, epochs=2, batch_size=256,
                              learning_rate=1e-2, device=device)
model = trainer.train()
result = {"arm": arm, "losses": [float(x) for x in LOSSES], "n_users": int(model.n_users), "n_items": int(model.n_items)}
probe_users = [1, 2, 3, 57, int(model.n_users)]
result["user_emb"] = [model.get_user_embedding(u, torch.device("cpu")).tolist() for u in probe_users]

if arm == "dropin":
    from src.training.build_index import IndexBuilder
    index_path = out_dir / "faiss.index"
    ib = IndexBuilder(model_path=str(model_path), data_dir=data_dir, index_output_path=str(index_path), embed_dim=64, n_lists=100,
                      n_probe=10)
    index = ib.build()
    result["index_stats"] = index.stats()
    from src.serving.recommender import RecommendationPipeline
    pipe = RecommendationPipeline(model_path=str(model_path), index_path=str(index_path), data_dir=data_dir, top_k_candidates=50,
                                  device="cpu")                      # run_pipeline.py / app.py pass cpu here
    pipe._load_model()
    pipe._load_index()
    served = {}
    for u in probe_users:
        emb = pipe._get_user_embedding(u)
        scores, ids = pipe.faiss_index.search(emb, k=pipe.top_k_candidates)
        served[str(u)] = {"scores": scores.tolist(), "ids": ids.tolist()}
    result["served"] = served
    result["unknown_user_embedding_is_none"] = pipe._get_user_embedding(int(model.n_users) + 1000) is None
    # exact retrieval over the same item embeddings (the index probes every list here: n_lists is cut to n_items // 39)
    import pandas as pd
    movies = pd.read_csv(Path(data_dir) / "movies.dat", sep="::", names=["item_id", "title", "genres"], engine="python", encoding="latin-1")
    item_ids, genre_matrix = ib._build_genre_matrix(movies)
    emb = pipe.model.get_item_embeddings(item_ids, genre_matrix, torch.device("cpu"))
    np.save(out_dir / "item_emb.npy", emb)
    np.save(out_dir / "item_ids.npy", np.asarray(item_ids))
print("RESULT " + json.dumps(result))
