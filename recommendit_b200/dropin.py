"""Make the unmodified reference callers use this package.

The reference imports its models as ``from src.models.two_tower import TwoTowerModel`` and
``from src.models.faiss_index import FAISSIndex`` (``src/training/train_embeddings.py:18``,
``src/training/build_index.py:15-16``, ``src/serving/recommender.py:18-20``, ``src/pipelines/run_pipeline.py``).
``install()`` registers this package's modules under those two names *before* the reference package is imported,
so every ``import`` in the reference tree resolves to the B200 implementation; nothing in the reference is edited.

    import recommendit_b200.dropin as dropin
    dropin.install()                      # then: from src.training.train_embeddings import EmbeddingTrainer

The reference picks ``device='cuda'`` on its own when a GPU is present (``train_embeddings.py:102-109``).
"""
import sys


def install() -> None:
    from . import faiss_index, two_tower
    for name in ("src.models.two_tower", "src.models.faiss_index"):
        if name in sys.modules and getattr(sys.modules[name], "__name__", "").startswith("src."):
            raise RuntimeError(f"{name} is already imported from the reference tree; call dropin.install() first")
    sys.modules["src.models.two_tower"] = two_tower
    sys.modules["src.models.faiss_index"] = faiss_index
    # `import src.models` executes the reference's own src/models/__init__.py, whose relative imports
    # (`from .two_tower import TwoTowerModel`) now hit the entries above.


def uninstall() -> None:
    for name in ("src.models.two_tower", "src.models.faiss_index"):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__name__", "").startswith("recommendit_b200"):
            del sys.modules[name]
