"""recommendit_b200 — the two-tower / IVF hot path of sarihammad/recommendit on B200 (sm_100a).

Python mirrors of the reference's model classes on top of ``librb200.so`` (C ABI in ``include/rb200.h``).
"""
from .two_tower import N_GENRES, ItemTower, TwoTowerModel, UserTower          # noqa: F401
from .faiss_index import FAISSIndex, flat_search, scores_nt, topk_merge        # noqa: F401
from .trainer import DataParallelBPRTrainer, FusedBPRTrainer                  # noqa: F401
from .producer import DeviceBatchProducer                                     # noqa: F401
from .serving import MicroBatcher, UserRecommender                                          # noqa: F401
from ._lib import RB200Error                                                  # noqa: F401

__all__ = ["TwoTowerModel", "UserTower", "ItemTower", "FAISSIndex", "FusedBPRTrainer", "DataParallelBPRTrainer", "N_GENRES", "RB200Error",
           "flat_search", "topk_merge", "DeviceBatchProducer", "UserRecommender", "MicroBatcher"]
