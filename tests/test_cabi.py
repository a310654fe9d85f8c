"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/rb200.h declares, its
structs match the ctypes mirrors, and the product refuses to compute without CUDA (no CPU fallback)."""
import ctypes
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "rb200.h"


@pytest.fixture(scope="module")
def lib():
    from recommendit_b200 import _lib
    if not _lib.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(rb200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_functions():
    names = declared_functions()
    assert len(names) >= 30 and "rb200_tower_fwd" in names and "rb200_ivf_search_run" in names


def test_library_exports_every_declared_symbol(lib):
    from recommendit_b200 import _lib
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    unbound = [n for n in declared_functions() if n not in _lib.SIGNATURES]
    assert not unbound, f"declared in rb200.h but not bound in _lib.SIGNATURES: {unbound}"
    extra = [n for n in _lib.SIGNATURES if n not in declared_functions()]
    assert not extra, f"bound but not declared: {extra}"


def test_symbols_are_plain_c(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", str(ROOT / "recommendit_b200" / "librb200.so")],
                         capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (rb200_\w+)", out))
    assert set(declared_functions()) <= exported


def test_struct_sizes_match(lib):
    from recommendit_b200 import _lib
    for which, cls in enumerate((_lib.TowerJob, _lib.TowerBwdJob, _lib.OptState, _lib.StepParams, _lib.StepViews, _lib.SumsqSeg)):
        assert lib.rb200_sizeof(which) == ctypes.sizeof(cls), cls.__name__
    assert lib.rb200_version() == 100


def test_argument_errors_are_reported_not_fatal(lib):
    rc = lib.rb200_bpr_pair(None, None, None, 4, 8, None, None, None, None, 1.0, None, 0, None)
    assert rc == -1 and b"bpr_pair" in lib.rb200_last_error()
    rc = lib.rb200_normalize_rows(None, 3, 4, 1e-8, None, None)
    assert rc == -1


def test_no_cpu_fallback():
    import recommendit_b200 as R
    model = R.TwoTowerModel(10, 10, 32, 64)
    with pytest.raises(R.RB200Error):
        model.user_tower(torch.tensor([1, 2]))
    with pytest.raises(R.RB200Error):
        model.bpr_loss(torch.randn(4, 32), torch.randn(4, 32), torch.randn(4, 32))
    with pytest.raises(R.RB200Error):
        model.in_batch_bpr_loss(torch.randn(4, 32), torch.randn(4, 32))
    if not torch.cuda.is_available():
        with pytest.raises(R.RB200Error):
            R.FAISSIndex(32, 4, 2).build_ivf_index(np.zeros((8, 32), np.float32), list(range(8)))
        with pytest.raises(R.RB200Error):
            R.FusedBPRTrainer(model)
        with pytest.raises(R.RB200Error):                      # the device-side batch producer has no host form either
            R.DeviceBatchProducer([1, 2], [3, 4], [5.0, 4.0], [3, 4, 5], 2)
        with pytest.raises(R.RB200Error):
            R.flat_search(torch.randn(2, 64), torch.randn(100, 64), 5)


def test_product_does_not_import_the_oracle():
    """Nothing under recommendit_b200/ may import or execute oracle/."""
    for f in (ROOT / "recommendit_b200").rglob("*.py"):
        txt = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
        assert "ivf_oracle" not in txt and "two_tower_oracle" not in txt, f
    for f in (ROOT / "recommendit_b200" / "csrc").glob("*.cu*"):
        assert "oracle" not in f.read_text().lower() or f.name == "README", f


def test_state_dict_surface_matches_reference_keys():
    import recommendit_b200 as R
    from oracle.two_tower_oracle import PARAM_KEYS
    m = R.TwoTowerModel(100, 200, embed_dim=32, hidden_dim=64)
    assert tuple(m.state_dict().keys()) == PARAM_KEYS
    assert [tuple(p.shape) for p in m.parameters()] == [(101, 32), (64, 32), (64,), (32, 64), (32,),
                                                         (201, 32), (64, 50), (64,), (32, 64), (32,)]
    assert (m.n_users, m.n_items, m.embed_dim) == (100, 200, 32)
    # padding row is initialised like every other row (reference runs xavier over the whole table)
    assert float(m.user_tower.embedding.weight[0].abs().sum()) > 0


def test_checkpoint_roundtrip_on_cpu(tmp_path):
    import recommendit_b200 as R
    m = R.TwoTowerModel(30, 40, embed_dim=32, hidden_dim=64)     # hidden 64: the case the reference's own load breaks on
    m._item_id_to_idx = {5: 0}
    p = tmp_path / "tt.pt"
    m.save(str(p))
    ck = torch.load(p, weights_only=False)
    assert {"state_dict", "n_users", "n_items", "embed_dim", "item_id_to_idx", "idx_to_item_id"} <= set(ck)
    m2 = R.TwoTowerModel.load(str(p))
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    assert m2._item_id_to_idx == {5: 0} and not m2.training


def test_faiss_index_host_surface(tmp_path):
    import recommendit_b200 as R
    idx = R.FAISSIndex(embed_dim=32, n_lists=10, n_probe=5)
    assert idx.stats() == {"status": "not built"}
    with pytest.raises(RuntimeError):
        idx.search(np.zeros(32, np.float32))
    with pytest.raises(RuntimeError):
        idx.batch_search(np.zeros((2, 32), np.float32))
    with pytest.raises(FileNotFoundError):
        R.FAISSIndex.load(str(tmp_path / "missing.index"))
    with pytest.raises(AssertionError):
        idx.build_ivf_index(np.zeros((4, 32), np.float64), [1, 2, 3, 4])
    with pytest.raises(AssertionError):
        idx.build_ivf_index(np.zeros((4, 16), np.float32), [1, 2, 3, 4])
    idx.set_n_probe(7)
    assert idx.n_probe == 7


@pytest.mark.skipif(not Path("/root/reference/src").exists(), reason="reference tree only exists in the build container")
def test_dropin_shim_rebinds_reference_callers():
    """With dropin.install() the UNMODIFIED reference callers import this package's classes."""
    code = r'''
import sys
sys.path.insert(0, "%s"); sys.path.insert(1, "/root/reference")
import recommendit_b200.dropin as d
d.install()
import recommendit_b200 as R
from src.training.train_embeddings import EmbeddingTrainer, TwoTowerModel as T1
from src.training.build_index import IndexBuilder, TwoTowerModel as T2, FAISSIndex as F2
import src.models as M
assert T1 is R.TwoTowerModel and T2 is R.TwoTowerModel and F2 is R.FAISSIndex
assert M.TwoTowerModel is R.TwoTowerModel and M.FAISSIndex is R.FAISSIndex
print("ok")
''' % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
