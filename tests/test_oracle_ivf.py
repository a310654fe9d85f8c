"""oracle/ivf_oracle.py self-consistency (the FAISS side is "parity unpinned": no faiss here).
Anchors: the reference's own test properties (tests/test_models.py:155-246 of the reference) and
IVF(nprobe = nlist) == exhaustive search.  CPU-only."""
import numpy as np
import pytest

from oracle import ivf_oracle as V


def _fixture(n=500, d=32, nlist=10, seed=123):
    # same generator as the reference fixture (tests/test_models.py:157-161)
    np.random.seed(seed)
    x = np.random.randn(n, d).astype(np.float32)
    x = x / np.linalg.norm(x, axis=1, keepdims=True)
    xn = V.normalize_rows(x)
    c = V.spherical_kmeans(xn, nlist)
    a = V.assign(xn, c)
    off, order = V.build_lists(a, nlist)
    return xn, c, off, order


def test_kmeans_centroids_are_unit_and_populated():
    xn, c, off, order = _fixture()
    assert np.allclose(np.linalg.norm(c, axis=1), 1.0, atol=1e-5)
    assert off[-1] == 500 and sorted(order.tolist()) == list(range(500))
    assert (np.diff(off) > 0).all()


def test_full_probe_equals_flat():
    xn, c, off, order = _fixture()
    q = V.normalize_rows(np.random.default_rng(0).standard_normal((17, 32)).astype(np.float32))
    s1, i1 = V.ivf_search(q, c, off, order, xn, nprobe=10, k=20)
    s2, i2 = V.flat_search(q, xn, k=20)
    V.assert_topk_equivalent(s1, i1, s2, i2)


def test_reference_properties():
    xn, c, off, order = _fixture()
    # self retrieval in top-5 with nprobe=5 (tests/test_models.py:189-196)
    s, i = V.ivf_search(xn[42:43], c, off, order, xn, nprobe=5, k=5)
    assert 42 in i[0]
    # descending (tests/test_models.py:198-206)
    q = V.normalize_rows(np.random.default_rng(1).standard_normal((4, 32)).astype(np.float32))
    s, i = V.ivf_search(q, c, off, order, xn, nprobe=5, k=20)
    assert (np.diff(s, axis=1) <= 0).all()
    # k larger than what the probed lists hold → -1 padding, wrapper drops it (:118-123)
    s, i = V.ivf_search(q[:1], c, off, order, xn, nprobe=2, k=500)
    assert (i[0] == -1).any() and (s[0][i[0] == -1] == V.NEG_SENTINEL).all()
    item_ids = np.arange(1, 501, dtype=np.int64)
    d, ids = V.wrapper_search(s[0], i[0], item_ids)
    assert len(ids) == (i[0] >= 0).sum() and ids.min() >= 1
    _, mapped = V.wrapper_batch_search(s, i, item_ids)
    assert (mapped[i < 0] == -1).all()


@pytest.mark.parametrize("nprobe,k", [(3, 10), (10, 500), (1, 7)])
def test_c_restatement_matches_numpy(nprobe, k):
    xn, c, off, order = _fixture(n=2000, d=64, nlist=16, seed=5)
    q = V.normalize_rows(np.random.default_rng(2).standard_normal((33, 64)).astype(np.float32))
    s1, i1 = V.ivf_search(q, c, off, order, xn, nprobe=nprobe, k=k)
    s2, i2 = V.ivf_search_c(q, c, off, xn[order], order, nprobe, k, threads=2)
    V.assert_topk_equivalent(s2, i2, s1, i1)
    f1 = V.flat_search(q, xn, k)
    f2 = V.flat_search_c(q, xn, k, threads=2)
    V.assert_topk_equivalent(f2[0], f2[1], f1[0], f1[1])


def test_exact_ties_are_ordered_by_scan_position():
    x = np.zeros((8, 4), np.float32); x[:, 0] = 1.0           # eight identical vectors
    c = np.array([[1, 0, 0, 0]], np.float32)
    off, order = V.build_lists(V.assign(x, c), 1)
    s, i = V.ivf_search(x[:1], c, off, order, x, nprobe=1, k=5)
    assert i[0].tolist() == [0, 1, 2, 3, 4]
    s2, i2 = V.ivf_search_c(x[:1], c, off, x[order], order, 1, 5, threads=1)
    assert i2[0].tolist() == [0, 1, 2, 3, 4]
