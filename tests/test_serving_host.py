"""Host logic of the request micro-batcher (recommendit_b200/serving.py::MicroBatcher) with a stand-in search function — CPU."""
import threading
import time

import numpy as np
import pytest

from recommendit_b200.serving import MicroBatcher


def _fake_search(calls):
    def fn(q, k):
        calls.append(q.shape[0])
        time.sleep(0.002)                                   # a "kernel" long enough for requests to pile up
        scores = np.tile(np.arange(k, 0, -1, dtype=np.float32), (q.shape[0], 1)) + q[:, :1]
        ids = np.tile(np.arange(k, dtype=np.int64), (q.shape[0], 1)) + (q[:, :1].astype(np.int64) * 1000)
        ids[:, k - 2:] = -1                                 # padding, dropped for the caller like FAISSIndex.search does
        return scores, ids
    return fn


def test_concurrent_requests_are_grouped_and_each_caller_gets_its_own_row():
    calls = []
    mb = MicroBatcher(_fake_search(calls), k=10, max_batch=16, max_wait_ms=20.0)
    out = {}

    def worker(u):
        out[u] = mb.search(np.full(4, float(u), np.float32))
    threads = [threading.Thread(target=worker, args=(u,)) for u in range(1, 41)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=10)
    mb.close()
    assert sorted(out) == list(range(1, 41))
    for u, (s, i) in out.items():
        assert len(i) == 8 and (i == np.arange(8) + u * 1000).all() and s[0] == 10 + u      # its own row, padding dropped
    assert sum(calls) == 40 and max(calls) <= 16 and len(calls) < 40                         # really batched
    assert mb.requests == 40 and mb.batches == len(calls)


def test_single_request_is_served_after_the_wait_and_errors_reach_the_caller():
    calls = []
    mb = MicroBatcher(_fake_search(calls), k=5, max_batch=8, max_wait_ms=1.0)
    t0 = time.perf_counter()
    s, i = mb.search(np.ones(3, np.float32))
    assert time.perf_counter() - t0 < 0.5 and calls == [1] and len(i) == 3
    mb.close()
    with pytest.raises(RuntimeError):
        mb.search(np.ones(3, np.float32))

    def boom(q, k):
        raise ValueError("search failed")
    mb2 = MicroBatcher(boom, k=5)
    with pytest.raises(ValueError, match="search failed"):
        mb2.search(np.ones(3, np.float32))
    with pytest.raises(ValueError):                        # the worker survived the failed batch
        mb2.search(np.ones(3, np.float32))
    mb2.close()
