"""A D = 64 index with fewer rows than one tensor-map box (128): `list_scan_pipe_kernel` declines it and the search runs on the
one-CTA-per-tile list scan (`list_scan_tc_kernel`), which the other IVF parity cases no longer reach.  Also reached explicitly,
in its own process, through RB200_IVF_PIPE=0 on a larger index.  (Last file of the suite on purpose: both forms were written after
the round's GPU budget was spent.)"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import ivf_oracle as V

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _case(n, nlist, nprobe, k, nq, seed):
    import recommendit_b200 as R
    rng = np.random.default_rng(seed)
    x = V.normalize_rows(rng.standard_normal((n, 64)).astype(np.float32))
    c = V.spherical_kmeans(x, nlist, seed=1234)
    q = V.normalize_rows(x[rng.integers(0, n, nq)] + 0.2 * rng.standard_normal((nq, 64)).astype(np.float32))
    idx = R.FAISSIndex(64, nlist, nprobe)
    idx.build_ivf_index(x, list(range(n)), centroids=c)
    a = np.empty(n, dtype=np.int64)                      # the oracle's lists from the GPU's own assignment (ties in `assign` aside)
    off_g, ids_g = idx.index.offsets.cpu().numpy(), idx.index.list_ids.cpu().numpy()
    for l in range(nlist):
        a[ids_g[off_g[l]:off_g[l + 1]]] = l
    off, order = V.build_lists(a, nlist)
    ok = V.coarse_probe_margin(q, c, nprobe) > 1e-5
    s_ref, i_ref = V.ivf_search(q, c, off, order, x, nprobe, min(k, n))
    s, ids = idx.batch_search(q, k)
    assert s.shape == (nq, min(k, n)) and (np.diff(s, axis=1) <= 0).all()
    V.assert_topk_equivalent(s[ok], ids[ok], s_ref[ok], i_ref[ok])


def test_index_smaller_than_one_box_of_rows():
    _case(n=100, nlist=4, nprobe=2, k=50, nq=32, seed=11)


def test_per_tile_list_scan_behind_its_knob():
    code = "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n" \
           "import test_gpu_zz_small_index as T\nT._case(n=9000, nlist=64, nprobe=8, k=500, nq=70, seed=12)\nprint('PER-TILE-OK')\n" \
           % (str(ROOT), str(ROOT / "tests"))
    out = subprocess.run([sys.executable, "-c", code], env={**os.environ, "RB200_IVF_PIPE": "0"}, capture_output=True, text=True,
                         timeout=300)
    assert "PER-TILE-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
