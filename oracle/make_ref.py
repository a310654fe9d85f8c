"""Recipe for ``oracle/_ref/``: a byte-for-byte copy of the reference's Python package (``/root/reference/src``).
TEST / BENCH INFRASTRUCTURE ONLY.

The reference is pure Python (PyTorch + faiss-cpu): there is nothing to compile, but ``/root/reference`` does not exist on
the GPU box.  ``oracle/_ref/`` is git-ignored (reference SOURCES never enter this repository's history) and NOT gpurun-ignored,
so the copy travels with the snapshot like the built ``.so`` files.  It serves
  * ``bench.py --impl reference`` / ``cpu_baseline``: the stock ``TwoTowerModel`` driven through the reference's own step body
    (``kind: "reference"``), and
  * ``tests/test_gpu_unchanged_callers.py``: the byte-identical ``EmbeddingTrainer`` / ``IndexBuilder`` /
    ``RecommendationPipeline`` running on the drop-in, and the stock CPU run they are compared with.
``__graft_entry__.build()`` runs this when ``/root/reference`` is present.  MANIFEST.json records the sha256 of every file.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
OUT = HERE / "_ref"


def available() -> bool:
    return (OUT / "src" / "models" / "two_tower.py").exists()


def path() -> str:
    """directory to put on sys.path so that ``import src.models.two_tower`` resolves to the copied reference"""
    return str(OUT)


def make(ref: Path = REF, out: Path = OUT) -> bool:
    src = ref / "src"
    if not src.is_dir():
        return False
    if (out / "src").exists():
        shutil.rmtree(out / "src")
    manifest = {}
    for f in sorted(src.rglob("*.py")):
        rel = f.relative_to(ref)
        dst = out / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        data = f.read_bytes()
        dst.write_bytes(data)
        manifest[str(rel)] = hashlib.sha256(data).hexdigest()
    (out / "MANIFEST.json").write_text(json.dumps({"source": str(ref), "files": manifest}, indent=1))
    return True


if __name__ == "__main__":
    ok = make()
    print("oracle/_ref:", "written" if ok else "reference tree not found", file=sys.stderr)
