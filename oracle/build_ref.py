"""Materialises ``oracle/_ref/`` -- the UNMODIFIED reference files of the hot path -- from /root/reference.
TEST INFRASTRUCTURE ONLY.

The reference (Clement25/SAS-VQA) is 100 % Python, so "building" it is a byte-for-byte copy of the few files the path
and its consumer live in, plus a manifest of their SHA-256 digests.  ``oracle/_ref/`` is git-ignored (no reference
source enters this repository's history) but NOT gpurun-ignored: it travels to the GPU box, where /root/reference does
not exist, so that

* ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg time the reference's OWN
  ``sample_representative_frames`` (src/preprocessing/datautils/utils.py:31-94) on the box's host cores
  (``cpu_baseline.kind == "reference"``), and
* ``tests/test_consumer_ref.py`` drives the reference's OWN ``VideoQADataset`` / ``GITVideoQACollator``
  (src/datasets/dataset_video_qa.py:17-108,323-406) over files this repo wrote.

Called from ``__graft_entry__.build()``; a no-op (keeping what is there) when /root/reference is absent.
Nothing under ``sas-vqa_b200/`` imports this or ``oracle/_ref``.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("SASVQA_REFERENCE_ROOT", "/root/reference")

# the sampler itself, its driver, the MIF step, the decode/preprocess pipeline; the consumer (datasets) and what it imports
FILES = [
    "src/__init__.py",
    "src/preprocessing/datautils/utils.py",
    "src/preprocessing/extract_features.py",
    "src/preprocessing/gen_sample.py",
    "src/preprocessing/prefetch_loader.py",
    "src/datasets/dataset_video_qa.py",
    "src/datasets/dataset_base.py",
    "src/datasets/data_utils.py",
    "src/datasets/decoder.py",
    "src/utils/basic_utils.py",
    "src/utils/load_save.py",
    "src/utils/logger.py",
]


def _sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def manifest_path() -> str:
    return os.path.join(REF_OUT, "MANIFEST.json")


def available() -> bool:
    """True when oracle/_ref holds every file of the manifest with the recorded digest (i.e. unmodified)."""
    try:
        man = json.load(open(manifest_path()))
    except (OSError, ValueError):
        return False
    for rel, digest in man.get("files", {}).items():
        p = os.path.join(REF_OUT, rel)
        if not os.path.isfile(p) or _sha256(p) != digest:
            return False
    return bool(man.get("files"))


def build(verbose: bool = False) -> str | None:
    """Copies FILES from the reference tree into oracle/_ref/ and writes the manifest.  Returns the output directory,
    or None when the reference tree is not present (the GPU box: the prebuilt oracle/_ref is used as it is)."""
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, FILES[1])):
        return REF_OUT if available() else None
    digests = {}
    for rel in FILES:
        src = os.path.join(REFERENCE_ROOT, rel)
        dst = os.path.join(REF_OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)                                 # bytes only: the files stay unmodified
        digests[rel] = _sha256(dst)
        assert digests[rel] == _sha256(src)
    with open(manifest_path(), "w") as f:
        json.dump({"source": "Clement25/SAS-VQA at " + REFERENCE_ROOT, "note": "byte-for-byte copies; git-ignored",
                   "files": digests}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(digests)} reference files materialised")
    return REF_OUT


if __name__ == "__main__":
    print(build(verbose=True))
