"""On-disk artefacts of the sampling stage, in the formats src/datasets and src/tasks read.

* ``sampled_frames``: float32 ``(N_videos, K, 3*IMG*IMG)`` rows, one per video in vidmapping
  order (reference writer: src/preprocessing/extract_features.py:77-79,96-97; reader:
  src/datasets/dataset_base.py:104, src/datasets/dataset_video_qa.py:53-56).  ``*.h5`` paths are
  written as HDF5: through h5py when it is importable, else by the hand-laid-out writer in
  ``hdf5_min.py`` (this image has no h5py / libhdf5); other paths get the same array as
  ``<name>.npy`` (memory-mapped).  ``open_sampled_frames`` reads all three back.
* ``vidmapping.json``: ``{video_id_without_ext: row}`` (extract_features.py:25-30).
* ``qa_winds_{split}.json``: the QA list with ``sampled_inds`` added, best first
  (src/preprocessing/gen_sample.py:90-94; read at src/tasks/run_video_qa.py:72,91-92).
* ``mdf_inds.json`` (additive, not in the reference): ``{video_id: [K indices]}`` so MDF picks
  are inspectable -- the reference never persists them.
"""
from __future__ import annotations

import json
import logging
import os

import numpy as np

from . import hdf5_min

try:  # pragma: no cover - h5py is absent in the build image
    import h5py  # type: ignore
except Exception:  # noqa: BLE001
    h5py = None

DATASET = "sampled_frames"


def generate_vidid_json(video_paths, json_outfile: str) -> dict:
    mapping = {}
    for row, path in enumerate(video_paths):
        mapping[path.split("/")[-1].split(".")[0]] = row
    with open(json_outfile, "w") as f:
        json.dump(mapping, f)
    return mapping


class SampledFramesWriter:
    """``with SampledFramesWriter(path, N, K) as w: w[i] = frames_KxCHW``"""

    def __init__(self, path: str, n_videos: int, K: int, img: int = 224, backend: str | None = None):
        self.shape = (n_videos, K, 3 * img * img)
        self.backend = backend or ("h5" if path.endswith((".h5", ".hdf5")) else "npy")
        if self.backend == "h5" and h5py is not None:
            self.path = path
            self._fd = h5py.File(path, "w")
            self._ds = self._fd.create_dataset(DATASET, self.shape)          # float32, like the reference
            self.backend_used = "h5py"
        elif self.backend == "h5":
            self.path = path
            self._fd = None
            self._ds = hdf5_min.create_dataset_file(path, DATASET, self.shape, np.float32)
            self.backend_used = "hdf5_min (h5py not importable; see the output contract in hdf5_min.py)"
        else:
            self.path = path if path.endswith(".npy") else path + ".npy"
            self._fd = None
            self._ds = np.lib.format.open_memmap(self.path, mode="w+", dtype=np.float32, shape=self.shape)
            self.backend_used = "npy"
        logging.getLogger("sasvqa_b200").info("sampled_frames %s -> %s written by %s", self.shape, self.path, self.backend_used)

    def __setitem__(self, row, frames) -> None:
        arr = frames.detach().cpu().numpy() if hasattr(frames, "detach") else np.asarray(frames)
        if isinstance(row, slice):
            self._ds[row] = arr.reshape(-1, self.shape[1], self.shape[2])
        else:
            self._ds[row] = arr.reshape(self.shape[1], self.shape[2])         # extract_features.py:96-97

    def close(self) -> None:
        if self._fd is not None:
            self._fd.close()
        elif self._ds is not None:
            self._ds.flush()
        self._ds = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def open_sampled_frames(path: str):
    """Row-indexable float32 array ``[N, K, 3*IMG*IMG]`` (what ``h5py.File(p,'r')['sampled_frames']`` gives)."""
    if path.endswith((".h5", ".hdf5")) and os.path.exists(path):
        if h5py is not None:
            return h5py.File(path, "r")[DATASET]
        return hdf5_min.open_datasets(path)[DATASET]
    return np.load(path if path.endswith(".npy") else path + ".npy", mmap_mode="r")


def write_sampled_inds(qa_samples: list, inds_per_sample, out_path: str) -> list:
    """qa_winds_{split}.json: each QA record gains 'sampled_inds' (list of ints, best first)."""
    out = []
    for sample, inds in zip(qa_samples, inds_per_sample):
        rec = dict(sample)
        rec["sampled_inds"] = [int(i) for i in inds]
        out.append(rec)
    with open(out_path, "w") as f:
        json.dump(out, f)
    return out


def write_mdf_inds(vidmapping: dict, indices, out_path: str) -> dict:
    idx = indices.detach().cpu().numpy() if hasattr(indices, "detach") else np.asarray(indices)
    rec = {vid: [int(i) for i in idx[row]] for vid, row in vidmapping.items()}
    with open(out_path, "w") as f:
        json.dump(rec, f)
    return rec


def collate_sampled_rows(rows, samp_policy: str, nframe: int, sampled_inds=None, img: int = 224):
    """Consumer-side view of the artefacts, restating the two sampling policies of the reference's collator that
    read them (src/datasets/dataset_video_qa.py:356-361): ``rows`` [B, K, 3*img*img] (rows of ``sampled_frames``
    looked up through vidmapping, dataset_video_qa.py:53-56) ->  [B, L, 3, img, img];
    'importance' keeps the first ``nframe`` rows (MDF order = importance order), 'question-caption' takes rows
    ``sampled_inds[:nframe]`` of each sample (the MIF indices of qa_winds_{split}.json)."""
    rows = np.asarray(rows)
    if samp_policy == "importance":
        picked = rows[:, :nframe]
    elif samp_policy == "question-caption":
        inds = np.asarray([list(s[:nframe]) for s in sampled_inds], dtype=np.int64)
        picked = rows[np.arange(rows.shape[0])[:, None], inds]
    else:
        raise ValueError("Sample strategy can only be chosen from ['importance', 'question-caption'] here")
    return picked.reshape(picked.shape[0], picked.shape[1], 3, img, img)
