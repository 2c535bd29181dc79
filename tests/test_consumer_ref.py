"""Row f1 with the REAL consumer: the reference's own ``VideoQADataset`` / ``GITVideoQACollator``
(src/datasets/dataset_video_qa.py:17-108,323-406 and src/datasets/dataset_base.py:104, executed unmodified from
/root/reference or oracle/_ref) read the artefacts this repo's writers produced -- the HDF5 ``sampled_frames`` file,
``vidmapping.json`` and ``qa_winds_{split}.json`` -- and their batches must equal ``writer.collate_sampled_rows``.

h5py / easydict / tensorboardX are absent from the image: tests/ref_import_shims.py provides ``h5py.File(p, 'r')[name]`` on top of the
repo's minimal HDF5 reader (the real h5py is used when importable) and inert stand-ins for the other two.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_loader
from sasvqa_b200 import writer

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="neither /root/reference nor oracle/_ref present")

IMG, K, N = 16, 6, 5


class _Encoding(dict):
    __getattr__ = dict.__getitem__


class _Processor:
    """Stand-in for the HF processor the collator tokenises with (no checkpoint offline): whitespace tokens -> ids."""

    def __call__(self, text, padding="longest", return_tensors="pt", **kw):
        rows = [[101] + [1000 + (hash(w) % 5000) for w in t.split()] + [102] for t in text]
        L = max(len(r) for r in rows)
        ids = torch.tensor([r + [0] * (L - len(r)) for r in rows])
        mask = torch.tensor([[1] * len(r) + [0] * (L - len(r)) for r in rows])
        return _Encoding(input_ids=ids, attention_mask=mask)


@pytest.fixture(scope="module")
def ref_consumer():
    import ref_import_shims
    used = ref_import_shims.install()
    root = ref_loader.source_root()
    sys.path.insert(0, root)
    try:
        import transformers
        orig = transformers.AutoProcessor.from_pretrained
        transformers.AutoProcessor.from_pretrained = staticmethod(lambda *a, **k: _Processor())   # the collator's left_processor
        from src.datasets import dataset_video_qa as dvq          # the reference module itself
        yield dvq, used
        transformers.AutoProcessor.from_pretrained = orig
    finally:
        sys.path.remove(root)
        for name in [m for m in sys.modules if m == "src" or m.startswith("src.")]:
            del sys.modules[name]


@pytest.fixture(scope="module")
def artefacts(tmp_path_factory):
    """What extraction + MIF leave on disk, written by the repo's own writers."""
    d = tmp_path_factory.mktemp("artefacts")
    rng = np.random.RandomState(0)
    rows = rng.randn(N, K, 3 * IMG * IMG).astype(np.float32)
    h5_path = str(d / "msvd_qa_video_feat.h5")
    with writer.SampledFramesWriter(h5_path, N, K, img=IMG) as w:
        for i in range(N):
            w[i] = torch.from_numpy(rows[i]).view(K, 3, IMG, IMG)              # extract_features.py:96-97
    video_paths = [f"/data/msvd/video/vid{i:03d}.avi" for i in (3, 0, 4, 1, 2)]   # shuffled like extract_features.py:162
    vidmap = writer.generate_vidid_json(video_paths, str(d / "vidmapping.json"))
    qa = [{"question": f"what is in clip {i} ?", "answer": "cat" if i % 2 else "dog", "video": f"vid{i:03d}.avi",
           "answer_type": "what"} for i in range(N)]
    inds = [list(rng.permutation(K)) for _ in range(N)]
    writer.write_sampled_inds(qa, inds, str(d / "qa_winds_train.json"))
    return dict(dir=d, rows=rows, h5=h5_path, vidmap=vidmap, inds=inds)


def _datalist(anno_path):
    """The msvd_qa branch of mk_tgif_qa_dataloader (src/tasks/run_video_qa.py:59-74), one example per video."""
    out = []
    for qid, raw in enumerate(json.load(open(anno_path))):
        d = dict(question=raw["question"], answer=raw["answer"], video_id=raw["video"].split(".")[0],
                 answer_type=raw["answer_type"], question_id=qid, sampled_inds=raw["sampled_inds"])
        out.append((d["video_id"], [d]))
    return out


@pytest.mark.parametrize("policy", ["importance", "question-caption"])
def test_reference_dataset_and_collator_read_our_artefacts(ref_consumer, artefacts, policy):
    dvq, used = ref_consumer
    datalist = _datalist(str(artefacts["dir"] / "qa_winds_train.json"))
    vidmap = json.load(open(artefacts["dir"] / "vidmapping.json"))
    ds = dvq.VideoQADataset(task_type="msvd_qa", datalist=datalist, tokenizer=None, img_hdf5_dir=artefacts["h5"],
                            ans2label={"cat": 0, "dog": 1}, vid2id=vidmap, is_train=False)
    assert len(ds) == N
    items = [ds[i] for i in range(N)]
    # row lookup through vidmapping (dataset_video_qa.py:53-56): video i sits in row vidmap[f"vid{i:03d}"]
    for i, it in enumerate(items):
        np.testing.assert_array_equal(np.asarray(it["vid"]), artefacts["rows"][vidmap[f"vid{i:03d}"]])
        assert it["sampled_inds"] == [int(v) for v in artefacts["inds"][i]]
    nframe = 3
    coll = dvq.GITVideoQACollator(processor=_Processor(), nframe=nframe, samp_policy=policy, img_size=IMG, task_type="msvd_qa")
    batch = coll.collate_batch(items)
    rows = np.stack([artefacts["rows"][vidmap[f"vid{i:03d}"]] for i in range(N)])
    want = writer.collate_sampled_rows(rows, policy, nframe, sampled_inds=artefacts["inds"], img=IMG)
    got = batch["visual_inputs"].numpy()
    assert got.shape == (N, nframe, 3, IMG, IMG)
    np.testing.assert_array_equal(got, want)
    assert batch["video_start_end"] == [i * nframe for i in range(N + 1)]
    assert used["h5py"] in ("real", "shim")


def test_file_is_what_the_reference_reader_expects(ref_consumer, artefacts):
    """dataset_base.py:104 verbatim: h5py.File(img_hdf5_dir, 'r')['sampled_frames'] -- shape, dtype, rows."""
    import h5py
    ds = h5py.File(artefacts["h5"], "r")["sampled_frames"]
    assert tuple(ds.shape) == (N, K, 3 * IMG * IMG) and ds.dtype == np.float32
    np.testing.assert_array_equal(np.asarray(ds[2]), artefacts["rows"][2])


def test_real_h5py_opens_the_writers_file(artefacts):
    """where the real h5py / libhdf5 is installed, it must open a file written by the hand-laid-out writer."""
    h5py = pytest.importorskip("h5py")
    if getattr(h5py, "__shim__", None):
        pytest.skip("only the hdf5_min-backed shim is importable here (no libhdf5 in the image)")
    from sasvqa_b200 import hdf5_min
    path = str(artefacts["dir"] / "by_hdf5_min.h5")
    mm = hdf5_min.create_dataset_file(path, "sampled_frames", artefacts["rows"].shape, np.float32)
    mm[:] = artefacts["rows"]
    mm.flush()
    with h5py.File(path, "r") as f:
        np.testing.assert_array_equal(np.asarray(f["sampled_frames"]), artefacts["rows"])
