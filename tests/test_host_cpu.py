"""CPU-side checks (run with -m "not gpu"): the C-ABI library loads and exports every symbol
the header declares, host logic matches the oracle, on-disk formats round-trip, the product
never touches oracle/, and the N>1 path works over gloo with world_size 2."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import sasvqa_b200
from sasvqa_b200 import _capi, ops, sampler, sharding, synth, writer
from oracle import mdf


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sasvqa.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sasvqa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = _declared_symbols()
    assert len(syms) >= 18
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sasvqa.h but not exported"
    assert set(syms) == set(_capi.SIGNATURES), "ctypes signature table out of sync with the header"
    assert _capi.lib().sasvqa_abi_version() == 1


def test_library_is_sm100a_blackwell_native():
    sass = subprocess.run(["cuobjdump", "-sass", _capi.LIB_PATH], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass.stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):      # tcgen05.mma, TMA load, tcgen05.ld
        assert mnemonic in sass.stdout, mnemonic


def test_one_backend_per_operation_no_runtime_switch():
    """no environment variable selects an alternate kernel inside the shipped library; the CUDA-core GEMM
    and mma.sync encoder attention used as A/B checks live in the test-only library."""
    blob = open(_capi.LIB_PATH, "rb").read()
    assert b"SASVQA_DEBUG" not in blob and b"gemm_simt" not in blob
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sas-vqa_b200", "csrc")):
        if os.path.basename(dirpath) == "check":
            continue
        for f in files:
            assert "getenv" not in open(os.path.join(dirpath, f)).read(), f
    test_lib = ctypes.CDLL(_capi.TEST_LIB_PATH)
    for name in _capi.TEST_SIGNATURES:
        assert hasattr(test_lib, name), name
    prod = ctypes.CDLL(_capi.LIB_PATH)
    assert not hasattr(prod, "sasvqa_check_gemm_simt") and not hasattr(prod, "sasvqa_check_attention_mma")


def test_reference_tree_materialised_unmodified():
    """oracle/_ref (what travels to the GPU box) holds byte-for-byte copies: every digest in its manifest matches the file
    beside it and, where /root/reference exists, the file it was copied from."""
    import hashlib
    from oracle import build_ref, ref_loader
    if not build_ref.available():
        pytest.skip("oracle/_ref not built (run python __graft_entry__.py where /root/reference exists)")
    man = json.load(open(build_ref.manifest_path()))
    assert "src/preprocessing/datautils/utils.py" in man["files"] and "src/datasets/dataset_video_qa.py" in man["files"]
    for rel, digest in man["files"].items():
        assert hashlib.sha256(open(os.path.join(build_ref.REF_OUT, rel), "rb").read()).hexdigest() == digest
        src = os.path.join(ref_loader.REFERENCE_ROOT, rel)
        if os.path.isfile(src):
            assert hashlib.sha256(open(src, "rb").read()).hexdigest() == digest, rel
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    assert tracked == "", "reference sources must never enter this repository's history"


def test_no_cpu_fallback_and_no_oracle_in_product():
    pkg = os.path.join(ROOT, "sas-vqa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
    with pytest.raises(_capi.SasvqaError):
        ops.preprocess_u8(torch.zeros(1, 224, 224, 3, dtype=torch.uint8))        # CPU tensor -> loud failure
    with pytest.raises(_capi.SasvqaError):
        ops.mdf_scores(torch.zeros(4, 768), 1)


def test_uniform_and_git6_index_math_match_oracle():
    for T in (8, 17, 64, 100, 333):
        for K in (1, 3, 8, 16):
            if K <= T:
                assert sampler.uniform_indices(T, K) == mdf.uniform_indices(T, K)
                fr = torch.arange(T).float().view(T, 1)
                assert sampler.sample_frames_uniform(fr, K).view(-1).long().tolist() == mdf.uniform_indices(T, K)
    np.random.seed(666)
    got = [sampler.sample_frame_indices(np.arange(T), 6, 4, T).tolist() for T in (40, 64, 300)]
    rng = np.random.RandomState(666)
    want = [mdf.git6_indices(T, 6, 4, rng=rng).tolist() for T in (40, 64, 300)]
    assert got == want


def test_empty_clip_contract():
    dc = {"Failure": 0, "Zeros": 0}
    sd = {}

    class Dummy(ops.FrameEncoder):      # no GPU needed: T == 0 never reaches the encoder
        def __init__(self):
            self._h = None
            self.device = torch.device("cpu")

    out = sampler.sample_representative_frames(torch.zeros(0, 3, 224, 224), Dummy(), 16, 8, dc)
    assert tuple(out.shape) == (16, 3, 224, 224) and float(out.abs().sum()) == 0 and dc["Zeros"] == 1
    with pytest.raises(TypeError):      # the reference dereferences debug_counter unconditionally here
        sampler.sample_representative_frames(torch.zeros(0, 3, 224, 224), Dummy(), 16, 8, None)


def test_state_dict_flattening():
    keys = synth.state_dict_keys()
    assert len(keys) == 199 - 0 - 0 or len(keys) == 5 + 12 * 16 + 2
    assert sum(int(np.prod(s)) for _, s in keys) == ops.NUM_PARAMS
    sd = {k: torch.full(s, float(i)) for i, (k, s) in enumerate(keys)}
    flat = ops.flatten_state_dict(sd)
    assert flat.numel() == ops.NUM_PARAMS and float(flat[0]) == 0.0 and float(flat[-1]) == len(keys) - 1
    bad = dict(sd)
    bad.pop(keys[7][0])
    with pytest.raises(KeyError):
        ops.flatten_state_dict(bad)


def test_hf_state_dict_key_order_matches():
    tr = pytest.importorskip("transformers")
    model = tr.GitVisionModel(tr.GitVisionConfig())
    hf_keys = [k for k in model.state_dict().keys() if "position_ids" not in k]
    assert hf_keys == [k for k, _ in synth.state_dict_keys()]


def test_writers_round_trip(tmp_path):
    mapping = writer.generate_vidid_json(["/d/video/vid12.avi", "/d/video/abc.mp4"], str(tmp_path / "vidmapping.json"))
    assert mapping == {"vid12": 0, "abc": 1} == json.load(open(tmp_path / "vidmapping.json"))
    K, img = 4, 8
    frames = torch.randn(2, K, 3, img, img)
    with writer.SampledFramesWriter(str(tmp_path / "msvd_qa_video_feat.h5"), 2, K, img=img, backend="npy") as w:
        w[0] = frames[0]
        w[1] = frames[1]
    ds = writer.open_sampled_frames(str(tmp_path / "msvd_qa_video_feat.h5"))
    assert ds.shape == (2, K, 3 * img * img) and ds.dtype == np.float32
    # consumer-side view (src/datasets/dataset_video_qa.py:53-56 + collator reshape)
    assert np.array_equal(np.asarray(ds[1]).reshape(K, 3, img, img), frames[1].numpy())
    qa = [{"question": "what", "video": 3, "answer": "x"}]
    out = writer.write_sampled_inds(qa, [[5, 1, 9]], str(tmp_path / "qa_winds_val.json"))
    assert json.load(open(tmp_path / "qa_winds_val.json")) == out and out[0]["sampled_inds"] == [5, 1, 9]
    rec = writer.write_mdf_inds(mapping, torch.tensor([[1, 2], [3, 4]]), str(tmp_path / "mdf_inds.json"))
    assert rec == {"vid12": [1, 2], "abc": [3, 4]}


def test_hdf5_sampled_frames_round_trip_and_reader_pinned_on_genuine_file(tmp_path):
    """The .h5 backend without h5py: hand-laid-out HDF5 (hdf5_min).  Its reader is first pinned on a file written
    by libhdf5 itself (ships with scipy's test data: doubles linspace(0, 2*pi, 9)), then reads what we wrote."""
    from sasvqa_b200 import hdf5_min
    try:
        import scipy.io.matlab
        genuine = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    except Exception:  # noqa: BLE001
        genuine = ""
    if os.path.exists(genuine):
        ds = hdf5_min.open_datasets(genuine)
        assert list(ds) == ["testdouble"] and ds["testdouble"].shape == (9, 1)
        assert np.allclose(np.asarray(ds["testdouble"]).ravel(), np.linspace(0, 2 * np.pi, 9))
    K, img = 4, 8
    frames = torch.randn(3, K, 3, img, img)
    path = str(tmp_path / "msvd_qa_video_feat.h5")
    with writer.SampledFramesWriter(path, 3, K, img=img) as w:             # extract_features.py:77-79
        for i in range(3):
            w[i] = frames[i].reshape(K, -1)                                 # extract_features.py:96-97
    raw = open(path, "rb").read(16)
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    ds = writer.open_sampled_frames(path)                                   # dataset_base.py:104
    assert ds.shape == (3, K, 3 * img * img) and ds.dtype == np.float32
    assert np.array_equal(np.asarray(ds), frames.reshape(3, K, -1).numpy())
    # the two collator policies that consume these rows (dataset_video_qa.py:356-361)
    imp = writer.collate_sampled_rows(ds[[0, 2]], "importance", 2, img=img)
    assert np.array_equal(imp, frames[[0, 2], :2].numpy())
    qc = writer.collate_sampled_rows(ds[[1, 2]], "question-caption", 3, [[3, 0, 2, 1], [1, 1, 0, 2]], img=img)
    assert np.array_equal(qc[0], frames[1][[3, 0, 2]].numpy()) and np.array_equal(qc[1], frames[2][[1, 1, 0]].numpy())


def test_shard_by_frames_balances_ragged_lists():
    import random
    rnd = random.Random(3)
    for world in (1, 2, 3, 8):
        for n in (0, 1, 5, 100):
            lens = [rnd.randint(0, 300) for _ in range(n)]
            spans = sharding.shard_by_frames(lens, world)
            assert len(spans) == world and spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1)) and all(s <= e for s, e in spans)
            if n == 100:                                         # frame totals within one longest clip of the ideal share
                loads = [sum(lens[s:e]) for s, e in spans]
                assert max(abs(l - sum(lens) / world) for l in loads) <= max(lens)
    assert sharding.shard_by_frames([100, 1, 1, 1, 1], 2) == [(0, 1), (1, 5)]      # by frames, not by clip count


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 10000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SASVQA_ROOT"])
from sasvqa_b200 import sharding
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
n, K = 7, 4
def make(s, e):
    return torch.arange(s, e, dtype=torch.uint8).view(-1, 1)          # stand-in "clips": carry their id
def fake_sampler(clips, model, K_, W_):                                 # deterministic per clip id
    ids = clips.view(-1).to(torch.int32)
    return dict(indices=ids[:, None] * 10 + torch.arange(K_, dtype=torch.int32)[None], status=ids % 2)
res = sharding.sample_mdf_sharded(make, n, None, K, 8, sampler=fake_sampler)
want = torch.arange(n, dtype=torch.int32)[:, None] * 10 + torch.arange(K, dtype=torch.int32)[None]
assert torch.equal(res["indices"], want), res["indices"]
assert torch.equal(res["status"], torch.arange(n, dtype=torch.int32) % 2)
assert res["shard"] == sharding.shard_range(n, rank, world)
# ---- ragged clip list sharded by frame count: tables gathered from slices of different sizes
lens = [5, 40, 3, 3, 3, 30, 9]
rag = [torch.full((t, 1, 1, 3), i, dtype=torch.uint8) for i, t in enumerate(lens)]     # clip i carries its id
def fake_ragged(clips_, model, K_, W_):
    ids = torch.tensor([int(c.flatten()[0]) if c.numel() else -1 for c in clips_], dtype=torch.int32)
    return dict(indices=ids[:, None] * 10 + torch.arange(K_, dtype=torch.int32)[None], status=ids % 3)
rr = sharding.sample_mdf_ragged_sharded(rag, None, K, 8, sampler=fake_ragged)
assert rr["spans"] == sharding.shard_by_frames(lens, world) and rr["spans"][0][1] != (len(lens) + 1) // 2   # not split by count
want_r = torch.arange(len(lens), dtype=torch.int32)[:, None] * 10 + torch.arange(K, dtype=torch.int32)[None]
assert torch.equal(rr["indices"], want_r) and torch.equal(rr["status"], torch.arange(len(lens), dtype=torch.int32) % 3)
# ---- MIF step sharded by QA sample: a stand-in scorer (deterministic scores from the token ids) on each rank
from sasvqa_b200 import synth
class FakeScorer:
    def select_captions_host(self, ids, tts, msk, n_samples, K_, ds_rate=1, label=0, want_scores=False, idx_out=None):
        T = ids.shape[0] // n_samples
        scores = ((ids * msk).sum(1) % 97).float().view(n_samples, T)
        idx = scores[:, ::ds_rate].topk(K_, dim=1)[1].to(torch.int32) * ds_rate
        return idx, None
tok = synth.SynthTokenizer(2048)
qa, caps = synth.make_qa_workload(9, 6, seed=5)
full = sharding.generate_inds_sharded(tok, FakeScorer(), qa, caps, 3, 2, samples_per_call=2)
assert [r["question"] for r in full] == [r["question"] for r in qa] and len(full) == 9
for smp, rec in zip(qa, full):
    c = caps["video%d" % smp["video"]]
    b = tok(text=[smp["question"]] * len(c), text_pair=c)
    sc = ((b["input_ids"] * b["attention_mask"]).sum(1) % 97).float()
    assert rec["sampled_inds"] == [2 * i for i in sc[::2].topk(3)[1].tolist()], (rec, sc)
dist.destroy_process_group()
print("OK", rank)
"""


def test_sharded_all_gather_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, SASVQA_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29653", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {r}" in o, o


def test_scorer_state_dict_flattening_and_hf_key_order():
    from sasvqa_b200 import scorer
    tr = pytest.importorskip("transformers")
    vocab, labels = 64, 2
    model = tr.BertForSequenceClassification(tr.BertConfig(vocab_size=vocab, num_labels=labels, num_hidden_layers=12))
    hf = [(k, tuple(v.shape)) for k, v in model.state_dict().items() if v.is_floating_point()]
    assert hf == [(k, tuple(s)) for k, s in synth.scorer_state_dict_keys(vocab, labels)]
    flat, v, l = scorer.flatten_scorer_state_dict(model.state_dict())
    assert (v, l) == (vocab, labels)
    assert flat.numel() == _capi.lib().sasvqa_scorer_num_params(vocab, labels)      # host arithmetic only, no GPU
    assert _capi.lib().sasvqa_scorer_num_params(28996, 2) == 108311810               # bert-base-cased, 2 labels
    bad = dict(model.state_dict())
    bad.pop("bert.pooler.dense.bias")
    with pytest.raises(KeyError):
        scorer.flatten_scorer_state_dict(bad)


def test_synth_tokenizer_follows_the_bert_pair_layout():
    tok = synth.SynthTokenizer(2048, max_length=16)
    out = tok(text=["what is it ?", "who"], text_pair=["a dog", "a b c d e f g h i j k l m n o p q r"])
    ids, tts, msk = out["input_ids"], out["token_type_ids"], out["attention_mask"]
    assert ids.dtype == torch.int64 and ids.shape == (2, 16)
    assert ids[0, :8].tolist()[0] == synth.BERT_CLS and ids[0, 5] == synth.BERT_SEP and ids[0, 8] == synth.BERT_SEP
    assert tts[0].tolist() == [0] * 6 + [1] * 3 + [0] * 7 and msk[0].tolist() == [1] * 9 + [0] * 7
    assert int(msk[1].sum()) == 16 and ids[1, 15] == synth.BERT_SEP          # truncated to max_length, SEP kept
    assert ids[0, 9:].eq(synth.BERT_PAD).all()
    single = tok(text=["a b"])
    assert single["input_ids"].shape == (1, 4) and single["token_type_ids"].sum() == 0


def test_scorer_needs_the_gpu_no_cpu_fallback():
    from sasvqa_b200 import scorer
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(Exception):
        scorer.CaptionScorer(synth.random_scorer_state_dict(vocab=32))


def test_plain_c_client_links_and_gets_error_codes(abi_check_exe):
    """tests/c_abi/abi_check.c (C99, knows only include/sasvqa.h): version, sizes and argument errors without a GPU."""
    out = subprocess.run([abi_check_exe], capture_output=True, text=True)
    assert out.returncode == 0 and "abi_check ok" in out.stdout, out.stdout + out.stderr


def test_git_decoder_state_dict_flattening_and_hf_key_order():
    from sasvqa_b200 import vqa
    tr = pytest.importorskip("transformers")
    cfg = tr.GitConfig(vocab_size=64, num_hidden_layers=2)
    model = tr.GitForCausalLM(cfg)
    hf = [(k, tuple(v.shape)) for k, v in model.state_dict().items()
          if v.is_floating_point() and not k.startswith(("git.image_encoder", "git.visual_projection", "git.img_temp"))]
    assert hf == [(k, tuple(s)) for k, s in synth.git_decoder_state_dict_keys(64, 2)]
    flat, vocab, n_layers = vqa.flatten_git_decoder_state_dict(model.state_dict())
    assert (vocab, n_layers) == (64, 2)
    assert flat.numel() == _capi.lib().sasvqa_git_decoder_num_params(64, 2)


def test_h264_pcm_writer_is_parseable():
    """tests/h264_pcm.py (the NVDEC known-answer stream) checked WITHOUT a decoder: NAL framing, emulation prevention, the
    SPS fields a parser reads the frame size from, and the raw samples of the first macroblock of every picture."""
    import h264_pcm
    rng = np.random.RandomState(3)
    T, H, W = 3, 32, 48
    y = rng.randint(0, 256, (T, H, W)).astype(np.uint8)
    u = rng.randint(0, 256, (T, H // 2, W // 2)).astype(np.uint8)
    v = rng.randint(0, 256, (T, H // 2, W // 2)).astype(np.uint8)
    y[0, :16, :16] = 0                                   # a macroblock of zeros: must be escaped (00 00 03)
    stream = h264_pcm.encode_i_pcm(y, u, v)
    nals = [n for n in stream.split(b"\x00\x00\x00\x01") if n]
    assert [n[0] & 0x1F for n in nals] == [7, 8] + [5] * T            # SPS, PPS, one IDR slice per frame
    for n in nals:                                       # no start-code emulation inside a NAL
        assert re.search(rb"\x00\x00[\x00-\x02]", n) is None

    def rbsp(nal):                                       # strip the header byte and the emulation-prevention bytes
        out, zeros = bytearray(), 0
        for b in nal[1:]:
            if zeros >= 2 and b == 3:
                zeros = 0
                continue
            out.append(b)
            zeros = zeros + 1 if b == 0 else 0
        return bytes(out)

    class Bits:
        def __init__(self, data):
            self.d, self.p = data, 0

        def u(self, n):
            val = 0
            for _ in range(n):
                val = (val << 1) | ((self.d[self.p >> 3] >> (7 - (self.p & 7))) & 1)
                self.p += 1
            return val

        def ue(self):
            z = 0
            while self.u(1) == 0:
                z += 1
            return (1 << z) - 1 + (self.u(z) if z else 0)

    sps = Bits(rbsp(nals[0]))
    assert sps.u(8) == 66 and sps.u(8) == 0xC0 and sps.u(8) == 40            # Baseline, level 4.0
    assert [sps.ue() for _ in range(4)] == [0, 0, 2, 1]                     # sps id, log2_max_frame_num-4, poc type 2, ref frames
    assert sps.u(1) == 0
    assert (sps.ue() + 1) * 16 == W and (sps.ue() + 1) * 16 == H
    for t in range(T):
        sl = Bits(rbsp(nals[2 + t]))
        assert sl.ue() == 0 and sl.ue() == 7 and sl.ue() == 0 and sl.u(4) == 0 and sl.ue() == t     # first_mb, I slice, pps, frame_num, idr_pic_id
        sl.u(2)                                                                # dec_ref_pic_marking flags
        assert sl.ue() == 0 and sl.ue() == 1                                   # slice_qp_delta (se(0) reads as ue 0), deblocking off
        assert sl.ue() == 25                                                   # mb_type I_PCM
        while sl.p & 7:
            assert sl.u(1) == 0                                                # pcm_alignment_zero_bit
        first_mb = np.frombuffer(sl.d[sl.p >> 3:(sl.p >> 3) + 256], dtype=np.uint8).reshape(16, 16)
        np.testing.assert_array_equal(first_mb, y[t, :16, :16])
    rgb = h264_pcm.yuv_to_rgb_bt601(y, u, v)
    assert rgb.shape == (T, H, W, 3) and rgb.dtype == np.uint8
