"""CPU oracle for the SAS-VQA frame-sampling hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package -- as the checker or as the timed CPU
baseline, never as the product.  The product (``sas-vqa_b200/``) never imports it and
fails loudly when its CUDA library is missing.

What it restates (reference = Clement25/SAS-VQA, read-only at /root/reference):

* ``oracle.mdf``  -- ``src/preprocessing/datautils/utils.py:29-109`` (MDF sampler, uniform
  sampler), ``src/preprocessing/extract_features.py:32-39`` (GIT-6 index sampler),
  ``src/preprocessing/gen_sample.py:83-90`` (MIF strided top-K).
* ``oracle.vit``  -- the third-party encoder the reference calls
  (``extract_features.py:145``: HF ``transformers`` ``GitVisionModel``; the reference pins no
  version, 5.5.0 is installed here; math at ``transformers/models/git/modeling_git.py:451-755``)
  and the HF ``CLIPImageProcessor`` arithmetic for 224x224 inputs
  (``src/preprocessing/prefetch_loader.py:74-75``).

* ``oracle.bert`` -- the MIF relevance model the reference calls (``gen_sample.py:79-88,160``: HF
  ``BertForSequenceClassification`` for a bert-base-cased checkpoint; math at
  ``transformers/models/bert/modeling_bert.py``) and the ``generate_inds`` loop body around it; pinned
  against HF itself with seeded random weights (``tests/golden/bert_scorer_hf.npz``).
* ``oracle.git`` -- the downstream video-QA forward on the sampled frames (``src/modeling/modeling.py:29-232``,
  ``MyGitModel`` / ``MyGitForCausalLM`` over HF ``GitModel``): inference logits of the text rows; pinned against HF
  ``GitForCausalLM`` with zeroed temporal embeddings (``tests/golden/git_vqa_hf.npz``).
* ``oracle.resize`` -- the image processor's shortest-edge bicubic resize + centre crop
  (``prefetch_loader.py:74-75`` -> HF ``CLIPImageProcessor`` -> ATen uint8 resampler).

Pinning: the reference ships no tests or golden vectors.  ``oracle/make_golden.py`` imports
the reference's own ``sample_representative_frames`` / ``sample_frames_uniform`` from
/root/reference (and HF ``GitVisionModel`` / ``CLIPImageProcessor``) in the build container
and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement
against those fixtures (and against the live reference when /root/reference exists).
Encoder outputs are pinned against HF with seeded random weights only -- no pretrained
checkpoint is reachable offline ("parity unpinned" for real checkpoints).
"""
