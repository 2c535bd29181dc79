"""fp32 restatement of the frozen frame encoder and the image-processor arithmetic.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference calls third-party code here: HF ``transformers`` ``GitVisionModel``
(``src/preprocessing/extract_features.py:145``; version not pinned by the reference, 5.5.0
installed).  The algorithm restated below is the published CLIP ViT-B/16 vision tower as
implemented in ``transformers/models/git/modeling_git.py``:

* embeddings (``:451-531``): Conv2d(3, 768, k=16, s=16, bias=False) patch embedding, a class
  token prepended, learned position embedding added -> 197 tokens;
* ``pre_layrnorm`` (``:742``), 12 pre-LN blocks (``:645-666``): LN1 -> MHA (12 heads x 64,
  scale 1/8, fp32 softmax, ``:556-575``) -> +residual -> LN2 -> fc1 -> quick_gelu
  ``x * sigmoid(1.702 x)`` -> fc2 -> +residual;
* ``post_layernorm`` over ALL tokens (``:751``); LN eps 1e-5.

``tests/test_oracle_golden.py`` checks this against HF itself (live when importable, and
through ``tests/golden/encoder_hf.npz``).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F

HIDDEN, HEADS, HEAD_DIM, LAYERS, PATCH, EPS = 768, 12, 64, 12, 16, 1e-5

IMAGE_MEAN = (0.48145466, 0.4578275, 0.40821073)
IMAGE_STD = (0.26862954, 0.26130258, 0.27577711)


def image_processor_224(u8_hwc: torch.Tensor) -> torch.Tensor:
    """``prefetch_loader.py:74-75`` for frames that are already 224x224 (shortest-edge
    resize and centre crop are identities): rescale by 1/255 then (x - mean) / std, fp32,
    HWC -> CHW.  Returns (T, 3, 224, 224)."""
    x = u8_hwc.permute(0, 3, 1, 2).to(torch.float32) * (1.0 / 255.0)
    mean = torch.tensor(IMAGE_MEAN, dtype=torch.float32, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(IMAGE_STD, dtype=torch.float32, device=x.device).view(1, 3, 1, 1)
    return (x - mean) / std


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(1.702 * x)


class VitOracle:
    """Callable like the reference's ``model``: ``model(frames).last_hidden_state``."""

    def __init__(self, state_dict: dict, dtype=torch.float32, device="cpu"):
        """``device``: where this plain-torch fp32 restatement runs.  The CPU is the pinned oracle (tests/golden); tests
        that need many clips run the SAME code on the GPU in strict fp32 (TF32 off) after checking it against the CPU."""
        self.w = {k: v.detach().to(device=device, dtype=dtype) for k, v in state_dict.items()}
        self.dtype = dtype

    def _ln(self, x, name):
        return F.layer_norm(x, (HIDDEN,), self.w[name + ".weight"], self.w[name + ".bias"], EPS)

    def embeddings(self, pixel_values: torch.Tensor) -> torch.Tensor:
        n = pixel_values.shape[0]
        g = pixel_values.shape[-1] // PATCH
        # conv with stride == kernel == patchify + matmul
        patches = pixel_values.to(self.dtype).reshape(n, 3, g, PATCH, g, PATCH).permute(0, 2, 4, 1, 3, 5)
        patches = patches.reshape(n, g * g, 3 * PATCH * PATCH)
        wp = self.w["vision_model.embeddings.patch_embedding.weight"].reshape(HIDDEN, -1)
        tok = patches @ wp.t()
        cls = self.w["vision_model.embeddings.class_embedding"].expand(n, 1, HIDDEN)
        x = torch.cat([cls, tok], dim=1)
        return x + self.w["vision_model.embeddings.position_embedding.weight"].unsqueeze(0)

    def layer(self, x: torch.Tensor, l: int) -> torch.Tensor:
        p = f"vision_model.encoder.layers.{l}."
        n, s, _ = x.shape
        h = self._ln(x, p + "layer_norm1")
        q = F.linear(h, self.w[p + "self_attn.q_proj.weight"], self.w[p + "self_attn.q_proj.bias"])
        k = F.linear(h, self.w[p + "self_attn.k_proj.weight"], self.w[p + "self_attn.k_proj.bias"])
        v = F.linear(h, self.w[p + "self_attn.v_proj.weight"], self.w[p + "self_attn.v_proj.bias"])
        q = q.view(n, s, HEADS, HEAD_DIM).transpose(1, 2)
        k = k.view(n, s, HEADS, HEAD_DIM).transpose(1, 2)
        v = v.view(n, s, HEADS, HEAD_DIM).transpose(1, 2)
        att = (q @ k.transpose(-1, -2)) * (HEAD_DIM ** -0.5)
        att = torch.softmax(att, dim=-1, dtype=torch.float32).to(q.dtype)
        o = (att @ v).transpose(1, 2).reshape(n, s, HIDDEN)
        x = x + F.linear(o, self.w[p + "self_attn.out_proj.weight"], self.w[p + "self_attn.out_proj.bias"])
        h = self._ln(x, p + "layer_norm2")
        h = quick_gelu(F.linear(h, self.w[p + "mlp.fc1.weight"], self.w[p + "mlp.fc1.bias"]))
        return x + F.linear(h, self.w[p + "mlp.fc2.weight"], self.w[p + "mlp.fc2.bias"])

    @torch.no_grad()
    def forward_hidden(self, pixel_values: torch.Tensor, n_layers: int = LAYERS, post_ln: bool = True):
        x = self._ln(self.embeddings(pixel_values), "vision_model.pre_layrnorm")
        for l in range(n_layers):
            x = self.layer(x, l)
        return self._ln(x, "vision_model.post_layernorm") if post_ln else x

    def __call__(self, pixel_values: torch.Tensor):
        return SimpleNamespace(last_hidden_state=self.forward_hidden(pixel_values))

    def features(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """Unit-norm pooled features (n, 768): token mean then L2 normalise (utils.py:44-47)."""
        return F.normalize(self.forward_hidden(pixel_values).mean(dim=1))


def hf_model_from_state_dict(state_dict: dict):
    """The reference's actual encoder class with our weights (needs ``transformers``)."""
    from transformers import GitVisionConfig, GitVisionModel
    model = GitVisionModel(GitVisionConfig())
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing), (missing, unexpected)
    return model.eval()


def visual_tokens(frames: torch.Tensor, encoder, projection_sd: dict | None) -> torch.Tensor:
    """The visual side of the reference's downstream forward, ``src/modeling/modeling.py:76-95``
    (``MyGitModel.forward``, 5-D ``pixel_values``): for every frame index ``image_encoder(pixel_values[:, f])
    .last_hidden_state``, concatenated along the sequence (the temporal embedding is commented out there,
    ``:87``), then ``visual_projection`` = HF ``GitProjection`` (``transformers/models/git/modeling_git.py``:
    ``nn.Sequential(nn.Linear(768, hidden), nn.LayerNorm(hidden, eps=1e-5))``).
    frames [B, K, 3, 224, 224] fp32 -> [B, K * 197, 768]; ``projection_sd`` None stops before the projection."""
    feats = [encoder(frames[:, f]).last_hidden_state for f in range(frames.shape[1])]
    x = torch.cat(feats, dim=1)
    if projection_sd is None:
        return x
    x = F.linear(x, projection_sd["visual_projection.0.weight"], projection_sd["visual_projection.0.bias"])
    return F.layer_norm(x, (HIDDEN,), projection_sd["visual_projection.1.weight"],
                        projection_sd["visual_projection.1.bias"], EPS)
