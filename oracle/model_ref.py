"""Whisper encoder/decoder oracle in PyTorch fp32 (test infrastructure; see oracle/__init__.py).

The reference runs this arithmetic inside opaque CoreML bundles
(`AudioEncoder.mlmodelc`, `TextDecoder.mlmodelc`;
`Sources/WhisperKit/Core/AudioEncoder.swift:50-63`,
`Sources/WhisperKit/Core/TextDecoder.swift:361-418`), so it is restated from the
published OpenAI Whisper model (whisper/model.py) using HuggingFace parameter
names, and cross-checked against `transformers.WhisperForConditionalGeneration`
in tests/test_oracle_hf.py.

Tensor IO the reference sees (`Sources/WhisperKit/Core/Models.swift`):
  encoder in  "melspectrogram_features" [1, nMels, 1, 3000]   (:909-933)
  encoder out "encoder_output_embeds"   [1, d, 1, 1500]        (:938-965)
  decoder in  input_ids[1], cache_length[1], key/value_cache [1, L*d, 1, 224],
              kv_cache_update_mask[1,224], decoder_key_padding_mask[1,224],
              encoder_output_embeds                           (:970-1034)
  decoder out logits [1,1,V], key/value_cache_updates [1, L*d, 1, 1],
              alignment_heads_weights [1,1500]                (:1037-1107)

Precision policy.  `policy` in {"fp32", "bf16", "f16"}: with a 16-bit policy the
oracle rounds to that type at exactly the points where the CUDA engine stores or
feeds 16-bit data (weights, LayerNorm outputs, Q/K/V, attention outputs, GELU
outputs, KV caches, mel); residual stream, accumulation, softmax statistics and
logits stay fp32.  That is the "same precision policy" twin of SURVEY.md 8(c).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class WhisperDims:
    n_mels: int
    d_model: int
    n_heads: int
    enc_layers: int
    dec_layers: int
    vocab: int
    n_audio_ctx: int = 1500
    n_text_ctx: int = 448

    @property
    def ffn(self):
        return 4 * self.d_model


# Variant table: external facts about OpenAI Whisper; the reference only keys on
# (logitsDim, encoderDim) in Utilities/ModelUtilities.swift:128-173.
VARIANTS = {
    "tiny.en": WhisperDims(80, 384, 6, 4, 4, 51864),
    "tiny": WhisperDims(80, 384, 6, 4, 4, 51865),
    "large-v3": WhisperDims(128, 1280, 20, 32, 32, 51866),
    "large-v3-turbo": WhisperDims(128, 1280, 20, 32, 4, 51866),
    "distil-large-v3": WhisperDims(128, 1280, 20, 32, 2, 51866),
    # toy shapes for fast CPU tests (not real checkpoints)
    "toy": WhisperDims(80, 128, 2, 2, 2, 1024),
    "toy128": WhisperDims(128, 256, 4, 2, 2, 2048),
    "toy512": WhisperDims(80, 512, 8, 2, 2, 2048),     # base-width (8 heads)
    "toy768": WhisperDims(80, 768, 12, 2, 2, 2048),    # small-width (12 heads: a head count that is not a power of two)
}


def round_to(x: torch.Tensor, policy: str) -> torch.Tensor:
    if policy == "bf16":
        return x.to(torch.bfloat16).to(torch.float32)
    if policy == "f16":
        return x.to(torch.float16).to(torch.float32)
    return x


def sinusoids(length: int, channels: int) -> torch.Tensor:
    """whisper/model.py sinusoids(): encoder positional embedding."""
    log_timescale_increment = math.log(10000) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2, dtype=torch.float32))
    scaled_time = torch.arange(length, dtype=torch.float32)[:, None] * inv_timescales[None, :]
    return torch.cat([scaled_time.sin(), scaled_time.cos()], dim=1)


def random_weights(dims: WhisperDims, seed: int = 0, policy: str = "bf16", std: float = 0.02,
                   device: str = "cpu", pool_size: int = 0) -> dict:
    """Seeded N(0, std) weights (LN gamma ~ 1 + N(0,0.02), small biases) under HF names.
    Values are pre-rounded to the 16-bit policy so CPU oracle and GPU engine see
    identical numbers.  conv1 is rounded to f16 (the engine feeds it f16 mel)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = {}
    pool = torch.randn(pool_size, generator=g, dtype=torch.float32) if pool_size else None
    cursor = [0]

    def rnd(*shape, s=std):
        if pool is None:
            return torch.randn(*shape, generator=g, dtype=torch.float32) * s
        # fast path for multi-billion-parameter CPU baselines: distinct memory per tensor, values drawn by
        # sliding a window over a fixed seeded pool (same distribution; timing does not depend on values)
        n = 1
        for v in shape:
            n *= v
        out = torch.empty(n, dtype=torch.float32)
        done = 0
        while done < n:
            off = (cursor[0] * 7919) % max(1, pool_size // 2)
            take = min(n - done, pool_size - off)
            out[done:done + take] = pool[off:off + take]
            done += take
            cursor[0] += 1
        return out.view(*shape) * s

    def lin(name, out_f, in_f, bias=True):
        w[name + ".weight"] = round_to(rnd(out_f, in_f), policy)
        if bias:
            w[name + ".bias"] = rnd(out_f)

    def ln(name, d):
        w[name + ".weight"] = 1.0 + rnd(d)
        w[name + ".bias"] = rnd(d)

    d = dims.d_model
    w["model.encoder.conv1.weight"] = round_to(rnd(d, dims.n_mels, 3, s=0.05), "f16" if policy != "fp32" else "fp32")
    w["model.encoder.conv1.bias"] = rnd(d)
    w["model.encoder.conv2.weight"] = round_to(rnd(d, d, 3), "f16" if policy != "fp32" else "fp32")  # f16 conv stem
    w["model.encoder.conv2.bias"] = rnd(d)
    w["model.encoder.embed_positions.weight"] = sinusoids(dims.n_audio_ctx, d)
    for i in range(dims.enc_layers):
        p = f"model.encoder.layers.{i}."
        lin(p + "self_attn.q_proj", d, d)
        lin(p + "self_attn.k_proj", d, d, bias=False)
        lin(p + "self_attn.v_proj", d, d)
        lin(p + "self_attn.out_proj", d, d)
        ln(p + "self_attn_layer_norm", d)
        lin(p + "fc1", dims.ffn, d)
        lin(p + "fc2", d, dims.ffn)
        ln(p + "final_layer_norm", d)
    ln("model.encoder.layer_norm", d)
    w["model.decoder.embed_tokens.weight"] = round_to(rnd(dims.vocab, d), policy)
    w["model.decoder.embed_positions.weight"] = rnd(dims.n_text_ctx, d)
    for i in range(dims.dec_layers):
        p = f"model.decoder.layers.{i}."
        for a in ("self_attn", "encoder_attn"):
            lin(p + a + ".q_proj", d, d)
            lin(p + a + ".k_proj", d, d, bias=False)
            lin(p + a + ".v_proj", d, d)
            lin(p + a + ".out_proj", d, d)
            ln(p + a + "_layer_norm", d)
        lin(p + "fc1", dims.ffn, d)
        lin(p + "fc2", d, dims.ffn)
        ln(p + "final_layer_norm", d)
    ln("model.decoder.layer_norm", d)
    if device != "cpu":
        w = {k: v.to(device) for k, v in w.items()}
    return w


def gelu(x):
    return F.gelu(x)  # exact erf GELU, as whisper/model.py nn.GELU()


class WhisperOracle:
    """Functional Whisper forward with explicit rounding points."""

    def __init__(self, dims: WhisperDims, weights: dict, policy: str = "bf16"):
        self.dims = dims
        self.w = weights
        self.policy = policy

    def r(self, x):
        return round_to(x, self.policy)

    def _ln(self, x, name):
        return F.layer_norm(x, (x.shape[-1],), self.w[name + ".weight"], self.w[name + ".bias"], 1e-5)

    def _lin(self, x, name):
        return F.linear(x, self.w[name + ".weight"], self.w.get(name + ".bias"))

    def _heads(self, x):
        b, t, d = x.shape
        h = self.dims.n_heads
        return x.view(b, t, h, d // h).transpose(1, 2)  # [b,h,t,dh]

    # ---------------- encoder ----------------
    def encode(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [B, nMels, 3000] (already f16-rounded log-mel) -> [B, 1500, d] fp32
        (final LayerNorm output, unrounded)."""
        w, d = self.w, self.dims
        x = F.conv1d(mel, w["model.encoder.conv1.weight"], w["model.encoder.conv1.bias"], padding=1)
        x = gelu(x)
        if self.policy != "fp32":
            x = round_to(x, "f16")  # the conv stem is f16 (mel is f16): conv1 output is stored as f16
        x = F.conv1d(x, w["model.encoder.conv2.weight"], w["model.encoder.conv2.bias"], stride=2, padding=1)
        x = gelu(x).transpose(1, 2) + w["model.encoder.embed_positions.weight"][None]
        scale = (d.d_model // d.n_heads) ** -0.5
        for i in range(d.enc_layers):
            p = f"model.encoder.layers.{i}."
            xn = self.r(self._ln(x, p + "self_attn_layer_norm"))
            q = self._heads(self.r(self._lin(xn, p + "self_attn.q_proj")))
            k = self._heads(self.r(self._lin(xn, p + "self_attn.k_proj")))
            v = self._heads(self.r(self._lin(xn, p + "self_attn.v_proj")))
            s = (q @ k.transpose(-1, -2)) * scale
            a = torch.softmax(s, dim=-1) @ v
            a = self.r(a.transpose(1, 2).reshape(x.shape))
            x = x + self._lin(a, p + "self_attn.out_proj")
            xn = self.r(self._ln(x, p + "final_layer_norm"))
            h = self.r(gelu(self._lin(xn, p + "fc1")))
            x = x + self._lin(h, p + "fc2")
        return self._ln(x, "model.encoder.layer_norm")

    # ---------------- decoder ----------------
    def cross_kv(self, enc: torch.Tensor):
        """Per-layer cross-attention K/V (rounded as the engine's cross-KV cache)."""
        encr = self.r(enc)
        out = []
        for i in range(self.dims.dec_layers):
            p = f"model.decoder.layers.{i}.encoder_attn."
            k = self._heads(self.r(self._lin(encr, p + "k_proj")))
            v = self._heads(self.r(self._lin(encr, p + "v_proj")))
            out.append((k, v))
        return out

    def new_cache(self, batch: int):
        return {"k": [None] * self.dims.dec_layers, "v": [None] * self.dims.dec_layers, "len": 0}

    def decode_step(self, tokens: torch.Tensor, pos: int, cache: dict, cross, align_heads=None):
        """One decoder step (mirrors TextDecoder.predictLogits, TextDecoder.swift:361-418):
        tokens [B] int64 at position `pos` -> logits [B, V] fp32.  Appends to cache.
        align_heads = [(layer, head), ...] additionally returns the decoder model's `alignment_heads_weights` output
        (TextDecoder.swift:310,414): the mean over those heads of the cross-attention softmax row, [B, T] rounded to Float16
        (FloatType) - openai-whisper's alignment heads (whisper/timing.py find_alignment, model.set_alignment_heads)."""
        align_acc = None
        w, d = self.w, self.dims
        scale = (d.d_model // d.n_heads) ** -0.5
        x = w["model.decoder.embed_tokens.weight"][tokens] + w["model.decoder.embed_positions.weight"][pos][None]
        x = x[:, None, :]  # [B,1,d]
        for i in range(d.dec_layers):
            p = f"model.decoder.layers.{i}."
            xn = self.r(self._ln(x, p + "self_attn_layer_norm"))
            q = self._heads(self._lin(xn, p + "self_attn.q_proj"))  # q stays fp32 in-kernel
            k = self._heads(self.r(self._lin(xn, p + "self_attn.k_proj")))
            v = self._heads(self.r(self._lin(xn, p + "self_attn.v_proj")))
            if cache["k"][i] is None:
                cache["k"][i], cache["v"][i] = k, v
            else:
                cache["k"][i] = torch.cat([cache["k"][i][:, :, :pos], k], dim=2)
                cache["v"][i] = torch.cat([cache["v"][i][:, :, :pos], v], dim=2)
            s = (q @ cache["k"][i].transpose(-1, -2)) * scale
            a = torch.softmax(s, dim=-1) @ cache["v"][i]
            a = self.r(a.transpose(1, 2).reshape(x.shape))
            x = x + self._lin(a, p + "self_attn.out_proj")
            xn = self.r(self._ln(x, p + "encoder_attn_layer_norm"))
            q = self._heads(self._lin(xn, p + "encoder_attn.q_proj"))
            ck, cv = cross[i]
            s = (q @ ck.transpose(-1, -2)) * scale
            pr = torch.softmax(s, dim=-1)
            if align_heads is not None:
                for (li, hi) in align_heads:
                    if li == i:
                        align_acc = pr[:, hi, 0] if align_acc is None else align_acc + pr[:, hi, 0]
            a = pr @ cv
            a = self.r(a.transpose(1, 2).reshape(x.shape))
            x = x + self._lin(a, p + "encoder_attn.out_proj")
            xn = self.r(self._ln(x, p + "final_layer_norm"))
            h = self.r(gelu(self._lin(xn, p + "fc1")))
            x = x + self._lin(h, p + "fc2")
        xn = self.r(self._ln(x, "model.decoder.layer_norm"))
        cache["len"] = pos + 1
        logits = F.linear(xn[:, 0], w["model.decoder.embed_tokens.weight"])
        if align_heads is not None:
            return logits, (align_acc / float(len(align_heads))).to(torch.float16).to(torch.float32)
        return logits


def to_hf_state_dict(weights: dict) -> dict:
    sd = dict(weights)
    sd["proj_out.weight"] = weights["model.decoder.embed_tokens.weight"]
    return sd


def weights_to_numpy(weights: dict) -> dict:
    return {k: v.detach().cpu().numpy() for k, v in weights.items()}
