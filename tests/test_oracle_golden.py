"""Pins the oracle (and the CPU replay of the mel kernel) to the committed golden vectors
(tests/golden/make_golden.py: the reference's jfk.wav through HF/openai-whisper log-mel; HF Whisper logits)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import mel_ref
from oracle import model_ref as M

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def jfk():
    z = np.load(os.path.join(GOLD, "jfk_logmel_hf.npz"))
    x = np.zeros(480000, np.float32)
    x[: len(z["pcm16"])] = z["pcm16"].astype(np.float32) / 32768.0
    return z, x, len(z["pcm16"])


@pytest.mark.parametrize("n_mels", [80, 128])
def test_oracle_logmel_matches_hf_on_jfk(n_mels):
    z, x, n = jfk()
    assert n == 176000  # 11.0 s clip
    mel = mel_ref.log_mel(x, n_mels)
    assert mel.shape == (n_mels, 3000)
    np.testing.assert_allclose(mel[:, ::8], z[f"mel{n_mels}_sub8"], atol=2e-5)
    st = z[f"mel{n_mels}_stats"]
    np.testing.assert_allclose([mel.mean(), mel.std(), mel.min(), mel.max()], st, atol=1e-5)


def test_pad_or_trim():
    a = np.arange(10, dtype=np.float32)
    np.testing.assert_array_equal(mel_ref.pad_or_trim(a, 2, 6), [2, 3, 4, 5, 6, 7])
    np.testing.assert_array_equal(mel_ref.pad_or_trim(a, 8, 6), [8, 9, 0, 0, 0, 0])
    assert mel_ref.pad_or_trim(a).shape == (480000,)


@pytest.mark.parametrize("n_mels", [80, 128])
def test_mel_kernel_cpu_replay_matches_oracle(n_mels):
    """The CUDA mel kernel's task functions (mel_core.cuh), replayed on the CPU in kernel order."""
    from whisperkit_b200 import build
    lib = ctypes.CDLL(build.build_hostcheck())
    z, x, n = jfk()
    for pcm, nv in ((x, n), (mel_ref.synthetic_pcm(5), 480000), (np.zeros(480000, np.float32), 480000)):
        out = np.zeros((n_mels, 3000), np.float32)
        lib.wk_hostcheck_mel(pcm.ctypes.data_as(ctypes.c_void_p), nv, n_mels, out.ctypes.data_as(ctypes.c_void_p))
        ref = mel_ref.log_mel(pcm, n_mels)
        assert np.abs(out - ref).max() <= 1e-4  # fixed-point code resolution 2.4e-4 / 4, before the f16 store


def test_oracle_model_matches_hf_golden():
    import torch
    z = np.load(os.path.join(GOLD, "toy_logits_hf.npz"))
    dims = M.VARIANTS["toy"]
    orc = M.WhisperOracle(dims, M.random_weights(dims, seed=1, policy="fp32"), "fp32")
    with torch.no_grad():
        enc = orc.encode(torch.from_numpy(z["mel"].astype(np.float32)))
        np.testing.assert_allclose(enc.numpy()[:, ::50], z["enc_sub"], atol=2e-5)
        cross, cache = orc.cross_kv(enc), orc.new_cache(1)
        toks = torch.from_numpy(z["tokens"])
        for t in range(toks.shape[1]):
            lg = orc.decode_step(toks[:, t], t, cache, cross)
            np.testing.assert_allclose(lg.numpy(), z["logits"][:, t], atol=2e-5)


def test_oracle_model_matches_hf_live():
    """Same check against transformers itself when it is importable (it is in this image)."""
    torch = pytest.importorskip("torch")
    tr = pytest.importorskip("transformers")
    dims = M.VARIANTS["toy128"]
    w = M.random_weights(dims, seed=9, policy="fp32")
    cfg = tr.WhisperConfig(vocab_size=dims.vocab, num_mel_bins=dims.n_mels, d_model=dims.d_model, encoder_layers=dims.enc_layers,
                           decoder_layers=dims.dec_layers, encoder_attention_heads=dims.n_heads, decoder_attention_heads=dims.n_heads,
                           encoder_ffn_dim=dims.ffn, decoder_ffn_dim=dims.ffn, max_source_positions=1500, max_target_positions=448,
                           activation_function="gelu", pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1,
                           suppress_tokens=None, begin_suppress_tokens=None)
    hf = tr.WhisperForConditionalGeneration(cfg).eval()
    hf.load_state_dict(M.to_hf_state_dict(w), strict=False)
    orc = M.WhisperOracle(dims, w, "fp32")
    mel = torch.randn(2, dims.n_mels, 3000)
    toks = torch.randint(0, dims.vocab, (2, 5))
    with torch.no_grad():
        enc_hf = hf.model.encoder(mel).last_hidden_state
        enc = orc.encode(mel)
        assert (enc - enc_hf).abs().max() < 2e-5
        out_hf = hf(encoder_outputs=(enc_hf,), decoder_input_ids=toks).logits
        cross, cache = orc.cross_kv(enc), orc.new_cache(2)
        for t in range(5):
            assert (orc.decode_step(toks[:, t], t, cache, cross) - out_hf[:, t]).abs().max() < 2e-5


def test_oracle_alignment_heads_match_hf_cross_attentions():
    """The oracle's `alignment_heads_weights` (mean over the alignment heads of the cross-attention softmax row, what the GPU path exports
    for word timestamps) against the cross-attention probabilities HuggingFace Whisper returns with output_attentions - the tensors
    openai-whisper's timing.py reads through its alignment-head hooks."""
    torch = pytest.importorskip("torch")
    tr = pytest.importorskip("transformers")
    dims = M.VARIANTS["toy128"]
    w = M.random_weights(dims, seed=4, policy="fp32", std=0.08)        # larger weights: attention rows far from uniform
    cfg = tr.WhisperConfig(vocab_size=dims.vocab, num_mel_bins=dims.n_mels, d_model=dims.d_model, encoder_layers=dims.enc_layers,
                           decoder_layers=dims.dec_layers, encoder_attention_heads=dims.n_heads, decoder_attention_heads=dims.n_heads,
                           encoder_ffn_dim=dims.ffn, decoder_ffn_dim=dims.ffn, max_source_positions=1500, max_target_positions=448,
                           activation_function="gelu", pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1,
                           suppress_tokens=None, begin_suppress_tokens=None, attn_implementation="eager")
    hf = tr.WhisperForConditionalGeneration(cfg).eval()
    hf.load_state_dict(M.to_hf_state_dict(w), strict=False)
    orc = M.WhisperOracle(dims, w, "fp32")
    mel = torch.randn(2, dims.n_mels, 3000)
    toks = torch.randint(0, dims.vocab, (2, 6))
    heads = [(0, 1), (1, 0), (1, 3)]
    with torch.no_grad():
        enc = orc.encode(mel)
        out = hf(encoder_outputs=(enc,), decoder_input_ids=toks, output_attentions=True)
        cross, cache = orc.cross_kv(enc), orc.new_cache(2)
        for t in range(6):
            _, al = orc.decode_step(toks[:, t], t, cache, cross, align_heads=heads)
            ref = sum(out.cross_attentions[l][:, h, t] for l, h in heads) / len(heads)       # [B, 1500]
            assert ref.max() > 5.0 / 1500                                                 # not a flat row
            assert (al - ref.to(torch.float16).to(torch.float32)).abs().max() <= 2e-6 + 1e-3 * ref.max()
            assert abs(float(al.sum(-1).mean()) - 1.0) < 2e-3
