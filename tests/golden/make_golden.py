"""Generates tests/golden/*.npz in THIS container (needs /root/reference and HuggingFace transformers).

  python tests/golden/make_golden.py

* jfk_logmel_hf.npz - the reference's own test clip `Tests/WhisperKitTests/Resources/jfk.wav` (11 s, 16 kHz mono s16;
  used by UnitTests.swift:676-693, FunctionalTests.swift:15-65) as int16 PCM, and the log-mel HF
  `WhisperFeatureExtractor` (== openai-whisper) computes for it at 80 and 128 mels, subsampled every 8th frame.
  The GPU box has no /root/reference, so the PCM travels inside the fixture.
* toy_logits_hf.npz - logits of a seeded random toy Whisper from `transformers.WhisperForConditionalGeneration`
  for a fixed mel / token prefix: pins oracle/model_ref.py to an independent implementation.
"""
import os
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    import torch
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    from transformers.models.whisper.feature_extraction_whisper import WhisperFeatureExtractor

    from oracle import model_ref as M

    with wave.open("/root/reference/Tests/WhisperKitTests/Resources/jfk.wav", "rb") as w:
        assert w.getframerate() == 16000 and w.getnchannels() == 1 and w.getsampwidth() == 2
        pcm16 = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).copy()
    x = np.zeros(480000, np.float32)
    x[: len(pcm16)] = pcm16.astype(np.float32) / 32768.0  # AVFoundation s16 -> f32 convention
    out = {"pcm16": pcm16}
    for nm in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=nm)
        mel = fe._np_extract_fbank_features(x[None], "cpu")[0]
        out[f"mel{nm}_sub8"] = mel[:, ::8].astype(np.float32)
        out[f"mel{nm}_stats"] = np.array([mel.mean(), mel.std(), mel.min(), mel.max()], np.float64)
    np.savez_compressed(os.path.join(HERE, "jfk_logmel_hf.npz"), **out)

    dims = M.VARIANTS["toy"]
    w_ = M.random_weights(dims, seed=1, policy="fp32")
    cfg = WhisperConfig(vocab_size=dims.vocab, num_mel_bins=dims.n_mels, d_model=dims.d_model, encoder_layers=dims.enc_layers,
                        decoder_layers=dims.dec_layers, encoder_attention_heads=dims.n_heads, decoder_attention_heads=dims.n_heads,
                        encoder_ffn_dim=dims.ffn, decoder_ffn_dim=dims.ffn, max_source_positions=1500, max_target_positions=448,
                        activation_function="gelu", pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1,
                        suppress_tokens=None, begin_suppress_tokens=None)
    hf = WhisperForConditionalGeneration(cfg).eval()
    hf.load_state_dict(M.to_hf_state_dict(w_), strict=False)
    g = torch.Generator().manual_seed(0)
    mel = torch.randn(1, dims.n_mels, 3000, generator=g).to(torch.float16).to(torch.float32)  # stored as f16
    toks = torch.randint(0, dims.vocab, (1, 8), generator=g)
    with torch.no_grad():
        enc = hf.model.encoder(mel).last_hidden_state
        logits = hf(encoder_outputs=(enc,), decoder_input_ids=toks).logits
    np.savez_compressed(os.path.join(HERE, "toy_logits_hf.npz"), mel=mel.numpy().astype(np.float16), tokens=toks.numpy(),
                        enc_sub=enc.numpy()[:, ::50].astype(np.float32), logits=logits.numpy().astype(np.float32))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
