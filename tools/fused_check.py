"""Bit-identity check of the two decode schedules: the fused phase chains (csrc/fused_chain.cu, WKB200_FUSED=1) against one launch per
phase (the default).  The arithmetic and its order are the same, only the launch structure differs, so tokens and logits must match bit
for bit - at toy widths and at large-v3 width (d 1280, 20 heads, vocabulary 51866; 3 decoder layers to keep it short).

    timeout 600 python tools/fused_check.py            # exits 0 on a match
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import numpy as np
    import whisperkit_b200 as wk
    from oracle import decode_ref as D
    from oracle import mel_ref
    out = {}
    cases = (("toy128", 3, None, "f16"), ("toy", 4, None, "bf16"),
             ("large-v3", 5, dict(enc_layers=2, dec_layers=3), "bf16"), ("large-v3", 64, dict(enc_layers=1, dec_layers=2), "f16"))
    for variant, B, cfg, dtype in cases:
        if cfg:
            st = wk.SpecialTokens(endToken=50257, englishToken=50259, noSpeechToken=50363, noTimestampsToken=50364, specialTokenBegin=50257,
                                  startOfPreviousToken=50362, startOfTranscriptToken=50258, timeTokenBegin=50365, transcribeToken=50360,
                                  translateToken=50359)
        else:
            st = wk.SpecialTokens.from_any(D.SpecialTokens.toy(1024 if variant == "toy" else 2048))
        model = wk.Model(variant, max_batch=B, dtype=dtype, config=cfg)
        model.init_random(seed=3)
        fe, enc, dec = wk.FeatureExtractor(model), wk.AudioEncoder(model), wk.TextDecoder(model, B)
        pcm = np.stack([mel_ref.synthetic_pcm(700 + (i % 6)) for i in range(B)])
        pcm[:, :1000] *= (1 + np.arange(B)[:, None] * 0.01)
        enc_t = enc.encodeFeatures(fe.logMelSpectrogram(pcm))
        o = wk.DecodingOptions(firstTokenLogProbThreshold=None, sampleLength=20, temperatureFallbackCount=0)
        prompt = dec.prefillDecoderInputs(o, st)
        res = dec.decodeText(enc_t, prompt, o, st)
        key = f"{variant}_{B}_{dtype}"
        out[key + "_tokens"] = np.array([r.tokens + [-1] * (40 - len(r.tokens)) for r in res])
        out[key + "_logits"] = dec.lastLogits()
        out[key + "_step"] = dec.predictLogits([7] * B, [21] * B)
        dec.close()
        model.close()
    np.savez(sys.argv[2], **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    import numpy as np
    import tempfile
    d = tempfile.mkdtemp()
    for name, env in (("base", {"WKB200_FUSED": "0"}), ("fused", {"WKB200_FUSED": "1"})):
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, __file__, "child", os.path.join(d, name + ".npz")], env=e, check=True, timeout=280)
    a, b = np.load(os.path.join(d, "base.npz")), np.load(os.path.join(d, "fused.npz"))
    ok = True
    for k in a.files:
        same = np.array_equal(a[k], b[k])
        print(k, "identical" if same else f"DIFFERENT (max abs diff {np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)).max():.3e})")
        ok &= same
    sys.exit(0 if ok else 1)
