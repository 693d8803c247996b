"""Bring-up check for the experimental fused decoder chains (csrc/fused_chain.cu, WKB200_FUSED=1): the same toy decode with and without
the fused path must give identical tokens and (bit for bit) identical logits - the arithmetic and its order are unchanged, only the
launch structure differs.  Run on a GPU box under a timeout (the kernel spins on grid barriers: a bug can hang it):

    timeout 120 python tools/fused_check.py            # exits 0 on a match
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
    for variant, B in (("toy128", 3), ("toy", 4)):
        st = wk.SpecialTokens.from_any(D.SpecialTokens.toy(1024 if variant == "toy" else 2048))
        kit = wk.WhisperKit(wk.WhisperKitConfig(model=variant, maxBatch=B, seed=3, specialTokens=st, dtype="f16"))
        pcm = np.stack([mel_ref.synthetic_pcm(700 + i) for i in range(B)])
        o = wk.DecodingOptions(firstTokenLogProbThreshold=None, sampleLength=20, temperatureFallbackCount=0)
        res = kit.transcribe(pcm, o)
        out[variant + "_tokens"] = np.array([r.tokens + [-1] * (40 - len(r.tokens)) for r in res])
        out[variant + "_logits"] = kit.textDecoder.lastLogits()
    np.savez(sys.argv[2], **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    import numpy as np
    import tempfile
    d = tempfile.mkdtemp()
    for name, env in (("base", {}), ("fused", {"WKB200_FUSED": "1"})):
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, __file__, "child", os.path.join(d, name + ".npz")], env=e, check=True, timeout=100)
    a, b = np.load(os.path.join(d, "base.npz")), np.load(os.path.join(d, "fused.npz"))
    ok = True
    for k in a.files:
        same = np.array_equal(a[k], b[k])
        print(k, "identical" if same else f"DIFFERENT (max abs diff {np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)).max():.3e})")
        ok &= same
    sys.exit(0 if ok else 1)
