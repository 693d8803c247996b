"""Beam search on the GPU against oracle/beam_ref.py - SELF-ORACLE parity: the reference's BeamSearchTokenSampler is an
unimplemented stub (TokenSampler.swift:254-290), so the specification is openai/whisper's BeamSearchDecoder inside WhisperKit's decodeText
loop, as restated in the oracle.  The oracle loop consumes the GPU decoder's own logits (predictLogits on explicit token prefixes), so token
IDs must match bit for bit; log-probs and the ranking score agree to 5e-4 under the f16 policy and 2e-3 under bf16.  Why not tighter: the
oracle's logits come from single-row decodes, whose cross-attention keeps q and P in f32, while the beam kernel carries them through the
tensor cores as hi + lo 16-bit pairs (~2^-17 relative).  That difference alone is invisible, but every attention output is then rounded
to the policy's 16-bit storage type, and a value that lands on the other side of a rounding boundary moves by one storage ulp (2^-8
relative for bf16, 2^-11 for f16) - about 1 % of the elements do - so two correct GPU paths agree only to the policy's own precision."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import whisperkit_b200 as wk  # noqa: E402
from oracle import beam_ref as BR  # noqa: E402
from oracle import decode_ref as D  # noqa: E402
from oracle import mel_ref  # noqa: E402


def _predictor(model, pcm_window, beam):
    """Logits of `beam` arbitrary token prefixes of ONE window, from the GPU decoder: the window is encoded `beam` times so that every
    row attends to its own copy of the cross K/V, and each call replays the prefixes from position 0."""
    fe, enc = wk.FeatureExtractor(model), wk.AudioEncoder(model)
    dec = wk.TextDecoder(model, beam)
    dec.bindEncoderOutput(enc.encodeFeatures(fe.logMelSpectrogram(np.repeat(pcm_window[None], beam, axis=0))))

    def predict(prefixes, tokenIndex):
        lg = None
        for t in range(tokenIndex + 1):
            lg = dec.predictLogits([p[t] for p in prefixes], [t] * beam)
        return lg
    return predict, dec


@pytest.mark.parametrize("variant,policy,beam,patience,without_ts", [("toy128", "f16", 5, 1.0, False), ("toy", "bf16", 3, 2.0, False),
                                                                     ("toy128", "f16", 2, 1.0, True)])
def test_beam_search_matches_the_oracle_on_gpu_logits(variant, policy, beam, patience, without_ts):
    vocab = 1024 if variant == "toy" else 2048
    st_o = D.SpecialTokens.toy(vocab)
    st = wk.SpecialTokens.from_any(st_o)
    n_win = 3
    kit = wk.WhisperKit(wk.WhisperKitConfig(model=variant, maxBatch=2 * beam, seed=17, specialTokens=st, dtype=policy))   # 2 windows in flight
    pcm = np.stack([mel_ref.synthetic_pcm(600 + i) for i in range(n_win)])
    kw = dict(firstTokenLogProbThreshold=None, sampleLength=22, withoutTimestamps=without_ts, temperatureFallbackCount=0,
              logProbThreshold=None, compressionRatioThreshold=None)
    o_gpu = wk.DecodingOptions(beamSize=beam, beamPatience=patience, **kw)
    o_ref = D.DecodingOptions(**kw)
    res = kit.transcribe(pcm, o_gpu)                     # 3 windows through 2 beam groups: the third is admitted when one ends
    prompt = kit.textDecoder.prefillDecoderInputs(o_gpu, st)
    greedy = kit.transcribe(pcm, wk.DecodingOptions(**kw))
    differs = 0
    for b in range(n_win):
        predict, dec = _predictor(kit.model, pcm[b], beam)
        ref = BR.decode_text_beam(predict, prompt, o_ref, st_o, True, beam, patience)
        dec.close()
        assert res[b].tokens == ref.tokens, (b, res[b].tokens, ref.tokens)
        atol = 2e-3 if policy == "bf16" else 5e-4
        np.testing.assert_allclose(res[b].tokenLogProbs, ref.tokenLogProbs, atol=atol)
        assert abs(res[b].avgLogProb - ref.avgLogProb) < atol and res[b].steps == ref.steps
        differs += res[b].tokens != greedy[b].tokens
    print(f"[{variant}/{policy} beam {beam} patience {patience}] windows whose beam result differs from greedy: {differs} of {n_win}")


def test_beam_rows_share_one_cross_kv_block_and_batch_independence():
    """5 windows through a session of 2 x 4 rows: every window's result equals the window decoded alone (beam groups are independent; the
    `beam` rows of a group read one shared cross K/V block)."""
    st = wk.SpecialTokens.from_any(D.SpecialTokens.toy(1024))
    kit = wk.WhisperKit(wk.WhisperKitConfig(model="toy", maxBatch=8, seed=23, specialTokens=st))
    pcm = np.stack([mel_ref.synthetic_pcm(650 + i) for i in range(5)])
    o = wk.DecodingOptions(beamSize=4, firstTokenLogProbThreshold=None, sampleLength=16, temperatureFallbackCount=0)
    allr = kit.transcribe(pcm, o)
    for b in range(5):
        alone = kit.transcribe(pcm[b], o)[0]
        assert alone.tokens == allr[b].tokens, b
    with pytest.raises(wk.WhisperError):
        kit.transcribe(pcm, wk.DecodingOptions(beamSize=4, wordTimestamps=True))
    with pytest.raises(wk.WhisperError):
        kit.transcribe(pcm, wk.DecodingOptions(beamSize=9))
