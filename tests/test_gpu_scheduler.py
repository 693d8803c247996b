"""The batched entry (wk_transcribe_windows_ex) as the reference's transcribeWithOptions (WhisperKit.swift:716-812): independent windows,
per-item options and prompts, per-item Result, the progress callback / early stop (TextDecoder.swift:724-762), windows ending at their
own length with their decode slot handed to the next window, and concurrent sessions on one model from several host threads."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import whisperkit_b200 as wk  # noqa: E402
from oracle import decode_ref as D  # noqa: E402
from oracle import mel_ref  # noqa: E402


def make_kit(slots, seed=4, model="toy", **kw):
    vocab = 1024 if model == "toy" else 2048
    return wk.WhisperKit(wk.WhisperKitConfig(model=model, maxBatch=slots, seed=seed,
                                             specialTokens=wk.SpecialTokens.from_any(D.SpecialTokens.toy(vocab)), **kw))


def base_opts(**kw):
    d = dict(firstTokenLogProbThreshold=None, sampleLength=20, temperatureFallbackCount=0)
    d.update(kw)
    return wk.DecodingOptions(**d)


def test_windows_with_their_own_options_retire_and_hand_over_slots():
    """7 windows through 3 decode slots, every window with its own sampleLength / timestamps / suppress list / prompt tokens: each result
    equals that window transcribed alone with its options (slots are re-used as soon as a window ends; encoder chunks of 2 and of 3
    windows overlap the running decode)."""
    kit = make_kit(3)
    n = 7
    pcm = np.stack([mel_ref.synthetic_pcm(70 + i) for i in range(n)])
    lens = [5, 30, 12, 3, 25, 9, 17]
    opts = [base_opts(sampleLength=lens[i], withoutTimestamps=(i % 3 == 1), suppressTokens=[3, 4, 5] if i % 2 else [],
                      promptTokens=[9, 8, 7] if i == 4 else None, suppressBlank=(i == 2)) for i in range(n)]
    alone = [kit.transcribe(pcm[i], opts[i])[0] for i in range(n)]
    assert len({r.steps for r in alone}) > 3            # the windows really end at different lengths
    for chunk in (0, 2, 3):
        got = kit.transcribe(pcm, opts, encoderChunk=chunk)
        for i in range(n):
            assert got[i].tokens == alone[i].tokens, (chunk, i)
            assert got[i].steps == alone[i].steps
            np.testing.assert_allclose(got[i].tokenLogProbs, alone[i].tokenLogProbs, atol=1e-5)
    # the same through decodeText on bound windows (per-window options, per-window prompts)
    fe, enc, dec = kit.featureExtractor, kit.audioEncoder, kit.textDecoder
    enc_t = enc.encodeFeatures(fe.logMelSpectrogram(pcm[:3]))
    prompts = [dec.prefillDecoderInputs(opts[i], kit.specialTokens) for i in range(3)]
    res = dec.decodeText(enc_t, prompts, opts[:3], kit.specialTokens)
    assert [r.tokens for r in res] == [alone[i].tokens for i in range(3)]


def test_per_window_result_isolates_a_bad_item():
    """One window with an unusable prompt fails alone (the reference captures per-item errors as Result, WhisperKit.swift:775-790); without
    a status array the call fails as a whole."""
    kit = make_kit(2)
    pcm = np.stack([mel_ref.synthetic_pcm(80 + i) for i in range(4)])
    o = base_opts(sampleLength=8)
    good = kit.transcribe(pcm, o)
    import ctypes as C
    from whisperkit_b200._lib import check, wk_decode_result
    from whisperkit_b200.api import make_batch_opts, _ptr
    prompts = [kit.textDecoder.prefillDecoderInputs(o, kit.specialTokens) for _ in range(4)]
    prompts[2] = [5, 99999]                                            # token outside the vocabulary
    status = (C.c_int32 * 4)()
    bo, keep = make_batch_opts(4, o, prompts, status=status)
    res = (wk_decode_result * 4)()
    st = kit.specialTokens.to_c()
    check(kit.model.lib.wk_transcribe_windows_ex(kit.model.handle, kit.textDecoder.handle, _ptr(pcm), 4, 480000, None, C.byref(st), C.byref(bo), res))
    assert list(status) == [0, 0, -4, 0]                               # prepareDecoderInputsFailed for window 2 only
    for i in (0, 1, 3):
        assert list(res[i].tokens[:res[i].n_tokens]) == good[i].tokens
    bo2, keep2 = make_batch_opts(4, o, prompts)
    with pytest.raises(wk.WhisperError) as ei:
        check(kit.model.lib.wk_transcribe_windows_ex(kit.model.handle, kit.textDecoder.handle, _ptr(pcm), 4, 480000, None, C.byref(st), C.byref(bo2), res))
    assert ei.value.case == "prepareDecoderInputsFailed"
    # samples_per_window out of range: audioProcessingFailed for that window only
    out = kit.transcribe(pcm, o, samplesPerWindow=[480000, 480001, 100, 0], returnErrors=True)
    assert isinstance(out[1], wk.WhisperError) and out[1].case == "audioProcessingFailed"
    assert out[0].tokens == good[0].tokens and not isinstance(out[2], wk.WhisperError)


def test_progress_callback_and_early_stop():
    """TranscriptionCallback: called with the live windows' tokens while they decode; returning False stops that window early
    (TextDecoder.swift:733-762) - its result is what it had, the other windows are untouched."""
    kit = make_kit(3)
    pcm = np.stack([mel_ref.synthetic_pcm(90 + i) for i in range(3)])
    o = base_opts(sampleLength=40)
    full = kit.transcribe(pcm, o)
    seen = {}

    def cb(window, tokens, avg):
        seen.setdefault(window, []).append(len(tokens))
        return not (window == 1 and len(tokens) >= 10)

    got = kit.transcribe(pcm, o, callback=cb, callbackEvery=4)
    assert got[0].tokens == full[0].tokens and got[2].tokens == full[2].tokens
    assert got[1].steps < full[1].steps and got[1].tokens[:-1] == full[1].tokens[:len(got[1].tokens) - 1]
    assert got[1].tokens[-1] == kit.specialTokens.endToken            # sampler.finalize appended EOT
    assert all(b > a for a, b in zip(seen[0], seen[0][1:])) and len(seen[0]) >= 3


def test_two_host_threads_share_one_model():
    """Up to concurrentWorkerCount tasks call the same protocol objects concurrently (WhisperKit.swift:735-791): one wk_model, one session
    per thread, each with its own streams, encoder workspace and KV caches.  Results equal the serial runs; the piecewise entry points
    (wk_mel / wk_encode, serialised inside the library) are hammered from both threads too."""
    kit = make_kit(4, model="toy128")
    model = kit.model
    st = kit.specialTokens
    pcm = [np.stack([mel_ref.synthetic_pcm(100 + 10 * t + i) for i in range(6)]) for t in range(2)]
    o = base_opts(sampleLength=16)
    serial = [kit.transcribe(pcm[t], o) for t in range(2)]
    serial_enc = [kit.audioEncoder.encodeFeatures(kit.featureExtractor.logMelSpectrogram(pcm[t][:2])).numpy() for t in range(2)]
    import ctypes as C
    from whisperkit_b200._lib import check, wk_decode_result
    from whisperkit_b200.api import make_batch_opts, _ptr
    decs = [wk.TextDecoder(model, 4) for _ in range(2)]
    out, errs = [None, None], []

    def worker(t):
        try:
            for rep in range(3):
                bo, keep = make_batch_opts(6, o, None)
                res = (wk_decode_result * 6)()
                stc = st.to_c()
                check(model.lib.wk_transcribe_windows_ex(model.handle, decs[t].handle, _ptr(pcm[t]), 6, 480000, None, C.byref(stc), C.byref(bo), res))
                out[t] = [list(r.tokens[:r.n_tokens]) for r in res]
                e = wk.AudioEncoder(model).encodeFeatures(wk.FeatureExtractor(model).logMelSpectrogram(pcm[t][:2])).numpy()
                np.testing.assert_array_equal(e, serial_enc[t])
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join(timeout=300)
    assert not errs, errs
    for t in range(2):
        assert out[t] == [r.tokens for r in serial[t]]
    for d_ in decs:
        d_.close()


def test_tensor_ownership_and_strided_readback():
    """A second logMelSpectrogram / encodeFeatures must not disturb a tensor the host still holds (the reference returns owned
    MLMultiArrays); readback into rows with padding goes through explicit strides."""
    kit = make_kit(2)
    fe, enc = kit.featureExtractor, kit.audioEncoder
    a = np.stack([mel_ref.synthetic_pcm(1), mel_ref.synthetic_pcm(2)])
    b = np.stack([mel_ref.synthetic_pcm(3), mel_ref.synthetic_pcm(4)])
    mel_a = fe.logMelSpectrogram(a)
    enc_a = enc.encodeFeatures(mel_a)
    ref_mel, ref_enc = mel_a.numpy(), enc_a.numpy()
    mel_b = fe.logMelSpectrogram(b)
    enc_b = enc.encodeFeatures(mel_b)
    assert np.abs(enc_b.numpy() - ref_enc).max() > 0
    np.testing.assert_array_equal(mel_a.numpy(), ref_mel)              # still the first window pair
    np.testing.assert_array_equal(enc_a.numpy(), ref_enc)
    np.testing.assert_array_equal(enc_a.numpy(row_pad=12), ref_enc)
    np.testing.assert_array_equal(mel_a.numpy(row_pad=8), ref_mel)
    dec = kit.textDecoder
    o = base_opts(sampleLength=6)
    prompt = dec.prefillDecoderInputs(o, kit.specialTokens)
    r1 = dec.decodeText(enc_a, prompt, o, kit.specialTokens)
    enc_a.close()                                                       # released while nothing reads it any more
    r2 = dec.decodeText(None, prompt, o, kit.specialTokens)            # cross K/V stays bound
    assert [x.tokens for x in r1] == [x.tokens for x in r2]


def test_language_option_is_resolved_through_the_tokenizer():
    """DecodingOptions.language goes through tokenizer.convertTokenToId like prefillDecoderInputs (TextDecoder.swift:181-186); without a
    tokenizer it is an error, not a silent <|en|>."""
    kit = make_kit(2)
    pcm = mel_ref.synthetic_pcm(5)
    with pytest.raises(wk.WhisperError):
        kit.transcribe(pcm, base_opts(language="xx"))
    from tests.test_gpu_pipeline import _toy_tokenizer
    kit.tokenizer = _toy_tokenizer(1024)
    xx = kit.tokenizer.convertTokenToId("<|xx|>")
    r_lang = kit.transcribe(pcm, base_opts(language="xx", sampleLength=4))[0]
    r_tok = kit.transcribe(pcm, base_opts(languageToken=xx, sampleLength=4))[0]
    assert r_lang.tokens == r_tok.tokens and r_lang.tokens[1] == xx
