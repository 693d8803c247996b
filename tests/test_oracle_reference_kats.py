"""Pins the oracle's host logic to the reference's own known-answer tests
(UnitTests.swift:1982-2115 filters, :816-878 DecodingFallback, :695-705 compression ratio)."""
import numpy as np
import pytest

from oracle import decode_ref as D
from tests import kat_vectors as K


def f16(v):
    # the reference's test helper builds f16 logits (TestUtils.swift:110-119)
    return np.array(v, dtype=np.float16).astype(np.float32)


def eq(a, b):
    np.testing.assert_array_equal(np.asarray(a, np.float32), f16(b))


@pytest.mark.parametrize("name,sup,logits,tokens,exp", K.SUPPRESS_TOKENS)
def test_suppress_tokens(name, sup, logits, tokens, exp):
    eq(D.SuppressTokensFilter(sup).filterLogits(f16(logits), tokens), exp)


@pytest.mark.parametrize("name,eot,ws,sb,logits,tokens,exp", K.SUPPRESS_BLANK)
def test_suppress_blank(name, eot, ws, sb, logits, tokens, exp):
    st = D.SpecialTokens.test_default(endToken=eot, whitespaceToken=ws)
    eq(D.SuppressBlankFilter(st, sb).filterLogits(f16(logits), tokens), exp)


@pytest.mark.parametrize("name,langs,dim,sb,logits,tokens,exp", K.LANGUAGE)
def test_language_filter(name, langs, dim, sb, logits, tokens, exp):
    eq(D.LanguageLogitsFilter(langs, dim, sb).filterLogits(f16(logits), tokens), exp)


@pytest.mark.parametrize("name,multi,sb,logits,tokens,exp", K.TIMESTAMP_RULES)
def test_timestamp_rules(name, multi, sb, logits, tokens, exp):
    st = D.SpecialTokens.test_default(**K.TS_SPECIAL)
    f = D.TimestampRulesFilter(st, sampleBegin=sb, maxInitialTimestampIndex=None, isModelMultilingual=multi)
    eq(f.filterLogits(f16(logits), tokens), exp)


def test_decoding_fallback_order():
    o = D.DecodingOptions(compressionRatioThreshold=-1.0, logProbThreshold=-1.0, noSpeechThreshold=-1.0)
    fb = D.DecodingFallback.make(o, True, 0, 0, -2.0)
    assert fb.fallbackReason == "firstTokenLogProbThreshold" and fb.needsFallback
    fb = D.DecodingFallback.make(o, False, 0, 0, -2.0)
    assert fb.fallbackReason == "silence" and not fb.needsFallback
    o = D.DecodingOptions(compressionRatioThreshold=-1.0, logProbThreshold=-1.0, noSpeechThreshold=0.0)
    fb = D.DecodingFallback.make(o, False, 0, 0, -2.0)
    assert fb.fallbackReason == "compressionRatioThreshold" and fb.needsFallback
    o = D.DecodingOptions(compressionRatioThreshold=0.0, logProbThreshold=-1.0, noSpeechThreshold=0.0)
    fb = D.DecodingFallback.make(o, False, 0, 0, -2.0)
    assert fb.fallbackReason == "logProbThreshold" and fb.needsFallback
    o = D.DecodingOptions(compressionRatioThreshold=0.0, logProbThreshold=0.0, noSpeechThreshold=0.0)
    assert D.DecodingFallback.make(o, False, 0, 0, 0) is None


def test_compression_ratio_ordering():
    u = D.compression_ratio(list(range(1, 11)))
    r = D.compression_ratio([1] * 10)
    rl = D.compression_ratio([1] * 20)
    assert u < r < rl


def test_filter_composition_order():
    # createLogitsFilters: custom -> SuppressBlank -> SuppressTokens -> TimestampRules (TextDecoder.swift:857-899)
    st = D.SpecialTokens()
    o = D.DecodingOptions(suppressBlank=True, suppressTokens=[1, 2, 60000])
    fs = D.createLogitsFilters(o, 0, 4, st, True, custom=["custom"])
    assert fs[0] == "custom"
    assert isinstance(fs[1], D.SuppressBlankFilter) and isinstance(fs[2], D.SuppressTokensFilter)
    assert isinstance(fs[3], D.TimestampRulesFilter)
    assert fs[2].suppressTokens == [1, 2]  # >= specialTokenBegin dropped
    assert fs[3].sampleBegin == 4 and fs[1].sampleBegin == 0
    fs = D.createLogitsFilters(D.DecodingOptions(withoutTimestamps=True), 0, 4, st, True)
    assert fs == []


def test_prefill_prompt():
    st = D.SpecialTokens()
    assert D.prefill_prompt(D.DecodingOptions(), st, True) == [50258, 50259, 50359, 50364]
    assert D.prefill_prompt(D.DecodingOptions(withoutTimestamps=True, task="translate"), st, True) == [50258, 50259, 50358, 50363]
    assert D.prefill_prompt(D.DecodingOptions(), st, False) == [50258, 50364]
    assert D.prefill_prompt(None, st, True) == [50258]
    p = D.prefill_prompt(D.DecodingOptions(promptTokens=[1, 2, 50300], prefixTokens=[7, 8]), st, True)
    assert p == [50361, 1, 2, 50258, 50259, 50359, 50364, 7, 8]


def test_decode_loop_quirks():
    """Scripted logits exercise the decodeText state machine (TextDecoder.swift:566-686)."""
    st = D.SpecialTokens.toy(64)  # sb=32: eot 32, sot 33, ts begin 41
    V = 64
    o = D.DecodingOptions(firstTokenLogProbThreshold=None, sampleLength=10)
    prompt = [st.startOfTranscriptToken, st.englishToken, st.transcribeToken, st.timeTokenBegin]
    calls = []

    def predict(tok, idx):
        calls.append((tok, idx))
        lg = np.full(V, -5.0, np.float32)
        script = {0: 5, 1: 6, 2: st.timeTokenBegin + 3, 3: 7, 4: 8, 5: st.endToken}
        lg[script[idx]] = 5.0
        return lg

    r = D.decode_text(predict, prompt, o, st, True)
    # model predicted a timestamp at the last prefill slot -> replaces forced <|0.00|> (:581-594)
    assert calls[:4] == [(prompt[0], 0), (prompt[1], 1), (prompt[2], 2), (st.timeTokenBegin + 3, 3)]
    assert r.currentTokens == [prompt[0], prompt[1], prompt[2], st.timeTokenBegin + 3, 7, 8]
    assert r.tokens == r.currentTokens + [st.endToken]
    assert r.steps == 6 and len(r.tokenLogProbs) == 7 and r.tokenLogProbs[-1] == 0.0
    # first-token threshold terminates at step 0 (:662-671)
    r2 = D.decode_text(lambda t, i: np.zeros(V, np.float32), prompt, D.DecodingOptions(), st, True)
    assert r2.steps == 1 and r2.isFirstTokenLogProbTooLow and r2.fallback.fallbackReason == "firstTokenLogProbThreshold"
    # sampleLength bound: loopCount = min(sampleLength, 223)
    r3 = D.decode_text(lambda t, i: np.eye(V, dtype=np.float32)[5] * 9, prompt,
                       D.DecodingOptions(firstTokenLogProbThreshold=None, sampleLength=7, withoutTimestamps=True), st, True)
    assert r3.steps == 7 and len(r3.currentTokens) == 4 + 4
