"""The beam-search oracle (oracle/beam_ref.py) on a scripted model: hand-checkable cases of the BeamSearchDecoder semantics it restates
(whisper/decoding.py) - a beam of width 1 is greedy, the finished list fills by score order up to maxCandidates = Int(beam * patience)
(TokenSampler.swift:266), the ranker divides by the number of sampled tokens."""
import numpy as np

from oracle import beam_ref as BR
from oracle import decode_ref as D

V = 16
ST = D.SpecialTokens.test_default(endToken=15, startOfTranscriptToken=14, specialTokenBegin=14, timeTokenBegin=100, noTimestampsToken=13,
                                  transcribeToken=12, translateToken=11, englishToken=10, whitespaceToken=9)
OPTS = dict(firstTokenLogProbThreshold=None, withoutTimestamps=True, sampleLength=8, logProbThreshold=None, compressionRatioThreshold=None)


def scripted(table):
    """logits depend only on the last token of each prefix: table[last] -> {token: logit}"""
    def predict(prefixes, idx):
        out = np.full((len(prefixes), V), -20.0, np.float32)
        for j, p in enumerate(prefixes):
            for t, v in table.get(p[idx], {}).items():
                out[j, t] = v
        return out
    return predict


def test_beam_one_is_greedy():
    table = {14: {1: 2.0, 2: 1.0}, 1: {3: 1.0, 15: 0.5}, 3: {15: 3.0}, 2: {15: 1.0}}
    o = D.DecodingOptions(**OPTS)
    prompt = [14]
    ref = D.decode_text(lambda tok, idx: scripted(table)([[0] * idx + [tok]], idx)[0], prompt, o, ST, False)
    got = BR.decode_text_beam(scripted(table), prompt, o, ST, False, 1)
    assert got.tokens == ref.tokens == [14, 1, 3, 15]
    np.testing.assert_allclose(got.tokenLogProbs, ref.tokenLogProbs, atol=1e-5)


def test_beam_finds_the_sequence_greedy_misses_and_ranks_by_mean_logprob():
    # greedy takes 1 (p .6) then is stuck with flat continuations; token 2 (p .4) leads to a certain continuation
    l6, l4 = np.log(0.6), np.log(0.4)
    table = {14: {1: l6, 2: l4}, 1: {3: np.log(0.25), 4: np.log(0.25), 5: np.log(0.25), 6: np.log(0.25)},
             2: {7: 0.0}, 7: {15: 0.0}, 3: {15: 0.0}, 4: {15: 0.0}, 5: {15: 0.0}, 6: {15: 0.0}}
    o = D.DecodingOptions(**OPTS)
    greedy = BR.decode_text_beam(scripted(table), [14], o, ST, False, 1)
    beam = BR.decode_text_beam(scripted(table), [14], o, ST, False, 3)
    assert greedy.tokens[:2] == [14, 1]
    assert beam.tokens == [14, 2, 7, 15]
    # sum of log-probs / sampled tokens before EOT: (log .4 + ~0) / 2 beats (log .6 + log .25) / 2
    assert abs(sum(beam.tokenLogProbs) - l4) < 1e-2


def test_patience_sets_the_number_of_finished_candidates():
    table = {14: {1: 0.0, 2: -0.1, 3: -0.2}, 1: {15: 0.0, 4: -3.0}, 2: {15: 0.0, 4: -3.0}, 3: {15: 0.0, 4: -3.0}, 4: {15: 0.0}}
    o = D.DecodingOptions(**OPTS)
    t1, t2 = [], []
    BR.decode_text_beam(scripted(table), [14], o, ST, False, 2, 1.0, trace=t1)
    BR.decode_text_beam(scripted(table), [14], o, ST, False, 2, 2.0, trace=t2)
    assert t1[-1]["finished"] == 2 and t2[-1]["finished"] == 4       # Int(2 * 1) and Int(2 * 2) finished sequences end the search
    assert len(t2) > len(t1)
