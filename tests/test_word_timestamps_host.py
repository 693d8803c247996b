"""Word-timestamp host logic, no GPU:
  * the oracle (oracle/words_ref.py) against the reference's own known-answer tests (UnitTests.swift:2336-2960);
  * libwkb200's C++ implementation (csrc/wordtiming.cu, through the C ABI) against the same KATs and, on random inputs,
    bit-for-bit against the oracle."""
import numpy as np
import pytest

from oracle import words_ref as W
from whisperkit_b200 import wordtiming as T

WT = W.WordTiming
SEEKER = T.WordTimingSeeker()


def c_words(ws):
    return [T.WordTiming(w.word, list(w.tokens), w.start, w.end, w.probability) for w in ws]


def same_words(got, ref, tol=0.0):
    assert [g.word for g in got] == [r.word for r in ref]
    assert [list(g.tokens) for g in got] == [list(r.tokens) for r in ref]
    for k in ("start", "end", "probability"):
        a, b = np.float32([getattr(g, k) for g in got]), np.float32([getattr(r, k) for r in ref])
        if tol:
            np.testing.assert_allclose(a, b, atol=tol)
        else:
            np.testing.assert_array_equal(a, b)


BOTH = [("oracle", W.merge_punctuations), ("c_abi", lambda al, p=None, a=None: SEEKER.mergePunctuations(c_words(al), p, a))]


def test_dtw_reference_kats():
    m = [[1.0, 1.0, 1.0], [5.0, 2.0, 1.0], [1.0, 5.0, 2.0]]          # testDynamicTimeWarpingSimpleMatrix
    exp = ([0, 1, 1, 2, 2], [0, 0, 1, 1, 2])
    assert W.dynamic_time_warping(m) == exp
    assert SEEKER.dynamicTimeWarping(m) == exp
    # testDynamicTimeWarpingLargeMatrix: 448 x 1500 Float16 random matrix -> path properties
    rng = np.random.default_rng(0)
    big = rng.random((448, 1500)).astype(np.float16)
    ti, tj = SEEKER.dynamicTimeWarping(big)
    assert (ti[0], tj[0]) == (0, 0) and (ti[-1], tj[-1]) == (447, 1499)
    d = np.diff(np.stack([ti, tj]), axis=1)
    assert set(np.unique(d)) <= {0, 1} and np.all(d.sum(0) >= 1)
    with pytest.raises(Exception):
        SEEKER.dynamicTimeWarping(np.zeros((3, 4, 5), np.float32))   # "Invalid alignment matrix shape"


def test_dtw_matches_oracle_bit_for_bit():
    rng = np.random.default_rng(1)
    for rows, cols, dt in ((1, 1, np.float32), (7, 3, np.float32), (3, 40, np.float16), (37, 150, np.float16), (60, 90, np.float32)):
        m = rng.random((rows, cols)).astype(dt)
        if rows == 37:
            m[:] = np.round(m * 4) / 4                                # many exact ties: the strict '<' tie rules decide
        assert SEEKER.dynamicTimeWarping(m) == W.dynamic_time_warping(m)


@pytest.mark.parametrize("name,merge", BOTH)
def test_merge_punctuations_reference_kats(name, merge):
    assert merge([]) == []                                            # testMergePunctuationsWithEmptyInput
    en = [WT("<|0.00|>", [50364], 0, 1, 1), WT(" Hello", [2425], 1, 2, 1), WT(",", [11], 2, 3, 1), WT(" world", [1002], 3, 4, 1),
          WT("!", [0], 4, 5, 1), WT("<|1.00|>", [50414], 5, 6, 1), WT("<|1.00|>", [50414], 6, 7, 1), WT(" This", [639], 7, 8, 1),
          WT(" is", [307], 8, 9, 1), WT(" a", [257], 9, 10, 1), WT(" test", [220, 31636], 10, 11, 1), WT(",", [11], 11, 12, 1),
          WT(" isn't", [1943, 380], 12, 13, 1), WT(" it", [309], 13, 14, 1), WT("?", [30], 14, 15, 1), WT("<|endoftext|>", [50257], 15, 16, 1)]
    exp = [WT("<|0.00|>", [50364], 0, 1, 1), WT(" Hello,", [2425, 11], 1, 2, 1), WT(" world!", [1002, 0], 3, 4, 1),
           WT("<|1.00|>", [50414], 5, 6, 1), WT("<|1.00|>", [50414], 6, 7, 1), WT(" This", [639], 7, 8, 1), WT(" is", [307], 8, 9, 1),
           WT(" a", [257], 9, 10, 1), WT(" test,", [220, 31636, 11], 10, 11, 1), WT(" isn't", [1943, 380], 12, 13, 1),
           WT(" it?", [309, 30], 13, 14, 1), WT("<|endoftext|>", [50257], 15, 16, 1)]
    same_words(merge(en, "\"'“¿([{-", "\"'.。,，!！?？:：”)]}、"), exp)     # testMergePunctuations
    es = [WT("<|notimestamps|>", [50363], 0, 1, 1), WT(" ¡", [24364], 0, 1, 1), WT("Hola", [48529], 1, 2, 1), WT(" Mundo", [376, 6043], 2, 3, 1),
          WT("!", [0], 3, 4, 1), WT(" Esta", [20547], 4, 5, 1), WT(" es", [785], 5, 6, 1), WT(" una", [2002], 6, 7, 1),
          WT(" prueba", [48241], 7, 8, 1), WT(",", [11], 8, 9, 1), WT(" ¿", [3841], 9, 10, 1), WT("no", [1771], 10, 11, 1),
          WT("?", [30], 11, 12, 1), WT("<|endoftext|>", [50257], 12, 13, 1)]
    exp = [WT("<|notimestamps|>", [50363], 0, 1, 1), WT(" ¡Hola", [24364, 48529], 1, 2, 1), WT(" Mundo!", [376, 6043, 0], 2, 3, 1),
           WT(" Esta", [20547], 4, 5, 1), WT(" es", [785], 5, 6, 1), WT(" una", [2002], 6, 7, 1), WT(" prueba,", [48241, 11], 7, 8, 1),
           WT(" ¿no?", [3841, 1771, 30], 10, 11, 1), WT("<|endoftext|>", [50257], 12, 13, 1)]
    same_words(merge(es), exp)                                        # testMergePunctuationsSpanish (default punctuation sets)
    es2 = [WT(" ¿", [1201], 0, 1, 1), WT("Que", [1202], 1, 2, 0.9), WT(" pasa", [1203], 2, 3, 1), WT(" mundo", [1204], 3, 4, 0.6),
           WT("?", [1205], 4, 5, 0.4)]
    exp = [WT(" ¿Que", [1201, 1202], 1, 2, 0.9), WT(" pasa", [1203], 2, 3, 1), WT(" mundo?", [1204, 1205], 3, 4, 0.6)]
    same_words(merge(es2), exp)                                       # testMergePunctuationsSpanishStartWithPrepend
    ja = [WT("<|0.00|>", [50364], 0, 1, 1), WT("こんにちは", [38088], 1, 2, 1), WT("、", [1231], 2, 3, 1), WT("世界", [24486], 3, 4, 1),
          WT("！", [171, 120, 223], 4, 5, 1), WT("これは", [25212], 5, 6, 1), WT("テ", [22985], 6, 7, 1), WT("スト", [40498], 7, 8, 1),
          WT("です", [4767], 8, 9, 1), WT("よね", [30346], 9, 10, 1), WT("？", [171, 120, 253], 10, 11, 1), WT("<|endoftext|>", [50257], 11, 12, 1)]
    exp = [WT("<|0.00|>", [50364], 0, 1, 1), WT("こんにちは、", [38088, 1231], 1, 2, 1), WT("世界！", [24486, 171, 120, 223], 3, 4, 1),
           WT("これは", [25212], 5, 6, 1), WT("テ", [22985], 6, 7, 1), WT("スト", [40498], 7, 8, 1), WT("です", [4767], 8, 9, 1),
           WT("よね？", [30346, 171, 120, 253], 9, 10, 1), WT("<|endoftext|>", [50257], 11, 12, 1)]
    same_words(merge(ja), exp)                                        # testMergePunctuationsJapanese


def _long_word_case():
    words = [WT(" The", [264], 0.5, 1.0, 1), WT(" first", [4589], 1.0, 2.0, 1), WT(" segment", [234], 2.0, 3.0, 1), WT(" with", [567], 3.0, 4.0, 1),
             WT(" a", [257], 4.0, 5.0, 1), WT(" long", [890], 5.0, 6.0, 1), WT(" ending", [123], 6.0, 35.0, 1), WT(".", [13], 35.0, 35.0, 1)]
    segs = [W.Segment(0.0, 6.0, [264, 4589, 234, 567, 257, 890], [0.0] * 6), W.Segment(6.5, 30.0, [123, 13], [0.0] * 2, id=1)]
    return words, segs


def _run_pipeline(impl, words, segs):
    if impl == "oracle":
        med, mx = W.calculate_word_duration_constraints(words)
        merged = W.merge_punctuations(W.truncate_long_words_at_sentence_boundaries(words, mx))
        upd = W.update_segments_with_word_timings(segs, merged, 0, 0.0, med, mx, 50257)
        return med, mx, [(s.start, s.end, s.words) for s in upd]
    med, mx = SEEKER.calculateWordDurationConstraints(c_words(words))
    merged = SEEKER.mergePunctuations(SEEKER.truncateLongWordsAtSentenceBoundaries(c_words(words), mx))
    return med, mx, SEEKER.updateSegmentsWithWordTimings(segs, merged, 0, 0.0, med, mx, 50257)


@pytest.mark.parametrize("impl", ["oracle", "c_abi"])
def test_long_word_durations_reference_kat(impl):
    """testLongWordDurations (UnitTests.swift:2760-2860)."""
    words, segs = _long_word_case()
    med, mx, upd = _run_pipeline(impl, words, segs)
    assert np.float32(med) == np.float32(0.7) and np.float32(mx) == np.float32(1.4)
    allw = [w for (_, _, ws) in upd for w in ws]
    assert len(upd) == 2
    assert abs(allw[-1].duration - mx) < 1e-4
    assert upd[-1][1] - upd[-1][0] <= 19.5
    i = [w.word for w in allw].index(" ending.")
    assert abs(allw[i].duration - mx) < 1e-4
    assert np.float32(allw[i].start) == np.float32(33.6)
    for a, b in zip(allw[:-1], allw[1:]):
        assert a.end <= b.start


@pytest.mark.parametrize("impl", ["oracle", "c_abi"])
def test_single_token_segment_reference_kat(impl):
    """testSingleTokenSegmentWordDuration (UnitTests.swift:2862-2925)."""
    words = [WT("<|notimestamps|>", [50363], 0, 0.5, 1), WT(" Hello", [314], 0.5, 20.5, 1), WT("<|endoftext|>", [50257], 20.5, 30, 1)]
    segs = [W.Segment(0.0, 30.0, [314], [0.0])]
    med, mx, upd = _run_pipeline(impl, words, segs)
    assert np.float32(med) == np.float32(0.7) and np.float32(mx) == np.float32(1.4)
    ws = upd[0][2]
    assert [w.word for w in ws] == [" Hello"] and ws[0].duration <= mx + 1e-6
    prev = 0.0
    for w in ws:
        assert w.start >= prev and w.duration <= mx + 1e-6
        prev = w.end


def _fake_split(tokens, special_begin=1000):
    """Stand-in for WhisperTokenizer.splitToWordTokens: token t spells chr(97 + t % 26); a token divisible by 3 starts a word with a space,
    t % 17 == 0 is a lone ',', special tokens are words of their own."""
    words, groups = [], []
    for t in tokens:
        if t >= special_begin:
            words.append(f"<|{t}|>"); groups.append([t])
        elif t % 17 == 0:
            words.append(","); groups.append([t])
        elif t % 3 == 0 or not words or groups[-1][0] >= special_begin:
            words.append(" " + chr(97 + t % 26)); groups.append([t])
        else:
            words[-1] += chr(97 + t % 26); groups[-1].append(t)
    return words, groups


def test_find_alignment_kat_and_oracle():
    """testFindAlignment (UnitTests.swift:2408-2482): probability = exp(mean log prob), monotone timings; plus C == oracle."""
    rng = np.random.default_rng(2)
    ids = [400, 370, 452, 7177, 6280, 11, 1029, 406, 437, 428, 1941, 393, 360, 337, 291, 11, 1029, 437, 291, 393, 360, 337, 428, 1941, 13]
    words, groups = _fake_split(ids, 50257)
    m = rng.random((len(ids), 300)).astype(np.float16)
    known = [-0.5, -1.0, -2.0, -0.1, -0.3]
    lps = [known[i % 5] for i in range(len(ids))]
    ref = W.find_alignment(words, groups, m, lps)
    got = SEEKER.findAlignment(words, groups, m, lps)
    same_words(got, ref, tol=1e-6)
    prev_end, k = -1.0, 0
    for w in got:
        assert w.word and w.tokens and w.start <= w.end and w.start >= prev_end
        assert abs(w.probability - np.exp(np.mean(lps[k:k + len(w.tokens)]))) < 1e-4
        prev_end, k = w.end, k + len(w.tokens)
    assert SEEKER.findAlignment(words[:1], groups[:1], m[:1], lps[:1]) == []      # wordTokens.count <= 1 -> []


def test_add_word_timestamps_matches_oracle_on_random_windows():
    rng = np.random.default_rng(3)
    SB, TT = 1000, 1100                                              # specialTokenBegin, timeTokenBegin of the stand-in vocabulary
    hits = 0
    for trial in range(40):
        segs, t0 = [], int(rng.integers(0, 50))
        for s in range(int(rng.integers(1, 4))):
            n = int(rng.integers(1, 12))
            t1 = t0 + int(rng.integers(20, 300))
            toks = [TT + t0] + [int(v) for v in rng.integers(1, SB, n)] + [TT + t1]
            if s == 0:
                toks = [SB + 1, SB + 2] + toks                        # <|startoftranscript|><|transcribe|> stand-ins
            segs.append(W.Segment(float(np.float32(t0) * np.float32(0.02)), float(np.float32(t1) * np.float32(0.02)), toks,
                                  [float(v) for v in -rng.random(len(toks))], id=s))
            t0 = t1
        rows = sum(len(s.tokens) for s in segs)
        cols = 400
        m = (rng.random((rows + 3, cols)) * 0.01).astype(np.float16)
        centre = np.sort(rng.integers(0, cols, rows))
        for r in range(rows):
            m[r, max(0, centre[r] - 2):centre[r] + 3] += np.float16(0.5)   # a monotone ridge like real alignment heads
        seek = int(rng.integers(0, 100000))
        last = float(np.float32(seek) / np.float32(16000))
        ref = W.add_word_timestamps(segs, m, lambda t: _fake_split(t, SB), seek, last, SB, decode=lambda t: "".join(chr(97 + v % 26) for v in t))
        got = SEEKER.addWordTimestamps(segs, m, lambda t: _fake_split(t, SB), seek, last, SB, decode=lambda t: "".join(chr(97 + v % 26) for v in t))
        assert len(got) == len(ref)
        for (gs, ge, gw), r in zip(got, ref):
            assert np.float32(gs) == np.float32(r.start) and np.float32(ge) == np.float32(r.end)
            same_words(gw, r.words, tol=1e-6)
            hits += len(gw)
    assert hits > 200
