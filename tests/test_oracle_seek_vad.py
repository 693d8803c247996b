"""Pins oracle/seek_ref.py to the reference's own tests: energy-VAD goldens on its jfk.wav (UnitTests.swift:2119-2190),
findLongestSilence cases (:2210-2241), prepareSeekClips semantics, and findSeekPointAndSegments behaviour cases."""
import os

import numpy as np

from oracle import seek_ref as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def jfk():
    z = np.load(os.path.join(GOLD, "jfk_logmel_hf.npz"))
    return z["pcm16"].astype(np.float32) / 32768.0


def test_energy_vad_jfk_goldens():
    x = jfk()
    vad = S.EnergyVAD()
    assert vad.voiceActivity([]) == []
    v = vad.voiceActivity(x)
    assert S.EnergyVAD.findLongestSilence(v) == (43, 54)
    assert vad.voiceActivityIndexToAudioSampleIndex(43) == 68800 and vad.voiceActivityIndexToAudioSampleIndex(54) == 86400
    big = S.EnergyVAD(frameLength=0.2, frameOverlap=0.1)
    clips = big.calculateNonSilentSeekClips(x)
    assert [c[0] for c in clips] == [3200, 51200, 83200, 128000, 169600]
    assert [c[1] for c in clips] == [35200, 70400, 121600, 166400, 176000]
    np.testing.assert_allclose(big.voiceActivityClipTimestamps(x), [0.2, 2.2, 3.2, 4.4, 5.2, 7.6, 8.0, 10.4, 10.6, 11.0], atol=1e-6)


def test_active_chunks_synthetic():
    v = S.EnergyVAD(frameLengthSamples=320)
    z, o = [0.0] * 1600, [1.0] * 1600
    assert v.calculateActiveChunks([]) == [] and v.calculateActiveChunks(z) == []
    assert v.calculateActiveChunks(o) == [(0, 1600)]
    assert v.calculateActiveChunks(z + o) == [(1600, 3200)]
    assert v.calculateActiveChunks([1.0] * 1601) == [(0, 1601)]
    assert v.calculateActiveChunks([1.0] * 1599) == [(0, 1599)]
    assert v.calculateActiveChunks([1.0] * 1599 + z) == [(0, 1600)]
    vo = S.EnergyVAD(frameLengthSamples=320, frameOverlapSamples=80)
    assert vo.calculateActiveChunks(z + o) == [(1280, 3200)]


def test_find_longest_silence_cases():
    f = S.EnergyVAD.findLongestSilence
    T, Fa = True, False
    assert f([]) is None and f([T]) is None and f([T, T, T, T, T]) is None
    assert f([Fa]) == (0, 1) and f([Fa, Fa]) == (0, 2) and f([T, Fa, Fa]) == (1, 3)
    assert f([Fa, Fa, T]) == (0, 2) and f([T, Fa, Fa, T]) == (1, 3)
    assert f([Fa, Fa, T, T, T, Fa, T, Fa, Fa, Fa, Fa, T, T]) == (7, 11)


def test_prepare_seek_clips():
    assert S.prepare_seek_clips([], 1000) == [(0, 1000)]
    assert S.prepare_seek_clips([1.0], 48000) == [(16000, 48000)]
    assert S.prepare_seek_clips([0.5, 1.5, 2.0], 48000) == [(8000, 24000), (32000, 48000)]


def test_vad_chunker_short_and_long():
    x = jfk()
    assert S.vad_chunk_all(x, 480000) == [(0, len(x))]
    chunks = S.vad_chunk_all(x, 80000)
    assert chunks[0][0] == 0 and chunks[-1][1] <= len(x)
    assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))       # contiguous
    assert all(e - s <= 80000 for s, e in chunks)
    # every split lands inside a silent 0.1 s frame of the second half of its span (middle of the longest silence there)
    v = S.EnergyVAD().voiceActivity(x)
    for s_, e_ in chunks[:-1]:
        assert not v[e_ // 1600] and e_ >= s_ + (min(len(x), s_ + 80000) - s_) // 2


def test_find_seek_point_and_segments_cases():
    TT = 100  # timeToken

    def run(tokens, **kw):
        lps = [-0.1 * i for i in range(len(tokens))]
        args = dict(noSpeechProb=0.0, avgLogProb=-0.5, compressionRatio=1.0, temperature=0.0, noSpeechThreshold=0.6,
                    logProbThreshold=-1.0, allSegmentsCount=3, currentSeek=32000, segmentSize=480000, sampleRate=16000, timeToken=TT)
        args.update(kw)
        return S.find_seek_point_and_segments(tokens, lps, **args)

    # two consecutive-timestamp pairs -> two segments, seek to the last timestamp
    seek, segs = run([1, 2, TT + 0, 5, 6, TT + 100, TT + 100, 7, TT + 250, TT + 250, 9])
    assert [s.id for s in segs] == [3, 4]
    assert [s.tokens for s in segs][0] == [1, 2, TT, 5, 6, TT + 100] and segs[1].tokens == [TT + 100, 7, TT + 250]
    np.testing.assert_allclose([segs[0].start, segs[0].end, segs[1].start, segs[1].end], [2.0, 4.0, 4.0, 7.0], atol=1e-5)
    assert seek == 32000 + 250 * 320       # last3 = [T, T, F]: seek to the last timestamp (5.00 s); the trailing text token is dropped
    # ending "... text <ts> EOT-like text": single timestamp ending
    seek, segs = run([TT + 0, 5, TT + 50, TT + 50, 6, TT + 120, 7])
    assert len(segs) == 2 and segs[1].tokens == [TT + 50, 6, TT + 120]
    assert seek == 32000 + int(np.float32(120) * np.float32(0.02) * np.float32(16000))
    # no consecutive timestamps: one segment over the window, duration from the last timestamp, seek += segmentSize
    seek, segs = run([1, TT + 0, 5, 6, TT + 77, 9])
    assert len(segs) == 1 and abs(segs[0].end - (2.0 + 77 * 0.02)) < 1e-5 and seek == 32000 + 480000
    # silence skip: noSpeechProb above threshold and low avg logprob
    seek, segs = run([1, 2, 3], noSpeechProb=0.9, avgLogProb=-2.0)
    assert segs is None and seek == 32000 + 480000
    seek, segs = run([1, 2, 3], noSpeechProb=0.9, avgLogProb=-0.5)   # confident -> not skipped
    assert segs is not None


def test_seek_loop_walks_windows():
    class R:
        def __init__(self, tokens):
            self.tokens, self.tokenLogProbs = tokens, [0.0] * len(tokens)
            self.avgLogProb, self.compressionRatio, self.temperature = -0.3, 1.0, 0.0
    TT = 50364
    calls = []

    def decode(seek, size):
        calls.append((seek, size))
        # every window: one sentence ending at 20.00 s with a consecutive-timestamp pair, then trailing text
        return R([50258, TT, 11, 12, TT + 1000, TT + 1000, 13, TT + 1100, 50257])

    segs, windows = S.seek_loop(16000 * 70, decode, timeToken=TT)
    assert windows[0] == (0, 480000) and windows[1][0] == int(np.float32(1100 * 0.02) * 16000)
    assert all(b[0] > a[0] for a, b in zip(windows, windows[1:]))
    assert [s.id for s in segs] == list(range(len(segs)))
