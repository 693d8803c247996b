"""The long-form host logic of libwkb200 (C++, csrc/longform.cu) through the C ABI, WITHOUT a GPU: reference goldens on
jfk.wav (UnitTests.swift:2119-2241) and differential tests against the oracle on random inputs."""
import os

import numpy as np
import pytest

from oracle import seek_ref as S
from whisperkit_b200 import longform as L
from whisperkit_b200.api import DecodingOptions

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def jfk():
    z = np.load(os.path.join(GOLD, "jfk_logmel_hf.npz"))
    return z["pcm16"].astype(np.float32) / 32768.0


def test_energy_vad_jfk_goldens_through_c_abi():
    x = jfk()
    vad = L.EnergyVAD()
    assert vad.voiceActivity(np.zeros(0, np.float32)) == []
    v = vad.voiceActivity(x)
    assert vad.findLongestSilence(v) == (43, 54)
    big = L.EnergyVAD(frameLength=0.2, frameOverlap=0.1)
    clips = big.calculateNonSilentSeekClips(x)
    assert [c[0] for c in clips] == [3200, 51200, 83200, 128000, 169600]
    assert [c[1] for c in clips] == [35200, 70400, 121600, 166400, 176000]
    z, o = np.zeros(1600, np.float32), np.ones(1600, np.float32)
    small = L.EnergyVAD(frameLengthSamples=320)
    assert small.calculateActiveChunks(z) == [] and small.calculateActiveChunks(o) == [(0, 1600)]
    assert small.calculateActiveChunks(np.concatenate([z, o])) == [(1600, 3200)]
    assert small.calculateActiveChunks(np.ones(1601, np.float32)) == [(0, 1601)]
    assert small.calculateActiveChunks(np.concatenate([np.ones(1599, np.float32), z])) == [(0, 1600)]
    assert L.EnergyVAD(frameLengthSamples=320, frameOverlapSamples=80).calculateActiveChunks(np.concatenate([z, o])) == [(1280, 3200)]
    f = vad.findLongestSilence
    assert f([]) is None and f([True, True]) is None and f([False]) == (0, 1)
    assert f([False, False, True, True, True, False, True, False, False, False, False, True, True]) == (7, 11)


def test_vad_and_chunker_match_oracle_on_random_audio():
    rng = np.random.default_rng(0)
    for trial in range(6):
        n = int(rng.integers(1, 400000))
        x = (rng.standard_normal(n) * 0.05).astype(np.float32)
        for a, b in zip(rng.integers(0, n, 6), rng.integers(1000, 40000, 6)):
            x[a:a + b] *= 0.01                                    # carve silences
        for fl, ov in ((1600, 0), (320, 80), (3200, 1600)):
            assert L.EnergyVAD(frameLengthSamples=fl, frameOverlapSamples=ov).voiceActivity(x) == \
                S.EnergyVAD(frameLengthSamples=fl, frameOverlapSamples=ov).voiceActivity(x)
            assert L.EnergyVAD(frameLengthSamples=fl, frameOverlapSamples=ov).calculateActiveChunks(x) == \
                S.EnergyVAD(frameLengthSamples=fl, frameOverlapSamples=ov).calculateActiveChunks(x)
        for mx, cts in ((80000, ()), (50000, (0.5,)), (120000, (1.0, 9.0, 12.0))):
            assert L.VADAudioChunker().chunkAll(x, mx, cts) == S.vad_chunk_all(x, mx, cts)
    x = jfk()
    assert L.VADAudioChunker().chunkAll(x, 480000) == [(0, len(x))]           # testVADAudioChunker: jfk is one chunk
    assert L.prepareSeekClips([], 1000) == [(0, 1000)] and L.prepareSeekClips([0.5, 1.5, 2.0], 48000) == [(8000, 24000), (32000, 48000)]


def test_find_seek_point_and_segments_matches_oracle():
    rng = np.random.default_rng(1)
    TT = 50364
    seeker = L.SegmentSeeker()
    n_multi = 0
    for trial in range(300):
        n = int(rng.integers(1, 40))
        toks = []
        for _ in range(n):
            r = rng.random()
            if r < 0.35:
                toks.append(int(TT + rng.integers(0, 1500)))
            elif r < 0.45 and toks and toks[-1] >= TT:
                toks.append(toks[-1])                              # consecutive timestamp pair
            else:
                toks.append(int(rng.integers(0, 50257)))
        lps = [float(v) for v in -rng.random(n)]
        kw = dict(noSpeechProb=float(rng.random()), avgLogProb=float(-2 * rng.random()), compressionRatio=1.3, temperature=0.2)
        nst = None if trial % 5 == 0 else 0.6
        lpt = None if trial % 7 == 0 else -1.0
        seek0, size = int(rng.integers(0, 10 ** 6)), int(rng.integers(16000, 480001))
        ref_seek, ref = S.find_seek_point_and_segments(toks, lps, noSpeechThreshold=nst, logProbThreshold=lpt, allSegmentsCount=trial,
                                                       currentSeek=seek0, segmentSize=size, sampleRate=16000, timeToken=TT, **kw)
        o = DecodingOptions(noSpeechThreshold=nst, logProbThreshold=lpt)
        got_seek, got = seeker.findSeekPointAndSegments(toks, lps, kw["avgLogProb"], kw["compressionRatio"], kw["temperature"], o, trial, seek0,
                                                        size, 16000, TT, noSpeechProb=kw["noSpeechProb"])
        assert got_seek == ref_seek
        assert (got is None) == (ref is None)
        if ref is not None:
            assert [g.tokens for g in got] == [r.tokens for r in ref] and [g.id for g in got] == [r.id for r in ref]
            np.testing.assert_array_equal(np.float32([g.start for g in got]), np.float32([r.start for r in ref]))
            np.testing.assert_array_equal(np.float32([g.end for g in got]), np.float32([r.end for r in ref]))
            n_multi += len(ref) > 1
    assert n_multi > 20
