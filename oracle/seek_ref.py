"""Windowing / seek / VAD oracle (test infrastructure; see oracle/__init__.py).

CPU restatement of the reference's host-side long-form logic:
  * SegmentSeeker.findSeekPointAndSegments   Sources/WhisperKit/Core/Text/SegmentSeeker.swift:41-189
  * DecodingOptions.prepareSeekClips         Sources/WhisperKit/Utilities/Extensions+Internal.swift:111-130
  * EnergyVAD.voiceActivity                  Sources/WhisperKit/Core/Audio/EnergyVAD.swift:41-56
    + calculateVoiceActivityInChunks         Sources/WhisperKit/Core/Audio/AudioProcessor.swift:674-702
  * VoiceActivityDetector.calculateActiveChunks / findLongestSilence / calculateNonSilentSeekClips
                                             Sources/WhisperKit/Core/Audio/VoiceActivityDetector.swift:52-159
  * VADAudioChunker.chunkAll                 Sources/WhisperKit/Core/Audio/AudioChunker.swift:53-107
  * the seek loop of TranscribeTask.run      Sources/WhisperKit/Core/TranscribeTask.swift:98-279
Swift `Float` arithmetic is mirrored with numpy float32.  Pinned by the reference's own tests
(UnitTests.swift:2119-2241, golden values on its jfk.wav) in tests/test_oracle_seek_vad.py.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

SAMPLE_RATE = 16000
SECONDS_PER_TIME_TOKEN = np.float32(0.02)  # WhisperKit.secondsPerTimeToken (WhisperKit.swift:38-40)
F = np.float32


@dataclass
class TranscriptionSegment:
    id: int
    seek: int
    start: float
    end: float
    tokens: List[int]
    tokenLogProbs: List[float]
    temperature: float = 0.0
    avgLogprob: float = 0.0
    compressionRatio: float = 0.0
    noSpeechProb: float = 0.0
    words: Optional[list] = None


def find_seek_point_and_segments(tokens: Sequence[int], tokenLogProbs: Sequence[float], noSpeechProb: float, avgLogProb: float,
                                 compressionRatio: float, temperature: float, noSpeechThreshold: Optional[float],
                                 logProbThreshold: Optional[float], allSegmentsCount: int, currentSeek: int, segmentSize: int,
                                 sampleRate: int, timeToken: int) -> Tuple[int, Optional[List[TranscriptionSegment]]]:
    """SegmentSeeker.swift:41-189."""
    seek = currentSeek
    timeOffset = F(seek) / F(sampleRate)
    if noSpeechThreshold is not None:
        shouldSkip = noSpeechProb > noSpeechThreshold
        if logProbThreshold is not None and avgLogProb > logProbThreshold:
            shouldSkip = False
        if shouldSkip:
            return seek + segmentSize, None
    segs: List[TranscriptionSegment] = []
    cur = list(tokens)
    lps = list(tokenLogProbs)
    isTs = [t >= timeToken for t in cur]
    last3 = isTs[-3:]
    singleTimestampEnding = last3 == [False, True, False]
    noTimestampEnding = last3 == [False, False, False]
    sliceIndexes = []
    prev = False
    for i, c in enumerate(isTs):
        if prev and c:
            sliceIndexes.append(i)
        prev = c

    def mk(tok, lp, start, end):
        return TranscriptionSegment(allSegmentsCount + len(segs), seek, float(start), float(end), list(tok), list(lp), temperature,
                                    avgLogProb, compressionRatio, noSpeechProb)

    if sliceIndexes:
        if singleTimestampEnding:
            sliceIndexes.append(max(i for i, v in enumerate(isTs) if v) + 1)
        elif noTimestampEnding:
            sliceIndexes.append(len(cur))
        lastSliceStart = 0
        for currentSliceEnd in sliceIndexes:
            st = cur[lastSliceStart:currentSliceEnd]
            sl = lps[lastSliceStart:currentSliceEnd]
            tts = [t for t in st if t >= timeToken]
            startS = F(tts[0] - timeToken) * SECONDS_PER_TIME_TOKEN
            endS = F(tts[-1] - timeToken) * SECONDS_PER_TIME_TOKEN
            segs.append(mk(st, sl, F(timeOffset + startS), F(timeOffset + endS)))
            lastSliceStart = currentSliceEnd
        if not noTimestampEnding:
            lastTimestampToken = cur[lastSliceStart - (1 if singleTimestampEnding else 0)] - timeToken
            lastTimestampSeconds = F(lastTimestampToken) * SECONDS_PER_TIME_TOKEN
            seek += int(F(lastTimestampSeconds * F(sampleRate)))
        else:
            seek += segmentSize
    else:
        durationSeconds = F(segmentSize) / F(sampleRate)
        tts = [t for t in cur if t > timeToken]
        if tts:
            durationSeconds = F(tts[-1] - timeToken) * SECONDS_PER_TIME_TOKEN
        segs.append(mk(cur, lps, timeOffset, F(timeOffset + durationSeconds)))
        seek += segmentSize
    return seek, segs


def prepare_seek_clips(clipTimestamps: Sequence[float], contentFrames: int) -> List[Tuple[int, int]]:
    """Extensions+Internal.swift:111-130 (Swift round = half away from zero)."""
    def swift_round(x):
        return int(np.floor(abs(x) + 0.5) * (1 if x >= 0 else -1))
    pts = [swift_round(float(F(t) * F(SAMPLE_RATE))) for t in clipTimestamps]
    if not pts:
        pts.append(0)
    if len(pts) % 2 == 1:
        pts.append(contentFrames)
    return [(pts[i], pts[i + 1] if i + 1 < len(pts) else contentFrames) for i in range(0, len(pts), 2)]


class EnergyVAD:
    """EnergyVAD.swift + VoiceActivityDetector.swift."""

    def __init__(self, sampleRate: int = SAMPLE_RATE, frameLength: float = 0.1, frameOverlap: float = 0.0,
                 energyThreshold: float = 0.02, frameLengthSamples: Optional[int] = None, frameOverlapSamples: Optional[int] = None):
        self.sampleRate = sampleRate
        self.frameLengthSamples = frameLengthSamples if frameLengthSamples is not None else int(F(frameLength) * F(sampleRate))
        self.frameOverlapSamples = frameOverlapSamples if frameOverlapSamples is not None else int(F(frameOverlap) * F(sampleRate))
        self.energyThreshold = F(energyThreshold)

    def voiceActivity(self, waveform) -> List[bool]:
        x = np.asarray(waveform, dtype=np.float32)
        n = len(x)
        count = int(np.ceil(n / self.frameLengthSamples)) if n else 0
        out = []
        for i in range(count):
            s = i * self.frameLengthSamples
            e = min(s + self.frameLengthSamples + self.frameOverlapSamples, n)
            chunk = x[s:e].astype(np.float64)
            rms = F(np.sqrt(np.mean(chunk * chunk))) if len(chunk) else F(0)   # vDSP_rmsqv
            out.append(bool(rms > self.energyThreshold))
        return out

    def calculateActiveChunks(self, waveform) -> List[Tuple[int, int]]:
        vad = self.voiceActivity(waveform)
        n = len(waveform)
        res: List[List[int]] = []
        cur = None
        for i, v in enumerate(vad):
            if v:
                s = i * self.frameLengthSamples
                e = min(s + self.frameLengthSamples, n)
                if cur is not None:
                    res[-1][1] = e
                else:
                    cur = s
                    res.append([s, e])
            else:
                cur = None
        return [(a, b) for a, b in res]

    def voiceActivityIndexToAudioSampleIndex(self, i: int) -> int:
        return i * self.frameLengthSamples

    @staticmethod
    def findLongestSilence(vad: Sequence[bool]) -> Optional[Tuple[int, int]]:
        best = None
        bestCount = 0
        i = 0
        while i < len(vad):
            if vad[i]:
                i += 1
            else:
                e = i
                while e < len(vad) and not vad[e]:
                    e += 1
                if e - i > bestCount:
                    bestCount = e - i
                    best = (i, e)
                i = e
        return best

    def voiceActivityClipTimestamps(self, waveform) -> List[float]:
        out = []
        for s, e in self.calculateActiveChunks(waveform):
            out += [float(F(s) / F(self.sampleRate)), float(F(e) / F(self.sampleRate))]
        return out

    def calculateNonSilentSeekClips(self, waveform) -> List[Tuple[int, int]]:
        return prepare_seek_clips(self.voiceActivityClipTimestamps(waveform), len(waveform))


def vad_chunk_all(audio, maxChunkLength: int, clipTimestamps: Sequence[float] = (), windowPadding: int = 16000,
                  vad: Optional[EnergyVAD] = None) -> List[Tuple[int, int]]:
    """VADAudioChunker.chunkAll (AudioChunker.swift:66-107) -> [(seekOffsetIndex, endIndex)]."""
    vad = vad or EnergyVAD()
    n = len(audio)
    if n <= maxChunkLength:
        return [(0, n)]
    out = []
    for clipStart, clipEnd in prepare_seek_clips(clipTimestamps, n):
        start = clipStart
        while start < clipEnd - windowPadding:
            if not (0 <= start < n):
                raise ValueError("startIndex is outside the buffer size")
            end = clipEnd
            if start + maxChunkLength < end:
                e2 = min(n, start + maxChunkLength)
                mid = start + (e2 - start) // 2
                va = vad.voiceActivity(audio[mid:e2])
                sil = vad.findLongestSilence(va)
                end = e2 if sil is None else mid + vad.voiceActivityIndexToAudioSampleIndex(sil[0] + (sil[1] - sil[0]) // 2)
            if end <= start:
                break
            out.append((start, end))
            start = end
    return out


def seek_loop(contentFrames: int, decode_window, clipTimestamps: Sequence[float] = (), windowClipTime: float = 1.0,
              windowSamples: int = 480000, timeToken: int = 50364, noSpeechThreshold: Optional[float] = 0.6,
              logProbThreshold: Optional[float] = -1.0, maxWindowSeek: Optional[int] = None, wordTimestamps=None):
    """The windowing loop of TranscribeTask.run (TranscribeTask.swift:98-279).
    decode_window(seek, segmentSize) -> object with tokens, tokenLogProbs, avgLogProb, compressionRatio, temperature.
    wordTimestamps = dict(alignment=fn(r) -> [rows, 1500] matrix of that window, split=fn(tokens) -> (words, wordTokens),
    decode=fn(tokens) -> str, specialTokenBegin=int) switches on the addWordTimestamps block (TranscribeTask.swift:197-239)."""
    from . import words_ref as WR
    allSegments: List[TranscriptionSegment] = []
    windows = []
    for clipStart, clipEnd in prepare_seek_clips(clipTimestamps, contentFrames):
        seek = clipStart
        windowPadding = int(F(windowClipTime) * F(SAMPLE_RATE))
        while seek < clipEnd - windowPadding and seek < contentFrames:   # 2nd term: guard shared with csrc/longform.cu (clip past the audio)
            segmentSize = min(windowSamples, contentFrames - seek, clipEnd - seek)
            r = decode_window(seek, segmentSize)
            windows.append((seek, segmentSize))
            previousSeek = seek
            newSeek, segs = find_seek_point_and_segments(r.tokens, r.tokenLogProbs, 0.0, r.avgLogProb, r.compressionRatio, r.temperature,
                                                         noSpeechThreshold, logProbThreshold, len(allSegments), seek, segmentSize,
                                                         SAMPLE_RATE, timeToken)
            seek = max(seek, newSeek)
            if wordTimestamps is not None:
                a = np.asarray(wordTimestamps["alignment"](r), dtype=np.float32)
                n_tok = len(r.tokens)
                if a.shape[0] < n_tok:                                # rows past the 224-row tensor read as zeros (see csrc/longform.cu)
                    a = np.concatenate([a, np.zeros((n_tok - a.shape[0], a.shape[1]), np.float32)])
                upd = WR.add_word_timestamps([WR.Segment(g.start, g.end, g.tokens, g.tokenLogProbs, id=g.id, seek=g.seek) for g in (segs or [])],
                                             a, wordTimestamps["split"], previousSeek, float(np.float64(previousSeek) / np.float64(SAMPLE_RATE)),
                                             wordTimestamps["specialTokenBegin"], decode=wordTimestamps.get("decode"))
                kept = []
                for g, u in zip(segs or [], upd):
                    g.start, g.end, g.words = u.start, u.end, u.words
                    if F(g.end) > F(g.start):                         # filter out zero length segments (:214)
                        kept.append(g)
                if segs is not None:
                    segs = kept
                if kept:
                    seek = max(seek, int(F(kept[-1].end) * F(SAMPLE_RATE)))
            if maxWindowSeek is not None:
                seek = min(seek, previousSeek + maxWindowSeek)
            if seek <= previousSeek:
                # termination guard shared with csrc/longform.cu (NOT in the reference, whose loop re-decodes the same window
                # forever when a window's last consecutive timestamp pair is <|0.00|><|0.00|>): always move on
                seek = previousSeek + segmentSize
            if segs is None:
                continue
            allSegments.extend(segs)
    return allSegments, windows
