"""Word-timestamp oracle (test infrastructure; see oracle/__init__.py).

CPU restatement of the reference's word-level timing path:
  * SegmentSeeker.dynamicTimeWarping / minCostAndTrace / backtrace   Sources/WhisperKit/Core/Text/SegmentSeeker.swift:195-276
  * mergePunctuations                                                SegmentSeeker.swift:278-338
  * findAlignment                                                    SegmentSeeker.swift:340-408
  * addWordTimestamps                                                SegmentSeeker.swift:410-496
  * calculateWordDurationConstraints / truncateLongWordsAtSentenceBoundaries   SegmentSeeker.swift:498-526
  * updateSegmentsWithWordTimings                                    SegmentSeeker.swift:528-659
  * Float.rounded(_:)                                                Sources/ArgmaxCore/FoundationExtensions.swift:10-13
The tokenizer (splitToWordTokens, Models.swift:1226-1306, which leans on Apple's NaturalLanguage) stays with the host: every
function here takes the words already split.  Swift `Float` arithmetic is mirrored with numpy float32.
Pinned by the reference's own tests (UnitTests.swift:2336-2960) in tests/test_oracle_words.py.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

F = np.float32
SECONDS_PER_TIME_TOKEN = F(0.02)
SAMPLE_RATE = 16000
DEFAULT_PREPEND = "\"'“¡¿([{-"          # Constants.defaultPrependPunctuations (Models.swift:1459)
DEFAULT_APPEND = "\"'.。,，!！?？:：”)]}、"  # Constants.defaultAppendPunctuations (Models.swift:1460)
SENTENCE_END_MARKS = [".", "。", "!", "！", "?", "？"]
# CharacterSet.whitespaces: Unicode general category Zs plus TAB
_WS = " \t               　"


@dataclass
class WordTiming:
    word: str
    tokens: List[int]
    start: float
    end: float
    probability: float

    @property
    def duration(self):
        return F(F(self.end) - F(self.start))


def swift_trim(s: str) -> str:
    return s.strip(_WS)


def swift_contains(hay: str, needle: str) -> bool:
    """String.contains(_ other: String): substring search; the empty string is contained (native Swift >= 5.7 semantics)."""
    return needle in hay


def rounded(x, places: int = 2):
    """Float.rounded(_ decimalPlaces:) - (self * 10^p).rounded() / 10^p, round half away from zero, all in Float."""
    div = F(10.0) ** F(places)
    v = F(F(x) * div)
    r = F(np.trunc(v))
    if abs(F(v - r)) >= F(0.5):            # v - trunc(v) is exact in Float
        r = F(r + (F(1) if v >= 0 else F(-1)))
    return F(r / div)


def dynamic_time_warping(matrix) -> Tuple[List[int], List[int]]:
    """SegmentSeeker.swift:195-276: cost in Double over -matrix; strict '<' tie rules: diagonal, then up, else left."""
    m = np.asarray(matrix)
    if m.ndim != 2:
        raise ValueError("Invalid alignment matrix shape")
    rows, cols = m.shape
    neg = -m.astype(np.float64)
    cost = np.full((rows + 1, cols + 1), np.inf)
    trace = np.full((rows + 1, cols + 1), -1, np.int8)
    cost[0, 0] = 0.0
    trace[0, 1:] = 2
    trace[1:, 0] = 1
    for r in range(1, rows + 1):
        prev, cur = cost[r - 1], cost[r]
        nr = neg[r - 1]
        for c in range(1, cols + 1):
            v = nr[c - 1]
            c0, c1, c2 = prev[c - 1] + v, prev[c] + v, cur[c - 1] + v
            if c0 < c1 and c0 < c2:
                cur[c], trace[r, c] = c0, 0
            elif c1 < c0 and c1 < c2:
                cur[c], trace[r, c] = c1, 1
            else:
                cur[c], trace[r, c] = c2, 2
    i, j = rows, cols
    ti, tj = [], []
    while i > 0 or j > 0:
        ti.append(i - 1)
        tj.append(j - 1)
        t = trace[i, j]
        if t == 0:
            i -= 1
            j -= 1
        elif t == 1:
            i -= 1
        elif t == 2:
            j -= 1
        else:
            break   # unreachable for a well-formed trace (the reference would spin here)
    return ti[::-1], tj[::-1]


def merge_punctuations(alignment: Sequence[WordTiming], prepended: str = DEFAULT_PREPEND, appended: str = DEFAULT_APPEND) -> List[WordTiming]:
    """SegmentSeeker.swift:278-338."""
    if not alignment:
        return []
    al = [replace(w, tokens=list(w.tokens)) for w in alignment]
    pre: List[WordTiming] = []
    if not swift_contains(prepended, swift_trim(al[0].word)):
        pre.append(al[0])
    for i in range(1, len(al)):
        cur, prev = replace(al[i], tokens=list(al[i].tokens)), al[i - 1]
        if prev.word and prev.word[0] in _WS and swift_contains(prepended, swift_trim(prev.word)):
            cur.word = prev.word + cur.word
            cur.tokens = list(prev.tokens) + list(cur.tokens)
            if not pre:
                pre.append(cur)
            else:
                pre[-1] = cur
        else:
            pre.append(cur)
    app: List[WordTiming] = []
    if pre:
        app.append(pre[0])
    for i in range(1, len(pre)):
        cur, prev = pre[i], replace(pre[i - 1], tokens=list(pre[i - 1].tokens))
        if not prev.word.endswith(" ") and swift_contains(appended, swift_trim(cur.word)):
            prev.word = prev.word + cur.word
            prev.tokens = list(prev.tokens) + list(cur.tokens)
            app[-1] = prev
        else:
            app.append(cur)
    return [w for w in app if w.word != "" and not swift_contains(appended, w.word) and not swift_contains(prepended, w.word)]


def find_alignment(words: Sequence[str], wordTokens: Sequence[Sequence[int]], alignmentWeights, tokenLogProbs: Sequence[float]) -> List[WordTiming]:
    """SegmentSeeker.swift:340-408; (words, wordTokens) = tokenizer.splitToWordTokens(tokenIds: wordTokenIds)."""
    textIdx, timeIdx = dynamic_time_warping(alignmentWeights)
    if len(wordTokens) <= 1:
        return []
    startTimes = [F(0.0)]
    endTimes = []
    cur = textIdx[0] if textIdx else 0
    for k in range(len(textIdx)):
        if textIdx[k] != cur:
            cur = textIdx[k]
            t = F(timeIdx[k]) * SECONDS_PER_TIME_TOKEN
            startTimes.append(t)
            endTimes.append(t)
    endTimes.append(F(timeIdx[-1] if timeIdx else 1500) * SECONDS_PER_TIME_TOKEN)
    out = []
    ci = 0
    for w, toks in zip(words, wordTokens):
        s0 = ci
        st = startTimes[ci]
        ci += len(toks) - 1
        en = endTimes[ci]
        ci += 1
        acc = F(0)
        for v in tokenLogProbs[s0:ci]:
            acc = F(acc + F(v))
        prob = F(np.exp(F(acc / F(ci - s0))))
        out.append(WordTiming(w, list(toks), float(st), float(en), float(prob)))
    return out


def calculate_word_duration_constraints(alignment: Sequence[WordTiming]) -> Tuple[float, float]:
    d = sorted(float(w.duration) for w in alignment if w.duration > 0)
    med = F(d[len(d) // 2]) if d else F(0.0)
    cm = F(min(F(0.7), med))
    return float(cm), float(F(cm * F(2)))


def truncate_long_words_at_sentence_boundaries(alignment: Sequence[WordTiming], maxDuration: float) -> List[WordTiming]:
    al = [replace(w, tokens=list(w.tokens)) for w in alignment]
    md = F(maxDuration)
    for i in range(1, len(al)):
        if al[i].duration > md:
            if al[i].word in SENTENCE_END_MARKS:
                al[i].end = float(F(F(al[i].start) + md))
            elif al[i - 1].word in SENTENCE_END_MARKS:
                al[i].start = float(F(F(al[i].end) - md))
    return al


@dataclass
class Segment:
    """TranscriptionSegment fields this path reads and writes."""
    start: float
    end: float
    tokens: List[int]
    tokenLogProbs: List[float] = field(default_factory=list)
    words: Optional[List[WordTiming]] = None
    id: int = 0
    seek: int = 0


def update_segments_with_word_timings(segments: Sequence[Segment], mergedAlignment: Sequence[WordTiming], seek: int, lastSpeechTimestamp: float,
                                      constrainedMedianDuration: float, maxDuration: float, specialTokenBegin: int,
                                      decode: Optional[Callable[[List[int]], str]] = None) -> List[Segment]:
    """SegmentSeeker.swift:528-659."""
    timeOffset = F(seek) / F(SAMPLE_RATE)
    cmd, md = F(constrainedMedianDuration), F(maxDuration)
    last = F(lastSpeechTimestamp)
    wordIndex = 0
    updated: List[Segment] = []
    for si, seg in enumerate(segments):
        saved = 0
        textTokens = [t for t in seg.tokens if t < specialTokenBegin]
        wis: List[WordTiming] = []
        # `for timing in mergedAlignment[wordIndex...] where savedTokens < textTokens.count`: the slice is fixed when the loop starts and
        # the where-clause SKIPS (does not stop) - skipped words do not advance wordIndex
        for timing in list(mergedAlignment[wordIndex:]):
            if not (saved < len(textTokens)):
                continue
            wordIndex += 1
            tt = [t for t in timing.tokens if t < specialTokenBegin]
            if not tt:
                continue
            word = timing.word
            if len(tt) < len(timing.tokens):
                word = decode(tt) if decode is not None else timing.word
            start = rounded(F(timeOffset + F(timing.start)))
            end = rounded(F(timeOffset + F(timing.end)))
            if F(end - start) < F(cmd / F(4)):
                if len(wis) >= 1:
                    pe = F(wis[-1].end)
                    if start > pe:
                        space = F(start - pe)
                        start = rounded(F(start - min(space, F(cmd / F(2)))))
                elif not wis and si > 0 and len(updated) > si - 1 and start > F(updated[si - 1].end):
                    space = F(start - F(updated[si - 1].end))
                    start = rounded(F(start - min(space, F(cmd / F(2)))))
            wis.append(WordTiming(word, tt, float(start), float(end), float(rounded(F(timing.probability)))))
            saved += len(tt)
        us = replace(seg, tokens=list(seg.tokens), tokenLogProbs=list(seg.tokenLogProbs))
        if wis:
            first = wis[0]
            pause = F(F(first.end) - last)
            firstTooLong = first.duration > md
            bothTooLong = len(wis) > 1 and F(F(wis[1].end) - F(first.start)) > F(md * F(2))
            if pause > F(cmd * F(4)) and (firstTooLong or bothTooLong):
                if len(wis) > 1 and wis[1].duration > md:
                    boundary = max(F(F(wis[1].end) / F(2)), F(F(wis[1].end) - md))
                    wis[0].end = float(boundary)
                    wis[1].start = float(boundary)
                wis[0].start = float(max(last, F(F(wis[0].end) - md)))
            if F(seg.start) < F(wis[0].end) and F(F(seg.start) - F(0.5)) > F(wis[0].start):
                wis[0].start = float(max(F(0), min(F(F(wis[0].end) - cmd), F(seg.start))))
            else:
                us.start = wis[0].start
            lw = wis[-1]
            if F(us.end) > F(lw.start) and F(F(seg.end) + F(0.5)) < F(lw.end):
                wis[-1].end = float(max(F(F(lw.start) + cmd), F(seg.end)))
            else:
                us.end = lw.end
            last = F(us.end)
        us.words = wis
        updated.append(us)
    return updated


def add_word_timestamps(segments: Sequence[Segment], alignmentWeights, split_to_word_tokens: Callable[[List[int]], Tuple[List[str], List[List[int]]]],
                        seek: int, lastSpeechTimestamp: float, specialTokenBegin: int, decode: Optional[Callable[[List[int]], str]] = None,
                        prependPunctuations: str = DEFAULT_PREPEND, appendPunctuations: str = DEFAULT_APPEND) -> List[Segment]:
    """SegmentSeeker.swift:410-496.  alignmentWeights rows are indexed by token position in the window (row i = token i)."""
    wordTokenIds: List[int] = []
    lps: List[float] = []
    idx: List[int] = []
    off = 0
    for seg in segments:
        for i, t in enumerate(seg.tokens):
            wordTokenIds.append(t)
            idx.append(i + off)
            lps.append(seg.tokenLogProbs[i])
        off += len(seg.tokens)
    w = np.asarray(alignmentWeights)
    filtered = w[idx] if idx else w[:0]
    words, wordTokens = split_to_word_tokens(wordTokenIds)
    alignment = find_alignment(words, wordTokens, filtered, lps) if len(idx) else []
    med, mx = calculate_word_duration_constraints(alignment)
    alignment = truncate_long_words_at_sentence_boundaries(alignment, mx)
    if alignment:
        alignment = merge_punctuations(alignment, prependPunctuations, appendPunctuations)
    return update_segments_with_word_timings(segments, alignment, seek, lastSpeechTimestamp, med, mx, specialTokenBegin, decode)
