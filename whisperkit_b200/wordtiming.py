"""Host mirror of the reference's word-timestamp pieces over the C ABI: SegmentSeeker's
dynamicTimeWarping / findAlignment / mergePunctuations / duration heuristics / updateSegmentsWithWordTimings / addWordTimestamps.
All logic lives in libwkb200.so (csrc/wordtiming.cu) and runs without a GPU; the tokenizer is the host's and is passed in as
callables (`split_to_word_tokens(tokens) -> (words, wordTokens)`, `decode(tokens) -> str`)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import DECODE_FN, SPLIT_FN, check, wk_segment, wk_tokenizer_hooks, wk_word


@dataclass
class WordTiming:
    """Models.swift:617-633."""
    word: str
    tokens: List[int]
    start: float
    end: float
    probability: float
    segment: int = -1

    @property
    def duration(self) -> float:
        return float(np.float32(self.end) - np.float32(self.start))


def _to_c(words: Sequence[WordTiming]):
    n = len(words)
    arr = (wk_word * max(1, n))()
    keep = []
    for i, w in enumerate(words):
        b = w.word.encode("utf-8")
        t = (C.c_int32 * max(1, len(w.tokens)))(*[int(v) for v in w.tokens])
        keep += [b, t]
        arr[i].word = b
        arr[i].tokens = C.cast(t, C.POINTER(C.c_int32))
        arr[i].n_tokens = len(w.tokens)
        arr[i].start, arr[i].end, arr[i].probability, arr[i].segment = float(w.start), float(w.end), float(w.probability), int(w.segment)
    return arr, n, keep


def _take(lib, h) -> List[WordTiming]:
    try:
        out = []
        w = wk_word()
        for i in range(lib.wk_words_count(h)):
            check(lib.wk_words_get(h, i, C.byref(w)))
            out.append(WordTiming(w.word.decode("utf-8"), [int(w.tokens[k]) for k in range(w.n_tokens)], float(w.start), float(w.end),
                                  float(w.probability), int(w.segment)))
        return out
    finally:
        lib.wk_words_free(h)


def _matrix(m):
    a = np.asarray(m)
    if a.dtype == np.float16:
        return np.ascontiguousarray(a), _lib.WK_DTYPE_F16
    return np.ascontiguousarray(a, dtype=np.float32), _lib.WK_DTYPE_F32


def make_hooks(split_to_word_tokens: Optional[Callable] = None, decode: Optional[Callable] = None):
    """wk_tokenizer_hooks over two Python callables; returns (struct, keepalive)."""
    def _split(user, toks, n, text, cap, counts, ccap):
        try:
            words, word_tokens = split_to_word_tokens([int(toks[i]) for i in range(n)])
            blob = b"".join(w.encode("utf-8") + b"\0" for w in words)
            if len(blob) > cap or len(words) > ccap:
                return -2
            C.memmove(text, blob, len(blob))
            for i, wt in enumerate(word_tokens):
                counts[i] = len(wt)
            return len(words)
        except Exception:  # noqa: BLE001 - must not propagate through the C frame
            return -1

    def _decode(user, toks, n, text, cap):
        try:
            b = decode([int(toks[i]) for i in range(n)]).encode("utf-8")
            if len(b) + 1 > cap:
                return -2
            C.memmove(text, b + b"\0", len(b) + 1)
            return len(b)
        except Exception:  # noqa: BLE001
            return -1

    h = wk_tokenizer_hooks()
    s = SPLIT_FN(_split) if split_to_word_tokens is not None else SPLIT_FN()
    d = DECODE_FN(_decode) if decode is not None else DECODE_FN()
    h.split_to_word_tokens, h.decode, h.user = s, d, None
    return h, (s, d)


class WordTimingSeeker:
    """The word-level half of SegmentSeeker (SegmentSeeker.swift:195-659)."""

    def __init__(self):
        self.lib = _lib.load()

    def dynamicTimeWarping(self, matrix) -> Tuple[List[int], List[int]]:
        a, dt = _matrix(matrix)
        if a.ndim != 2:
            raise _lib.WhisperError(-1, "Invalid alignment matrix shape")
        rows, cols = a.shape
        cap = rows + cols + 1
        ti, tj, n = (C.c_int32 * cap)(), (C.c_int32 * cap)(), C.c_int32()
        check(self.lib.wk_dtw(C.c_void_p(a.ctypes.data), dt, rows, cols, cols, ti, tj, cap, C.byref(n)))
        return list(ti[: n.value]), list(tj[: n.value])

    def findAlignment(self, words: Sequence[str], wordTokens: Sequence[Sequence[int]], alignmentWeights, tokenLogProbs: Sequence[float]) -> List[WordTiming]:
        a, dt = _matrix(alignmentWeights)
        arr, n, keep = _to_c([WordTiming(w, list(t), 0, 0, 0) for w, t in zip(words, wordTokens)])
        lp = (C.c_float * max(1, len(tokenLogProbs)))(*[float(v) for v in tokenLogProbs])
        h = C.c_void_p()
        check(self.lib.wk_find_alignment(arr, n, C.c_void_p(a.ctypes.data), dt, a.shape[0], a.shape[1], a.shape[1], lp, len(tokenLogProbs), C.byref(h)))
        return _take(self.lib, h)

    def mergePunctuations(self, alignment: Sequence[WordTiming], prepended: Optional[str] = None, appended: Optional[str] = None) -> List[WordTiming]:
        arr, n, keep = _to_c(alignment)
        h = C.c_void_p()
        check(self.lib.wk_merge_punctuations(arr, n, None if prepended is None else prepended.encode("utf-8"),
                                             None if appended is None else appended.encode("utf-8"), C.byref(h)))
        return _take(self.lib, h)

    def calculateWordDurationConstraints(self, alignment: Sequence[WordTiming]) -> Tuple[float, float]:
        arr, n, keep = _to_c(alignment)
        med, mx = C.c_float(), C.c_float()
        check(self.lib.wk_word_duration_constraints(arr, n, C.byref(med), C.byref(mx)))
        return med.value, mx.value

    def truncateLongWordsAtSentenceBoundaries(self, alignment: Sequence[WordTiming], maxDuration: float) -> List[WordTiming]:
        arr, n, keep = _to_c(alignment)
        h = C.c_void_p()
        check(self.lib.wk_truncate_long_words(arr, n, maxDuration, C.byref(h)))
        return _take(self.lib, h)

    @staticmethod
    def _segs_to_c(segments):
        n = len(segments)
        arr = (wk_segment * max(1, n))()
        toks, lps = [], []
        for i, s in enumerate(segments):
            arr[i].id, arr[i].seek, arr[i].start, arr[i].end = int(getattr(s, "id", 0)), int(getattr(s, "seek", 0)), float(s.start), float(s.end)
            arr[i].token_offset, arr[i].n_tokens = len(toks), len(s.tokens)
            toks += [int(t) for t in s.tokens]
            lp = list(getattr(s, "tokenLogProbs", []) or [])
            lps += [float(v) for v in lp] + [0.0] * (len(s.tokens) - len(lp))
        return arr, n, (C.c_int32 * max(1, len(toks)))(*toks), (C.c_float * max(1, len(lps)))(*lps)

    def updateSegmentsWithWordTimings(self, segments, mergedAlignment: Sequence[WordTiming], seek: int, lastSpeechTimestamp: float,
                                      constrainedMedianDuration: float, maxDuration: float, specialTokenBegin: int,
                                      decode: Optional[Callable] = None):
        """Returns [(start, end, words)] per segment."""
        sarr, ns, toks, _ = self._segs_to_c(segments)
        warr, nw, keep = _to_c(mergedAlignment)
        hooks, keep2 = make_hooks(None, decode)
        h = C.c_void_p()
        check(self.lib.wk_update_segments_with_word_timings(sarr, ns, toks, warr, nw, seek, lastSpeechTimestamp, constrainedMedianDuration,
                                                            maxDuration, specialTokenBegin, C.byref(hooks), C.byref(h)))
        words = _take(self.lib, h)
        return [(float(sarr[i].start), float(sarr[i].end), [w for w in words if w.segment == i]) for i in range(ns)]

    def addWordTimestamps(self, segments, alignmentWeights, split_to_word_tokens: Callable, seek: int, lastSpeechTimestamp: float,
                          specialTokenBegin: int, decode: Optional[Callable] = None, prependPunctuations: Optional[str] = None,
                          appendPunctuations: Optional[str] = None):
        """Returns [(start, end, words)] per segment."""
        sarr, ns, toks, lps = self._segs_to_c(segments)
        a, dt = _matrix(alignmentWeights)
        hooks, keep2 = make_hooks(split_to_word_tokens, decode)
        h = C.c_void_p()
        check(self.lib.wk_add_word_timestamps(sarr, ns, toks, lps, C.c_void_p(a.ctypes.data), dt, a.shape[0], a.shape[1], a.shape[1], C.byref(hooks),
                                              seek, lastSpeechTimestamp, specialTokenBegin,
                                              None if prependPunctuations is None else prependPunctuations.encode("utf-8"),
                                              None if appendPunctuations is None else appendPunctuations.encode("utf-8"), C.byref(h)))
        words = _take(self.lib, h)
        return [(float(sarr[i].start), float(sarr[i].end), [w for w in words if w.segment == i]) for i in range(ns)]
