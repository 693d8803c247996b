"""Decode-loop / logits-filter / sampler oracle (test infrastructure; see oracle/__init__.py).

Line-by-line CPU restatement of the reference's host logic:
  * LogitsFiltering impls     Sources/WhisperKit/Core/Text/LogitsFilter.swift:8-276
  * GreedyTokenSampler        Sources/WhisperKit/Core/Text/TokenSampler.swift:29-252
  * createLogitsFilters       Sources/WhisperKit/Core/TextDecoder.swift:857-899
  * prefillDecoderInputs      Sources/WhisperKit/Core/TextDecoder.swift:163-216
  * decodeText                Sources/WhisperKit/Core/TextDecoder.swift:541-855
  * DecodingFallback          Sources/WhisperKit/Core/Models.swift:357-381
  * compressionRatio          Sources/WhisperKit/Utilities/TextUtilities.swift:14-28
  * DecodingOptions defaults  Sources/WhisperKit/Core/Configurations.swift:184-246

Precision note: the reference filters/samples on f16 logits (FloatType = Float16,
ArgmaxCore/FloatType.swift:9-13) via Apple BNNS; this oracle works on the float32
logits it is handed (the engine's logits are fp32) - decisions are identical except
on exact f16 ties, which the known-answer tests do not contain.
"""
from __future__ import annotations

import math
import struct
import zlib
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

NEG_INF = -np.inf
MAX_TOKEN_CONTEXT = 448 // 2  # Models.swift:1334 Constants.maxTokenContext


@dataclass
class SpecialTokens:
    """Models.swift:1111-1149; defaults Models.swift:1309-1322 (base multilingual vocab)."""
    endToken: int = 50257
    englishToken: int = 50259
    noSpeechToken: int = 50362
    noTimestampsToken: int = 50363
    specialTokenBegin: int = 50257
    startOfPreviousToken: int = 50361
    startOfTranscriptToken: int = 50258
    timeTokenBegin: int = 50364
    transcribeToken: int = 50359
    translateToken: int = 50358
    whitespaceToken: int = 220

    @staticmethod
    def test_default(**kw) -> "SpecialTokens":
        """Tests/WhisperKitTests/TestUtils.swift:327-355: every id defaults to 0."""
        base = dict(endToken=0, englishToken=0, noSpeechToken=0, noTimestampsToken=0, specialTokenBegin=0,
                    startOfPreviousToken=0, startOfTranscriptToken=0, timeTokenBegin=0, transcribeToken=0,
                    translateToken=0, whitespaceToken=0)
        base.update(kw)
        return SpecialTokens(**base)

    @staticmethod
    def large_v3() -> "SpecialTokens":
        """Whisper large-v3 vocabulary (51866): one extra language token shifts the tail by 1."""
        return SpecialTokens(endToken=50257, englishToken=50259, noSpeechToken=50363, noTimestampsToken=50364,
                             specialTokenBegin=50257, startOfPreviousToken=50362, startOfTranscriptToken=50258,
                             timeTokenBegin=50365, transcribeToken=50360, translateToken=50359, whitespaceToken=220)

    @staticmethod
    def english_only() -> "SpecialTokens":
        """Whisper *.en vocabulary (51864, gpt2 tokenizer)."""
        return SpecialTokens(endToken=50256, englishToken=50258, noSpeechToken=50361, noTimestampsToken=50362,
                             specialTokenBegin=50256, startOfPreviousToken=50360, startOfTranscriptToken=50257,
                             timeTokenBegin=50363, transcribeToken=50358, translateToken=50357, whitespaceToken=220)

    @staticmethod
    def toy(vocab: int) -> "SpecialTokens":
        """Scaled-down layout for toy vocabularies: text < special < timestamps."""
        sb = vocab // 2
        return SpecialTokens(endToken=sb, englishToken=sb + 2, noSpeechToken=sb + 7, noTimestampsToken=sb + 8,
                             specialTokenBegin=sb, startOfPreviousToken=sb + 6, startOfTranscriptToken=sb + 1,
                             timeTokenBegin=sb + 9, transcribeToken=sb + 4, translateToken=sb + 3, whitespaceToken=3)


@dataclass
class DecodingOptions:
    """Configurations.swift:155-247 (fields on the hot path)."""
    task: str = "transcribe"
    language: Optional[str] = None
    temperature: float = 0.0
    temperatureIncrementOnFallback: float = 0.2
    temperatureFallbackCount: int = 5
    sampleLength: int = MAX_TOKEN_CONTEXT
    topK: int = 5
    usePrefillPrompt: bool = True
    skipSpecialTokens: bool = False
    withoutTimestamps: bool = False
    maxInitialTimestamp: Optional[float] = None
    promptTokens: Optional[List[int]] = None
    prefixTokens: Optional[List[int]] = None
    suppressBlank: bool = False
    suppressTokens: List[int] = field(default_factory=list)
    compressionRatioThreshold: Optional[float] = 2.4
    logProbThreshold: Optional[float] = -1.0
    firstTokenLogProbThreshold: Optional[float] = -1.5
    noSpeechThreshold: Optional[float] = 0.6


# ----------------------------------------------------------------------------------
# Logits filters (operate in place on a 1-D float array, like MLMultiArray [1,1,V])
# ----------------------------------------------------------------------------------
class SuppressTokensFilter:
    """LogitsFilter.swift:12-25."""

    def __init__(self, suppressTokens: Sequence[int]):
        self.suppressTokens = list(suppressTokens)

    def filterLogits(self, logits: np.ndarray, tokens: Sequence[int]) -> np.ndarray:
        for t in self.suppressTokens:
            logits[t] = NEG_INF
        return logits


class SuppressBlankFilter:
    """LogitsFilter.swift:27-51."""

    def __init__(self, specialTokens: SpecialTokens, sampleBegin: int):
        self.st = specialTokens
        self.sampleBegin = sampleBegin

    def filterLogits(self, logits, tokens):
        if len(tokens) != self.sampleBegin:
            return logits
        logits[self.st.whitespaceToken] = NEG_INF
        logits[self.st.endToken] = NEG_INF
        return logits


def log_softmax(x: np.ndarray) -> np.ndarray:
    m = np.max(x)
    if not np.isfinite(m):
        return x - m  # all -inf -> nan, as BNNS would
    e = np.exp((x - m).astype(np.float32))
    return (x - m) - np.log(np.sum(e, dtype=np.float32))


def logsumexp(x: np.ndarray) -> float:
    m = np.max(x)
    if m == NEG_INF:
        return NEG_INF
    return float(m + np.log(np.sum(np.exp((x - m).astype(np.float32)), dtype=np.float32)))


class TimestampRulesFilter:
    """LogitsFilter.swift:54-243."""

    def __init__(self, specialTokens: SpecialTokens, sampleBegin: int, maxInitialTimestampIndex: Optional[int],
                 isModelMultilingual: bool):
        self.st = specialTokens
        self.sampleBegin = sampleBegin
        self.maxInitialTimestampIndex = maxInitialTimestampIndex
        self.isModelMultilingual = isModelMultilingual

    def _sampleBegin(self, tokens) -> Optional[int]:
        # LogitsFilter.swift:131-142
        if self.isModelMultilingual:
            for i, t in enumerate(list(tokens)[:3]):
                if t == self.st.transcribeToken or t == self.st.translateToken:
                    return max(i + 1, self.sampleBegin)
            return None
        return self.sampleBegin

    def filterLogits(self, logits, tokens):
        tokens = list(tokens)
        sb = self._sampleBegin(tokens)
        if sb is None or not (sb <= len(tokens)):
            return logits  # :73-78
        st = self.st
        logits[st.noTimestampsToken] = NEG_INF  # :81
        if len(tokens) > sb:  # :83
            sampled = tokens[sb:]
            lastWasTimestamp = len(sampled) >= 1 and sampled[-1] >= st.timeTokenBegin
            penultimateWasTimestamp = len(sampled) < 2 or sampled[-2] >= st.timeTokenBegin
            if lastWasTimestamp:
                if penultimateWasTimestamp:
                    logits[st.timeTokenBegin:] = NEG_INF  # :90
                else:
                    logits[: st.endToken] = NEG_INF  # :93
            timestamps = [t for t in sampled if t >= st.timeTokenBegin]
            if timestamps:
                lastTimestamp = timestamps[-1]
                timestampLast = lastTimestamp if (lastWasTimestamp and not penultimateWasTimestamp) else lastTimestamp + 1
                logits[st.timeTokenBegin:timestampLast] = NEG_INF  # :108
        # "force initial timestamp" rule is commented out in the reference (:112-122)
        if self._sumOfProbabilityOverTimestampsIsAboveAnyOtherToken(logits, st.timeTokenBegin):
            logits[: st.timeTokenBegin] = NEG_INF  # :125-127
        return logits

    @staticmethod
    def _sumOfProbabilityOverTimestampsIsAboveAnyOtherToken(logits, timeTokenBegin) -> bool:
        # LogitsFilter.swift:144-242: logSoftmax over V, logSumExp over [ts:], max over [:ts], strict >
        if timeTokenBegin >= len(logits) or timeTokenBegin <= 0:
            # degenerate partitions: BNNS on an empty vector fails -> reference returns false
            if timeTokenBegin <= 0:
                return False
            return False
        logprobs = log_softmax(np.asarray(logits, dtype=np.float32))
        timestampLogProb = logsumexp(logprobs[timeTokenBegin:])
        maxTextTokenLogProb = float(np.max(logprobs[:timeTokenBegin]))
        if math.isnan(timestampLogProb) or math.isnan(maxTextTokenLogProb):
            return False
        return timestampLogProb > maxTextTokenLogProb


class LanguageLogitsFilter:
    """LogitsFilter.swift:245-276."""

    def __init__(self, allLanguageTokens, logitsDim: int, sampleBegin: int):
        self.allLanguageTokens = set(allLanguageTokens)
        self.logitsDim = logitsDim
        self.sampleBegin = sampleBegin

    def filterLogits(self, logits, tokens):
        if not (len(tokens) >= self.sampleBegin):
            return logits
        mask = np.ones(self.logitsDim, dtype=bool)
        mask[list(self.allLanguageTokens)] = False
        logits[mask] = NEG_INF
        return logits


def createLogitsFilters(options: DecodingOptions, prefilledIndex: int, initialPromptIndex: int,
                        st: SpecialTokens, isModelMultilingual: bool, custom=None, secondsPerTimeToken=0.02):
    """TextDecoder.swift:857-899 (order: custom, SuppressBlank, SuppressTokens, TimestampRules)."""
    allFilters = list(custom or [])
    if options.suppressBlank:
        allFilters.append(SuppressBlankFilter(st, sampleBegin=prefilledIndex))
    if options.suppressTokens:
        allFilters.append(SuppressTokensFilter([t for t in options.suppressTokens if t < st.specialTokenBegin]))
    if not options.withoutTimestamps:
        mi = int(options.maxInitialTimestamp / secondsPerTimeToken) if options.maxInitialTimestamp is not None else None
        allFilters.append(TimestampRulesFilter(st, sampleBegin=initialPromptIndex, maxInitialTimestampIndex=mi,
                                               isModelMultilingual=isModelMultilingual))
    return allFilters


# ----------------------------------------------------------------------------------
# Sampler
# ----------------------------------------------------------------------------------
@dataclass
class SamplingResult:
    tokens: List[int]
    logProbs: List[float]
    completed: bool


class GreedyTokenSampler:
    """TokenSampler.swift:29-252.  temperature == 0 -> argmax; else top-k multinomial
    (non-deterministic in the reference: Float.random, :61/:169; here `rng` is injectable)."""

    def __init__(self, temperature: float, eotToken: int, decodingOptions: DecodingOptions, rng=None):
        self.temperature = temperature
        self.eotToken = eotToken
        self.decodingOptions = decodingOptions
        self.rng = rng or np.random.default_rng(0)

    def _sample(self, logits: np.ndarray):
        x = np.asarray(logits, dtype=np.float32)
        if self.temperature != 0.0:
            x = x / np.float32(self.temperature)
        m = np.max(x)
        e = np.exp(x - m)
        probs = e / np.sum(e, dtype=np.float32)  # "always softmax once" (:55,:133)
        if self.temperature != 0.0:
            k = self.decodingOptions.topK
            idx = np.argsort(-probs, kind="stable")[:k]
            vals = probs[idx]
            rnd = float(np.sum(vals)) * float(self.rng.random())
            acc, chosen = 0.0, 0
            for i in range(len(vals)):
                acc += float(vals[i])
                if rnd < acc:
                    chosen = i
                    break
            tok = int(idx[chosen])
        else:
            tok = int(np.argmax(x))  # first maximal index
        return tok, float(np.log(probs[tok]))

    def update(self, tokens, logits, logProbs) -> SamplingResult:
        tok, lp = self._sample(logits)
        return SamplingResult(list(tokens) + [tok], list(logProbs) + [lp], tok == self.eotToken)  # :215-240

    def finalize(self, tokens, logProbs) -> SamplingResult:
        tokens, logProbs = list(tokens), list(logProbs)
        if not tokens or tokens[-1] != self.eotToken:  # :242-251
            tokens.append(self.eotToken)
            logProbs.append(0.0)
        return SamplingResult(tokens, logProbs, True)


# ----------------------------------------------------------------------------------
# Prompt, fallback, compression ratio
# ----------------------------------------------------------------------------------
def prefill_prompt(options: Optional[DecodingOptions], st: SpecialTokens, isModelMultilingual: bool,
                   languageToken: Optional[int] = None) -> List[int]:
    """prefillDecoderInputs (TextDecoder.swift:163-216).  `languageToken` stands in for
    tokenizer.convertTokenToId("<|lang|>") (default: englishToken)."""
    prefill = [st.startOfTranscriptToken]
    if options is not None:
        if isModelMultilingual:
            prefill.append(languageToken if languageToken is not None else st.englishToken)
            prefill.append(st.translateToken if options.task == "translate" else st.transcribeToken)
        prefill.append(st.noTimestampsToken if options.withoutTimestamps else st.timeTokenBegin)
        if options.promptTokens is not None:
            maxPromptLen = (MAX_TOKEN_CONTEXT // 2) - 1
            trimmed = [t for t in options.promptTokens[-maxPromptLen:] if t < st.specialTokenBegin]
            prefill = [st.startOfPreviousToken] + trimmed + prefill
        if options.prefixTokens is not None:
            trimmed = [t for t in options.prefixTokens[-(MAX_TOKEN_CONTEXT // 2):] if t < st.specialTokenBegin]
            prefill.extend(trimmed)
    return prefill


def compression_ratio(tokens: Sequence[int]) -> float:
    """TextUtilities.compressionRatio(of: [Int]) (TextUtilities.swift:14-28):
    len(Int32 LE bytes) / len(zlib(bytes)).  NSData.compressed(using: .zlib) emits a raw
    DEFLATE stream (no zlib header), level 5."""
    data = b"".join(struct.pack("<i", int(t)) for t in tokens)
    if not data:
        return float("inf")  # NSData compression of empty data throws -> infinity
    co = zlib.compressobj(5, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    return len(data) / len(comp)


@dataclass
class DecodingFallback:
    needsFallback: bool
    fallbackReason: str

    @staticmethod
    def make(options: DecodingOptions, isFirstTokenLogProbTooLow: bool, noSpeechProb: float,
             compressionRatio: float, avgLogProb: float) -> Optional["DecodingFallback"]:
        """Models.swift:357-381 (order matters)."""
        if isFirstTokenLogProbTooLow:
            return DecodingFallback(True, "firstTokenLogProbThreshold")
        if options.noSpeechThreshold is not None and noSpeechProb > options.noSpeechThreshold:
            return DecodingFallback(False, "silence")
        if options.compressionRatioThreshold is not None and compressionRatio > options.compressionRatioThreshold:
            return DecodingFallback(True, "compressionRatioThreshold")
        if options.logProbThreshold is not None and avgLogProb < options.logProbThreshold:
            return DecodingFallback(True, "logProbThreshold")
        return None


@dataclass
class DecodingResult:
    tokens: List[int]
    tokenLogProbs: List[float]
    avgLogProb: float
    compressionRatio: float
    temperature: float
    fallback: Optional[DecodingFallback]
    # extras for parity checks (not in the reference struct)
    currentTokens: List[int] = field(default_factory=list)
    logProbs: List[float] = field(default_factory=list)
    steps: int = 0
    isFirstTokenLogProbTooLow: bool = False
    stepLogits: List[np.ndarray] = field(default_factory=list)
    stepMargins: List[float] = field(default_factory=list)


def decode_text(predict_logits: Callable[[int, int], np.ndarray], initialPrompt: Sequence[int],
                options: DecodingOptions, st: SpecialTokens, isModelMultilingual: bool,
                sampler: Optional[GreedyTokenSampler] = None, custom_filters=None,
                prefilledIndex: int = 0, keep_logits: bool = False) -> DecodingResult:
    """decodeText (TextDecoder.swift:541-855).  `predict_logits(token, tokenIndex)` is the
    model call (predictLogits, :616-626); it must update its own KV cache at `tokenIndex`."""
    sampler = sampler or GreedyTokenSampler(options.temperature, st.endToken, options)
    initialPromptIndex = len(initialPrompt)
    currentTokens = list(initialPrompt)
    nextToken = initialPrompt[-1]
    logProbs = [0.0] * len(currentTokens)
    filters = createLogitsFilters(options, prefilledIndex, initialPromptIndex, st, isModelMultilingual, custom_filters)
    loopCount = min(options.sampleLength, MAX_TOKEN_CONTEXT - 1)  # :566
    isFirstTokenLogProbTooLow = False
    steps = 0
    stepLogits, stepMargins = [], []
    for tokenIndex in range(prefilledIndex, loopCount):
        isPrefill = tokenIndex < initialPromptIndex - 1
        isLastPrefillToken = tokenIndex == initialPromptIndex - 1
        isFirstToken = tokenIndex == prefilledIndex
        if tokenIndex < initialPromptIndex:  # :581-594
            isTimestampToken = currentTokens[tokenIndex] >= st.timeTokenBegin
            modelPredictedTimestamp = nextToken >= st.timeTokenBegin
            if not (isLastPrefillToken and isTimestampToken and modelPredictedTimestamp):
                nextToken = currentTokens[tokenIndex]
            else:
                currentTokens[tokenIndex] = nextToken
        logits = np.array(predict_logits(nextToken, tokenIndex), dtype=np.float32).reshape(-1)
        if keep_logits:
            stepLogits.append(logits.copy())
        for f in filters:  # :641-643
            logits = f.filterLogits(logits, currentTokens)
        res = sampler.update(currentTokens, logits, logProbs)  # :652
        nextToken = res.tokens[-1]
        nextTokenLogProb = res.logProbs[-1]
        if keep_logits:
            srt = np.sort(logits[np.isfinite(logits)])
            stepMargins.append(float(srt[-1] - srt[-2]) if len(srt) > 1 else float("inf"))
        steps += 1
        isFirstTokenLogProbTooLow = bool(
            isFirstToken and options.firstTokenLogProbThreshold is not None
            and nextTokenLogProb < options.firstTokenLogProbThreshold)  # :662-667
        isSegmentCompleted = res.completed or len(currentTokens) >= MAX_TOKEN_CONTEXT - 1 or isFirstTokenLogProbTooLow
        if isSegmentCompleted:
            break  # sampled token NOT appended (:673-678)
        if not isPrefill:
            currentTokens.append(nextToken)
            logProbs.append(nextTokenLogProb)
    fin = sampler.finalize(currentTokens, logProbs)
    segmentTokens, segmentLogProbs = fin.tokens, fin.logProbs
    startIndex = segmentTokens.index(st.startOfTranscriptToken) if st.startOfTranscriptToken in segmentTokens else 0
    endIndex = segmentTokens.index(st.endToken) if st.endToken in segmentTokens else len(segmentTokens)
    filteredTokens = segmentTokens[startIndex:endIndex + 1]
    filteredLogProbs = segmentLogProbs[startIndex:endIndex + 1]
    s = np.float32(0.0)
    for v in filteredLogProbs:  # Float reduce(0,+) in order (:785)
        s = np.float32(s + np.float32(v))
    avgLogProbs = float(s / np.float32(len(filteredLogProbs)))
    wordTokens = [t for t in filteredTokens if t < st.specialTokenBegin]
    finalCompressionRatio = compression_ratio(wordTokens)
    temperature = round(float(np.float16(sampler.temperature)), 3)  # :796-800
    fb = DecodingFallback.make(options, isFirstTokenLogProbTooLow, 0.0, finalCompressionRatio, avgLogProbs)
    return DecodingResult(filteredTokens, filteredLogProbs, avgLogProbs, finalCompressionRatio, temperature, fb,
                          currentTokens=currentTokens, logProbs=logProbs, steps=steps,
                          isFirstTokenLogProbTooLow=isFirstTokenLogProbTooLow,
                          stepLogits=stepLogits, stepMargins=stepMargins)
