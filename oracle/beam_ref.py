"""Beam-search decoding oracle (test infrastructure; see oracle/__init__.py).  PARITY: SELF-ORACLE ONLY.

The reference has no beam search: `BeamSearchTokenSampler.update/finalize` are `fatalError("Not implemented")` stubs
(Sources/WhisperKit/Core/Text/TokenSampler.swift:254-290) and `DecodingOptions` has no beam field (Configurations.swift:155-183).  The only
semantics the stub fixes are `maxCandidates = Int(Float(beamSize) * patience)` and `patience = 1` (:261-275).  SURVEY.md section 8(f) row 2
therefore asks for openai/whisper's `BeamSearchDecoder` (whisper/decoding.py, external, restated here from its published algorithm)
placed inside WhisperKit's decodeText loop (TextDecoder.swift:541-855, restated in oracle/decode_ref.py):

  * prompt tokens are still pushed one step at a time with filters + sampling applied and the sample discarded (all beams are identical
    copies during prefill, so the greedy bookkeeping of beam 0 stands for all of them, including the model-predicted first timestamp
    rule and the first-token log-prob threshold);
  * from the last prompt slot on, every step ranks the candidates `sum_logprob[beam] + log_softmax(filtered logits)[token]` of the
    top (beamSize + 1) tokens of every beam (whisper/decoding.py BeamSearchDecoder.update): walking them best first, a candidate ending in
    EOT joins the finished list (while it holds fewer than maxCandidates), any other becomes one of the next beamSize beams; identical
    beams (the first ranking step) count once, as the dict keyed by the token sequence does in the original;
  * the window completes when maxCandidates sequences have finished, at 223 tokens, at sampleLength steps (loop bound), or on the
    first-token rule; `finalize` tops the finished list up with the live beams (best sum first) to beamSize entries, and
    MaximumLikelihoodRanker with length_penalty = None picks argmax(sum_logprob / number of sampled tokens) (divisor at least 1: the
    original divides by zero for an empty sequence);
  * the chosen sequence then goes through decodeText's own result assembly (EOT appended with log-prob 0 like sampler.finalize, SOT..EOT
    slice, avgLogProb over the slice, compressionRatio, DecodingFallback).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np

from .decode_ref import (MAX_TOKEN_CONTEXT, DecodingFallback, DecodingOptions, DecodingResult, SpecialTokens, compression_ratio,
                         createLogitsFilters)


def log_softmax_f32(row: np.ndarray) -> np.ndarray:
    x = row.astype(np.float32)
    m = np.max(x)
    if not np.isfinite(m):
        return np.full_like(x, -np.inf)
    lse = m + np.log(np.sum(np.exp((x - m).astype(np.float64)))).astype(np.float32)
    return (x - np.float32(lse)).astype(np.float32)


def top_candidates(logprobs: np.ndarray, k: int):
    """The k best (token, logprob) pairs, best first; equal values keep the lower token id first."""
    idx = np.argsort(-logprobs, kind="stable")[:k]
    return [(int(i), np.float32(logprobs[i])) for i in idx if np.isfinite(logprobs[i])]


def decode_text_beam(predict_logits: Callable[[List[List[int]], int], np.ndarray], initialPrompt: Sequence[int], options: DecodingOptions,
                     st: SpecialTokens, isModelMultilingual: bool, beamSize: int, patience: float = 1.0, trace: list = None) -> DecodingResult:
    """`predict_logits(prefixes, tokenIndex)` -> logits [beamSize, V]: row j is the decoder output at position tokenIndex after feeding
    prefixes[j][0..tokenIndex] (the model side keeps or rebuilds each beam's KV cache; the original's rearrange_kv_cache)."""
    maxCandidates = int(np.float32(beamSize) * np.float32(patience))          # TokenSampler.swift:266
    assert beamSize >= 1 and maxCandidates >= 1
    P = len(initialPrompt)
    beams = [list(initialPrompt) for _ in range(beamSize)]
    beam_lps = [[0.0] * P for _ in range(beamSize)]
    sums = [np.float32(0.0)] * beamSize
    nextToken = initialPrompt[-1]
    filters = createLogitsFilters(options, 0, P, st, isModelMultilingual, None)
    loopCount = min(options.sampleLength, MAX_TOKEN_CONTEXT - 1)
    finished = []                      # (tokens incl. EOT, per-token logprobs (EOT -> 0), score), insertion ordered
    firstLow = False
    steps = 0
    for tokenIndex in range(0, loopCount):
        isPrefill = tokenIndex < P - 1
        if tokenIndex < P:                                                        # TextDecoder.swift:581-594, identical for every beam
            cur = beams[0][tokenIndex]
            if tokenIndex == P - 1 and cur >= st.timeTokenBegin and nextToken >= st.timeTokenBegin:
                for b in beams:
                    b[tokenIndex] = nextToken
        logits = np.asarray(predict_logits([list(b) for b in beams], tokenIndex), dtype=np.float32)
        steps += 1
        considered = range(beamSize) if tokenIndex > P - 1 else range(1)         # identical beams count once
        cand = []                                                                 # (score, beam, token, logprob) in insertion order
        greedy = None
        for j in considered:
            row = logits[j].copy()
            for f in filters:
                row = f.filterLogits(row, beams[j])
            lp = log_softmax_f32(row)
            tops = top_candidates(lp, beamSize + 1)
            if j == 0:
                greedy = tops[0] if tops else (st.endToken, np.float32(-np.inf))
            for tok, v in tops:
                cand.append((np.float32(sums[j] + v), j, tok, v))
        firstLow = bool(tokenIndex == 0 and options.firstTokenLogProbThreshold is not None and greedy[1] < options.firstTokenLogProbThreshold)
        nextToken = greedy[0]
        if isPrefill:
            if greedy[0] == st.endToken or firstLow:                             # EOT sampled during prefill terminates (:668-671)
                break
            continue
        if len(beams[0]) >= MAX_TOKEN_CONTEXT - 1 or firstLow:
            break
        order = sorted(range(len(cand)), key=lambda i: -cand[i][0])               # stable: ties keep (beam, rank) order
        new_beams, new_lps, new_sums = [], [], []
        for i in order:
            score, j, tok, v = cand[i]
            if tok == st.endToken:
                if len(finished) < maxCandidates:
                    finished.append((beams[j] + [tok], beam_lps[j] + [0.0], score))
            else:
                new_beams.append(beams[j] + [tok])
                new_lps.append(beam_lps[j] + [float(v)])
                new_sums.append(score)
                if len(new_beams) == beamSize:
                    break
        if trace is not None:
            trace.append(dict(tokenIndex=tokenIndex, beams=[list(b) for b in new_beams], sums=[float(x) for x in new_sums], finished=len(finished)))
        beams, beam_lps, sums = new_beams, new_lps, new_sums
        while len(beams) < beamSize:                                              # cannot happen with beamSize + 1 candidates per beam
            beams.append(list(beams[-1])); beam_lps.append(list(beam_lps[-1])); sums.append(np.float32(-np.inf))
        nextToken = beams[0][-1]
        if len(finished) >= maxCandidates:
            break
    # finalize: not enough finished sequences -> the live beams, best sum first (BeamSearchDecoder.finalize)
    if len(finished) < beamSize:
        for j in sorted(range(len(beams)), key=lambda i: -sums[i]):
            finished.append((beams[j] + [st.endToken], beam_lps[j] + [0.0], sums[j]))
            if len(finished) >= beamSize:
                break

    def rank(entry):
        toks, _, score = entry
        length = max(len(toks) - P - 1, 1)                                        # sampled tokens before EOT
        return np.float32(score) / np.float32(length)
    best = max(range(len(finished)), key=lambda i: (rank(finished[i]), -i))
    segmentTokens, segmentLogProbs, _ = finished[best]
    startIndex = segmentTokens.index(st.startOfTranscriptToken) if st.startOfTranscriptToken in segmentTokens else 0
    endIndex = segmentTokens.index(st.endToken) if st.endToken in segmentTokens else len(segmentTokens)
    filteredTokens = segmentTokens[startIndex:endIndex + 1]
    filteredLogProbs = segmentLogProbs[startIndex:endIndex + 1]
    s = np.float32(0.0)
    for v in filteredLogProbs:
        s = np.float32(s + np.float32(v))
    avg = float(s / np.float32(len(filteredLogProbs)))
    ratio = compression_ratio([t for t in filteredTokens if t < st.specialTokenBegin])
    fb = DecodingFallback.make(options, firstLow, 0.0, ratio, avg)
    return DecodingResult(filteredTokens, filteredLogProbs, avg, ratio, round(float(np.float16(options.temperature)), 3), fb,
                          currentTokens=segmentTokens[:-1], logProbs=segmentLogProbs[:-1], steps=steps, isFirstTokenLogProbTooLow=firstLow)
