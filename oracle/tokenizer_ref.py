"""Tokenizer oracle (test infrastructure; see oracle/__init__.py): WhisperTokenizerWrapper's word splitting restated over any
`decode(tokens) -> str` (Sources/WhisperKit/Core/Models.swift:1224-1306).  `decode` itself is pinned in the tests against the
HuggingFace `tokenizers` library (the implementation swift-transformers mirrors).  Apple-only pieces are restated as in
csrc/tokenizer.cu: CharacterSet.punctuationCharacters = Unicode category P*, NLLanguageRecognizer -> script majority."""
from __future__ import annotations

import unicodedata
from typing import Callable, List, Sequence, Tuple

REPLACEMENT = "�"
_WS = " \t               　"


def prefers_unicode_split(text: str) -> bool:
    cjk = other = 0
    for ch in text:
        cp = ord(ch)
        target = (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2EBEF or 0x3040 <= cp <= 0x30FF or
                  0x0E00 <= cp <= 0x0E7F or 0x0E80 <= cp <= 0x0EFF or 0x1000 <= cp <= 0x109F)
        letter = ("a" <= ch <= "z") or ("A" <= ch <= "Z") or (cp >= 0xC0 and not unicodedata.category(ch).startswith("P") and ch not in _WS
                                                               and ch != REPLACEMENT)
        if target:
            cjk += 1
        elif letter:
            other += 1
    return cjk > other


def split_tokens_on_unicode(decode: Callable[[List[int]], str], tokens: Sequence[int]) -> Tuple[List[str], List[List[int]]]:
    full = decode(list(tokens)).encode("utf-8")
    words, groups, cur = [], [], []
    for t in tokens:
        cur.append(t)
        dec = decode(cur)
        at = dec.encode("utf-8").find(REPLACEMENT.encode("utf-8"))
        in_full = at >= 0 and full[at:at + 3] == REPLACEMENT.encode("utf-8")      # Models.swift:1238-1241 (offset from the start, see tokenizer.cu)
        if at < 0 or in_full:
            words.append(dec)
            groups.append(cur)
            cur = []
    return words, groups


def split_tokens_on_spaces(decode, tokens: Sequence[int], special_begin: int) -> Tuple[List[str], List[List[int]]]:
    sub, subg = split_tokens_on_unicode(decode, tokens)
    words: List[str] = []
    groups: List[List[int]] = []
    for w, g in zip(sub, subg):
        special = g[0] >= special_begin
        with_space = w.startswith(" ")
        stripped = w.strip(_WS)
        punctuation = len(stripped) == 1 and unicodedata.category(stripped).startswith("P")
        if special or with_space or punctuation or not words:
            words.append(w)
            groups.append(list(g))
        else:
            words[-1] += w
            groups[-1] += g
    return words, groups


def split_to_word_tokens(decode, tokens: Sequence[int], special_begin: int) -> Tuple[List[str], List[List[int]]]:
    plain = decode([t for t in tokens if t < special_begin])
    if prefers_unicode_split(plain):
        return split_tokens_on_unicode(decode, tokens)
    return split_tokens_on_spaces(decode, tokens, special_begin)
