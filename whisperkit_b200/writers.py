"""ResultWriting (Sources/WhisperKit/Utilities/ResultWriter.swift): formatTime, WriteSRT, WriteVTT, WriteJSON over the C ABI
(csrc/writers.cu).  `result` is a longform.TranscriptionResult (segments with start / end / text and optional word timings)."""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import asdict, is_dataclass

from . import _lib


def formatTime(seconds: float, alwaysIncludeHours: bool, decimalMarker: str) -> str:
    buf = C.create_string_buffer(64)
    n = _lib.load().wk_format_time(float(seconds), int(alwaysIncludeHours), decimalMarker.encode(), buf, 64)
    if n < 0:
        raise ValueError("formatTime: buffer too small")
    return buf.value.decode()


def _cues(result):
    """One cue per word when a segment has word timings, else the segment itself (ResultWriter.swift:79-92)."""
    out = []
    for g in result.segments:
        if getattr(g, "words", None):
            out += [(w.start, w.end, w.word) for w in g.words]
        else:
            out.append((g.start, g.end, g.text))
    return out


def _render(fn_name: str, result) -> str:
    cues = _cues(result)
    n = len(cues)
    st = (C.c_float * max(1, n))(*[c[0] for c in cues])
    en = (C.c_float * max(1, n))(*[c[1] for c in cues])
    tx = (C.c_char_p * max(1, n))(*[c[2].encode("utf-8") for c in cues])
    fn = getattr(_lib.load(), fn_name)
    cap = 256 + sum(len(c[2].encode("utf-8")) + 64 for c in cues)
    buf = C.create_string_buffer(cap)
    r = fn(st, en, tx, n, buf, cap)
    if r < 0:
        raise ValueError(f"{fn_name}: buffer too small")
    return buf.value.decode("utf-8")


class _Writer:
    ext = ""

    def __init__(self, outputDir: str):
        self.outputDir = outputDir

    def render(self, result) -> str:
        raise NotImplementedError

    def write(self, result, to: str, options=None) -> str:
        path = os.path.join(self.outputDir, f"{to}.{self.ext}")
        with open(path, "w", encoding="utf-8") as f:
            f.write(self.render(result))
        return path


class WriteSRT(_Writer):
    ext = "srt"

    def render(self, result) -> str:
        return _render("wk_write_srt", result)


class WriteVTT(_Writer):
    ext = "vtt"

    def render(self, result) -> str:
        return _render("wk_write_vtt", result)


class WriteJSON(_Writer):
    """WriteJSON (ResultWriter.swift:40-68): the Codable TranscriptionResult, pretty-printed."""
    ext = "json"

    def render(self, result) -> str:
        def enc(o):
            if is_dataclass(o):
                return asdict(o)
            if hasattr(o, "__dict__"):
                return dict(o.__dict__)
            raise TypeError(type(o))
        return json.dumps(result, default=enc, indent=2, ensure_ascii=False)
