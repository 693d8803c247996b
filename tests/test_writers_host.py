"""Result writers (ResultWriter.swift:12-134) through the C ABI.  The reference holds no test vectors for them, so the expected strings
are the Swift expressions evaluated in float32 (numpy) by hand: Float arithmetic, Int() truncation, %02d / %03d."""
from types import SimpleNamespace as NS

from whisperkit_b200 import writers as W


def test_format_time_matches_the_swift_expressions():
    assert W.formatTime(0.0, True, ",") == "00:00:00,000"
    assert W.formatTime(3.5, False, ".") == "00:03.500"
    assert W.formatTime(75.25, False, ".") == "01:15.250"
    assert W.formatTime(3725.125, False, ".") == "01:02:05.125"      # hours appear once non-zero
    assert W.formatTime(59.999, True, ",") == "00:00:59,999"         # Float(59.999) - 59 = 0.99900055 -> Int(999.0005) = 999
    assert W.formatTime(10.48, True, ",") == "00:00:10,479"          # 10.48 is not exact in Float: .4799995 * 1000 truncates


def test_srt_and_vtt_bodies():
    words = [NS(word=" And", start=0.0, end=0.32, tokens=[1], probability=1.0), NS(word=" so", start=0.32, end=0.5, tokens=[2], probability=1.0)]
    res = NS(text="And so my fellow", segments=[NS(start=0.0, end=0.5, text="<|0.00|> And so<|0.50|>", words=words),
                                                 NS(start=30.0, end=3661.5, text=" my fellow", words=None)])
    srt = W.WriteSRT("/tmp").render(res)
    assert srt == ("1\n00:00:00,000 --> 00:00:00,320\n And\n\n2\n00:00:00,320 --> 00:00:00,500\n so\n\n"
                   "3\n00:00:30,000 --> 01:01:01,500\n my fellow\n\n")
    vtt = W.WriteVTT("/tmp").render(res)
    assert vtt == ("WEBVTT\n\n00:00.000 --> 00:00.320\n And\n\n00:00.320 --> 00:00.500\n so\n\n00:30.000 --> 01:01:01.500\n my fellow\n\n")
    js = W.WriteJSON("/tmp").render(res)
    assert '"text": "And so my fellow"' in js and '"start": 30.0' in js


def test_write_files(tmp_path):
    res = NS(text="x", segments=[NS(start=1.0, end=2.0, text="x", words=None)])
    for cls, ext in ((W.WriteSRT, "srt"), (W.WriteVTT, "vtt"), (W.WriteJSON, "json")):
        p = cls(str(tmp_path)).write(res, "out")
        assert p.endswith("out." + ext) and open(p, encoding="utf-8").read()
