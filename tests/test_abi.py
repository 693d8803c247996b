"""The C-ABI library loads on a CPU-only box, exports every symbol include/wkb200.h declares, and refuses to
compute without a GPU (no CPU fallback)."""
import os
import re

import pytest

import whisperkit_b200 as wk
from whisperkit_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "wkb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wk_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = wk.load()
    names = header_symbols()
    assert len(names) >= 30
    bound = {n for n, _, _ in _lib.SYMBOLS}
    for n in names:
        assert hasattr(lib, n), f"{n} declared in wkb200.h but not exported"
        assert n in bound, f"{n} has no ctypes prototype"
    assert lib.wk_version().startswith(b"wkb200")


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_lib.wk_model_config) == 40
    assert C.sizeof(_lib.wk_special_tokens) == 44
    assert C.sizeof(_lib.wk_decode_result) == 4 + 226 * 4 + 226 * 4 + 8 * 4
    assert _lib.wk_decode_opts.seed.offset % 8 == 0


def test_default_configs():
    import ctypes as C
    lib = wk.load()
    c = _lib.wk_model_config()
    lib.wk_default_config(b"large-v3", C.byref(c))
    assert (c.n_mels, c.d_model, c.n_heads, c.enc_layers, c.dec_layers, c.vocab) == (128, 1280, 20, 32, 32, 51866)
    lib.wk_default_config(b"tiny.en", C.byref(c))
    assert (c.n_mels, c.d_model, c.n_heads, c.enc_layers, c.dec_layers, c.vocab) == (80, 384, 6, 4, 4, 51864)
    lib.wk_default_config(b"large-v3-turbo", C.byref(c))
    assert c.dec_layers == 4
    lib.wk_default_config(b"distil-large-v3", C.byref(c))
    assert c.dec_layers == 2


def test_no_cpu_fallback():
    lib = wk.load()
    if lib.wk_device_available():
        pytest.skip("GPU present")
    with pytest.raises(wk.WhisperError) as e:
        wk.Model("toy")
    assert e.value.case == "modelsUnavailable"


def test_variant_table_reference_kat():
    """testTokenizerModelVariantDetection (UnitTests.swift:1101-1124) through wk_detect_variant, and wk_default_config for every variant."""
    import ctypes as C
    from whisperkit_b200 import _lib
    lib = _lib.load()
    kat = [(51865, 384, "tiny"), (51864, 384, "tiny.en"), (51865, 512, "base"), (51864, 512, "base.en"), (51865, 768, "small"), (51864, 768, "small.en"),
           (51865, 1024, "medium"), (51864, 1024, "medium.en"), (51865, 1280, "large-v2"), (51866, 1280, "large-v3")]
    for logits, enc, name in kat:
        v, repo, ml = C.c_char_p(), C.c_char_p(), C.c_int32()
        assert lib.wk_detect_variant(logits, enc, C.byref(v), C.byref(repo), C.byref(ml)) == 0
        assert v.value.decode() == name and repo.value.decode() == "openai/whisper-" + name
        assert bool(ml.value) == (not name.endswith(".en"))
        cfg = _lib.wk_model_config()
        lib.wk_default_config(name.encode(), C.byref(cfg))
        assert (cfg.vocab, cfg.d_model) == (logits, enc) and cfg.d_model == cfg.n_heads * 64 and cfg.n_mels == (128 if name == "large-v3" else 80)
    v = C.c_char_p()
    lib.wk_detect_variant(12345, 999, C.byref(v), None, None)
    assert v.value.decode() == "base"                                   # unrecognised vocabulary size -> base
