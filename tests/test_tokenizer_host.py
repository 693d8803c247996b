"""Decode side of the tokenizer (csrc/tokenizer.cu through the C ABI), no GPU:
  * decode() against the HuggingFace `tokenizers` library (byte-level BPE + added tokens, the implementation swift-transformers mirrors)
    on a tokenizer.json written by that library, including invalid UTF-8 splits and skipSpecialTokens;
  * splitToWordTokens against the oracle restatement (oracle/tokenizer_ref.py) and the words of the reference's own word-timing tests;
  * special-token lookups with the reference's defaults."""
import json
import os

import numpy as np
import pytest

tokenizers = pytest.importorskip("tokenizers")

from oracle import tokenizer_ref as TR  # noqa: E402
from whisperkit_b200.tokenizer import WhisperTokenizer  # noqa: E402

CORPUS = ["And so my fellow Americans, ask not what your country can do for you, ask what you can do for your country.",
          "Hello, world! This is a test, isn't it?", "¡Hola Mundo! Esta es una prueba, ¿no?", "こんにちは、世界！これはテストですよね？",
          "สวัสดีชาวโลก", "naïve café — “quoted” text… 3.14 § 42", "emoji 🙂👍🏽 and tabs\tand  double  spaces"]
SPECIALS = ["<|endoftext|>", "<|startoftranscript|>", "<|en|>", "<|ja|>", "<|translate|>", "<|transcribe|>", "<|startofprev|>", "<|nospeech|>",
            "<|notimestamps|>"] + [f"<|{i * 0.02:.2f}|>" for i in range(0, 101)]


@pytest.fixture(scope="module")
def toks(tmp_path_factory):
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, trainers
    hf = Tokenizer(models.BPE())
    hf.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    hf.decoder = decoders.ByteLevel()
    tr = trainers.BpeTrainer(vocab_size=600, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), special_tokens=[])
    hf.train_from_iterator(CORPUS * 10, tr)
    hf.add_special_tokens(SPECIALS)
    d = tmp_path_factory.mktemp("tok")
    hf.save(str(d / "tokenizer.json"))
    return hf, WhisperTokenizer(str(d)), str(d)


def hf_decode(hf, ids, skip=False):
    # tokenizers' own clean-up is off; apply the reference's cleanUp rules (Tokenizer.swift:434-447) on top
    s = hf.decode(list(ids), skip_special_tokens=skip)
    for a, b in ((" .", "."), (" ?", "?"), (" !", "!"), (" ,", ","), (" ' ", "'"), (" n't", "n't"), (" 'm", "'m"), (" 's", "'s"), (" 've", "'ve"),
                 (" 're", "'re")):
        s = s.replace(a, b)
    return s


def test_decode_matches_huggingface_tokenizers(toks):
    hf, wt, _ = toks
    assert wt.vocabSize == hf.get_vocab_size()
    for text in CORPUS:
        ids = hf.encode(" " + text).ids
        assert wt.decode(ids) == hf_decode(hf, ids)
        for cut in range(1, len(ids)):                       # every prefix: multi-byte characters split across tokens -> U+FFFD
            assert wt.decode(ids[:cut]) == hf_decode(hf, ids[:cut]), (text, cut)
    rng = np.random.default_rng(0)
    V = hf.get_vocab_size()
    for _ in range(300):                                     # random ids: arbitrary (mostly invalid) byte sequences + added tokens
        ids = [int(v) for v in rng.integers(0, V, int(rng.integers(1, 30)))]
        assert wt.decode(ids) == hf_decode(hf, ids)
        assert wt.decode(ids, skipSpecialTokens=True) == hf_decode(hf, ids, skip=True)
    assert wt.decode([]) == "" and wt.decode([V + 5, -1]) == ""          # unknown ids are dropped (compactMap)


def test_special_tokens_and_defaults(toks):
    hf, wt, _ = toks
    st = wt.specialTokens
    assert st.endToken == hf.token_to_id("<|endoftext|>") == st.specialTokenBegin
    assert st.startOfTranscriptToken == hf.token_to_id("<|startoftranscript|>") and st.timeTokenBegin == hf.token_to_id("<|0.00|>")
    assert st.englishToken == hf.token_to_id("<|en|>") and st.noTimestampsToken == hf.token_to_id("<|notimestamps|>")
    assert wt.convertTokenToId("<|ja|>") == hf.token_to_id("<|ja|>") and wt.convertTokenToId("<|zz|>") is None
    bare = WhisperTokenizer(tokens=["a", "b"], ids=[0, 1])
    d = bare.specialTokens                                    # Models.swift:1309-1322 defaults when the vocabulary lacks the tokens
    assert (d.endToken, d.startOfTranscriptToken, d.englishToken, d.transcribeToken, d.translateToken, d.noSpeechToken, d.noTimestampsToken,
            d.timeTokenBegin, d.startOfPreviousToken, d.whitespaceToken, d.specialTokenBegin) == \
        (50257, 50258, 50259, 50359, 50358, 50362, 50363, 50364, 50361, 220, 50257)


def test_vocab_json_plus_added_tokens_layout(toks, tmp_path):
    hf, wt, d = toks
    data = json.load(open(os.path.join(d, "tokenizer.json"), encoding="utf-8"))
    json.dump(data["model"]["vocab"], open(tmp_path / "vocab.json", "w", encoding="utf-8"), ensure_ascii=True)     # \\uXXXX escapes exercised
    json.dump({a["content"]: a["id"] for a in data["added_tokens"]}, open(tmp_path / "added_tokens.json", "w", encoding="utf-8"))
    wt2 = WhisperTokenizer(str(tmp_path))
    ids = hf.encode(" " + CORPUS[3]).ids + [hf.token_to_id("<|endoftext|>")]
    assert wt2.decode(ids) == wt.decode(ids) == hf_decode(hf, ids)
    with pytest.raises(Exception):
        WhisperTokenizer(str(tmp_path / "missing"))


def test_split_to_word_tokens_matches_oracle(toks):
    hf, wt, _ = toks
    sb = wt.specialTokens.specialTokenBegin
    sot, t0, t1, eot = (hf.token_to_id(s) for s in ("<|startoftranscript|>", "<|0.00|>", "<|1.00|>", "<|endoftext|>"))
    dec = lambda ids: wt.decode(ids)                          # noqa: E731
    for text in CORPUS:
        body = hf.encode(" " + text).ids
        ids = [sot, t0] + body + [t1, eot]
        words, groups = wt.splitToWordTokens(ids)
        rw, rg = TR.split_to_word_tokens(dec, ids, sb)
        assert words == rw and groups == rg, text
        assert sum(groups, []) == ids and "".join(words) != ""
    # English: leading-space words, punctuation split off, specials on their own (the shape mergePunctuations' KATs start from)
    ids = [t0] + hf.encode(" Hello, world! This is a test, isn't it?").ids + [eot]
    words, groups = wt.splitToWordTokens(ids)
    assert words[0] == "<|0.00|>" and words[-1] == "<|endoftext|>"
    assert [w for w in words if w.strip() in (",", "!", "?")] == [",", "!", ",", "?"]
    assert " Hello" in words and " world" in words and " isn't" in words
    # Japanese: script majority -> unicode split, every word is whole characters
    ids = hf.encode("こんにちは、世界！これはテストですよね？").ids
    words, groups = wt.splitToWordTokens(ids)
    assert "".join(words) == "こんにちは、世界！これはテストですよね？" and all("�" not in w for w in words)
    rng = np.random.default_rng(1)
    V = hf.get_vocab_size()
    for _ in range(100):
        ids = [int(v) for v in rng.integers(0, V, int(rng.integers(1, 25)))]
        rw, rg = TR.split_to_word_tokens(dec, ids, sb)
        assert wt.splitToWordTokens(ids) == ([w.replace(chr(0), "") for w in rw], rg)   # a NUL byte cannot travel in the C-string layout


def test_word_timestamps_with_builtin_tokenizer_hooks(toks):
    """wk_add_word_timestamps driven by the library's own tokenizer hooks (no host callbacks) == the same call through Python callables."""
    import ctypes as C
    from oracle import words_ref as W
    from whisperkit_b200 import wordtiming as T
    from whisperkit_b200._lib import check
    hf, wt, _ = toks
    sb = wt.specialTokens.specialTokenBegin
    sot, t0, t1, eot = (hf.token_to_id(s) for s in ("<|startoftranscript|>", "<|0.00|>", "<|2.00|>", "<|endoftext|>"))
    body = hf.encode(" Hello, world! This is a test.").ids
    toks_ = [sot, t0] + body + [t1, eot]
    seg = W.Segment(0.0, 2.0, toks_, [-0.1] * len(toks_))
    rng = np.random.default_rng(2)
    m = (rng.random((len(toks_), 150)) * 0.01).astype(np.float32)
    for r, c in enumerate(np.sort(rng.integers(0, 150, len(toks_)))):
        m[r, max(0, c - 1):c + 2] += 0.5
    seeker = T.WordTimingSeeker()
    via_py = seeker.addWordTimestamps([seg], m, wt.splitToWordTokens, 0, 0.0, sb, decode=wt.decode)
    sarr, ns, tk, lps = seeker._segs_to_c([seg])
    hooks = wt.hooks()
    h = C.c_void_p()
    check(seeker.lib.wk_add_word_timestamps(sarr, ns, tk, lps, C.c_void_p(m.ctypes.data), 0, m.shape[0], m.shape[1], m.shape[1], C.byref(hooks), 0, 0.0,
                                            sb, None, None, C.byref(h)))
    via_c = T._take(seeker.lib, h)
    assert [w.word for w in via_c] == [w.word for w in via_py[0][2]] and len(via_c) >= 4
    assert [w.tokens for w in via_c] == [w.tokens for w in via_py[0][2]]
    np.testing.assert_array_equal(np.float32([w.start for w in via_c]), np.float32([w.start for w in via_py[0][2]]))
    ref = W.add_word_timestamps([seg], m, lambda t: TR.split_to_word_tokens(wt.decode, t, sb), 0, 0.0, sb, decode=wt.decode)
    assert [w.word for w in ref[0].words] == [w.word for w in via_c]


def test_encode_matches_huggingface_tokenizers(toks):
    """text -> ids (pre-tokenizer pattern + byte alphabet + BPE merges + verbatim added tokens) against `tokenizers` on the tokenizer.json
    it wrote, without the post-processor on either side."""
    hf, wt, _ = toks
    cases = CORPUS + ["", " ", "  ", "a", " a", "a ", "  leading and trailing  ", "tabs\t\tand\nnewlines\n\n x", "it's we're I'll they'd don't 'tis 'Twas",
                      "num83r5 1234567 3.14159 ½ ²", "mixed-CASE_snake_case::path/to/file.txt", "<|startoftranscript|><|en|> hello<|0.00|> world<|endoftext|>",
                      "x<|notimestamps|>y <|nospeech|>", "😀 emoji 👩‍👩‍👧‍👦 zwj", "한국어 Русский العربية עברית ไทย", " nbsp emspace　ideographic",
                      "a  b   c    d", "trailing space after punctuation !  ?", "'", "''s", "' s", "x's'", "<|", "<|not a token|>"]
    for text in cases:
        assert wt.encode(text) == hf.encode(text, add_special_tokens=False).ids, repr(text)
        assert wt.decode(wt.encode(text)) == hf_decode(hf, hf.encode(text, add_special_tokens=False).ids)
    rng = np.random.default_rng(5)
    alphabet = list(" \t\n'.,!?-_:;()[]{}<>|/\\\"0123456789") + list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ") + \
        list("áéíóúüñçßøåœ") + list("こんにちは世界テスト") + list("привет") + list("🙂👍") + ["'s", "'re", "'ll", " ", "  ", "<|en|>", "<|0.00|>"]
    for _ in range(400):
        text = "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), int(rng.integers(1, 40))))
        assert wt.encode(text) == hf.encode(text, add_special_tokens=False).ids, repr(text)


def test_encode_from_vocab_json_and_merges_txt(toks, tmp_path):
    hf, wt, d = toks
    data = json.load(open(os.path.join(d, "tokenizer.json"), encoding="utf-8"))
    json.dump(data["model"]["vocab"], open(tmp_path / "vocab.json", "w", encoding="utf-8"))
    json.dump({a["content"]: a["id"] for a in data["added_tokens"]}, open(tmp_path / "added_tokens.json", "w", encoding="utf-8"))
    with open(tmp_path / "merges.txt", "w", encoding="utf-8") as f:
        f.write("#version: 0.2\n")
        for m in data["model"]["merges"]:
            f.write((m if isinstance(m, str) else " ".join(m)) + "\n")
    wt2 = WhisperTokenizer(str(tmp_path))
    for text in CORPUS:
        assert wt2.encode(" " + text) == hf.encode(" " + text, add_special_tokens=False).ids
