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


def _gpt2_alphabet():
    bs = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b); cs.append(256 + n); n += 1
    return bs, {b: chr(c) for b, c in zip(bs, cs)}


def _reference_mini_vocab():
    """The slice of the real Whisper (multilingual tiny) vocabulary that the reference's own tokenizer tests reveal
    (UnitTests.swift:1326-1375, 2494-2652): ids 0..255 are the GPT-2 byte alphabet in its canonical order (so 0 = '!', 11 = ',',
    30 = '?', 220 = ' ', and 171,120,223 = EF BC 81 = '！' exactly as the tests list them), plus the multi-byte tokens at their real ids."""
    order, alpha = _gpt2_alphabet()
    toks, ids, flags = [], [], []
    for i, b in enumerate(order):
        toks.append(alpha[b]); ids.append(i); flags.append(0)
    known = {2425: " Hello", 1002: " world", 639: " This", 307: " is", 257: " a", 31636: "test", 1943: " isn", 380: "'t", 309: " it",
             24364: "¡", 48529: "Hola", 376: " M", 6043: "undo", 20547: " Esta", 785: " es", 2002: " una", 48241: " prueba", 3841: " ¿", 1771: "no",
             38088: "こんにちは", 1231: "、", 24486: "世界", 25212: "これは", 22985: "テ", 40498: "スト", 4767: "です", 30346: "よね"}
    for i, text in known.items():
        toks.append("".join(alpha[b] for b in text.encode("utf-8"))); ids.append(i); flags.append(0)
    for i, name in {50257: "<|endoftext|>", 50258: "<|startoftranscript|>", 50363: "<|notimestamps|>", 50364: "<|0.00|>", 50414: "<|1.00|>"}.items():
        toks.append(name); ids.append(i); flags.append(3)
    return WhisperTokenizer(tokens=toks, ids=ids, flags=flags)


def test_split_to_word_tokens_reference_kats():
    """testSplitToWordTokens / ...Spanish / ...Japanese (UnitTests.swift:1326-1375) on the vocabulary slice those tests reveal."""
    wt = _reference_mini_vocab()
    assert wt.specialTokens.specialTokenBegin == 50257 and wt.specialTokens.whitespaceToken == 220   # " " is not a vocabulary string -> default 220
    ids = [50364, 2425, 11, 1002, 0, 50414, 50414, 639, 307, 257, 220, 31636, 11, 1943, 380, 309, 30, 50257]
    words, groups = wt.splitToWordTokens(ids)
    assert words == ["<|0.00|>", " Hello", ",", " world", "!", "<|1.00|>", "<|1.00|>", " This", " is", " a", " test", ",", " isn't", " it", "?", "<|endoftext|>"]
    assert groups == [[50364], [2425], [11], [1002], [0], [50414], [50414], [639], [307], [257], [220, 31636], [11], [1943, 380], [309], [30], [50257]]
    ids = [50363, 24364, 48529, 376, 6043, 0, 20547, 785, 2002, 48241, 11, 3841, 1771, 30, 50257]
    words, groups = wt.splitToWordTokens(ids)
    assert words == ["<|notimestamps|>", "¡Hola", " Mundo", "!", " Esta", " es", " una", " prueba", ",", " ¿no", "?", "<|endoftext|>"]
    assert groups == [[50363], [24364, 48529], [376, 6043], [0], [20547], [785], [2002], [48241], [11], [3841, 1771], [30], [50257]]
    ids = [50364, 38088, 1231, 24486, 171, 120, 223, 25212, 22985, 40498, 4767, 30346, 171, 120, 253, 50257]
    words, groups = wt.splitToWordTokens(ids)
    assert words == ["<|0.00|>", "こんにちは", "、", "世界", "！", "これは", "テ", "スト", "です", "よね", "？", "<|endoftext|>"]
    assert groups == [[50364], [38088], [1231], [24486], [171, 120, 223], [25212], [22985], [40498], [4767], [30346], [171, 120, 253], [50257]]
    # the same vocabulary reproduces the inputs of the reference's mergePunctuations tests from token ids alone
    assert wt.decode([2425, 11, 1002, 0]) == " Hello, world!" and wt.decode([3841, 1771, 30]) == " ¿no?"


def test_tokenizer_output_reference_kat():
    """testTokenizerOutput (UnitTests.swift:1288-1297): the large-v3 ids decode to the jfk sentence, special tokens verbatim.  The token
    strings come from the reference's word-timestamp goldens (UnitTests.swift:2703-2725), ids 50364 / 50889 are the large-v3 specials."""
    _, alpha = _gpt2_alphabet()
    words = {400: " And", 370: " so", 452: " my", 7177: " fellow", 6280: " Americans", 1029: " ask", 406: " not", 437: " what", 428: " your", 1941: " country",
             393: " can", 360: " do", 337: " for", 291: " you", 13: "."}
    toks = ["".join(alpha[b] for b in w.encode("utf-8")) for w in words.values()] + ["<|notimestamps|>", "<|10.48|>"]
    wt = WhisperTokenizer(tokens=toks, ids=list(words) + [50364, 50889], flags=[0] * len(words) + [3, 3])
    ids = [50364, 400, 370, 452, 7177, 6280, 1029, 406, 437, 428, 1941, 393, 360, 337, 291, 1029, 437, 291, 393, 360, 337, 428, 1941, 13, 50889]
    assert wt.decode(ids) == "<|notimestamps|> And so my fellow Americans ask not what your country can do for you ask what you can do for your country.<|10.48|>"
    assert wt.decode(ids, skipSpecialTokens=True) == " And so my fellow Americans ask not what your country can do for you ask what you can do for your country."
