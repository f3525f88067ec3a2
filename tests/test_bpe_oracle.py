"""Pin the C BPE oracle (oracle/bpe_oracle.c) to the reference: golden files written by the live
reference (HF tokenizers under FIGBPE) and, where the `tokenizers` wheel is importable, the
library itself on fresh corpora.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle.bpe_oracle import OracleBPE, pretokenize

BPE_CASES = ["bpe_d14", "bpe_d14_small", "bpe_v1000"]


def ref_files(name):
    d = os.path.join(GOLDEN, f"{name}_pretrained", "bpe_tokenizer")
    return (open(os.path.join(d, "vocab.json"), encoding="utf-8").read(),
            open(os.path.join(d, "merges.txt"), encoding="utf-8").read())


@pytest.mark.parametrize("name", BPE_CASES)
def test_train_reproduces_reference_files(name):
    g = load_golden(name)
    vocab_json, merges_txt = ref_files(name)
    o = OracleBPE.train(g["corpus_bins"].astype(np.int64), int(g["bpe_vocab_size"]))
    assert (o.min_token, o.max_token) == (int(g["min_token"]), int(g["max_token"]))
    assert o.merges_txt() == merges_txt
    assert o.vocab_json() == vocab_json                       # byte-identical vocab.json
    assert json.loads(vocab_json) == o.vocab_dict()


@pytest.mark.parametrize("name", BPE_CASES)
def test_encode_decode_match_reference(name):
    g = load_golden(name)
    vocab_json, merges_txt = ref_files(name)
    o = OracleBPE.from_strings(json.loads(vocab_json), merges_txt.splitlines()[1:])
    ids_ref = np.split(g["ids_flat"], np.cumsum(g["ids_len"])[:-1])
    mn = int(g["min_token"])
    for row, want in zip(g["mp_tokens"], ids_ref):
        got = o.encode(row - mn)
        assert got == want.tolist()
        assert np.array_equal(o.decode(got) + mn, row)
    assert np.array_equal(g["bpe_to_mp"], g["mp_tokens"])


def test_pretokenizer_examples():
    """SURVEY.md Appendix A.2 examples (probed against the library)."""
    def pieces(s):
        cp = [ord(c) for c in s]
        ws = pretokenize(cp)
        idx = [i for i, w in enumerate(ws) if w] + [len(cp)]
        return [s[a:b] for a, b in zip(idx[:-1], idx[1:])]
    assert pieces("ab  cd") == ["ab", " ", " cd"]
    assert pieces("ab \n cd") == ["ab", " \n", " cd"]
    assert pieces(" 12ab") == [" 12", "ab"]
    assert pieces("x'll") == ["x", "'ll"]
    assert pieces("\x00\x01'tA") == ["\x00\x01'", "tA"]
    assert pieces("a\x85b\xa0") == ["a", "\x85", "b", "\xa0"]
    assert pieces("  ") == ["  "]
    assert pieces("a\u0416\u0661b \u2003x") == ["a\u0416", "\u0661", "b", " ", "\u2003", "x"]      # beyond Latin-1


def test_against_live_library():
    tokenizers = pytest.importorskip("tokenizers")
    from tokenizers import ByteLevelBPETokenizer, pre_tokenizers
    from tokenizers.trainers import BpeTrainer
    rng = np.random.default_rng(7)
    pt = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=True)
    for trial in range(900):
        n = int(rng.integers(1, 50))
        cp = (rng.integers(0, 256, n) if trial % 3 == 1 else rng.integers(0, 0xD800, n) if trial % 3 == 2
              else rng.choice([32, 39, 115, 116, 114, 101, 108, 65, 48, 10, 160, 33], n))
        starts = np.zeros(n, np.uint8)
        for _, (a, _b) in pt.pre_tokenize_str("".join(map(chr, cp))):
            starts[a] = 1
        assert np.array_equal(pretokenize(cp), starts), cp.tolist()
    for bins, vs in ((rng.integers(0, 256, (600, 140)), 500), (rng.integers(30, 100, (300, 70)), 420),
                     (rng.choice([97, 98, 99], (100, 40)), 300), (rng.integers(0, 1000, (400, 100)), 1500),
                     (rng.integers(1500, 4000, (200, 60)), 3000)):
        mn, mx = int(bins.min()), int(bins.max())
        hf = ByteLevelBPETokenizer()
        trainer = BpeTrainer(vocab_size=vs, min_frequency=2, show_progress=False, special_tokens=[],
                             initial_alphabet=[chr(i) for i in range(mx - mn + 1)], max_token_length=10000)
        hf._tokenizer.train_from_iterator(["".join(map(chr, r - mn)) for r in bins], trainer=trainer)
        o = OracleBPE.train(bins, vs)
        assert o.vocab_dict() == hf.get_vocab()
        model = json.loads(hf._tokenizer.to_str())["model"]
        hf_merges = [m if isinstance(m, str) else " ".join(m) for m in model["merges"]]
        assert o.merges_lines() == hf_merges
        for r in bins[:50]:
            assert o.encode(r - mn) == hf.encode("".join(map(chr, r - mn)), add_special_tokens=False).ids


def test_ragged_and_special_tokens_against_live_library():
    """FIGBPE's remaining arguments (beast/beast_bpe_trainer.py:46, 53, 76-98): sequences of unequal length and
    BpeTrainer special tokens (ids first, duplicates skipped, a single-character special token keeps its id when the
    alphabet holds the same character) — oracle vs the library: vocabulary and merges."""
    pytest.importorskip("tokenizers")
    from tokenizers import ByteLevelBPETokenizer
    from tokenizers.trainers import BpeTrainer
    rng = np.random.default_rng(17)
    cases = [
        ([rng.integers(40, 100, int(rng.integers(1, 60))) for _ in range(300)], 300, []),
        ([rng.integers(0, 256, int(rng.integers(5, 140))) for _ in range(400)], 600, ["<pad>", "<eos>", "<pad>"]),
        ([rng.integers(30, 90, 50) for _ in range(200)], 330, ["<s>", chr(5), "ab"]),
        ([rng.integers(0, 900, int(rng.integers(2, 80))) for _ in range(300)], 1300, ["<unk>"]),
    ]
    for rows, vs, special in cases:
        mn = min(int(r.min()) for r in rows)
        mx = max(int(r.max()) for r in rows)
        hf = ByteLevelBPETokenizer()
        trainer = BpeTrainer(vocab_size=vs, min_frequency=2, show_progress=False, special_tokens=special,
                             initial_alphabet=[chr(i) for i in range(mx - mn + 1)], max_token_length=10000)
        hf._tokenizer.train_from_iterator(["".join(map(chr, r - mn)) for r in rows], trainer=trainer)
        o = OracleBPE.train_ragged(rows, vs, special_tokens=special)
        assert (o.min_token, o.max_token) == (mn, mx)
        assert o.vocab_dict() == hf.get_vocab()
        model = json.loads(hf._tokenizer.to_str())["model"]
        assert o.merges_lines() == [m if isinstance(m, str) else " ".join(m) for m in model["merges"]]
    # equal-length input through the ragged entry point is the plain trainer
    bins = rng.integers(0, 256, (200, 60))
    a, b = OracleBPE.train(bins, 400), OracleBPE.train_ragged(list(bins), 400)
    assert a.merges_txt() == b.merges_txt() and a.vocab_json() == b.vocab_json()


def test_pretoken_starts_are_a_local_rule():
    """The warp-per-sequence encode kernel (csrc/bpe.cu: base_start / token_start) finds token starts
    with a rule that looks at most three codepoints back and one ahead.  Restated here and checked
    against the sequential matcher of the oracle on adversarial strings."""
    from oracle.bpe_oracle import pretokenize, unicode_classes
    cls = unicode_classes(0x3000)
    S = 3

    def base_start(c, i):
        n = len(c)
        if i == 0:
            return True
        k, kp = cls[c[i]], cls[c[i - 1]]
        if k != S:
            if kp == k:
                return False
            if kp == S:
                return c[i - 1] != 32
            return True
        if kp != S:
            return True
        return i + 1 < n and cls[c[i + 1]] != S

    def contraction_len(c, i):
        n = len(c)
        if i + 1 < n:
            d = chr(c[i + 1])
            if d in "stmd":
                return 2
            if i + 2 < n and d + chr(c[i + 2]) in ("re", "ve", "ll"):
                return 3
        return 0

    def token_start(c, i):
        for back in range(1, min(3, i) + 1):
            j = i - back
            if c[j] == 39:
                ln = contraction_len(c, j)
                if ln > 0 and base_start(c, j):
                    if back < ln:
                        return False
                    if back == ln:
                        return True
        return base_start(c, i)

    rng = np.random.default_rng(0)
    alphabet = [39, 39, 39, 32, 32, 9, 10, 160] + [ord(ch) for ch in "stmdrevla12!?"] + [200, 178, 0x3b1, 0x660, 0x2003]
    for _ in range(20000):
        c = [int(v) for v in rng.choice(alphabet, int(rng.integers(1, 14)))]
        want = pretokenize(np.asarray(c, dtype=np.uint16)).astype(bool).tolist()
        assert [token_start(c, i) for i in range(len(c))] == want, c
    for _ in range(300):
        c = [int(v) for v in rng.integers(0, 256, 140)]
        want = pretokenize(np.asarray(c, dtype=np.uint16)).astype(bool).tolist()
        assert [token_start(c, i) for i in range(len(c))] == want
