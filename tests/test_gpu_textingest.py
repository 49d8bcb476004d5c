"""Device text ingest: dpk_tokenize_* against Python's str.split(), and the word-count pipeline
through DparkContext with the device tokeniser against the same pipeline run row-wise (the path that is itself checked
against the reference's output files in test_gpu_rdd.py).  -m gpu."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def nv():
    from dpark_b200 import _native
    return _native


def _tokens(data: bytes):
    d = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda() if data else torch.zeros(0, dtype=torch.uint8, device="cuda")
    starts, lens, ok = nv().tokenize(d)
    if not ok:
        return None
    s, l = starts.cpu().tolist(), lens.cpu().tolist()
    return [data[a:a + b] for a, b in zip(s, l)]


WS = b" \t\n\r\x0b\x0c\x1c\x1d\x1e\x1f"


@pytest.mark.parametrize("case", [b"", b" ", b"\n\n\n", b"a", b"a b", b" a  b ", b"word", b"x" * 5000, b"a\x1cb\x1dc\x1ed\x1fe",
                                  b"tab\tsep\rcr\x0bvt\x0cff end", b"a" * 15 + b" " + b"b" * 16 + b"\n" + b"c" * 17,
                                  b"\x00nul\x00 is\x7fnot space"])
def test_tokenize_edge_cases(case):
    assert _tokens(case) == [w.encode("ascii") for w in case.decode("ascii").split()]


@pytest.mark.parametrize("seed,n", [(1, 4095), (2, 4096), (3, 4097), (4, 70001), (5, 1 << 20)])
def test_tokenize_random_ascii_matches_str_split(seed, n):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(bytes(range(0x21, 0x7f)) * 3 + WS * 4 + b"\x00\x01\x7f", dtype=np.uint8)
    data = alphabet[rng.integers(0, len(alphabet), n)].tobytes()
    got = _tokens(data)
    assert got == [w.encode("ascii") for w in data.decode("ascii").split()]


def test_tokenize_reports_high_bytes():
    assert _tokens("café au lait".encode("utf-8")) is None
    assert _tokens(b"plain" + b" " * 9000 + b"\xa0") is None


def test_gather_bytes_selected_rows():
    data = b"zero one  two three\nfour"
    d = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
    starts, lens, ok = nv().tokenize(d)
    assert ok
    idx = torch.tensor([4, 0, 2, 2], dtype=torch.int64, device="cuda")
    out, off = nv().gather_bytes(d, starts, lens, idx)
    raw, o = out.cpu().numpy().tobytes(), off.cpu().tolist()
    assert [raw[o[i]:o[i + 1]] for i in range(4)] == [b"four", b"zero", b"two", b"two"]
    out, off = nv().gather_bytes(d, starts, lens)
    assert out.cpu().numpy().tobytes() == b"zeroonetwothreefour" and off.cpu().tolist() == [0, 4, 7, 10, 15, 19]


def ctx():
    sys.argv = [sys.argv[0]]
    from dpark_b200 import DparkContext
    return DparkContext("local")


def fm(x):
    for w in x.strip().split():
        yield (w, 1)


def _wc(dc, path, out, split_size):
    from dpark_b200.rdd import TextFileRDD
    (dc.textFile(path, splitSize=split_size) if split_size else dc.textFile(path)) \
        .flatMap(fm).reduceByKey(lambda x, y: x + y, numSplits=6) \
        .map(lambda x: " ".join(list(map(str, x)))).saveAsTextFile(out, overwrite=False)
    assert isinstance(dc.textFile(path), TextFileRDD)
    return {fn: sorted(open(os.path.join(out, fn)).read().splitlines()) for fn in sorted(os.listdir(out))}


@pytest.mark.parametrize("split_size", [0, 10_000])
def test_wc_with_device_tokeniser_equals_the_row_wise_pipeline(tmp_path, split_size, monkeypatch):
    """The same files, byte for byte per partition, whether the tokens are made by the user's Python generator or by
    dpk_tokenize on the device -- and the device path really ran."""
    from dpark_b200 import engine, textingest
    rng = np.random.default_rng(11)
    vocab = ["w%d" % i for i in range(3000)] + ["x" * 40, "a", "B", "~!@", "0"]
    seps = [" ", "  ", "\t", " \t ", "\x0c", "\x1c"]
    lines = []
    for _ in range(20000):
        k = int(rng.integers(0, 12))
        words = [vocab[int(i)] for i in (rng.zipf(1.3, k) - 1) % len(vocab)]
        line = "".join(w + seps[int(rng.integers(0, len(seps)))] for w in words)
        lines.append(("  " + line) if rng.random() < 0.1 else line)
    inp = tmp_path / "in.txt"
    inp.write_text("\n".join(lines) + ("\n" if split_size else ""), encoding="ascii")
    calls = []
    real = textingest.reduce_tokens

    def spy(*a, **kw):
        r = real(*a, **kw)
        calls.append(r is not None)
        return r

    monkeypatch.setattr(textingest, "reduce_tokens", spy)
    got = _wc(ctx(), str(inp), str(tmp_path / "dev"), split_size)
    assert calls == [True]
    monkeypatch.setattr(engine, "TEXT_INGEST", False)
    want = _wc(ctx(), str(inp), str(tmp_path / "rows"), split_size)
    assert calls == [True]
    assert got == want and sum(len(v) for v in got.values()) == len(set(" ".join(lines).split()))


def test_wc_non_ascii_text_takes_the_row_wise_path(tmp_path, monkeypatch):
    from dpark_b200 import textingest
    inp = tmp_path / "in.txt"
    inp.write_text("café thé noir\ncafé au lait\n", encoding="utf-8")
    calls = []
    real = textingest.reduce_tokens
    monkeypatch.setattr(textingest, "reduce_tokens", lambda *a, **kw: calls.append(real(*a, **kw)) or calls[-1])
    got = ctx().textFile(str(inp)).flatMap(fm).reduceByKey(lambda x, y: x + y, numSplits=3).collectAsMap()
    assert calls == [None]
    assert got == {"café": 2, "thé": 1, "noir": 1, "au": 1, "lait": 1}
