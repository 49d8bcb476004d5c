"""Host logic of the device text ingest (dpark_b200/textingest.py): which user functions are recognised as the
word-count tokeniser (structurally, never by probing), and which bytes a split owns.  No GPU needed."""
import os
import sys

import pytest

from dpark_b200 import textingest as ti


def fm(x):
    for w in x.strip().split():
        yield (w, 1)


def fm_other_names(line):
    for word in line.split():
        yield (word, 1)


def fm_sep(x):
    for w in x.strip().split(","):
        yield (w, 1)


def fm_two(x):
    for w in x.strip().split():
        yield (w, 2)


def fm_lower(x):
    for w in x.strip().split():
        yield (w.lower(), 1)


def fm_default(x, n=1):
    for w in x.strip().split():
        yield (w, 1)


ONE = 1


def fm_global(x):
    for w in x.strip().split():
        yield (w, ONE)


def ctx():
    sys.argv = [sys.argv[0]]
    from dpark_b200 import DparkContext
    return DparkContext("local")


def test_only_exact_tokenisers_are_recognised(tmp_path):
    p = tmp_path / "t.txt"
    p.write_text("a b\n")
    dc = ctx()
    tf = dc.textFile(str(p))
    yes = [tf.flatMap(fm), tf.flatMap(fm_other_names), tf.flatMap(lambda x: [(w, 1) for w in x.split()]),
           tf.flatMap(lambda l: l.split()).map(lambda w: (w, 1)),
           tf.flatMap(lambda l: l.strip().split()).map(lambda w: (w, 1))]
    for r in yes:
        assert ti.recognize(r) is tf
    k = 1
    no = [tf.flatMap(fm_sep), tf.flatMap(fm_two), tf.flatMap(fm_lower), tf.flatMap(fm_default), tf.flatMap(fm_global),
          tf.flatMap(lambda l: l.split()).map(lambda w: (w, 1.0)),
          tf.flatMap(lambda l: l.split()).map(lambda w: (w, k)),                 # closure
          tf.flatMap(lambda l: l.split(None, 1)).map(lambda w: (w, 1)),
          tf.flatMap(lambda l: l.split()).filter(lambda w: w).map(lambda w: (w, 1)),
          tf.map(lambda l: l.upper()).flatMap(fm),                                # something between file and tokeniser
          dc.makeRDD(["a b"], 1).flatMap(fm)]
    for r in no:
        assert ti.recognize(r) is None


def test_a_subclassed_text_rdd_is_left_alone(tmp_path):
    from dpark_b200.rdd import TextFileRDD

    class Mine(TextFileRDD):
        pass

    p = tmp_path / "t.txt"
    p.write_text("a b\n")
    dc = ctx()
    assert ti.recognize(Mine(dc, str(p)).flatMap(fm)) is None


@pytest.mark.parametrize("split_size", [1, 3, 7, 16, 1000])
def test_owned_ranges_are_the_lines_a_split_yields(tmp_path, split_size):
    """owned_range == the bytes of the lines TextFileRDD.compute yields for the split (a line belongs to the split it
    starts in), for splits smaller than a line, empty lines, a missing final newline."""
    from dpark_b200.rdd import TextFileRDD
    body = b"alpha beta\n\n\ngamma\nd\n" + b"x" * 40 + b"\nlast line without newline"
    p = tmp_path / "t.txt"
    p.write_bytes(body)
    dc = ctx()
    tf = TextFileRDD(dc, str(p), splitSize=split_size)
    size = os.path.getsize(str(p))
    prev_end = 0
    for sp in tf.splits:
        a, b = ti.owned_range(str(p), sp.begin, sp.end, size)
        assert a == prev_end or a >= prev_end
        lines = list(tf.compute(sp))
        chunk = body[a:b]
        want = chunk.decode().split("\n")
        if chunk.endswith(b"\n"):
            want = want[:-1]
        assert (lines == want) if chunk else (lines == [])
        prev_end = b
    assert prev_end == size


def _hostcheck():
    import ctypes as C
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "tests", "_hostcheck.so")
    if not os.path.exists(path):
        pytest.skip("hostcheck not built")
    L = C.CDLL(path)
    L.hc_tokenize.restype = C.c_int64
    L.hc_tokenize.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    return L


def _hc_tokens(L, data):
    import numpy as np
    buf = np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(1, np.uint8)
    # the kernel's vector path needs the 16-byte alignment a device buffer has: give the host copy the same
    raw = np.zeros(len(buf) + 64, np.uint8)
    off = (-raw.ctypes.data) % 16
    view = raw[off:off + len(buf)]
    view[:] = buf
    starts = np.zeros(len(data) + 1, np.int64)
    lens = np.zeros(len(data) + 1, np.int64)
    m = L.hc_tokenize(view.ctypes.data, len(data), starts.ctypes.data, lens.ctypes.data)
    if m < 0:
        return None
    return [data[a:a + b] for a, b in zip(starts[:m].tolist(), lens[:m].tolist())]


def test_tokeniser_arithmetic_equals_str_split_on_the_cpu():
    """The product's __host__ __device__ tokeniser arithmetic (dpk_common.cuh tok_ws / tok_starts16, the per-thread
    step of k_tok_count / k_tok_emit) run on the CPU through tests/hostcheck.cu: every ASCII byte is whitespace exactly
    when Python's str.split() treats it as such, and the tokens of random ASCII text are str.split()'s."""
    import numpy as np
    L = _hostcheck()
    for b in range(128):
        s = b"a" + bytes([b]) + b"b"
        assert _hc_tokens(L, s) == [w.encode("ascii") for w in s.decode("ascii").split()], b
    for case in (b"", b" ", b"a", b" a  b ", b"x" * 100, b"a" * 15 + b" " + b"b" * 16 + b"\n" + b"c" * 17):
        assert _hc_tokens(L, case) == [w.encode("ascii") for w in case.decode("ascii").split()]
    rng = np.random.default_rng(3)
    alphabet = np.frombuffer(bytes(range(0x21, 0x7f)) * 3 + b" \t\n\r\x0b\x0c\x1c\x1d\x1e\x1f" * 4 + b"\x00\x01\x7f", dtype=np.uint8)
    for n in (17, 4095, 4096, 4097, 100003):
        data = alphabet[rng.integers(0, len(alphabet), n)].tobytes()
        assert _hc_tokens(L, data) == [w.encode("ascii") for w in data.decode("ascii").split()]
    assert _hc_tokens(L, "café au lait".encode("utf-8")) is None


def test_owned_ranges_and_tokens_equal_what_the_reference_hands_out(tmp_path):
    """tests/golden/textfile_cases.json was written by the REAL reference (make_textfile_golden.py): for five split sizes,
    the lines every split of `textFile(path, splitSize)` yields and the tokens wc.py's flatMap makes of them.  The
    product must (a) cut the same splits, (b) own exactly the bytes of those lines, and (c) tokenise that byte range --
    the device kernels' arithmetic, run here on the CPU -- into the same tokens in the same order."""
    import json
    from dpark_b200.rdd import TextFileRDD
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "textfile_cases.json")) as f:
        gold = json.load(f)
    body = gold["text"].encode("ascii")
    p = tmp_path / "in.txt"
    p.write_bytes(body)
    L = _hostcheck()
    dc = ctx()
    for case in gold["cases"]:
        tf = TextFileRDD(dc, str(p), splitSize=case["split_size"])
        assert [[sp.begin, sp.end] for sp in tf.splits] == case["ranges"]
        prev = 0
        for sp, want_lines, want_tokens in zip(tf.splits, case["lines"], case["tokens"]):
            a, b = ti.owned_range(str(p), sp.begin, sp.end, len(body))
            chunk = body[a:b]
            got_lines = chunk.decode("ascii").split("\n")
            if chunk.endswith(b"\n"):
                got_lines = got_lines[:-1]
            assert (got_lines if chunk else []) == want_lines
            assert list(tf.compute(sp)) == want_lines
            assert _hc_tokens(L, chunk) == [w.encode("ascii") for w in want_tokens]
            assert a >= prev
            prev = b
        assert prev == len(body)


def test_long_ranges_are_cut_at_line_starts(tmp_path):
    """cut_pieces: a byte range longer than the piece limit is tokenised in pieces; every cut is a line start, the pieces
    tile the range, and a line longer than the limit stays whole."""
    body = b"".join(b"line %d has some words\n" % i for i in range(200)) + b"x" * 500 + b"\n" + b"tail"
    p = tmp_path / "t.txt"
    p.write_bytes(body)
    size = len(body)
    for limit in (1, 10, 64, 300, 10 ** 6):
        pieces = ti.cut_pieces(str(p), 0, size, size, limit)
        assert pieces[0][0] == 0 and pieces[-1][1] == size
        for (a0, b0), (a1, b1) in zip(pieces[:-1], pieces[1:]):
            assert b0 == a1 and body[a1 - 1:a1] == b"\n"
        assert all(b0 > a0 for a0, b0 in pieces)
        tokens = [w for a0, b0 in pieces for w in body[a0:b0].split()]
        assert tokens == body.split()
    assert ti.cut_pieces(str(p), 0, size, size, 10 ** 6) == [[0, size]]
