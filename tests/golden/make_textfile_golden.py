#!/usr/bin/env python
"""Golden fixture for the device text ingest, captured FROM THE REAL REFERENCE.

Runs only in the build container (needs /root/reference).  Builds the scratch copy of the reference exactly as
make_golden.py does, writes a seeded ASCII text with every whitespace byte str.split() knows, empty lines, long lines
and no final newline, and records -- for several split sizes -- what the REFERENCE's `textFile(path, splitSize=...)`
hands out: the lines of every split (TextFileRDD.compute, dpark/rdd.py:1672-1711) and the rows the tokenising flatMap of
examples/wc.py (`for w in x.strip().split(): yield (w, 1)`) makes of them.  tests/test_textingest_host.py checks the
product's owned byte ranges and its tokeniser arithmetic (run on the CPU) against this file.

    python tests/golden/make_textfile_golden.py        # rewrites tests/golden/textfile_cases.json
"""
import json
import os
import random
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def make_text(seed=20260923):
    rng = random.Random(seed)
    vocab = ["w%d" % i for i in range(200)] + ["x" * 37, "a", "B", "~!@#", "0", "tab", "\x00nul", "\x7fdel"]
    seps = [" ", "  ", "\t", " \t ", "\x0b", "\x0c", "\x1c", "\x1d", "\x1e", "\x1f", "\r"]
    lines = []
    for i in range(400):
        k = rng.choice([0, 0, 1, 2, 3, 5, 8, 13, 40])
        words = [rng.choice(vocab) for _ in range(k)]
        line = "".join(w + rng.choice(seps) for w in words)
        if rng.random() < 0.15:
            line = "  " + line
        if rng.random() < 0.1:
            line += "\r"                      # CRLF files
        lines.append(line)
    return "\n".join(lines) + "\nlast line without newline"


def fm(x):
    for w in x.strip().split():
        yield (w, 1)


def main():
    scratch = tempfile.mkdtemp(prefix="dpark_ref_")
    try:
        mg.build_reference(scratch)
        mg.bootstrap(scratch)
        import dpark  # the REFERENCE (scratch copy)
        dc = dpark.DparkContext("local")
        text = make_text()
        path = os.path.join(scratch, "in.txt")
        with open(path, "wb") as f:
            f.write(text.encode("ascii"))
        cases = []
        for split_size in (7, 64, 1000, 4096, 1 << 20):
            rdd = dc.textFile(path, splitSize=split_size)
            lines = [list(rdd.iterator(sp)) for sp in rdd.splits]
            toks = [[w for w, one in rdd.flatMap(fm).iterator(sp)] for sp in rdd.splits]
            assert all(one == 1 for sp in rdd.splits for w, one in rdd.flatMap(fm).iterator(sp))
            cases.append({"split_size": split_size, "ranges": [[sp.begin, sp.end] for sp in rdd.splits],
                          "lines": lines, "tokens": toks})
        dc.stop()
        out = {"text": text, "cases": cases,
               "how": "reference textFile(path, splitSize) -> lines per split; flatMap(wc.py's fm) -> tokens per split"}
        with open(os.path.join(HERE, "textfile_cases.json"), "w") as f:
            json.dump(out, f)
        print("wrote textfile_cases.json: %d bytes of text, %d cases, %d tokens"
              % (len(text), len(cases), sum(len(t) for t in cases[0]["tokens"])))
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
