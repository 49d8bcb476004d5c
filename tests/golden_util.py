"""Decode the typed JSON written by tests/golden/make_golden.py."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dec(o):
    if o is None or isinstance(o, int):
        return o
    if "f" in o:
        return float.fromhex(o["f"])
    if "b" in o:
        return bytes.fromhex(o["b"])
    if "s" in o:
        return "".join(chr(c) for c in o["s"])
    if "tu" in o:
        return tuple(dec(x) for x in o["tu"])
    if "l" in o:
        return [dec(x) for x in o["l"]]
    if "np" in o:
        return np.dtype(o["np"]).type(eval(o["v"], {"inf": float("inf"), "nan": float("nan")}))
    if "bool" in o:
        return bool(o["bool"])
    raise ValueError(o)


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def canon(parts):
    """Canonical form of a list of per-partition {k: v} dicts / item lists, the
    same one make_golden.py stores: per partition, items sorted by their JSON."""
    from golden.make_golden import enc  # noqa: the encoder is the contract
    out = []
    for part in parts:
        items = part.items() if isinstance(part, dict) else part
        out.append(sorted(([enc(k), enc(v)] for k, v in items), key=json.dumps))
    return out


def split_rows(rows, sizes):
    out, i = [], 0
    for s in sizes:
        out.append(rows[i:i + s])
        i += s
    return out
