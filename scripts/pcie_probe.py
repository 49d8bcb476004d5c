#!/usr/bin/env python
"""What the host link gives: pinned H2D alone, D2H alone, and both at once on two streams (GB/s).
The e2e leg of bench.py moves 1.6 GB in and 1.56 GB out per batch; this bounds it."""
import time

import torch


def main():
    n = 200_000_000          # int64 -> 1.6 GB
    h_in = torch.empty(n, dtype=torch.int64).pin_memory()
    h_out = torch.empty(n, dtype=torch.int64).pin_memory()
    d_in = torch.empty(n, dtype=torch.int64, device="cuda")
    d_out = torch.ones(n, dtype=torch.int64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    gb = n * 8 / 1e9

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d(chunks=1):
        with torch.cuda.stream(s1):
            per = n // chunks
            for c in range(chunks):
                d_in[c * per:(c + 1) * per].copy_(h_in[c * per:(c + 1) * per], non_blocking=True)

    def d2h(chunks=1):
        with torch.cuda.stream(s2):
            per = n // chunks
            for c in range(chunks):
                h_out[c * per:(c + 1) * per].copy_(d_out[c * per:(c + 1) * per], non_blocking=True)

    for chunks in (1, 16):
        t = timed(lambda: h2d(chunks))
        print("H2D alone   chunks=%-3d %.1f GB/s" % (chunks, gb / t))
        t = timed(lambda: d2h(chunks))
        print("D2H alone   chunks=%-3d %.1f GB/s" % (chunks, gb / t))
        t = timed(lambda: (h2d(chunks), d2h(chunks)))
        print("both        chunks=%-3d %.1f GB/s per direction, %.1f total" % (chunks, gb / t, 2 * gb / t))


if __name__ == "__main__":
    main()
