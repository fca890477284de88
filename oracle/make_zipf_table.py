#!/usr/bin/env python3
"""Regenerate the Zipf inverse-CDF table used by the RF-1 synthetic corpora (oracle/SPEC.md).

TEST INFRASTRUCTURE / data-generation script: not imported by the product.

Entry r (0 <= r < 65536) is the smallest vocab id v in [0, 50000) whose cumulative Zipf(s=1.07)
probability exceeds (r + 0.5) / 65536.  All arithmetic is 50-digit `decimal`, so the output does
not depend on the platform's libm.  The committed copy is
rag_foundation_b200/data/zipf_vocab_u16.bin; `python oracle/make_zipf_table.py --check` compares.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import struct
import sys
from decimal import Decimal, getcontext

VOCAB = 50_000
ENTRIES = 65_536
S = Decimal("1.07")
HERE = os.path.dirname(os.path.abspath(__file__))
TABLE_PATH = os.path.join(HERE, "..", "rag_foundation_b200", "data", "zipf_vocab_u16.bin")


def build() -> bytes:
    getcontext().prec = 50
    weights = [Decimal(v + 1) ** (-S) for v in range(VOCAB)]
    total = sum(weights)
    out = []
    v = 0
    cum = weights[0]
    for r in range(ENTRIES):
        target = (Decimal(r) + Decimal("0.5")) / Decimal(ENTRIES) * total
        while cum <= target and v < VOCAB - 1:
            v += 1
            cum += weights[v]
        out.append(v)
    return struct.pack("<%dH" % ENTRIES, *out)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    blob = build()
    digest = hashlib.sha256(blob).hexdigest()
    if args.check:
        with open(TABLE_PATH, "rb") as f:
            have = f.read()
        ok = have == blob
        print("zipf table", "matches" if ok else "DIFFERS", digest)
        return 0 if ok else 1
    os.makedirs(os.path.dirname(TABLE_PATH), exist_ok=True)
    with open(TABLE_PATH, "wb") as f:
        f.write(blob)
    print("wrote", os.path.normpath(TABLE_PATH), len(blob), "bytes sha256", digest)
    return 0


if __name__ == "__main__":
    sys.exit(main())
