#!/usr/bin/env python
"""Binary DRAT (what the engine's proof stream and the reference's default -proof mode write) -> text DRAT, the
conversion cuPROOF::writeClause does for -proofnonbinary (src/gpu/proof.cu:123-157).

    python tools/drat_text.py proof.drat > proof.txt        # or: ... | drat-trim formula.cnf /dev/stdin
"""
import sys


def to_text(raw: bytes) -> str:
    out = []
    i, n = 0, len(raw)
    while i < n:
        kind = raw[i]
        if kind not in (0x61, 0x64):
            raise ValueError(f"byte {i}: expected 'a' or 'd', found {kind:#x}")
        i += 1
        lits = []
        while True:
            if i >= n:
                raise ValueError("truncated line")
            if raw[i] == 0:
                i += 1
                break
            v, shift = 0, 0
            while True:
                b = raw[i]
                i += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            lits.append(-(v >> 1) if v & 1 else v >> 1)       # literal = 2 var + sign (constants.hpp:72-80)
        out.append(("d " if kind == 0x64 else "") + " ".join(map(str, lits)) + " 0")
    return "\n".join(out) + ("\n" if out else "")


if __name__ == "__main__":
    data = open(sys.argv[1], "rb").read() if len(sys.argv) > 1 else sys.stdin.buffer.read()
    sys.stdout.write(to_text(data))
