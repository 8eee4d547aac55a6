#!/usr/bin/env python3
"""Extract the reference's known-answer vectors for the verify path into tests/golden/reference_kat.json.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tools/extract_golden.py
Sources (data constants only, no code):
  * src/multisig.rs:544-735   `multisig_transcript_known_answer` byte arrays
  * tests/serde.rs:34-142     base58 strings produced from StdRng::seed_from_u64(2321)
"""
import json
import os
import re
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_kat.json")


def multisig_arrays():
    src = open(os.path.join(REF, "src/multisig.rs")).read()
    body = src[src.index("fn multisig_transcript_known_answer"):]
    out = {}
    for m in re.finditer(r"const (\w+): (\[\[u8; 32\]; 3\]|\[u8; (?:32|64)\]) = (\[.*?\]);", body, re.S):
        name, _ty, lit = m.groups()
        vals = [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", lit)]
        if _ty.startswith("[["):
            out[name] = [bytes(vals[i * 32:(i + 1) * 32]).hex() for i in range(3)]
        else:
            out[name] = bytes(vals).hex()
    out["_inputs"] = {"sk": [3, 5, 7], "r": [11, 13, 17], "s": [19, 23, 29], "m": 31}
    out["_source"] = "src/multisig.rs:544-735"
    return out


def serde_strings():
    src = open(os.path.join(REF, "tests/serde.rs")).read()
    out = {}
    for m in re.finditer(r"fn (serde_\w+)\(\).*?\{(.*?)\n\}", src, re.S):
        name, body = m.groups()
        s = re.search(r'"\\"([1-9A-HJ-NP-Za-km-z]+)\\""', body)
        if s:
            out[name] = s.group(1)
    out["_seed"] = 2321
    out["_source"] = "tests/serde.rs:34-142"
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present; golden file is already committed")
    data = {"multisig_kat": multisig_arrays(), "serde_kat": serde_strings()}
    with open(OUT, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)
    print("wrote", os.path.normpath(OUT), {k: len(v) for k, v in data.items()})


if __name__ == "__main__":
    main()
