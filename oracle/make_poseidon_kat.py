#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/poseidon_pallas_kat.json: the known answers the reference itself holds
for the Poseidon constant generation (circuits/src/poseidon/primitives/fp.rs: ROUND_CONSTANTS, MDS, MDS_INV for the Pallas
base field, t = 3, R_F = 8, R_P = 56; its own test p128pow5t3.rs:116-148 checks the Grain generator against them).
Runs only where /root/reference exists; the fixture travels."""
import hashlib
import json
import os
import re
import sys

SRC = "/root/reference/circuits/src/poseidon/primitives/fp.rs"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse(text, name, rows, cols):
    body = text[text.index("const " + name + ":"):]
    vals = []
    for m in re.finditer(r"from_raw\(\[\s*((?:0x[0-9a-f_]+,\s*){4})\]\)", body):
        limbs = [int(x.replace("_", ""), 16) for x in re.findall(r"0x[0-9a-f_]+", m.group(1))]
        vals.append(sum(l << (64 * i) for i, l in enumerate(limbs)))
        if len(vals) == rows * cols:
            break
    assert len(vals) == rows * cols, (name, len(vals))
    return [vals[r * cols:(r + 1) * cols] for r in range(rows)]


def byte_arrays(text, count):
    """the first `count` 32-byte arrays (little-endian field elements) of `text` as integers"""
    out = []
    for m in re.finditer(r"\[((?:\s*0x[0-9a-f]{2},){32})\s*\]", text):
        out.append(int.from_bytes(bytes(int(x, 16) for x in re.findall(r"0x[0-9a-f]{2}", m.group(1))), "little"))
        if len(out) == count:
            break
    assert len(out) == count
    return out


def main():
    if not os.path.exists(SRC):
        sys.exit("needs " + SRC)
    text = open(SRC).read()
    rc, mds, inv = parse(text, "ROUND_CONSTANTS", 64, 3), parse(text, "MDS", 3, 3), parse(text, "MDS_INV", 3, 3)
    digest = hashlib.sha256(",".join(hex(v) for row in rc for v in row).encode()).hexdigest()
    out = {"source": "circuits/src/poseidon/primitives/fp.rs (Pallas base field, t=3, R_F=8, R_P=56)",
           "modulus": hex(0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001), "num_bits": 255,
           "round_constants_first": [[hex(v) for v in row] for row in rc[:2]],
           "round_constants_last": [hex(v) for v in rc[-1]],
           "round_constants_sha256": digest,
           "mds": [[hex(v) for v in row] for row in mds], "mds_inv": [[hex(v) for v in row] for row in inv]}
    tv = open(os.path.join(os.path.dirname(SRC), "test_vectors.rs")).read()
    fp = tv[tv.index("pub(crate) mod fp"):tv.index("pub(crate) mod fq")]
    perm = byte_arrays(fp[fp.index("fn permute()"):fp.index("fn hash()")], 12)    # two vectors: 3 in + 3 out each
    hsh = byte_arrays(fp[fp.index("fn hash()"):], 6)                              # two vectors: 2 in + 1 out each
    out["permute_vectors"] = [{"initial": [hex(v) for v in perm[i:i + 3]], "final": [hex(v) for v in perm[i + 3:i + 6]]} for i in (0, 6)]
    out["hash_vectors"] = [{"input": [hex(v) for v in hsh[i:i + 2]], "output": hex(hsh[i + 2])} for i in (0, 3)]
    out["vectors_source"] = "circuits/src/poseidon/primitives/test_vectors.rs (mod fp: permute :16, hash :420)"
    path = os.path.join(ROOT, "tests", "golden", "poseidon_pallas_kat.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path)


if __name__ == "__main__":
    main()
