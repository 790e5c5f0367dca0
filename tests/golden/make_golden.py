"""Generates the committed golden vectors from the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile). Run in the build container only:  python tests/golden/make_golden.py

Small cases are stored whole (<name>.in / <name>.crs2); larger ones as (recipe, sizes, sha256) in golden.json,
their inputs being regenerated from the recipe by tests/golden_cases.py."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from golden_cases import CASES, SMALL, make_input  # noqa: E402
from oracle_lib import Reference  # noqa: E402


def main():
    ref = Reference()
    meta = {}
    for name in CASES:
        data = make_input(name)
        img = ref.compress(data)
        for kind in ("simple", "fast", "table"):
            if len(img) - 1040 >= 16 or kind == "simple":  # Fast/Table read 8 bytes past tiny payloads
                assert ref.decompress(img, kind) == data, (name, kind)
        meta[name] = {"n": len(data), "crs2_bytes": len(img), "sha256_in": hashlib.sha256(data).hexdigest(),
                      "sha256_crs2": hashlib.sha256(img).hexdigest()}
        if name in SMALL:
            with open(os.path.join(HERE, name + ".in"), "wb") as f:
                f.write(data)
            with open(os.path.join(HERE, name + ".crs2"), "wb") as f:
                f.write(img)
        print(name, meta[name])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
