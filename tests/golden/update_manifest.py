"""Rewrites tests/golden/MANIFEST.json (name -> SHA-256, size, generator script) after a generator has been re-run:
   python tests/golden/make_golden.py && python tests/golden/make_reference_fixtures.py && python tests/golden/update_manifest.py"""
import hashlib
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_MADE = {"default_m1_k2_p32.npz", "tiny_m1_k2.npz", "tiny_m2_k3.npz"}   # make_golden.py; ref_*.npz: the reference's own code


def main():
    out = {}
    for name in sorted(os.listdir(HERE)):
        if not name.endswith(".npz"):
            continue
        with open(os.path.join(HERE, name), "rb") as f:
            blob = f.read()
        out[name] = {"sha256": hashlib.sha256(blob).hexdigest(), "bytes": len(blob),
                     "generator": "make_golden.py" if name in ORACLE_MADE else "make_reference_fixtures.py"}
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
