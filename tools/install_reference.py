#!/usr/bin/env python
"""Install the UNMODIFIED reference hot path into baseline/_ref/ (git-ignored, travels with gpurun).

    python tools/install_reference.py [--src /root/reference]

The reference is a script tree without setup.py / pyproject.toml, so `pip install --target` has nothing to
build (DESIGN.md section 2 records the attempt); the importable packages of the path -- models/, decoding/,
misc/, config/ -- are copied byte for byte instead, and MANIFEST.json records the sha256 of every file so that
`verify()` (used by bench.py --impl reference and the tests) can prove the copy is unmodified.  Nothing under
baseline/_ref/ is product source: the product never imports it, only the CPU baseline / reference arm of
bench.py and the tests do.
"""
import argparse
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("models", "decoding", "misc", "config")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def install(src="/root/reference", dst=DST):
    if not os.path.isfile(os.path.join(src, "models", "seq2seq.py")):
        raise SystemExit("reference not found under %s" % src)
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for pkg in PACKAGES:
        out = os.path.join(dst, pkg)
        if os.path.isdir(out):
            shutil.rmtree(out)
        os.makedirs(out)
        for name in sorted(os.listdir(os.path.join(src, pkg))):
            s = os.path.join(src, pkg, name)
            if not os.path.isfile(s) or name.endswith("~") or name.endswith(".pyc"):
                continue
            shutil.copyfile(s, os.path.join(out, name))
            manifest["%s/%s" % (pkg, name)] = _sha(s)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


def verify(dst=DST):
    """True when every file listed in MANIFEST.json is present with the recorded digest."""
    path = os.path.join(dst, "MANIFEST.json")
    if not os.path.isfile(path):
        return False
    files = json.load(open(path))["files"]
    return bool(files) and all(os.path.isfile(os.path.join(dst, k)) and _sha(os.path.join(dst, k)) == v
                               for k, v in files.items())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    m = install(a.src)
    print("installed %d reference files into %s (verify: %s)" % (len(m), DST, verify()))
