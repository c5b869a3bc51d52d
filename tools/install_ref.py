#!/usr/bin/env python
"""Install the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` so that it travels to the GPU box
with the gpurun snapshot (``/root/reference`` does not exist there).

    python tools/install_ref.py [--src /root/reference] [--force]

The reference is pure Python with no build system (no setup.py / pyproject), so "installing" it is copying its
importable packages -- cost_volume/, model/, loss/, configure/, tools/, dataset/ -- byte for byte.  Nothing
under baseline/_ref/ is product source: it is used by the `-m gpu` full-model parity tests (patched vs unpatched
forward on the same device) and by ``bench.py --impl reference`` (the reference's own CPU path).  A MANIFEST with
the sha256 of every file is written so a test can check the copy is unmodified.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("cost_volume", "model", "loss", "configure", "tools", "dataset")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def install(src: str = "/root/reference", force: bool = False, verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(src, "model")):
        raise FileNotFoundError(f"no reference checkout at {src}")
    manifest_path = os.path.join(DEST, "MANIFEST.json")
    if os.path.exists(manifest_path) and not force:
        return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    manifest = {}
    for pkg in PACKAGES:
        s = os.path.join(src, pkg)
        if not os.path.isdir(s):
            continue
        for dirpath, dirnames, filenames in os.walk(s):
            dirnames[:] = [d for d in dirnames if d != "__pycache__"]
            for fn in filenames:
                if fn.endswith((".pyc", ".pyo")):
                    continue
                a = os.path.join(dirpath, fn)
                rel = os.path.relpath(a, src)
                b = os.path.join(DEST, rel)
                os.makedirs(os.path.dirname(b), exist_ok=True)
                shutil.copyfile(a, b)
                os.chmod(b, 0o644)
                manifest[rel] = _sha(b)
    with open(manifest_path, "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"installed {len(manifest)} reference files into {DEST}")
    return DEST


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    install(a.src, a.force)


if __name__ == "__main__":
    sys.exit(main())
