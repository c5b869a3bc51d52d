"""TEST / BASELINE INFRASTRUCTURE ONLY -- locate and import the UNMODIFIED reference.

The reference is pure Python.  ``tools/install_ref.py`` copies it byte for byte into the git-ignored
``baseline/_ref/`` (which gpurun ships to the GPU box; ``/root/reference`` does not exist there).  The
`-m gpu` full-model parity tests and ``bench.py --impl reference`` import it through this module; the product
package never does.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "MANIFEST.json")) and os.path.isdir(os.path.join(REF_DIR, "model"))


def verify_unmodified() -> int:
    """sha256 of every installed file against the manifest written at install time; returns the file count."""
    man = json.load(open(os.path.join(REF_DIR, "MANIFEST.json")))["files"]
    for rel, want in man.items():
        with open(os.path.join(REF_DIR, rel), "rb") as f:
            got = hashlib.sha256(f.read()).hexdigest()
        if got != want:
            raise RuntimeError(f"baseline/_ref/{rel} differs from the installed reference")
    return len(man)


def load():
    """Put baseline/_ref first on sys.path and import the reference's packages.  Returns a namespace with
    ``model``, ``loss``, ``cost_volume`` modules and ``config(name)`` to read configure/*.json."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python tools/install_ref.py` in the build container "
                           "(needs /root/reference); gpurun ships the directory to the GPU box")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)

    class Ref:
        pass

    ref = Ref()
    ref.dir = REF_DIR
    ref.model = importlib.import_module("model")
    ref.loss = importlib.import_module("loss")
    ref.loss_loss = importlib.import_module("loss.loss")
    ref.cv_concatenate = importlib.import_module("cost_volume.concatenate")
    ref.cv_interweave = importlib.import_module("cost_volume.interweave")
    ref.cv_inner_product = importlib.import_module("cost_volume.inner_product")
    ref.cv_groupwise = importlib.import_module("cost_volume.groupwise")
    ref.v4 = importlib.import_module("model.mobile_stereo_net_v4")
    ref.config = lambda name: json.load(open(os.path.join(REF_DIR, "configure", name)))
    return ref
