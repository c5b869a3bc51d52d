import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files if k != "meta"}
    meta = json.loads(str(z["meta"]))
    return d, meta


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def round_to(a, dname):
    """Round an fp32 array to fp16/bf16 (round-to-nearest-even) and return it as fp32."""
    a = np.asarray(a, dtype=np.float32)
    if dname == "fp32":
        return a
    if dname == "fp16":
        return a.astype(np.float16).astype(np.float32)
    assert dname == "bf16"
    u = a.view(np.uint32).astype(np.uint64)
    nan = np.isnan(a)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    out = u.astype(np.uint32).view(np.float32).copy()
    out[nan] = np.nan
    return out
