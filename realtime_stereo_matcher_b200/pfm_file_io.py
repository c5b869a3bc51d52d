"""Mirror of the reference's ``tools/pfm_file_io.py`` (:6-77) on the library's host-side PFM entry points
(``rsm_pfm_write`` / ``rsm_pfm_read``): same function names, arguments, return values and exceptions.  Tensors
(CPU or CUDA) are accepted next to numpy arrays; ``write_disparity_pfm`` is the call ``test_stereo.py:133`` makes
(rows stored bottom-up, scale 1) without the flipped copy."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L

__all__ = ["read_pfm_file", "write_pfm_file", "write_disparity_pfm"]


def _as_host_array(image):
    if hasattr(image, "detach"):                      # torch tensor
        image = image.detach().cpu().numpy()
    return np.asarray(image)


def write_pfm_file(pfm_file, image, scale=1, flip=False):
    """tools/pfm_file_io.py:47-77.  ``flip=True`` stores the rows bottom-up (== passing ``np.flipud(image)``)."""
    image = _as_host_array(image)
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if len(image.shape) == 3 and image.shape[2] == 3:
        channels = 3
    elif len(image.shape) == 2 or len(image.shape) == 3 and image.shape[2] == 1:
        channels = 1
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    image = np.ascontiguousarray(image, dtype=np.float32)      # native byte order, C order (what tofile writes)
    rc = L.load().rsm_pfm_write(os.fsencode(pfm_file), image.ctypes.data, image.shape[0], image.shape[1], channels,
                                float(scale), 1 if flip else 0)
    L.check(rc, "rsm_pfm_write")


def write_disparity_pfm(pfm_file, disparity):
    """test_stereo.py:133: ``write_pfm_file(file, np.flipud(flow_pr), 1.0)`` for an (H,W) map."""
    write_pfm_file(pfm_file, disparity, 1.0, flip=True)


def read_pfm_file(pfm_file):
    """tools/pfm_file_io.py:6-44: returns ``(data, scale)``; data is (H,W) or (H,W,3) float32 with the first row on
    top (the file stores rows bottom-up), scale is the absolute value stored in the header."""
    lib = L.load()
    h, w, off = C.c_int64(), C.c_int64(), C.c_int64()
    ch, scale = C.c_int(), C.c_double()
    rc = lib.rsm_pfm_read_header(os.fsencode(pfm_file), C.addressof(h), C.addressof(w), C.addressof(ch), C.addressof(scale),
                                 C.addressof(off))
    if rc:
        msg = lib.rsm_last_error(rc).decode()
        for known in ("Not a PFM file.", "Malformed PFM header."):
            if msg.startswith(known):
                raise Exception(known)
        if "cannot open" in msg:
            raise FileNotFoundError(msg)
        L.check(rc, "rsm_pfm_read_header")
    shape = (h.value, w.value, 3) if ch.value == 3 else (h.value, w.value)
    data = np.empty(shape, dtype=np.float32)
    L.check(lib.rsm_pfm_read(os.fsencode(pfm_file), data.ctypes.data, h.value, w.value, ch.value, 1), "rsm_pfm_read")
    return data, abs(scale.value)
