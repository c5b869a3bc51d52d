"""PFM reader / writer (tools/pfm_file_io.py:6-77) through the library's host entry points: byte-identical files
and identical read-back against fixtures produced by the reference's own module.  Runs without a GPU."""
import struct

import numpy as np
import pytest
import torch

from golden_io import load


@pytest.fixture(scope="module")
def pfm():
    import realtime_stereo_matcher_b200 as rsm
    rsm.load_library()
    return rsm


def test_files_are_byte_identical_and_read_back(pfm, tmp_path):
    g, m = load("pfm_files")
    for k, case in enumerate(m["cases"]):
        path = tmp_path / f"w{k}.pfm"
        pfm.write_pfm_file(str(path), g[f"img{k}"], case["scale"])
        assert path.read_bytes() == g[f"bytes{k}"].tobytes()
        data, scale = pfm.read_pfm_file(str(path))
        assert data.dtype == np.float32 and data.shape == g[f"read{k}"].shape
        assert np.array_equal(data, g[f"read{k}"]) and scale == case["read_scale"]
    path = tmp_path / "disp.pfm"
    pfm.write_disparity_pfm(str(path), torch.from_numpy(g["disp"]))              # test_stereo.py:133
    assert path.read_bytes() == g["disp_bytes"].tobytes()
    data, scale = pfm.read_pfm_file(str(path))
    assert np.array_equal(data, g["disp"]) and scale == 1.0


def test_big_endian_file(pfm, tmp_path):
    img = np.arange(12, dtype=np.float32).reshape(3, 4) - 5.5
    path = tmp_path / "be.pfm"
    path.write_bytes(b"Pf\n4 3\n1.000000\n" + b"".join(struct.pack(">f", v) for v in img.ravel()))
    data, scale = pfm.read_pfm_file(str(path))
    assert np.array_equal(data, np.flipud(img)) and scale == 1.0


def test_error_behaviour(pfm, tmp_path):
    with pytest.raises(Exception, match="Image dtype must be float32"):
        pfm.write_pfm_file(str(tmp_path / "x.pfm"), np.zeros((2, 2), np.float64))
    with pytest.raises(Exception, match="H x W x 3, H x W x 1 or H x W"):
        pfm.write_pfm_file(str(tmp_path / "x.pfm"), np.zeros((2, 2, 2), np.float32))
    bad = tmp_path / "bad.pfm"
    bad.write_bytes(b"P6\n2 2\n255\n")
    with pytest.raises(Exception, match="Not a PFM file"):
        pfm.read_pfm_file(str(bad))
    bad.write_bytes(b"Pf\nwide high\n-1.0\n")
    with pytest.raises(Exception, match="Malformed PFM header"):
        pfm.read_pfm_file(str(bad))
    bad.write_bytes(b"Pf\n4 4\n-1.0\n" + b"\\0" * 8)
    with pytest.raises(RuntimeError, match="shorter than its header"):
        pfm.read_pfm_file(str(bad))
    with pytest.raises(FileNotFoundError):
        pfm.read_pfm_file(str(tmp_path / "missing.pfm"))
    with pytest.raises(RuntimeError, match="cannot open for writing"):
        pfm.write_pfm_file(str(tmp_path / "no_such_dir" / "x.pfm"), np.zeros((2, 2), np.float32))
