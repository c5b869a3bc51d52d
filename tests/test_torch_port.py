"""CPU suite: the torch-CPU port timed by bench.py's reference arm reproduces the golden vectors
(bit for bit on fp32: it issues the same ATen ops as the reference)."""
import numpy as np
import pytest
import torch

from golden_io import load, names
from oracle import torch_port as tp


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("name", [n for n in names("vol_") if n.endswith("fp32")])
def test_volumes(name):
    g, m = load(name)
    l, r, d, ng = t(g["left"]), t(g["right"]), m["D"], m["G"]
    for op, out in (("concat", tp.concat_volume(l, r, d)), ("interweave", tp.interweave(l, r)),
                    ("inner", tp.inner_product_volume(l, r, d)),
                    ("corr_mean", tp.inner_product_volume(l, r, d, mean=True)),
                    ("groupwise", tp.groupwise_volume(l, r, ng, d)),
                    ("difference", tp.difference_volume(l, r, d))):
        np.testing.assert_array_equal(out.numpy(), g[f"{op}.out"], err_msg=op)


def test_regression_and_tail():
    g, m = load("regress_wide_fp32")
    np.testing.assert_allclose(tp.soft_argmax(t(g["cost"])).numpy(), g["e"], atol=2e-5)  # thread-count dependent sum order
    np.testing.assert_array_equal(tp.hard_argmin(t(g["cost"])).numpy(), g["argmin"])
    g, m = load("tail_v4like")
    np.testing.assert_allclose(tp.v4_tail(t(g["cost"]), m["D"], m["H"], m["W"]).numpy(), g["pred"], atol=1e-4)
