"""Drop-in for the reference's ``cost_volume`` package: same module and class names, same
constructor / forward signatures, output layouts and error behaviour; the work runs in
librsm_b200.so (sm_100a) instead of D slice-assign loops of ATen ops."""
from .concatenate import TorchConcatenateCost
from .groupwise import TorchGroupwiseCost
from .inner_product import TorchInnerProductCost
from .interweave import TorchInterweaveCost

__all__ = ["TorchConcatenateCost", "TorchGroupwiseCost", "TorchInnerProductCost", "TorchInterweaveCost"]
