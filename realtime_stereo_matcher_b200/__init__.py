"""realtime_stereo_matcher_b200 -- B200 (sm_100a) cost-volume construction and disparity
regression for babiking/realtime_stereo_matcher, behind the reference's own call signatures.

Host code is PyTorch (device memory, streams, autograd, torch.distributed); every op is one call
into the hand-written CUDA library librsm_b200.so through the C ABI of include/rsm.h.  There is
no CPU fallback: the ops raise on non-CUDA tensors or when the library has not been built.
"""
from . import cost_volume, functional, model_functions
from ._lib import LIB_PATH, load as load_library
from .cost_volume import (TorchConcatenateCost, TorchGroupwiseCost, TorchInnerProductCost,
                          TorchInterweaveCost)
from .functional import (concat_volume, difference_volume, expectation, finalize_disparity, flow_map_metrics,
                         groupwise_pointwise,
                         groupwise_volume, hard_argmax, hard_argmin, inner_product_regress,
                         inner_product_volume, interweave, prepare_input, regress, sequence_loss_term,
                         shift_interweave_volume,
                         soft_argmax, upsample_regress, v4_cost_volume, warp_by_flow_map)
from .loss import SequenceLoss, build_loss_function, get_flow_map_metrics
from .model_functions import (disparity_regression_dispnetc, disparity_regression_v4, interweave_tensors,
                              make_correlation_volume, make_cost_volume, softmax_regression, v4_head)
from .patch import patch_reference, unpatch_reference
from .pfm_file_io import read_pfm_file, write_disparity_pfm, write_pfm_file
from .pipeline import HostPipeline
from .sharding import all_gather_metrics, bind_host_to_device, shard_range

__version__ = "0.1.0"
