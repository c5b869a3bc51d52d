"""Mirror of the reference's ``loss/loss.py`` on the device kernels: same names, constructor and
call signatures.  ``SequenceLoss`` launches two small kernels per prediction and never synchronises unless
``check_finite`` (the reference's asserts, loss.py:60,66-67) is on, in which case all predictions are tested with ONE
device read at the end; ``get_flow_map_metrics`` reads its seven numbers back with one copy instead of seven
``.item()`` calls (train_stereo.py:174)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as F_rsm

__all__ = ["SequenceLoss", "get_flow_map_metrics", "build_loss_function"]

_METRIC_KEYS = ("epe", "0.5px", "1px", "3px", "5px", "min", "max")


def get_flow_map_metrics(flow_gt, flow_pred, flow_valid):
    """loss/loss.py:6-22: end-point-error statistics over ``flow_valid >= 0.5`` and min / max of the first batch
    item, as a dict of Python floats."""
    values = F_rsm.flow_map_metrics(flow_gt, flow_pred, flow_valid).tolist()      # the only synchronisation
    return {k: float(torch.tensor(v, dtype=torch.float32)) for k, v in zip(_METRIC_KEYS, values)}


class SequenceLoss(nn.Module):
    """loss/loss.py:25-81.  ``check_finite=False`` drops the reference's NaN / Inf asserts and with them every
    host synchronisation of the loss."""

    def __init__(self, loss_gamma=0.9, max_flow_magnitude=700, check_finite=True, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.loss_gamma = loss_gamma
        self.max_flow_magnitude = max_flow_magnitude
        self.check_finite = check_finite

    def forward(self, flow_preds, flow_gt, flow_valid):
        n_preds = len(flow_preds)
        assert n_preds >= 1, f"empty flow predictions ({n_preds})!"
        assert flow_valid.unsqueeze(1).shape == flow_gt.shape, [flow_valid.unsqueeze(1).shape, flow_gt.shape]
        flow_loss = 0.0
        stats = []
        for i in range(n_preds):
            i_weight = self.loss_gamma ** (n_preds - 1 - i)
            mean, st = F_rsm.sequence_loss_term(flow_preds[i], flow_gt, flow_valid, self.max_flow_magnitude,
                                                smooth=(i == n_preds - 1))
            stats.append(st)
            flow_loss += i_weight * mean
        if self.check_finite:
            bad = torch.stack(stats)[:, 2:].sum(dim=0).tolist()                   # one read for every prediction
            assert bad[0] == 0, "non-finite values in the flow predictions"
            assert bad[1] == 0, "infinite ground truth inside the valid mask"
        return flow_loss


def build_loss_function(loss_config):
    """loss/__init__.py:4-10."""
    loss_type = loss_config["type"]
    if loss_type == "SequenceLoss":
        return SequenceLoss(**loss_config["parameters"])
    raise NotImplementedError(f"invalid loss type: {loss_type}!")
