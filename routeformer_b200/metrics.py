"""ADE / FDE and the future-discounted loss on the CUDA library.

Drop-ins for routeformer/score/error.py:10-51 and routeformer/losses/future_discounted_mse.py:9-95.
Note the reference's `fde` is NOT a final-step error: it is the Frobenius norm of the whole [T,2] error of the LAST
batch element (error.py:51 indexes dim 0).  That definition is kept, since the 4-decimal parity contract is on it.
"""
from __future__ import annotations

from typing import Dict, Union

import torch
import torch.nn as nn

from . import functional as Fn
from . import ops


def _as3d(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).reshape(-1, t.shape[-2], 2) if t.dim() != 3 else t.to(torch.float32)


def ade(predicted_trajectory: torch.Tensor, ground_truth_trajectory: torch.Tensor) -> torch.Tensor:
    assert predicted_trajectory.shape == ground_truth_trajectory.shape, "Predicted and ground truth trajectories must be of the same shape"
    res, _ = ops.ade_fde(_as3d(predicted_trajectory), _as3d(ground_truth_trajectory))
    return res[0]


def fde(predicted_trajectory: torch.Tensor, ground_truth_trajectory: torch.Tensor) -> torch.Tensor:
    assert predicted_trajectory.shape == ground_truth_trajectory.shape, "Predicted and ground truth trajectories must be of the same shape"
    res, _ = ops.ade_fde(_as3d(predicted_trajectory), _as3d(ground_truth_trajectory))
    return res[1]


def ade_fde_per_sample(pred: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
    """[B,2] = per-clip (ADE, FDE) exactly as the per-sample loop of experiments/full_comparison.py:667-674 computes them,
    in one launch."""
    _, ps = ops.ade_fde(_as3d(pred), _as3d(truth), per_sample=True)
    return ps


class FutureDiscountedLoss(nn.Module):
    current_epoch = 0

    def __init__(self, discount_factor: Union[float, Dict[int, float]] = 0.9, epsilon: float = None, loss_function: str = "mse"):
        super().__init__()
        self.current_discount_factor = discount_factor if isinstance(discount_factor, float) else discount_factor[0]
        self.discount_factor_dict = discount_factor if isinstance(discount_factor, dict) else {}
        self.epsilon = epsilon
        self.loss_function = loss_function
        if loss_function not in ["mae", "mse", "smooth_l1"]:
            raise ValueError(f"Unknown loss function {loss_function}")

    def forward(self, y_pred: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
        epoch = self.current_epoch
        if epoch in self.discount_factor_dict:
            self.current_discount_factor = self.discount_factor_dict[epoch]
        assert y_pred.dim() >= 2
        eps = 0.0 if self.epsilon is None else float(self.epsilon)
        if self.loss_function != "smooth_l1" and self.epsilon is None:
            raise TypeError("epsilon must be set for the mse / mae losses")  # the reference fails the same way (abs(error) < None)
        return Fn.DiscountedLoss.apply(y_pred.to(torch.float32), y_true.to(torch.float32), float(self.current_discount_factor), eps,
                                       self.loss_function)
