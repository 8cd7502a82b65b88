"""Fused loss + metric epilogue of the eigenvalue head.

The reference computes, per batch (TRAIN_FINAL.py:262-263, 340-341):
    loss = criterion(normalizer.denormalize_eigenvalue(pred), normalizer.denormalize_eigenvalue(batch.y))
    mape += MAPE_error(pred, batch.y, prediction_type, normalizer).item()
i.e. two uploads of the scaler constants (Normalizer.py:209-210) and a host sync per metric.  Here one
kernel (`bg_eigen_loss`) produces the RelativeErrorLoss, the MAPE and d loss / d pred, and keeps the epoch's
running sums on the device; the host reads them once per epoch.
"""
from __future__ import annotations

import torch

from . import capi
from .engine import _stream


class _EigenLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, y, scale, center, eps, accum):
        if not pred.is_cuda:
            raise RuntimeError("buckgnn_b200.loss runs on CUDA tensors only")
        p = pred.detach().reshape(-1).to(torch.float32).contiguous()
        t = y.detach().reshape(-1).to(torch.float32).contiguous()
        if p.numel() != t.numel():
            raise ValueError("pred and target must have the same number of elements")
        out = torch.empty(2, dtype=torch.float32, device=p.device)
        dpred = torch.empty_like(p)
        capi.eigen_loss(p.data_ptr(), t.data_ptr(), p.numel(), scale, center, eps, out.data_ptr(), dpred.data_ptr(),
                        None if accum is None else accum.data_ptr(), _stream())
        ctx.save_for_backward(dpred)
        ctx.shape = pred.shape
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, grad_loss, _grad_out):
        (dpred,) = ctx.saved_tensors
        return (dpred * grad_loss).reshape(ctx.shape), None, None, None, None, None


class EigenvalueRelativeLoss(torch.nn.Module):
    """`RelativeErrorLoss` on de-normalised eigenvalues + MAPE, one kernel, no host sync.

        crit = EigenvalueRelativeLoss(scale=normalizer.eigenvalue_scaler.scale_[0], center=...center_[0])
        loss = crit(pred, batch.y); loss.backward()
        ...
        mean_loss, mean_mape = crit.epoch_means()      # one read-back per epoch
    """

    def __init__(self, scale: float = 1.0, center: float = 0.0, epsilon: float = 1e-8):
        super().__init__()
        self.scale, self.center, self.epsilon = float(scale), float(center), float(epsilon)
        self._accum = None
        self.last_mape = None

    def forward(self, pred, target):
        if self._accum is None or self._accum.device != pred.device:
            self._accum = torch.zeros(3, dtype=torch.float32, device=pred.device)
        loss, out = _EigenLossFn.apply(pred, target, self.scale, self.center, self.epsilon, self._accum)
        self.last_mape = out[1]
        return loss

    def epoch_means(self, reset: bool = True):
        if self._accum is None:
            return float("nan"), float("nan")
        s = self._accum.tolist()
        if reset:
            self._accum.zero_()
        n = max(s[2], 1.0)
        return s[0] / n, s[1] / n
