"""Segmentation loss and IoU metric on the device -- no per-step ``.cpu()`` / ``.item()``.

``seg_loss`` restates ``SemanticTorchpackTrainer.calc_loss`` (FusionTransformer/modules/SemanticTorchpackTrainer.py:
70-108) for one modality: ``F.cross_entropy(logits, labels.long(), weight=class_weights)`` mixed, when
``lambda_xm > 0``, with the cross-modal ``F.kl_div(log_softmax(logits), softmax(teacher.detach()), 'none').sum(1)
.mean()`` as ``(1 - lambda_xm) * CE + lambda_xm * KL``.  Forward and the gradient with respect to the logits come
out of one libft3d pass (csrc/loss.cu).  ``SegIoU`` mirrors FusionTransformer/models/metric.py:26-82 with the
confusion matrix kept on the GPU (the reference moves logits and labels to the CPU on every step, metric.py:43-44).
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import lib

__all__ = ["seg_loss", "SegIoU"]

_WS = {}


def _workspace(device):
    key = (device.index, ops._stream())
    ws = _WS.get(key)
    if ws is None:
        ws = _WS[key] = torch.empty(int(lib().seg_loss_workspace()), dtype=torch.uint8, device=device)
    return ws


class _SegLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, weight, teacher, lambda_xm, ignore_index):
        if not logits.is_cuda:
            raise RuntimeError("seg_loss: CUDA tensors only (libft3d has no CPU path)")
        logits = ops._chk(logits, torch.float32, "logits")
        labels = ops._chk(labels, torch.int64, "labels")
        n, c = logits.shape
        if labels.shape != (n,):
            raise ValueError("seg_loss: labels must be [n]")
        if weight is not None:
            weight = ops._chk(weight, torch.float32, "weight")
        if teacher is not None:
            teacher = ops._chk(teacher.detach(), torch.float32, "teacher")
            if teacher.shape != logits.shape:
                raise ValueError("seg_loss: teacher logits must have the shape of the logits")
        out = torch.empty(4, dtype=torch.float32, device=logits.device)
        grad = torch.empty_like(logits)
        ws = _workspace(logits.device)
        lib().seg_loss(logits.data_ptr(), labels.data_ptr(), n, c, int(ignore_index), ops._p(weight), ops._p(teacher),
                       float(lambda_xm), ops._valid(n), out.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(),
                       ops._stream())
        ctx.save_for_backward(grad)
        ctx.terms = out
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None


def seg_loss(logits, labels, weight=None, teacher=None, lambda_xm: float = 0.0, ignore_index: int = -100):
    """-> 0-dim loss tensor on the device (differentiable with respect to ``logits`` only, as in the reference, where
    the teacher is detached).  ``labels`` may be any integer dtype (the reference casts with ``.long()``)."""
    if labels.dtype != torch.int64:
        labels = labels.long()
    return _SegLoss.apply(logits, labels, weight, teacher, float(lambda_xm), int(ignore_index))


class SegIoU:
    """metric.py:26-82 with the confusion matrix accumulated by one histogram kernel on the device."""

    def __init__(self, num_classes, ignore_index=0, name="seg_iou"):
        self.num_classes = num_classes
        self.ignore_index = ignore_index
        self.mat = None
        self.name = name

    def update(self, seg_logit: torch.Tensor, seg_label: torch.Tensor):
        n = self.num_classes
        logit = ops._chk(seg_logit.detach(), torch.float32, "seg_logit")
        label = seg_label.detach()
        if label.dtype != torch.int64:
            label = label.long()
        label = ops._chk(label, torch.int64, "seg_label")
        if logit.shape != (label.shape[0], n):
            raise ValueError("SegIoU: logits must be [num_points, num_classes]")
        if self.mat is None:
            self.mat = torch.zeros((n, n), dtype=torch.int64, device=logit.device)
        lib().confusion_update(logit.data_ptr(), label.data_ptr(), logit.shape[0], n, int(self.ignore_index),
                               ops._valid(logit.shape[0]), self.mat.data_ptr(), ops._stream())

    def update_dict(self, preds, labels):
        key = "lidar_seg_logit" if "3d" in self.name else "img_seg_logit" if "2d" in self.name else "lidar_seg_logit"
        self.update(preds[key], labels["seg_label"])

    def reset(self):
        self.mat = None

    @property
    def iou(self):
        h = self.mat.float()
        return torch.diag(h) / (h.sum(1) + h.sum(0) - torch.diag(h))

    @property
    def global_avg(self):
        return self.iou.mean().item()          # the only host read: when somebody asks for the number

    avg = global_avg

    def __str__(self):
        return "{iou:.4f}".format(iou=self.iou.mean().item())

    @property
    def summary_str(self):
        return str(self)
