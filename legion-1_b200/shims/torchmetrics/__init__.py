"""Accuracy('multiclass', num_classes=..) / Accuracy(): running top-1 accuracy with torchmetrics' call protocol."""
import torch


class Accuracy:
    def __init__(self, task=None, num_classes=None, **kw):
        self.correct = torch.zeros((), dtype=torch.long)
        self.total = torch.zeros((), dtype=torch.long)

    def to(self, device):
        self.correct, self.total = self.correct.to(device), self.total.to(device)
        return self

    def __call__(self, preds, target):
        if preds.dim() > 1:
            preds = preds.argmax(dim=1)
        ok = (preds == target).sum()
        self.correct = self.correct + ok
        self.total = self.total + target.numel()
        return ok.float() / max(1, target.numel())

    update = __call__

    def compute(self):
        return self.correct.float() / self.total.clamp(min=1).float()

    def reset(self):
        self.correct.zero_()
        self.total.zero_()
