"""`torchmetrics.Accuracy(task="multiclass", top_k=1)` as /eval.py of the reference uses it (eval.py:31, 47-55): running
top-1 accuracy, synchronised over the default process group at compute()."""
import torch


class Accuracy:
    def __init__(self, task="multiclass", num_classes=None, top_k=1, dist_sync_on_step=False, process_group=None, **kw):
        if task != "multiclass" or top_k != 1:
            raise NotImplementedError("compat torchmetrics.Accuracy: multiclass top-1 only")
        self.correct = torch.zeros((), dtype=torch.float64)
        self.total = torch.zeros((), dtype=torch.float64)

    def to(self, device):
        self.correct, self.total = self.correct.to(device), self.total.to(device)
        return self

    def forward(self, preds, target):
        if preds.ndim == target.ndim + 1:
            preds = preds.argmax(dim=-1)
        if target.ndim == 2:
            target = target.argmax(dim=-1)
        hit = (preds == target).sum().to(self.correct.dtype)
        self.correct += hit
        self.total += target.numel()
        return hit / max(1, target.numel())

    __call__ = forward
    update = forward

    def compute(self):
        c, t = self.correct.clone(), self.total.clone()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(c)
            torch.distributed.all_reduce(t)
        return (c / t.clamp(min=1)).float()

    def reset(self):
        self.correct.zero_()
        self.total.zero_()
