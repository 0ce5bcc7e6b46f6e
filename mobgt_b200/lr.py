"""PolynomialDecayLR — same schedule as the reference's graphormer/lr.py:7-34 (linear warm-up to `lr` over
`warmup_updates`, then polynomial decay to `end_lr` at `tot_updates`)."""
from torch.optim.lr_scheduler import LRScheduler


class PolynomialDecayLR(LRScheduler):
    def __init__(self, optimizer, warmup_updates, tot_updates, lr, end_lr, power, last_epoch=-1):
        self.warmup_updates = warmup_updates
        self.tot_updates = tot_updates
        self.lr = lr
        self.end_lr = end_lr
        self.power = power
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        if self._step_count <= self.warmup_updates:                     # lr.py:19-21
            lr = self._step_count / float(self.warmup_updates) * self.lr
        elif self._step_count >= self.tot_updates:                      # lr.py:22-23
            lr = self.end_lr
        else:                                                           # lr.py:24-30
            pct_remaining = 1 - (self._step_count - self.warmup_updates) / (self.tot_updates - self.warmup_updates)
            lr = (self.lr - self.end_lr) * pct_remaining ** self.power + self.end_lr
        return [lr for _ in self.optimizer.param_groups]
