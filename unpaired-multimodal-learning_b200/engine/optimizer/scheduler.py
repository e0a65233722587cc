"""Learning-rate schedules (host-side scalars handed to the update kernels).

Same factory and semantics as the reference's ``engine/optimizer/scheduler.py:84-143``: a "cosine" or
"linear" decay over ``max_iter`` steps, optionally preceded by ``warmup_iter`` steps of "constant" or
"linear" warm-up.  The reference wraps torch's ``CosineAnnealingLR``/``LambdaLR`` in a warm-up
scheduler that only starts stepping the wrapped one after the warm-up, so the decay is evaluated at
``step - warmup_iter``; linear warm-up emits ``warmup_lr`` at step 0 and ``base*step/warmup`` after
(so 1e-5, 2e-5, 4e-5 ... for base 1e-3).  Evaluated in closed form here.
"""
import math

AVAI_SCHEDS = ["cosine", "linear"]
AVAI_WARMUP_SCHEDS = ["constant", "linear"]


class LRSchedule:
    """Drop-in for the torch scheduler objects: ``step()``, ``get_last_lr()``, ``last_epoch``."""

    def __init__(self, optimizer, kind, warmup_iter, max_iter, warmup_type, warmup_lr):
        self.optimizer = optimizer
        self.kind, self.warmup_iter, self.max_iter = kind, int(warmup_iter), float(max_iter)
        self.warmup_type, self.warmup_lr = warmup_type, warmup_lr
        self.base_lrs = [g["lr"] for g in optimizer.param_groups]
        self.last_epoch = 0
        self._apply()

    def lr_at(self, step, base_lr):
        if self.warmup_iter > 0 and step < self.warmup_iter:
            if self.warmup_type == "constant":
                return float(self.warmup_lr)
            return float(self.warmup_lr) if step == 0 else base_lr * step / self.warmup_iter
        t = step - self.warmup_iter if self.warmup_iter > 0 else step
        if self.kind == "cosine":
            return base_lr * (1.0 + math.cos(math.pi * t / self.max_iter)) / 2.0
        return base_lr * (1.0 - t / self.max_iter)

    def _apply(self):
        self._last_lr = [self.lr_at(self.last_epoch, b) for b in self.base_lrs]
        for g, lr in zip(self.optimizer.param_groups, self._last_lr):
            g["lr"] = lr

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else int(epoch)
        self._apply()

    def get_last_lr(self):
        return list(self._last_lr)

    def state_dict(self):
        return {"last_epoch": self.last_epoch, "base_lrs": list(self.base_lrs)}

    def load_state_dict(self, sd):
        self.last_epoch, self.base_lrs = sd["last_epoch"], list(sd["base_lrs"])
        self._apply()


def build_lr_scheduler(optimizer, lr_scheduler, warmup_iter, max_iter, warmup_type=None, warmup_lr=None,
                       verbose=False):
    if verbose:
        print(f"Building scheduler: {lr_scheduler} with warmup: {warmup_type}")
    if lr_scheduler not in AVAI_SCHEDS:
        raise ValueError(f"scheduler must be one of {AVAI_SCHEDS}, but got {lr_scheduler}")
    if warmup_iter > 0 and warmup_type not in AVAI_WARMUP_SCHEDS:
        raise ValueError(f"warmup_type must be one of {AVAI_WARMUP_SCHEDS}, but got {warmup_type}")
    return LRSchedule(optimizer, lr_scheduler, warmup_iter, max_iter, warmup_type, warmup_lr)
