"""Optimizer factory (reference: engine/optimizer/optim.py:15-71).

``build_optimizer(params, name, lr, weight_decay)`` returns a ``FusedOptimizer`` with the torch
optimizer surface the reference loop uses (``param_groups``, ``zero_grad``, ``step``, ``state_dict``).
The update rules are torch.optim's AdamW (decoupled decay) / Adam (L2) / SGD(momentum=0.9, no
nesterov, L2), betas (0.9, 0.999), eps 1e-8 - executed by the fused CUDA kernels, either stand-alone
from ``p.grad`` (``step()``) or inside the dW GEMM epilogue / split-K reduction (``engine/trainer.py``).
"""
import torch

from ... import ops

AVAI_OPTIMS = ["adam", "sgd", "adamw"]
ADAM_BETAS = (0.9, 0.999)
MOMENTUM = 0.9
SGD_NESTEROV = False


class FusedOptimizer:
    def __init__(self, params, name, lr, weight_decay, betas=ADAM_BETAS, eps=1e-8, momentum=MOMENTUM):
        params = list(params)
        if params and isinstance(params[0], dict):
            groups = [dict(g) for g in params]
            for g in groups:
                g["params"] = list(g["params"])
        else:
            groups = [{"params": params}]
        if not any(g["params"] for g in groups):
            raise ValueError("optimizer got an empty parameter list")
        for g in groups:
            g.setdefault("lr", lr)
            g.setdefault("weight_decay", weight_decay)
            g.setdefault("betas", betas)
            g.setdefault("eps", eps)
            g.setdefault("momentum", momentum)
        self.name, self.param_groups, self.state = name, groups, {}
        self.defaults = dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps, momentum=momentum)

    # ---- state shared with the fused trainer -----------------------------------------------
    def slot(self, p):
        """(m, v) buffers and the step counter of a parameter, created on first use (like torch)."""
        st = self.state.get(p)
        if st is None:
            st = {"step": 0, "m": torch.zeros_like(p.data, memory_format=torch.contiguous_format),
                  "v": None if self.name == "sgd" else torch.zeros_like(p.data)}
            self.state[p] = st
        return st

    def group_of(self, p):
        for g in self.param_groups:
            if any(q is p for q in g["params"]):
                return g
        raise KeyError("parameter not in optimizer")

    def update_struct(self, p):
        """Advance the parameter's step count and describe its update for a fused GEMM epilogue."""
        g, st = self.group_of(p), self.slot(p)
        st["step"] += 1
        return ops.make_update(self.name, g["lr"], st["step"], st["m"], st["v"], weight_decay=g["weight_decay"],
                               betas=g["betas"], eps=g["eps"], momentum=g["momentum"])

    # ---- torch.optim surface ----------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None

    def step(self):
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                self.apply(p, p.grad)

    def apply(self, p, grad, grad2=None, grad2_weight=0.0, shadow=None):
        g, st = self.group_of(p), self.slot(p)
        st["step"] += 1
        data = p.data
        if self.name == "sgd":
            ops.sgd_step(data, grad, st["m"], lr=g["lr"], step=st["step"], momentum=g["momentum"],
                         weight_decay=g["weight_decay"], g2=grad2, g2_weight=grad2_weight, shadow=shadow)
        else:
            ops.adamw_step(data, grad, st["m"], st["v"], lr=g["lr"], step=st["step"], weight_decay=g["weight_decay"],
                           betas=g["betas"], eps=g["eps"], decoupled=(self.name == "adamw"), g2=grad2,
                           g2_weight=grad2_weight, shadow=shadow)

    def state_dict(self):
        return {"name": self.name,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                "state": [{"step": s["step"], "m": s["m"], "v": s["v"]} for s in self.state.values()]}


def build_optimizer(params_groups, name, lr, weight_decay):
    assert name in AVAI_OPTIMS, f"Optimizer {name} not found; available optimizers = {AVAI_OPTIMS}"
    return FusedOptimizer(params_groups, name, lr, weight_decay)
