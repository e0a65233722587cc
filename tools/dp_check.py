"""Data-parallel parity on real GPUs (run under torchrun, one rank per GPU):
the DP run (global batches sliced over the ranks, NCCL all-reduce of dW inside the step launcher) must give
the weights of a single-process run over the same global batches.  Prints DP_CHECK_OK on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uml_b200  # noqa: F401,E402
from uml_b200 import finetune as ft  # noqa: E402
from uml_b200.engine.datasets.utils import BankLoader, FeatureBank  # noqa: E402
from uml_b200.engine.models.head import UML, UMLClip  # noqa: E402
from uml_b200.engine.optimizer.optim import build_optimizer  # noqa: E402
from uml_b200.engine.optimizer.scheduler import build_lr_scheduler  # noqa: E402
from uml_b200.engine.trainer import StepEngine  # noqa: E402


def run(world, rank, dev, prec, B, D, C, steps, banks, adapter_dv=0):
    (xi, yi), (xt, yt) = banks
    ib, tb = FeatureBank(xi, yi, dev), FeatureBank(xt, yt, dev)
    torch.manual_seed(1)
    if adapter_dv:   # UML with the linear adapter img_proj (Dv -> D) and learnable temperatures (preset "linear")
        model = UML(f"synthetic:{adapter_dv}", D, C, learnable_temp=True)
        model.to(dev)
    else:
        model = UMLClip(f"synthetic:{D}", C, logit_scale_init=4.60517)
        model.to(dev)
        model.zero_shot_init(tb)
        model.to(dev)
    opt = build_optimizer(model.parameters(), "adamw", 1e-3, 0.01)
    sch = build_lr_scheduler(opt, "cosine", 5, 100, warmup_type="linear", warmup_lr=1e-5)
    eng = StepEngine(model, opt, dev, -(-B // world), -(-B // world), log_slots=steps + 1, precision=prec, world_size=world)
    il, tl = BankLoader(ib, B, shuffle=True), BankLoader(tb, B, shuffle=True)
    torch.manual_seed(7)  # same seed everywhere -> same global index stream
    ii, ti = iter(il), iter(tl)
    batches, lrs = [], []
    for _ in range(steps):
        a, ii = ft.fetch_next(il, ii)
        b, ti = ft.fetch_next(tl, ti)
        batches.append((a, b))
        lrs.append(sch.get_last_lr()[0])
        sch.step()
    if world > 1:
        eng.run(batches, 0.5, lrs, slot0=0)          # slices each global batch for this rank
    else:
        saved = eng.world
        eng.run(batches, 0.5, lrs, slot0=0)
    torch.cuda.synchronize()
    if world > 1 and eng.dW is not None:
        ds = [torch.empty_like(eng.dW) for _ in range(world)]
        dist.all_gather(ds, eng.dW)
        if rank == 0:
            print(f"    [{prec}] last all-reduced dW identical on all ranks: {all(torch.equal(ds[0], d) for d in ds)}")
    w = model.head.weight.detach().clone()
    if adapter_dv:
        w = torch.cat([w.flatten(), model.img_proj.weight.detach().flatten(), model.img_scale.detach().view(1),
                       model.txt_scale.detach().view(1)])
    return w, eng.read_log(list(range(steps)))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for prec, B, D, C, steps, tol, dv in (("fp32", 30, 512, 100, 12, 2e-5, 0), ("bf16", 2051, 768, 1000, 12, 2e-3, 0),
                                          ("bf16", 2051, 512, 100, 8, 5e-3, 256)):
        g = torch.Generator().manual_seed(3)
        banks = ((torch.randn(5000, dv or D, generator=g) * (0.05 if dv else 1.0), torch.randint(0, C, (5000,), generator=g)),
                 (torch.randn(3001, D, generator=g) * (0.05 if dv else 1.0), torch.arange(3001) % C))
        w_dp, log_dp = run(world, rank, dev, prec, B, D, C, steps, banks, dv)
        w_1, log_1 = run(1, 0, dev, prec, B, D, C, steps, banks, dv)
        err = float((w_dp - w_1).norm() / w_1.norm())
        # every rank must hold the same weights
        ws = [torch.empty_like(w_dp) for _ in range(world)]
        dist.all_gather(ws, w_dp)
        same = all(torch.equal(ws[0], w) for w in ws)
        if not same and rank == 0:
            d = (ws[0] - ws[1]).abs()
            print(f"    rank0 vs rank1: max |dW| {float(d.max()):.3e}, differing elements {int((d > 0).sum())} of {d.numel()}, "
                  f"rows touched {int((d > 0).any(dim=1).sum())}")
        if rank == 0:
            print(f"[{prec}{' adapter' if dv else ''}] B={B} world={world}: |W_dp - W_single| / |W| = {err:.3e} (tol {tol}); ranks identical: {same}; "
                  f"local loss step0 {log_dp[0]} vs global {log_1[0]}")
        ok = ok and err < tol and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_CHECK_OK" if int(flag.item()) else "DP_CHECK_FAILED")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
