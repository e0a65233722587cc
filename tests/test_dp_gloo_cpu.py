"""Data-parallel bookkeeping on CPU with a real world_size-2 process group (gloo): every rank derives
the same global index stream, takes its slice, weights its local mean-gradient with
``dp_loss_weights`` and the all-reduced sum equals the single-process gradient of the global batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import uml_b200  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import uml_oracle as O
        from uml_b200 import finetune as ft
        from uml_b200.engine.datasets.utils import BankLoader, FeatureBank
        from uml_b200.engine.trainer import dp_loss_weights

        g = torch.Generator().manual_seed(0)
        C, D = 7, 12
        xi, yi = torch.randn(50, D, generator=g), torch.randint(0, C, (50,), generator=g)
        xt, yt = torch.randn(33, D, generator=g), torch.randint(0, C, (33,), generator=g)
        W = torch.randn(C, D, generator=g)
        torch.manual_seed(11)  # same seed on every rank -> same global permutation
        il = BankLoader(FeatureBank(xi, yi, "cpu"), 9, shuffle=True)
        tl = BankLoader(FeatureBank(xt, yt, "cpu"), 9, shuffle=True)
        ii, ti = iter(il), iter(tl)
        ok = True
        for _ in range(8):  # crosses epoch tails (short batches) for both loaders
            gi, ii = ft.fetch_next(il, ii)
            gt, ti = ft.fetch_next(tl, ti)
            li, lt = ft._local_slice(gi, rank, world), ft._local_slice(gt, rank, world)
            wi, wt = dp_loss_weights(li.n, lt.n, 0.5, gi.n, gt.n)
            st = O.HeadState(head=W.clone(), img_scale=3.0, txt_scale=3.0)
            dW = torch.zeros_like(W)
            if li.n:
                s, gr = O.uml_step_grads(st, xi[li.host_idx], yi[li.host_idx], None, None, 0.0)
                dW += wi * gr["head.weight"]
            if lt.n:
                s, gr = O.uml_step_grads(st, None, None, xt[lt.host_idx], yt[lt.host_idx], 1.0)
                dW += wt * gr["head.weight"]
            dist.all_reduce(dW)
            _, full = O.uml_step_grads(st, xi[gi.host_idx], yi[gi.host_idx], xt[gt.host_idx], yt[gt.host_idx], 0.5)
            ok = ok and torch.allclose(dW, full["head.weight"], rtol=1e-5, atol=1e-6)
            # the ranks' slices tile the global batch
            sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([li.n]))
            ok = ok and int(sum(sizes)) == gi.n
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_per_rank_sharded_loaders_cover_their_shards_every_epoch():
    """shard_bank + BankLoader(shard_of=...): equal strided shards, every epoch of a rank is a permutation of its
    shard (with and without the sampler thread), ranks are decorrelated, batches carry the global row count."""
    from uml_b200 import finetune as ft
    from uml_b200.engine.datasets.utils import BankLoader, shard_bank
    feats, labels = torch.arange(1003 * 2, dtype=torch.float32).view(1003, 2), torch.arange(1003)
    for threshold in (16, 10 ** 9):
        banks = [shard_bank(feats, labels, r, 4, "cpu") for r in range(4)]
        assert all(len(b) == 250 for b in banks)
        assert torch.equal(banks[1].labels, labels[1::4][:250])
        torch.manual_seed(5)
        loaders = [BankLoader(b, 64, shuffle=True, shard_of=(r, 4)) for r, b in enumerate(banks)]
        for ld in loaders:
            ld.async_min_rows = threshold
        its, seen = [iter(ld) for ld in loaders], [[] for _ in loaders]
        for _ in range(12):
            for r, ld in enumerate(loaders):
                b, its[r] = ft.fetch_next(ld, its[r])
                assert b.global_n == b.n * 4
                seen[r].append(b.host_idx.clone())
        for r in range(4):
            idx = torch.cat(seen[r])
            for e in range(3):
                assert sorted(idx[e * 250:(e + 1) * 250].tolist()) == list(range(250))
            assert not torch.equal(idx[:250], idx[250:500])
        assert not torch.equal(torch.cat(seen[0]), torch.cat(seen[1]))
