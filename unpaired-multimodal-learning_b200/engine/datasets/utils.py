"""Feature-bank datasets and index loaders (L1 of the hot path).

Replaces the reference's ``engine/datasets/utils.py`` (TextTensorDataset :48-107, DatasetWrapper
:153-174) and the ``torch.utils.data.DataLoader`` objects built in ``finetune.py:370-383``.

The reference loader fetches rows one by one on the host, stacks them and copies the batch to the
device every step.  Here a bank lives in HBM for the whole run and a loader only produces *index
batches*; the rows are gathered on the device by the CUDA kernels.  What is preserved bit-exactly is
the ORDER in which rows are visited: ``BankLoader`` draws from the global torch CPU generator exactly
when and how ``DataLoader(shuffle=True)`` + ``RandomSampler`` do (see ``BankLoader.__iter__``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


def get_few_shot_setup_name(train_shot, seed):
    """``shot_{k}-seed_{s}`` - names few-shot splits and result directories (reference utils.py:9-12)."""
    return "shot_{}-seed_{}".format(train_shot, seed)


class TextTensorDataset(torch.utils.data.Dataset):
    """Text feature rows with optional per-class subsampling / averaging.

    Same contract as the reference class of this name (engine/datasets/utils.py:48-107):
    ``n_shots=None`` keeps every row, an int keeps that many random rows per class (one global-RNG
    ``randperm`` per class, classes in ``torch.unique`` order), ``'average'`` replaces each class by
    its mean row.  Exposes ``input_tensor``, ``label_tensor``, ``eot_indices``."""

    def __init__(self, input_tensor, label_tensor, eot_indices, n_shots=None, class_order=None, class_starts=None):
        """``class_order`` / ``class_starts`` (optional): the class-sorted row index a v2 bank file carries
        (``features.write_bank_v2``: row ids sorted by class, stable; offsets per class id).  Selection and averaging
        walk that index - one pass over the rows - instead of building a boolean mask per class; without it the index is
        computed here once.  The rows themselves may live on the device: they are gathered / reduced there."""
        if n_shots is None:
            picked = (input_tensor, label_tensor, eot_indices)
        elif isinstance(n_shots, int) and not isinstance(n_shots, bool):
            picked = self._subsample(input_tensor, label_tensor, eot_indices, n_shots, class_order, class_starts)
            print(f"=> Using {n_shots} text shots per class, with total of {picked[1].shape[0]} samples")
        elif isinstance(n_shots, str) and n_shots.lower() == "average":
            picked = self._class_means(input_tensor, label_tensor, eot_indices, class_order, class_starts)
            print(f"=> Averaging text features per class, with total of {picked[1].shape[0]} samples")
        else:
            raise ValueError("n_shots must be an int, None, or 'average'")
        self.input_tensor, self.label_tensor, self.eot_indices = picked

    @staticmethod
    def _class_index(labels, class_order, class_starts):
        """(order, starts) on the host: ``order[starts[c]:starts[c + 1]]`` = the rows of class c in ascending row order -
        what ``(labels == c).nonzero()`` yields in the reference (engine/datasets/utils.py:78-79)."""
        if class_order is not None and class_starts is not None:
            return class_order.to("cpu", torch.int64), class_starts.to("cpu", torch.int64)
        lab = labels.to("cpu", torch.int64)
        n_classes = int(lab.max()) + 1 if lab.numel() else 0
        starts = torch.zeros(n_classes + 1, dtype=torch.int64)
        if lab.numel():
            starts[1:] = torch.cumsum(torch.bincount(lab, minlength=n_classes), 0)
        return torch.argsort(lab, stable=True), starts

    @staticmethod
    def _subsample(feats, labels, eot, k, class_order=None, class_starts=None):
        order, starts = TextTensorDataset._class_index(labels, class_order, class_starts)
        counts = (starts[1:] - starts[:-1]).tolist()
        lo = starts.tolist()
        chosen = []
        for cls, cnt in enumerate(counts):  # classes in ascending id = torch.unique order; absent classes draw nothing
            if cnt == 0:
                continue
            perm = torch.randperm(cnt)      # global generator, one draw per present class - the reference's RNG protocol
            chosen.append(order[lo[cls]:lo[cls] + cnt][perm[: min(k, cnt)]])
        chosen = torch.cat(chosen) if chosen else torch.zeros(0, dtype=torch.int64)
        if isinstance(feats, list):
            feats = [feats[i] for i in chosen.tolist()]
        else:
            feats = feats[chosen.to(feats.device)]  # a device-resident bank is gathered on the device
        return feats, labels[chosen.to(labels.device)], eot[chosen.to(eot.device)]

    @staticmethod
    def _class_means(feats, labels, eot, class_order=None, class_starts=None):
        order, starts = TextTensorDataset._class_index(labels, class_order, class_starts)
        counts = starts[1:] - starts[:-1]
        classes = torch.nonzero(counts > 0, as_tuple=True)[0]
        dev = feats.device
        lab_dev = labels.to(dev, torch.int64)
        # one segmented sum over the rows (on the device when the bank lives there) instead of a boolean mask per class
        sums = torch.zeros(counts.numel(), feats.shape[1], dtype=feats.dtype, device=dev).index_add_(0, lab_dev, feats)
        means = sums[classes.to(dev)] / counts[classes].to(dev).unsqueeze(1).to(feats.dtype)
        first = order[starts[:-1][classes]]
        return means, classes.to(labels.device).to(labels.dtype), eot[first.to(eot.device)]

    def __getitem__(self, i):
        return self.input_tensor[i], self.label_tensor[i], self.eot_indices[i]

    def __len__(self):
        t = self.input_tensor
        return t.size(0) if isinstance(t, torch.Tensor) else len(t)


class FeatureBank:
    """An ``[N, D]`` fp32 feature matrix plus ``[N]`` int64 labels resident in HBM.

    ``features.py`` of the reference writes exactly these two tensors per split
    (features.py:152-184, 225-248); nothing in the reference's finetune.py reads the image ones back -
    this class is what does."""

    def __init__(self, features: torch.Tensor, labels: torch.Tensor, device="cuda"):
        if features.dim() != 2 or labels.dim() != 1 or features.shape[0] != labels.shape[0]:
            raise ValueError("FeatureBank: expected features [N, D] and labels [N]")
        self.features = features.detach().to(device=device, dtype=torch.float32).contiguous()
        self.labels = labels.detach().to(device=device, dtype=torch.int64).contiguous()

    @classmethod
    def from_text_dataset(cls, ds: TextTensorDataset, device="cuda"):
        return cls(ds.input_tensor, ds.label_tensor, device)

    def __len__(self):
        return self.features.shape[0]

    def bf16(self):
        """bf16 copy of the rows (tensor-core eval operand), converted once on the device and cached."""
        if getattr(self, "_bf16", None) is None:
            from ... import ops
            self._bf16 = ops.cast_bf16(self.features)
        return self._bf16

    def labels32(self):
        if getattr(self, "_labels32", None) is None:
            self._labels32 = self.labels.to(torch.int32)
        return self._labels32

    @property
    def dim(self):
        return self.features.shape[1]

    @property
    def device(self):
        return self.features.device


@dataclass
class IndexBatch:
    """What a BankLoader yields: ``n`` row indices into ``bank`` (a device int64 view), or
    ``idx=None`` meaning the dense range ``[start, start+n)`` for sequential (eval) loaders."""
    bank: FeatureBank
    idx: Optional[torch.Tensor]
    n: int
    start: int = 0
    host_idx: Optional[torch.Tensor] = None  # the same indices on the host (tests / tracing)
    global_n: Optional[int] = None           # set by per-rank (sharded) loaders: rows of the step over ALL ranks
    # a CUDA event recorded (on the stream the indices were uploaded on) after ``idx`` became valid in device memory, and
    # its position in the process-wide order of such events: with it the step launcher may read ``idx`` from another
    # stream without waiting for everything else that was enqueued before the call (uml_linear_step_args.idx_ready)
    ready: Optional[object] = None
    ready_seq: int = 0


def local_slice(batch: "IndexBatch", rank: int, world: int) -> "IndexBatch":
    """A data-parallel rank's contiguous share of a global batch (sizes differ by at most one row)."""
    if world == 1:
        return batch
    lo, hi = (batch.n * rank) // world, (batch.n * (rank + 1)) // world
    idx = batch.idx[lo:hi] if batch.idx is not None else None
    host = batch.host_idx[lo:hi] if batch.host_idx is not None else None
    return IndexBatch(batch.bank, idx, hi - lo, batch.start + lo, host, ready=batch.ready, ready_seq=batch.ready_seq)


_READY_SEQ = [0]


class _ReadyRing:
    """Events that vouch for uploaded index ranges (IndexBatch.ready), recycled: an event that is recorded again while an
    older batch still points at it only makes that batch's consumer wait for a LATER point of the same stream."""

    def __init__(self, n=32):
        self.events, self.pos, self.n = [], 0, n

    def mark(self):
        if not torch.cuda.is_available():
            return None, 0
        if len(self.events) < self.n:
            self.events.append(torch.cuda.Event())
        ev = self.events[self.pos % len(self.events)] if len(self.events) == self.n else self.events[-1]
        self.pos += 1
        ev.record()
        _READY_SEQ[0] += 1
        return ev, _READY_SEQ[0]


def mark_ready(batch: "IndexBatch", ring: Optional["_ReadyRing"] = None) -> "IndexBatch":
    """Record 'idx is valid from here on' on the current stream for a batch whose indices were just written there."""
    ring = ring or _DEFAULT_RING
    batch.ready, batch.ready_seq = ring.mark()
    return batch


_DEFAULT_RING = _ReadyRing(64)


def shard_bank(features: torch.Tensor, labels: torch.Tensor, rank: int, world: int, device="cuda") -> "FeatureBank":
    """Rank ``rank``'s row shard of a bank for the per-rank sampler of data-parallel runs: rows rank, rank+world,
    ... (strided, so class-sorted banks stay mixed), all shards cut to the same length N // world (the last
    N % world rows are dropped, as DistributedSampler(drop_last=True) does).  Each rank then holds and shuffles only
    its own rows: the epoch permutation - a sequential Fisher-Yates on the host - costs 1/world per rank instead
    of capping every rank at the single global sampler's ~200 M rows/s."""
    n = features.shape[0] // world
    return FeatureBank(features[rank::world][:n], labels[rank::world][:n], device)


def _native_randperm(seed: int, n: int, pin: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torch.randperm(n, generator=Generator().manual_seed(seed)) computed by uml_randperm_i64."""
    from ..._lib import check, load
    if n >= (2 ** 32 - 1) // 20:  # torch switches to a 64-bit draw there; not a bank size this path sees
        return torch.randperm(n, generator=torch.Generator().manual_seed(seed))
    if out is None:
        out = torch.empty(n, dtype=torch.int64, pin_memory=pin)
    check(load().uml_randperm_i64(seed & (2 ** 64 - 1), n, out.data_ptr()))
    return out


class _PinnedRing:
    """Pinned host buffers a loader cycles through from epoch to epoch.  Allocating pinned memory per epoch is
    far too slow for a loader whose epoch is two steps long (cudaHostAlloc; torch's caching host allocator cannot
    recycle a block whose asynchronous copies are still queued), so the buffers live as long as the loader and
    a CUDA event per buffer says when the copies that read it have drained.  The ring is deep enough that the
    host can run ``RUN_AHEAD`` steps ahead of the GPU without ever waiting on such an event."""

    RUN_AHEAD = 128  # steps the host may be ahead of the device (two 64-step chunks)

    def __init__(self, n: int, steps_per_epoch: int):
        self.n = n
        depth = min(80, max(3, -(-self.RUN_AHEAD // max(1, steps_per_epoch)) + 2))
        # ONE pinned allocation, sliced: cudaHostAlloc costs milliseconds per call whatever the size
        arena = torch.empty(depth * max(n, 1), dtype=torch.int64, pin_memory=True)
        self.bufs = [arena[i * n:(i + 1) * n] for i in range(depth)]
        self.events = [None] * depth
        self.turn = -1

    def acquire(self) -> torch.Tensor:
        """Next buffer of the ring (its ``_ring_slot`` attribute names the slot to release later)."""
        self.turn = (self.turn + 1) % len(self.bufs)
        ev = self.events[self.turn]
        if ev is not None:
            ev.synchronize()  # copies issued len(bufs) epochs ago; complete unless the host is very far ahead
        buf = self.bufs[self.turn]
        buf._ring_slot = self.turn
        return buf

    def release(self, buf):
        """Call after the last asynchronous copy out of ``buf`` has been enqueued."""
        slot = buf._ring_slot
        ev = self.events[slot]
        if ev is None:
            ev = self.events[slot] = torch.cuda.Event()
        ev.record()


def _draw_int64(generator=None) -> int:
    return int(torch.empty((), dtype=torch.int64).random_(generator=generator).item())


class _SamplerThread:
    """One long-lived daemon thread that runs permutation jobs (each job is a single C call that releases the GIL).
    Spawning a Python thread per epoch costs 0.1-0.2 ms of main-thread time - too much when an epoch is five
    steps long (8-GPU shards)."""

    _inst = None

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        import queue
        import threading
        self.q = queue.SimpleQueue()
        threading.Thread(target=self._loop, daemon=True, name="uml-sampler").start()

    def _loop(self):
        while True:
            fn, args = self.q.get()
            fn(*args)

    def submit(self, fn, *args):
        self.q.put((fn, args))


class _EpochPerm:
    """One epoch's permutation, produced incrementally by uml_randperm_begin / uml_randperm_advance.

    Iteration i of torch's Fisher-Yates makes element i final, so a batch only needs the prefix that covers it.
    Small permutations are finished on the spot; large ones (the 1.28 M-row ImageNet bank costs ~7 ms of
    sequential host work) are advanced in chunks by a daemon thread while the training loop consumes the prefix -
    ``wait(upto)`` blocks only if the loop catches up with the generator."""

    CHUNK = 32768

    def __init__(self, seed: int, n: int, out: torch.Tensor, threaded: bool, prefilled: bool = False,
                 next_out: Optional[torch.Tensor] = None, queued: bool = False):
        import ctypes as C
        import threading
        from ..._lib import check, load
        self.n, self.out = n, out
        self.ready = 0
        self.next_out = None
        self._lib, self._check = load(), check
        if n >= (2 ** 32 - 1) // 20:  # torch switches to a 64-bit draw there; not a bank size this path sees
            out.copy_(torch.randperm(n, generator=torch.Generator().manual_seed(seed)))
            self.ready = n
            return
        self._state = (C.c_ubyte * 3072)()
        if not threaded:
            check(self._lib.uml_randperm_begin(self._state, seed & (2 ** 64 - 1), n, out.data_ptr()))
            check(self._lib.uml_randperm_advance(self._state, n))
            self.ready = n
            return
        # ONE C call is the whole thread body (progress is published through the state block), so the generator
        # never waits for the interpreter lock between chunks while the training loop runs Python code
        self._seed = seed & (2 ** 64 - 1)
        self._threaded = True
        self.next_out = next_out  # the thread leaves the identity there for the next epoch (its seed is not known yet)
        args = (self._state, self._seed, n, out.data_ptr(), self.CHUNK, int(prefilled),
                next_out.data_ptr() if next_out is not None else None)
        if queued:   # shard mode: jobs are prepared an epoch ahead, FIFO order on one persistent thread is fine
            _SamplerThread.get().submit(self._lib.uml_randperm_run, *args)
        else:        # the training loop waits for this one right away: its own thread, never queued behind another
            threading.Thread(target=self._lib.uml_randperm_run, name="uml-sampler", daemon=True, args=args).start()

    def next_filled(self) -> bool:
        return self.next_out is not None and bool(self._lib.uml_randperm_next_filled(self._state))

    def wait(self, upto: int):
        """Returns once out[0:upto] is final."""
        if self.ready >= upto:
            return
        self._check(self._lib.uml_randperm_wait(self._state, upto))
        self.ready = upto


class BankLoader:
    """``DataLoader(dataset, batch_size, shuffle, drop_last, num_workers, generator)`` over a bank.

    RNG protocol (what makes the sampler order bit-exact with the reference):
      * ``iter(loader)`` draws one int64 "base seed" from ``generator`` (global CPU generator when
        None) - every DataLoader iterator does, shuffled or not (so ``validate`` consumes RNG too);
      * a shuffled loader then draws the sampler seed and a ``torch.randperm(n)`` from a fresh
        generator seeded with it - lazily at the first ``next()`` when ``num_workers == 0``, but
        already inside ``iter()`` when ``num_workers > 0`` (worker loaders prefetch in their
        constructor).  With an explicit ``generator`` the permutation is drawn from it directly and a
        second, discarded permutation is drawn when the epoch ends (``RandomSampler.__iter__`` tail).
    ``num_workers`` therefore only selects the protocol; no worker processes exist.

    ``upload`` chooses how index batches reach the device: "epoch" copies the whole permutation once
    per epoch and yields views of it; "step" copies each batch from pinned host memory when it is
    fetched."""

    def __init__(self, bank: FeatureBank, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                 num_workers: int = 0, generator: Optional[torch.Generator] = None, upload: str = "epoch",
                 pin_memory: bool = True, shard_of: Optional[tuple] = None, rng: Optional[torch.Generator] = None):
        """``rng``: a generator that stands in for the GLOBAL default CPU generator in the protocol above (base seeds and
        sampler seeds are drawn from it) - lets several independent runs live in one process, each with the index
        stream it would have had alone after ``torch.manual_seed`` (sweep-level batching, ``engine/sweep.py``).
        ``shard_of=(rank, world)``: the bank is this rank's shard (``shard_bank``) of a data-parallel run and
        ``batch_size`` the PER-RANK batch; batches are tagged with the global row count and every rank's sampler
        seed is decorrelated by its rank.  The index stream is then no longer the single-process reference's."""
        if upload not in ("epoch", "step"):
            raise ValueError("upload must be 'epoch' or 'step'")
        self.bank, self.batch_size, self.shuffle = bank, int(batch_size), bool(shuffle)
        self.drop_last, self.num_workers, self.generator = bool(drop_last), int(num_workers), generator
        self.upload = upload
        self.dataset = bank
        self.shard_of = shard_of
        self.rng = rng
        self._seed_mix = 0 if shard_of is None else ((shard_of[0] + 1) * 0x9E3779B97F4A7C15) & (2 ** 63 - 1)
        self._dev_ring, self._dev_turn = None, 0  # device copies of the permutation (large banks), alternating
        self._ring = None   # pinned permutation buffers (CUDA banks only), created at the first shuffled epoch
        self._live = None   # the iterator whose permutation currently occupies the ring's buffer
        self.async_min_rows = 65536  # permutations at least this long are produced by a sampler thread
        self._prepared = None        # shard mode: the next epoch's permutation, already being generated
        self._ready_ring = _ReadyRing()    # events after the index uploads (IndexBatch.ready)
        self._shard_base, self._shard_epoch = 0, 0

    def _upload(self, host):
        """Host (pinned) index range -> device, asynchronously on the current stream.  (Measured on B200: a copy stream of
        the loader's own - first with a tensor and an event per call, then with a preallocated ring of device buffers - was
        never faster end to end: at the throughput batch the public train() call is bounded by the reference-exact
        sequential sampler, ~225 M indices/s on the GPU box's host, not by the 24 us of PCIe time per step.)"""
        return host.to(self.bank.device, non_blocking=True)

    def __len__(self):
        n = len(self.bank)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _shard_seed(self) -> int:
        """Seed of the next epoch of a per-rank loader: splitmix64 of (base, epoch counter)."""
        self._shard_epoch += 1
        z = (self._shard_base + 0x9E3779B97F4A7C15 * self._shard_epoch) & (2 ** 64 - 1)
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        return (z ^ (z >> 31)) & (2 ** 63 - 1)

    def __iter__(self):
        return _BankIter(self)


class _BankIter:
    def __init__(self, loader: BankLoader):
        self.l = loader
        self.n = len(loader.bank)
        self.pos = 0
        self.perm_host = None
        self.perm_dev = None
        self.perm = None
        self.uploaded = 0
        self.mark = (None, 0)  # event + order of the last upload into perm_dev (IndexBatch.ready)
        self.tail_drawn = False
        if loader.shard_of is None:
            # base seed (the reference DataLoader's protocol)
            _draw_int64(loader.generator if loader.generator is not None else loader.rng)
        if loader.shuffle and loader.num_workers > 0:
            self._draw()

    def _next_buffer(self):
        l = self.l
        if l.bank.device.type == "cuda":
            if l._ring is None:
                l._ring = _PinnedRing(self.n, len(l))
            return l._ring.acquire()
        return torch.empty(self.n, dtype=torch.int64)

    def _draw(self):
        l = self.l
        if l._ring is not None and l._live is not None and l._live.perm_host is not None:
            l._ring.release(l._live.perm_host)  # the previous epoch's copies are all enqueued by now
        self.perm = None
        if l.generator is None:
            # Fresh generator seeded from the global stream: the native sampler restates torch.randperm for this
            # case bit-exactly (csrc/sampler.cu), straight into pinned memory, incrementally for large banks.
            threaded = self.n >= l.async_min_rows
            if l.shard_of is not None:
                # Per-rank sampler: its seeding protocol is ours (the index stream is not the single-process
                # reference's anyway), so the seed of epoch e+1 is drawn when epoch e STARTS and that permutation is
                # generated in the background during epoch e - an epoch boundary then costs nothing, which matters
                # when a shard's epoch is only a handful of steps long (8 GPUs: 4.7 steps).
                self.perm, l._prepared = l._prepared, None
                if self.perm is None:
                    l._shard_base = _draw_int64(None) ^ l._seed_mix  # ONE draw from the global generator per loader
                    self.perm = _EpochPerm(l._shard_seed(), self.n, self._next_buffer(), threaded)
                l._prepared = _EpochPerm(l._shard_seed(), self.n, self._next_buffer(), threaded, queued=True)
                self.perm_host = self.perm.out
            else:
                seed = _draw_int64(l.rng)
                prev = l._live.perm if l._live is not None else None
                if threaded and prev is not None and prev.next_filled():
                    buf, prefilled = prev.next_out, True  # the previous epoch's thread left the identity in this buffer
                else:
                    buf, prefilled = self._next_buffer(), False
                nxt = self._next_buffer() if threaded else None
                self.perm = _EpochPerm(seed, self.n, buf, threaded, prefilled, nxt)
                self.perm_host = buf
        else:
            self.perm_host = torch.randperm(self.n, generator=l.generator, out=self._next_buffer())
        l._live = self
        if l.upload == "epoch":
            if self.perm is not None and self.perm.ready < self.n:
                # still being generated: the device copy is filled batch by batch as the prefix becomes final.
                # Two device buffers per loader, alternating: a fresh 10 MB allocation at an epoch boundary can
                # mean a cudaMalloc (milliseconds, synchronising); stream order makes the reuse safe.
                if l._dev_ring is None:
                    l._dev_ring = [torch.empty(self.n, dtype=torch.int64, device=l.bank.device) for _ in range(2)]
                l._dev_turn ^= 1
                self.perm_dev = l._dev_ring[l._dev_turn]
                self.uploaded = 0
            else:
                # asynchronous copy from pinned memory on the current stream: a pageable source would make the host
                # wait for every kernel already enqueued and lose its lead over the GPU once per epoch
                self.perm_dev = self.perm_host.to(l.bank.device, non_blocking=True)
                self.uploaded = self.n
                self.mark = l._ready_ring.mark() if self.perm_dev.is_cuda else (None, 0)

    def __iter__(self):
        return self

    def batches_left(self) -> int:
        """Batches this epoch can still yield without re-iterating (0 when exhausted)."""
        l = self.l
        left = self.n - self.pos
        if left <= 0:
            return 0
        return left // l.batch_size if l.drop_last else -(-left // l.batch_size)

    def take_run(self, k: int):
        """Advance over the next k batches of THIS epoch (k <= batches_left()) without materialising them: returns
        ``(perm_dev, start, total)`` - the device copy of the epoch's permutation, final and uploaded through
        ``start + total``.  For callers that hand the kernels a permutation pointer plus a position (sweep batching)."""
        l = self.l
        if not l.shuffle or l.upload != "epoch":
            raise ValueError("take_run needs a shuffled loader with upload='epoch'")
        if self.perm_host is None:
            self._draw()
        start = self.pos
        total = min(k * l.batch_size, self.n - start)
        if l.drop_last:
            total -= total % l.batch_size
        if self.perm is not None:
            self.perm.wait(start + total)
        if self.uploaded < start + total:
            self.perm_dev[self.uploaded:start + total].copy_(self.perm_host[self.uploaded:start + total], non_blocking=True)
            self.uploaded = start + total
            self.mark = l._ready_ring.mark() if self.perm_dev.is_cuda else (None, 0)
        self.pos = start + total
        return self.perm_dev, start, total

    def take_chunk(self, k: int):
        """The next k batches of THIS epoch (k <= batches_left()) with one wait on the sampler and - for
        ``upload="step"`` - ONE host->device copy of the contiguous index range instead of one per batch."""
        l = self.l
        if not l.shuffle or l.upload != "step":
            return [next(self) for _ in range(k)]
        if self.perm_host is None:
            self._draw()
        start = self.pos
        total = min(k * l.batch_size, self.n - start)
        if l.drop_last:
            total -= total % l.batch_size
        if self.perm is not None:
            self.perm.wait(start + total)
        host = self.perm_host[start:start + total]
        dev = l._upload(host)
        ev, seq = l._ready_ring.mark() if dev.is_cuda else (None, 0)
        out, off = [], 0
        while off < total:
            take = min(l.batch_size, total - off)
            gn = take * l.shard_of[1] if l.shard_of is not None else None
            out.append(IndexBatch(l.bank, dev[off:off + take], take, start + off, host[off:off + take], gn, ev, seq))
            off += take
        self.pos = start + total
        return out

    def __next__(self) -> IndexBatch:
        l = self.l
        if l.shuffle and self.perm_host is None:
            self._draw()
        left = self.n - self.pos
        if left <= 0 or (l.drop_last and left < l.batch_size):
            if l.shuffle and l.generator is not None and not self.tail_drawn:
                torch.randperm(self.n, generator=l.generator)  # RandomSampler's trailing empty slice
                self.tail_drawn = True
            raise StopIteration
        take = min(l.batch_size, left)
        start = self.pos
        self.pos += take
        gn = take * l.shard_of[1] if l.shard_of is not None else None  # equal shards: every rank has `take` rows
        if not l.shuffle:
            return IndexBatch(l.bank, None, take, start, None, gn)
        if self.perm is not None:
            self.perm.wait(start + take)
        host = self.perm_host[start:start + take]
        if l.upload == "epoch":
            if self.uploaded < start + take:
                self.perm_dev[self.uploaded:start + take].copy_(self.perm_host[self.uploaded:start + take], non_blocking=True)
                self.uploaded = start + take
                self.mark = l._ready_ring.mark() if self.perm_dev.is_cuda else (None, 0)
            dev = self.perm_dev[start:start + take]
            ev, seq = self.mark
        else:
            dev = l._upload(host)
            ev, seq = l._ready_ring.mark() if dev.is_cuda else (None, 0)
        return IndexBatch(l.bank, dev, take, start, host, gn, ev, seq)
