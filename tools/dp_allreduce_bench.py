"""Latency of the in-launcher gradient all-reduce (uml_dp_allreduce_f32 -> ncclAllReduce) for the head gradient
(768 k floats = 3 MB).  Run under torchrun; NCCL environment variables select algorithm / protocol."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uml_b200  # noqa: F401,E402
from uml_b200 import _lib  # noqa: E402
from uml_b200.engine.trainer import ensure_dp_comm, ensure_dp_p2p  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ensure_dp_comm()
    lib = _lib.load()
    for n in (768_000, 8_120_000):
        buf = torch.ones(n, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(20):
            _lib.check(lib.uml_dp_allreduce_f32(buf.data_ptr(), n, st))
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            _lib.check(lib.uml_dp_allreduce_f32(buf.data_ptr(), n, st))
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 200 * 1e3
        t = torch.tensor([us], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world {world} n {n}: {float(t):.1f} us per all-reduce  "
                  f"[NCCL_ALGO={os.environ.get('NCCL_ALGO')} NCCL_PROTO={os.environ.get('NCCL_PROTO')} "
                  f"NVLS={os.environ.get('NCCL_NVLS_ENABLE')}]", flush=True)
    # the peer-memory two-shot all-reduce (csrc/dp.cu) on the same sizes
    ensure_dp_p2p(8_120_000)
    st = torch.cuda.current_stream().cuda_stream
    for n in (768_000, 8_120_000):
        for _ in range(20):
            _lib.check(lib.uml_dp_allreduce_p2p(n, st))
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            _lib.check(lib.uml_dp_allreduce_p2p(n, st))
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world {world} n {n}: {float(t):.1f} us per peer-memory all-reduce; failed flag {lib.uml_dp_p2p_failed()}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
