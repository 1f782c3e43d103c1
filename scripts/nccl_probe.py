"""Latency of the per-trip collective: all-reduce of the (n x B + B) partial block scores (0.32 MB at the headline size).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/nccl_probe.py"""
import os, torch, torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for nelem in (4 * 10_000 + 4, 1 << 10):
    t = torch.ones(nelem, dtype=torch.float64, device=dev)
    for _ in range(20):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    for _ in range(200):
        dist.all_reduce(t)
    e1.record()
    torch.cuda.synchronize()
    if dist.get_rank() == 0:
        print(f"all_reduce of {nelem * 8 / 1e6:.3f} MB over {dist.get_world_size()} GPUs: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call", flush=True)
dist.destroy_process_group()
