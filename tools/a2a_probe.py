"""Times torch.distributed.all_to_all_single on int64 records (per-rank payload like stage A's exchange)."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 108_000_000
send = torch.arange(n, dtype=torch.int64, device="cuda")
per = n // world
for mode in ("equal", "splits"):
    recv = torch.empty(per * world, dtype=torch.int64, device="cuda")
    ts = []
    for it in range(6):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if mode == "equal":
            dist.all_to_all_single(recv, send[: per * world])
        else:
            dist.all_to_all_single(recv, send[: per * world], output_split_sizes=[per] * world, input_split_sizes=[per] * world)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    if rank == 0:
        print(mode, world, "ranks", [round(x, 3) for x in ts], "ms;", round(per * (world - 1) * 8 / 1e9 / (min(ts) * 1e-3), 1), "GB/s out per rank")
dist.destroy_process_group()
