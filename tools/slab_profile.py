"""Developer tool: kernel-level timeline summary (torch.profiler / CUPTI) of the multi-GPU slab step.
torchrun --nproc-per-node N tools/slab_profile.py [nz] -- rank 0 prints the device-time table of 3 steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cabanapic_b200 as cp  # noqa: E402
from cabanapic_b200 import decks, dist as cdist  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nz = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    d = decks.uniform_plasma(256, 256, nz, 64)
    k, _, we = d.consts()
    r = cdist.make_runner(d, k, we, rank, world, local, mode="slab", fp_mode=cp.FP_STRICT)
    r.setup()
    r.step(4, -1)
    dist.barrier(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        ev[0].record()
        for _ in range(3):
            r.stepper.step(fused=True)
        ev[1].record()
        torch.cuda.synchronize()
    if rank == 0:
        print(f"3 steps: {ev[0].elapsed_time(ev[1]):.3f} ms on the device", flush=True)
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70), flush=True)
        # timeline of one step: device-side gaps
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        last_end = t0
        for e in evs:
            st, en = e.time_range.start - t0, e.time_range.end - t0
            if st > 40000:      # first step only (us)
                break
            gap = e.time_range.start - last_end
            print(f"{st:9.1f} us  +{en - st:8.1f}  gap {gap:7.1f}  {e.name[:90]}")
            last_end = max(last_end, e.time_range.end)
    dist.barrier()
    r.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
