"""Developer tool: kernel-level timeline (torch.profiler / CUPTI) of the NATIVE multi-GPU slab step (cpic_mgpu_step,
eager launches so that every kernel shows).  torchrun --nproc-per-node N tools/slab_profile_native.py [nx ny nz]
-- rank 0 prints the device-time table of 4 steps and the timeline of the first of them."""
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
    g = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else [256, 256, 256]
    d = decks.uniform_plasma(g[0], g[1], g[2], 64)
    k, _, we = d.consts()
    r = cdist.make_runner(d, k, we, rank, world, local, mode="slab", fp_mode=cp.FP_STRICT)
    r.use_graph = False
    r.setup()
    r.use_graph = False
    r.step(4, -1)
    dist.barrier(); torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        r.m.step(k, 4, -1, use_graph=False)
        r.m.sync()
        torch.cuda.synchronize()
    if rank == 0:
        print(f"transport: {r.transport}; 4 steps: {r.m.ctx.last_ms(3):.3f} ms on the device", flush=True)
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=80), flush=True)
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        last_end = t0
        npush = 0
        for e in evs:
            if "k_push3" in e.name:
                npush += 1
                if npush > 1:
                    break
            st, en = e.time_range.start - t0, e.time_range.end - t0
            gap = e.time_range.start - last_end
            print(f"{st:9.1f} us  +{en - st:8.1f}  gap {gap:7.1f}  {e.name[:100]}")
            last_end = max(last_end, e.time_range.end)
    dist.barrier()
    r.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
