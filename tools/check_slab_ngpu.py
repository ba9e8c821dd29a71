"""Multi-GPU parity check on real GPUs + NCCL (the CPU suite covers the choreography with gloo and an oracle
engine; this covers the CUDA path): the N-rank z-slab run -- eager and CUDA-graph replay -- against the single-GPU
cpic_step of the same uniform thermal plasma.  Compared after `steps` steps: total particle count (exact), the
per-cell particle histogram of the whole box (equal up to the rare borderline crossing: summation order differs),
field energies (rtol 1e-3), E and cB fields (1e-4 of scale).
    torchrun --nproc-per-node N tools/check_slab_ngpu.py [nx ny nz nppc steps]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cabanapic_b200 as cp  # noqa: E402
from cabanapic_b200 import decks, dist as cdist  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    a = [int(v) for v in sys.argv[1:6]] if len(sys.argv) >= 6 else [48, 40, 32, 16, 12]
    nx, ny, nz, nppc, steps = a
    d = decks.uniform_plasma(nx, ny, nz, nppc)
    k, _, we = d.consts()
    gx, gy = nx + 2, ny + 2
    ok = True
    for use_graph in (False, True):
        r = cdist.make_runner(d, k, we, rank, world, local, mode="slab", fp_mode=cp.FP_STRICT)
        r.setup()
        r.step(2, -1)
        if use_graph:
            r.prepare_timed(-1)
        r.step(steps - 2, -1)
        with r._on_stream():
            n_local = r.eng.ctx.num_particles
            p = r.eng.ctx.download_particles()
            f = r.eng.ctx.download_fields()
            e_loc = r.eng.ctx.energies()
        z0, nzl = cdist.slab_ranges(nz, world)[rank]
        # local cell -> global cell (z re-based), histogram over the global box
        c = p["cell"].astype(np.int64)
        cg = c + z0 * gx * gy
        hist = torch.from_numpy(np.bincount(cg, minlength=gx * gy * (nz + 2)).astype(np.int64)).cuda()
        dist.all_reduce(hist)
        tot = torch.tensor([n_local, e_loc[0], e_loc[1]], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot)
        # interior planes of E and cB of this slab, gathered on rank 0
        fl = torch.from_numpy(f[:6].reshape(6, nzl + 2, gy, gx)[:, 1:-1].copy()).cuda()
        parts = [torch.empty((6, n, gy, gx), dtype=fl.dtype, device="cuda") for _, n in cdist.slab_ranges(nz, world)] if rank == 0 else None
        dist.gather(fl, parts, dst=0)
        used_graph = r.used_graph
        r.close()
        if rank == 0:
            with cp.Context(nx, ny, nz, 1, max_particles=d.num_particles, real=np.float32, device=local) as ref:
                ref.init_uniform_plasma(0, d.num_particles, nx, ny, nz, nppc, weight=we)
                ref.upload_fields(d.initial_fields())
                ref.step(k, steps, cp.SORT_FUSED, False)
                pr = ref.download_particles()
                fr = ref.download_fields()
                er = ref.energies()
            hr = np.bincount(pr["cell"].astype(np.int64), minlength=gx * gy * (nz + 2))
            h = hist.cpu().numpy()
            fg = torch.cat(parts, dim=1).cpu().numpy()
            frr = fr[:6].reshape(6, nz + 2, gy, gx)[:, 1:-1]
            scale = np.abs(frr).reshape(6, -1).max(axis=1).reshape(6, 1, 1, 1) + 1e-30
            ferr = float((np.abs(fg - frr) / scale)[:, :, 1:-1, 1:-1].max())
            res = {"ranks": world, "graph": bool(used_graph), "particles": int(tot[0].item()), "particles_ref": int(len(pr["cell"])),
                   "cells_with_different_count": int((h != hr).sum()), "of_cells": int(nx * ny * nz),
                   "e_energy": tot[1].item(), "e_energy_ref": er[0], "b_energy": tot[2].item(), "b_energy_ref": er[1],
                   "max_field_err_of_scale": ferr}
            good = (res["particles"] == res["particles_ref"] and res["cells_with_different_count"] <= 0.001 * res["of_cells"]
                    and abs(res["e_energy"] - er[0]) <= 1e-3 * er[0] and abs(res["b_energy"] - er[1]) <= 1e-3 * max(er[1], 1e-300)
                    and ferr < 1e-4 and (bool(used_graph) == use_graph))
            ok = ok and good
            print(("OK   " if good else "FAIL ") + str(res), flush=True)
    if rank == 0:
        print("slab parity:", "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
