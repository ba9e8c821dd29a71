#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for P in 1 0; do
CPIC_MGPU_P2P=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2971$P \
   tools/slab_profile_native.py 256 256 64 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | cut -c1-220 > gpurun_out/c13_timeline_p$P.log
done
tail -70 gpurun_out/c13_timeline_p1.log
