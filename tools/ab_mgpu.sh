#!/bin/bash
# Same-box A/B of two builds of the library on 4 GPUs (gpurun --gpus 4 -- bash tools/ab_mgpu.sh): the in-tree library
# against tools/variants/libnxb_<tag>.so (built from another commit: git worktree add /tmp/x <commit>; python -m
# nixis_b200.build there; copy the .so), two rounds each, per-sweep time of the exchanging loop (tools/mgpu_time.py).
for r in 1 2; do
for tag in v10d tree; do
  if [ $tag = tree ]; then export NXB_SO=$PWD/nixis_b200/libnixis_b200.so; else export NXB_SO=$PWD/tools/variants/libnxb_$tag.so; fi
  echo "== $tag round $r"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 tools/mgpu_time.py 2>&1 | grep "exchanging loop {}\|compute-only [0-9]\|rank 1:" | cut -c1-200
done; done
