for r in 1 2; do
for tag in v10d tree; do
  if [ $tag = tree ]; then export NXB_SO=$PWD/nixis_b200/libnixis_b200.so; else export NXB_SO=$PWD/tools/variants/libnxb_$tag.so; fi
  echo "== $tag round $r"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 tools/mgpu_time.py 2>&1 | grep "exchanging loop {}\|compute-only [0-9]\|rank 1:" | cut -c1-200
done; done
