"""Multi-GPU parity on real devices (needs >= 2 GPUs on the box, skipped otherwise): the sharded
pipeline -- fBm, height assembly with its all-reduced min/max and power summaries, erosion with each
halo transport (fused in-kernel NVLink put, stand-alone put kernel, NCCL p2p) -- must give bit for
bit what one GPU gives (SURVEY 8d tolerance (v)).  The host-side partition logic is covered on CPU
with gloo in tests/test_partition.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_sharded_bit_identical_to_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    env = dict(os.environ, MGPU_SKIP_TIMING="1", MASTER_ADDR="127.0.0.1", MGPU_CHECK_K="700")   # k=700: most tiles affine
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ok = [ln for ln in r.stdout.splitlines() if "bit-identical h/w/s = (True, True, True)" in ln]
    assert len(ok) == 9, r.stdout[-2000:]          # 3 transports x 3 mesh sizes
