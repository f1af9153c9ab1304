"""Multi-GPU parity on real devices: the sharded pipeline -- fBm, height assembly with its all-reduced
min/max and power summaries, erosion with each halo transport (fused in-kernel peer stores, stand-alone
put kernel, NCCL p2p) -- must give bit for bit what one GPU gives (SURVEY 8d tolerance (v)).

Two set-ups:
  * one GPU, TWO PROCESSES sharing it (always runs): the process group is gloo, the state buffers are
    mapped between the processes with CUDA IPC, and the fused exchange runs exactly as it does over
    NVLink -- peer stores from inside the sweep, system-scope fences, last-CTA flag raise, the
    flag-wait kernel, the C-side sweep loop with programmatic dependent launch;
  * two GPUs (skipped on a one-GPU box): NCCL process group, torch symmetric memory.
The host-side partition logic is covered on CPU with gloo in tests/test_partition.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port, extra_env):
    env = dict(os.environ, MGPU_SKIP_TIMING="1", MASTER_ADDR="127.0.0.1", **extra_env)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return [ln for ln in r.stdout.splitlines() if "bit-identical h/w/s = (True, True, True)" in ln], r.stdout


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_bit_identical_two_processes_one_gpu(world):
    """Peer stores, flags and waits of the fused exchange on ONE device (gloo + CUDA IPC)."""
    ok, out = _run(world, 29541 + world, {"MGPU_SAME_DEVICE": "1", "MGPU_CHECK_K": "300", "NXB_HALO_STRICT": "1"})
    assert len(ok) == 6, out[-3000:]              # 2 transports (fused, nvlink) x 3 mesh sizes
    assert all("peer memory ipc" in ln for ln in ok)


@pytest.mark.parametrize("world", [2])
def test_sharded_bit_identical_to_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    ok, out = _run(world, 29533, {"MGPU_CHECK_K": "700"})   # k=700: most tiles affine
    assert len(ok) == 9, out[-3000:]               # 3 transports x 3 mesh sizes
