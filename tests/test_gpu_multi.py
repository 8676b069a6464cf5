"""k-GPU row slabs == 1 GPU, bit for bit (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

from conftest import free_port

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("halo", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_equal_single_gpu_bitwise(world, halo):
    """halo = p2p: rows pushed into the neighbours' IPC-mapped ghost rows + stream mem-op flags;
    halo = nccl: ncclSend/ncclRecv groups.  Both must reproduce the single-GPU run bit for bit."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tools", "slab_check.py")]
    env = dict(os.environ, FLUID_HALO=halo, SLAB_ITERS="50", SLAB_H="1024", SLAB_HD="2048")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert "SLAB_CHECK ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
