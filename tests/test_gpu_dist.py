"""GPU multi-rank parity (needs >= 2 GPUs on the box; skipped otherwise): the fused peer-memory halo exchange of the
sharded SpMV (bmsp_spmv_halo) vs the NCCL exchange and vs the single-GPU product.  One process per GPU via torchrun."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_spmv_peer_memory(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for name in ("poisson", "clustered", "rmat", "uniform", "one_directional"):
        assert f"DIST_OK {name}" in r.stdout, r.stdout[-2000:]
