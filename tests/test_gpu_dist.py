"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): the distributed path
(m-distributed alm, ring-distributed map) against the CPU oracle, through
sharp_execute_mpi_fortran / the fused IQU entry point, host and device buffers -- once with the
exchange fused into the kernels (peer stores over NVLink, the default) and once with the NCCL
all-to-all fallback."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_distributed_matches_oracle(shtlib, cpu_oracle, p2p):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_check.py")]
    env = dict(os.environ, CMDR_SHT_P2P=p2p)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_CHECK_OK" in out.stdout
