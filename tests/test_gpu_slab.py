"""Multi-GPU x-slab path: N ranks exchanging one halo column per step over NCCL must equal the monolithic
oracle (bit-exact in strict arithmetic).  Needs >= 2 GPUs; skipped on a single-GPU box (the decomposition logic
itself is covered on CPU by tests/test_slab_cpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("kernel", ["auto", "tma"])
@pytest.mark.parametrize("world", [2, 4])
def test_slabs_over_nccl_equal_monolithic_oracle(world, kernel):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world * 7 + (1 if kernel == "tma" else 0)),
           os.path.join(ROOT, "tests", "slab_worker.py"), kernel]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("SLAB-OK") == 2, res.stdout[-2000:]
    if kernel == "auto":   # the register kernel takes the peer-memory halo path on NVLink boxes; extras: viz fields, bounce-back
        assert "halo=peer" in res.stdout and "SLAB-EXTRAS-OK" in res.stdout, res.stdout[-2000:]


@pytest.mark.parametrize("world", [2])
def test_slabs_over_nccl_fallback_path(world):
    """The grouped ncclSend/ncclRecv exchange stays as the fallback of the peer-memory path: force it."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "slab_worker.py"), "auto"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, LBM2D_HALO="nccl"))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("SLAB-OK") == 2 and "halo=nccl" in res.stdout and "SLAB-EXTRAS-OK" in res.stdout


def test_configs3_full_grid_on_slabs():
    """BASELINE configs[3] at its stated size, 32768x8192 random obstacles, on 2 x-slabs (tests/slab_full_worker.py)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "slab_full_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "SLAB-FULL-OK" in res.stdout and "halo=peer" in res.stdout, res.stdout[-2000:]
