"""Row bands on SEVERAL GPUs (one process per GPU under torchrun, peer-mapped halos with the exchange fused into the spatial
pass) reproduce the single-GPU frame bit for bit.  Skipped on a box with one GPU; the CPU-side plumbing is covered by
tests/test_bands_gloo.py and the band arithmetic by the single-GPU band tests (tests/test_gpu_full_size.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, extra, port):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_bands_multi_gpu.py")] + extra
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "== single GPU: True" in r.stdout and "False" not in r.stdout, r.stdout[-3000:]


def test_two_gpu_bands_equal_the_single_gpu_frame():
    _run(2, ["--width", "960", "--height", "540", "--frames", "4"], 29541)


def test_two_gpu_bands_with_light_edits():
    _run(2, ["--width", "640", "--height", "360", "--frames", "5", "--edit-lights"], 29542)


def test_two_gpu_bands_equal_rows_and_nccl_transport():
    _run(2, ["--width", "640", "--height", "360", "--frames", "3", "--equal-rows", "--halo", "nccl"], 29543)


def test_four_gpu_bands_c3_many_lights():
    _run(4, ["--width", "768", "--height", "432", "--frames", "3", "--config", "c3"], 29544)
