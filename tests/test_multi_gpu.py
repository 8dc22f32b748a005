"""Multi-GPU parity on hardware (skipped with fewer than 2 devices): bench.py's sharded decode of a grid on 2 GPUs must give
the same bytes as on 1 GPU, and both must match what the unmodified reference decodes (bench.py's own in-run check)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(n, port):
    base = ["bench.py", "--gpus", str(n), "--steps", "1", "--warmup", "1", "--tiles-side", "4", "--ref-sample-side", "4", "--no-ctb64"]
    cmd = [sys.executable] + base if n == 1 else [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                                                  "--master-port", str(port)] + base
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


@pytest.mark.gpu
def test_two_gpus_equal_one_gpu_equal_reference(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    one = run_bench(1, 29631)
    two = run_bench(2, 29632)
    assert one["parity_checked"] and two["parity_checked"], (one.get("parity"), two.get("parity"))
    assert one["rgb_md5"] == two["rgb_md5"]
    assert two["n_gpus"] == 2


@pytest.mark.gpu
def test_bench_checks_itself_against_the_reference(cuda):
    """1 GPU: the bench line carries parity_checked = true (GPU RGB == heif_decode_image of the unmodified reference)."""
    one = run_bench(1, 29633)
    assert one["parity_checked"] is True, one.get("parity")
    assert one["parity"]["mismatching_bytes"] == 0
