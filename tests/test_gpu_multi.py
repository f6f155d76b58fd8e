"""Multi-GPU correctness on hardware (needs >= 2 visible GPUs: `gpurun --gpus 2 -- python -m pytest tests -m gpu`;
skipped on a single-GPU box).  Host-side logic of the same path is covered on CPU by the world-size-2 gloo test in
test_host_api.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_two_gpu_data_parallel_step_equals_one_gpu_step_on_the_concatenated_batch(dtype):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = str(29600 + os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", port, os.path.join(ROOT, "tests", "ddp_worker.py"), dtype]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0 and "ddp_worker ok" in out.stdout, out.stdout[-3000:]
