"""One process driving several GPUs (b200_init(0) + b200_compress_batch over one shared list): the shape caesiumclt's
start_compression would call on an 8-GPU box.  Needs >= 2 visible devices (skipped on the single-GPU test box); runs in a child
process because the library binds its devices once per process."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_one_process_shards_a_batch_over_all_devices():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "inprocess_multi.py"), "384"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    rec = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert rec["devices"] == torch.cuda.device_count()
    assert rec["sample_equals_oracle"]
    jobs = rec["jobs_per_device"]
    assert all(j > 0 for j in jobs), jobs
    assert max(jobs) <= 2 * min(jobs) + 4, f"unbalanced sharding: {jobs}"
