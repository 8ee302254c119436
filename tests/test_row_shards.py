"""Row-sharded single-graph solver (BASELINE configs[3]): parity against the oracle.

The ranks' kernels exchange halo rows and packed scalars among themselves, so the protocol
can be exercised on ONE GPU (shards as co-resident kernels on separate streams) as well as on
real peers under torchrun; each run happens in a worker process (tests/rowshard_worker.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "rowshard_worker.py")


def _run(cmd, timeout=600):
    env = dict(os.environ, SIGSDP_SHARD_TIMEOUT_S="15")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert out.returncode == 0 and "ROWSHARD OK" in out.stdout, (out.stdout[-3000:], out.stderr[-3000:])
    return out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("case,nranks,dtype,nit,tiling", [
    ("n300_z16_d32", 2, "f64", 24, -1),     # the benchmarked two-chunk kernels (D = 32 fp64)
    ("n300_z16_d32", 4, "f64", 12, 32),     # more ranks than neighbours: rows pushed to several peers
    ("n500_z8_d64", 3, "f32", 12, -1),      # fp32 two-chunk kernels, uneven split
    ("n500_z13", 2, "f64", 12, -1),         # D = 26: one-chunk staged kernels
    ("n300_z10", 2, "f64", 12, 0),          # direct-gather kernels, rows cut anywhere
    ("cfg3", 4, "f64", 4, -1),              # BASELINE cfg3 size (19,845 nodes)
])
def test_row_shards_on_one_gpu_match_oracle(case, nranks, dtype, nit, tiling):
    _run([sys.executable, WORKER, "group", case, str(nranks), dtype, str(nit), str(tiling)])


@pytest.mark.gpu
def test_row_shards_two_processes_torchrun():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29500 + os.getpid() % 2000
    for case, dtype, nit in (("n300_z16_d32", "f64", 16), ("cfg3", "f64", 4)):
        _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
              "--master-port", str(port), WORKER, "dist", case, dtype, str(nit)])
