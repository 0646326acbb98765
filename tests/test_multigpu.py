"""
Multi-rank coverage.

* CPU (gloo, world_size 2): the host-side plumbing of the sharded path --
  interval blocks, column ownership, rendezvous of the NCCL id through
  torch.distributed.
* GPU (needs >= 2 GPUs, skipped otherwise): torchrun of tests/mg_worker.py --
  sharded E_cost and SCG against the single-GPU result.
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]

_GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["MFVGP_ROOT"])
from meanfieldvargp_b200.parallel import shard_intervals, owned_columns
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
M = 16
L = M - 1
blocks = [None] * world
dist.all_gather_object(blocks, shard_intervals(L, rank, world))
assert blocks[0][0] == 0 and blocks[-1][1] == L
assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
# every column of x is owned by exactly one rank
cols = [None] * world
dist.all_gather_object(cols, owned_columns(M, *blocks[rank]))
allm = np.concatenate([c[0] for c in cols]); alls = np.concatenate([c[1] for c in cols])
assert np.array_equal(np.sort(allm), np.arange(3 * M + 4))
assert np.array_equal(np.sort(alls), np.arange(2 * M + 3))
# the id made by rank 0 reaches every rank unchanged
box = [os.urandom(128) if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
got = [None] * world
dist.all_gather_object(got, box[0])
assert all(g == got[0] for g in got) and len(got[0]) == 128
dist.barrier()
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
'''


def _torchrun(nproc, script_args, env_extra=None, timeout=300):
    env = dict(os.environ)
    env["MFVGP_ROOT"] = str(ROOT)
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + os.getpid() % 2000)] + script_args
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)


def test_sharding_host_logic_gloo_world2(tmp_path):
    script = tmp_path / "gloo_worker.py"
    script.write_text(_GLOO_WORKER)
    res = _torchrun(2, [str(script)])
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "GLOO_OK" in res.stdout


@pytest.mark.gpu
def test_sharded_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 4 if n >= 4 else 2
    res = _torchrun(world, [str(ROOT / "tests" / "mg_worker.py")], timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert f"MG_OK {world}" in res.stdout


@pytest.mark.gpu
def test_status_bits_survive_the_rank_combine():
    """a status bit raised on ONE rank (here MFVGP_ST_SYNC_TIMEOUT, injected on rank 1) reaches
    the status word of every rank"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    res = _torchrun(2, [str(ROOT / "tests" / "mg_worker.py")],
                    env_extra={"MG_MODE": "inject", "MFVGP_INJECT_STATUS": "32@1"}, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "MG_INJECT_OK 2" in res.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"MFVGP_SCG_FUSED": "2"}, {"MFVGP_SCG_FUSED": "0"},
                                 {"MFVGP_XCHG": "nccl"}, {"MFVGP_WALK_TAIL": "0"}])
def test_sharded_variants_match_single_gpu(env):
    """the same worker under the other exchange / control-loop variants: two-kernel SCG folds
    over peer memory, host-driven SCG, the NCCL all-gather exchange, the unfused walk tail"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    res = _torchrun(2, [str(ROOT / "tests" / "mg_worker.py")], env_extra=env, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "MG_OK 2" in res.stdout
