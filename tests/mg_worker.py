"""
Worker for the multi-GPU tests (launched by torchrun, one rank per GPU):
sharded E_cost / find_minimum against the single-GPU evaluation of the same
problem on the same device.  Prints "MG_OK" on success.
"""
import contextlib
import io
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, rel_err, synthetic_l96  # noqa: E402
from meanfieldvargp_b200.parallel import ShardedFreeEnergy  # noqa: E402


def active_columns(fe, D, M):
    n0, n1 = fe.intervals
    L = M - 1
    last = n1 == L
    Nm, Ns = 3 * M + 4, 2 * M + 3
    cm = np.arange(3 * n0, Nm if last else 3 * n1 + 1)
    cs = np.arange(2 * n0, Ns if last else 2 * n1 + 1)
    idx = np.concatenate([(np.arange(D)[:, None] * Nm + cm[None, :]).ravel(),
                          D * Nm + (np.arange(D)[:, None] * Ns + cs[None, :]).ravel()])
    return idx


def inject_status(rank, world, local):
    """MFVGP_INJECT_STATUS=32@1: rank 1 alone raises MFVGP_ST_SYNC_TIMEOUT in every evaluation
    (what the walk kernel does when a mailbox wait gives up); every rank must see the bit in
    the combined status word -- E_cost with and without gradient, and the SCG loop."""
    from meanfieldvargp_b200 import MfvgpError
    assert os.environ["MFVGP_INJECT_STATUS"] == "32@1"
    for path in ("", "walk"):
        g = synthetic_l96(67, 9, seed=5)
        if path:
            os.environ["MFVGP_PATH"] = path
        sh = ShardedFreeEnergy(g, rank, world, device=local)
        os.environ.pop("MFVGP_PATH", None)
        x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
        for want in (True, False):
            sh.E_cost(x, output_gradients=want)
            try:
                sh.local.check_last()
                raise AssertionError("status bit 32 was dropped on rank %d" % rank)
            except MfvgpError as ex:
                assert "MFVGP_ST_SYNC_TIMEOUT" in str(ex)
            assert sh.local.last_status & 32
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                sh.local.find_minimum(x, max_iter=10)
            raise AssertionError("SCG dropped status bit 32 on rank %d" % rank)
        except MfvgpError as ex:
            assert "MFVGP_ST_SYNC_TIMEOUT" in str(ex)
        dist.barrier()
    if rank == 0:
        print("MG_INJECT_OK", world)
    dist.destroy_process_group()


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if os.environ.get("MG_MODE") == "inject":
        return inject_status(rank, world, local)
    # sharded path: default kernels (small problems: coop kernel + assemble + edge exchange)
    # and the walk kernel (what the D=8192 long-horizon problem runs)
    for D, M, path in ((67, 9, ""), (130, 13, ""), (67, 9, "walk"), (130, 41, "walk")):
        g = synthetic_l96(D, M, seed=77 + D)
        os.environ.pop("MFVGP_PATH", None)
        single = make_free_energy(g, device=local)
        x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
        F1, g1 = single.E_cost(x)
        if path:
            os.environ["MFVGP_PATH"] = path
        sh = ShardedFreeEnergy(g, rank, world, device=local)
        os.environ.pop("MFVGP_PATH", None)
        assert sh.local.describe()["path"] == (path or "coop")
        Fs, gs = sh.E_cost(x)
        idx = active_columns(sh.local, D, M)
        assert abs(Fs.item() - F1.item()) <= 1e-13 * abs(F1.item()), (Fs.item(), F1.item())
        err = rel_err(gs.cpu().numpy()[idx], g1.cpu().numpy()[idx])
        assert err <= 1e-13, err
        # value only
        Fv = sh.E_cost(x, output_gradients=False)
        assert abs(Fv.item() - F1.item()) <= 1e-13 * abs(F1.item())
        # SCG: same branch sequence and iterates on the rank's own columns
        with contextlib.redirect_stdout(io.StringIO()):
            x1, fx1, st1 = single.find_minimum(x, max_iter=12)
            xs, fxs, sts = sh.local.find_minimum(x, max_iter=12)
        assert st1["nit"] == sts["nit"] and st1["func_eval"] == sts["func_eval"]
        assert abs(fx1 - fxs) <= 1e-10 * abs(fx1), (fx1, fxs)
        assert rel_err(sts["fx"], st1["fx"]) <= 1e-10
        e2 = rel_err(xs.cpu().numpy()[idx], x1.cpu().numpy()[idx])
        assert e2 <= 1e-8, e2
        dist.barrier()
    # soak: the peer-memory exchange (csrc/peer.cuh, walk_tail.cuh) 150 times over -- the same bits
    # every time on every rank, and identical bits on both owners of a shared column
    g = synthetic_l96(130, 41, seed=9)
    os.environ["MFVGP_PATH"] = "walk"
    sh = ShardedFreeEnergy(g, rank, world, device=local)
    os.environ.pop("MFVGP_PATH", None)
    assert sh.local.describe()["xchg"] in ("peer", "nccl")
    x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
    F0, g0 = sh.E_cost(x)
    F0, g0 = F0.clone(), g0.clone()
    for k in range(150):
        F, gk = sh.E_cost(x)
        if k % 3 == 0:
            sh.E_cost(x, output_gradients=False)          # records-only exchanges in between
    idx = active_columns(sh.local, 130, 41)
    assert torch.equal(gk[idx].view(torch.int64), g0[idx].view(torch.int64)) and F.item() == F0.item()
    n0, n1 = sh.local.intervals
    edge = torch.stack((gk[:130 * 127].view(130, 127)[:, 3 * n0], gk[:130 * 127].view(130, 127)[:, min(3 * n1, 126)]))
    got = [torch.empty_like(edge) for _ in range(world)]
    dist.all_gather(got, edge)
    for r in range(world - 1):                            # right edge of r == left edge of r + 1
        assert torch.equal(got[r][1].view(torch.int64), got[r + 1][0].view(torch.int64)), r
    Fs = [torch.empty_like(F.reshape(1)) for _ in range(world)]
    dist.all_gather(Fs, F.reshape(1))
    assert all(f.item() == F.item() for f in Fs)
    assert sh.local.check_last() == 0
    dist.barrier()
    # the other systems shard the same way (one CTA per interval kernels + assemble + exchange)
    from helpers import load_golden
    for name in ("DW", "L63_T5"):
        g = load_golden("scg", name)
        single = make_free_energy(g, device=local)
        x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
        F1, g1 = single.E_cost(x)
        sh = ShardedFreeEnergy(g, rank, world, device=local)
        Fs, gs = sh.E_cost(x)
        Dn, Mn = sh.local.dim_D, sh.local.num_M
        idx = active_columns(sh.local, Dn, Mn)
        assert abs(Fs.item() - F1.item()) <= 1e-13 * abs(F1.item()), (name, Fs.item(), F1.item())
        assert rel_err(gs.cpu().numpy()[idx], g1.cpu().numpy()[idx]) <= 1e-13
        t1, s1 = single.grad_params(x)
        ts, ss = sh.grad_params(x)
        assert rel_err(ts.cpu().numpy(), t1.cpu().numpy()) <= 1e-13
        assert rel_err(ss.cpu().numpy(), s1.cpu().numpy()) <= 1e-13
        dist.barrier()
    if rank == 0:
        print("MG_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
