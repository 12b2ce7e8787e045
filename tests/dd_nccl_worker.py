"""torchrun worker for test_dd.py::test_dd_nccl_processes: a decomposed box over WORLD_SIZE processes
(one GPU each, NCCL send/recv between bricks) against the single-domain run on rank 0's GPU.
usage: torchrun --nproc-per-node N tests/dd_nccl_worker.py gx gy gz nx steps [cluster_n]   (cluster_n = 4 or 8: the clusterpair scheme)"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    gx, gy, gz, nx, steps = [int(v) for v in sys.argv[1:6]]
    cn = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = importlib.import_module("md-bench_b200")
    uid = [m.dd_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    # clusterpair at DP rel 1e-10 needs a skin no atom outruns between two rebuilds (see test_cp_dd_matches_single_domain)
    extra = dict(skin=0.9) if cn else {}
    P = m.default_params(nx=nx, ny=nx, nz=nx, **extra)
    d = m.Decomposition(P, (gx, gy, gz), nprocs=world, proc=rank, nccl_id=uid[0], device=local, cluster_n=cn)
    n = d.createAtom()
    d.setup(adjust=True)
    rec, _ = d.run(steps)
    tags, x = d.get("x")
    _, v = d.get("v")
    parts = [None] * world
    dist.all_gather_object(parts, (tags, x, v, rec))
    ok = True
    if rank == 0:
        tg = np.concatenate([p[0] for p in parts]); xx = np.concatenate([p[1] for p in parts]); vv = np.concatenate([p[2] for p in parts])
        if cn:
            s = m.ClusterSimulation(m.default_params(nx=nx, ny=nx, nz=nx, **extra), cluster_n=cn, device=local)
            s.createAtom(); s.setup(adjust=True)
            rs, _ = s.run(steps)
            xs, ts = s.atoms("x", tags=True)
            vs = s.atoms("v")
            sx, sv = np.empty_like(xs), np.empty_like(vs)
            sx[ts], sv[ts] = xs, vs
            box = (4.0 / 0.8442) ** (1.0 / 3.0) * nx
        else:
            s = m.Simulation(m.default_params(nx=nx, ny=nx, nz=nx, layout=m.SOA), device=local)
            s.createAtom(); s.setup(adjust=True)
            rs, _ = s.run(steps)
            box = s.neighborParams()["xprd"]
            sx, sv = s.get("x"), s.get("v")
        dx = xx - sx[tg]
        dx -= box * np.round(dx / box)
        ok = (len(tg) == n and np.array_equal(np.sort(tg), np.arange(n))
              and np.abs(dx).max() <= 1e-10 * box and np.abs(vv - sv[tg]).max() <= 1e-9 * np.abs(sv).max()
              and np.abs(rec[:, 1:] - rs[:, 1:]).max() <= 1e-10 * np.abs(rs[:, 1:]).max()
              and all(np.array_equal(p[3], rec) for p in parts))
        print("DD_NCCL_%s world=%d grid=%dx%dx%d atoms=%d max|dx|=%.3e T=%.12f (single %.12f)"
              % ("OK" if ok else "FAIL", world, gx, gy, gz, n, np.abs(dx).max(), rec[-1][1], rs[-1][1]), flush=True)
        s.close()
    d.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
