"""Tuning: the sharded MPPI call's tail on real peers (torchrun, one rank per GPU, exchange over peer memory).
Per rank, in its own globaltimer: when the merger CTAs saw all partials, merged, received every rank's words, updated.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/mppi_stages_mgpu.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B2N_MPPI_DEBUG_TIMES"] = "1"
import _pkg  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
pkg = _pkg.load()
prm = pkg.synthetic.SHIPPED
Kl = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = Kl * world
m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]), prm["lambda_"],
             prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], 0.64, 0.01, Kl, rollout_offset=rank * Kl, rollouts_total=K, device=local)
mine = torch.frombuffer(bytearray(m.p2pExport(world)), dtype=torch.uint8).cuda()
hs = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
dist.all_gather(hs, mine)
m.p2pInit(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in hs))
m.setStateRing(16)
m.seed(42)
m.setWaypoint(pkg.Pose(theta=1.5707, x=1.0, y=0.0))
pose = pkg.Pose(theta=0.0, x=0.0, y=0.0)
names = ["loop start", "loop end", "partial written", "counted in", "merger: all in", "merger: merged", "merger: updated", "merger: received"]
lines = []
for mode in ("synchronous", "pipelined"):
    for rep in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        if mode == "synchronous":
            for _ in range(10):
                m.newControls(pose)
        else:
            for _ in range(200):
                m.enqueue(pose)
            m.wait()
        d = m.debugTimes().astype(np.int64)
        T = m.steps
        t0 = d[:-T, 0].min()
        mg = d[-T:]
        row = {"mode": mode, "rep": rep}
        for j in (1, 2, 3):
            row[names[j]] = (d[:-T, j].max() - t0) / 1e3
        for j in (4, 5, 7, 6):
            row[names[j]] = ((mg[:, j].min() - t0) / 1e3, (np.median(mg[:, j]) - t0) / 1e3, (mg[:, j].max() - t0) / 1e3)
        row["prev updated (max) -> this loop start (min)"] = (t0 - mg[:, 12].max()) / 1e3
        row["period (loop start to loop start, median)"] = float(np.median(d[:-T, 0] - d[:-T, 11])) / 1e3
        row["merged->received (median over steps)"] = float(np.median(mg[:, 7] - mg[:, 5])) / 1e3
        row["received->updated"] = float(np.median(mg[:, 6] - mg[:, 7])) / 1e3
        lines.append(row)
# pipelined rate
for n in (2000,):
    dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        m.enqueue(pose)
    m.wait()
    us = (time.perf_counter() - t) / n * 1e6
out = [None] * world
dist.all_gather_object(out, (rank, us, lines))
if rank == 0:
    for r, us, ls in out:
        print("rank %d: %.2f us per pipelined call (host clock)" % (r, us))
        for row in ls:
            print("  ", {k: (tuple(round(x, 2) for x in v) if isinstance(v, tuple) else (round(v, 2) if isinstance(v, float) else v)) for k, v in row.items()})
dist.barrier()
m.close()
dist.destroy_process_group()
