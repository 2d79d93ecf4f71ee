"""Worker of tests/test_multi_gpu.py (launched by torch.distributed.run, one rank per GPU, NCCL).

Sharded MPPI and sharded RBPF through the C ABI against the UNSHARDED CPU oracle on the same global noise streams:
MPPI controls / plan; RBPF weights, poses, ancestors, N_eff and the best particle's exported map after every scan,
with resampling steps that migrate particles between GPUs.  Rank 0 prints a JSON verdict.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import _oracle as orc  # noqa: E402
import _pkg  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    pkg = _pkg.load()
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(pkg.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    uid_bytes = bytes(uid.cpu().numpy().tobytes())
    out = {}

    # ---- MPPI: K rollouts over `world` ranks; the [T][6] partial travels (a) with one ncclAllGather per call,
    # (b) inside the update kernel over NVLink peer memory (CUDA IPC) -------------------------------------------------
    K, hor, dt = 4096, 0.64, 0.01
    prm = orc.SHIPPED
    Kl = K // world
    for mode in ("nccl", "p2p"):
        m = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]), prm["lambda_"],
                     prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], hor, dt, Kl, rollout_offset=rank * Kl, rollouts_total=K, device=local)
        if mode == "nccl":
            m.commInit(rank, world, uid_bytes)
        else:
            mine = torch.frombuffer(bytearray(m.p2pExport(world)), dtype=torch.uint8).cuda()
            hs = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(hs, mine)
            m.p2pInit(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in hs))
        m.seed(42)
        m.setWaypoint(pkg.Pose(theta=1.5707, x=1.0, y=0.0))
        o = orc.OracleMppi(hor, dt, K)
        o.noise_philox(42)
        o.setWaypoint(1.0, 0.0, 1.5707)
        pose, worst = (0.0, 0.0, 0.0), 0.0
        for _ in range(6):
            v = m.newControls(pkg.Pose(theta=pose[2], x=pose[0], y=pose[1]))
            c = o.newControls(*pose)
            worst = max(worst, max(abs(v.ul - c[0]), abs(v.ur - c[1])) / max(1e-3, abs(c[0]), abs(c[1])))
            pose = orc.unicycle_step(pose, c[0], c[1], dt)
        plan_err = float(np.max(np.abs(m.plan() - o.get()["plan"]) / np.maximum(np.abs(o.get()["plan"]), 1e-3)))
        # a burst of queued calls: the exchange must stay in step without host synchronisation
        for _ in range(50):
            m.enqueue(pkg.Pose(theta=0.0, x=0.0, y=0.0))
        m.wait()
        plans = [torch.zeros(2 * m.steps, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(plans, torch.from_numpy(m.plan().reshape(-1).copy()).cuda())
        replicated = all(bool(torch.equal(plans[0], q)) for q in plans)
        out["mppi_" + mode] = {"controls_rel_err": worst, "plan_rel_err": plan_err, "plan_replicated_bitwise": replicated}
        dist.barrier()
        m.close()

    # ---- RBPF: N particles over `world` ranks: weights allgather, identical walk, migration -----------------------
    N, scans = 64, 6
    Nl = N // world
    poses, twists = orc.circle_path(scans)
    q = dict(init_pose=tuple(poses[0]), motion_noise=(2e-3, 1e-3, 1e-3))
    for mode in ("nccl", "p2p"):
        uid2 = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid2.copy_(torch.frombuffer(bytearray(pkg.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid2, 0)
        rng = np.random.default_rng(5)
        f = pkg.bmapping.make_filter(orc.pf_params(num_particles=Nl, **q), particle_offset=rank * Nl, particles_total=N, device=local)
        f.commInit(rank, world, bytes(uid2.cpu().numpy().tobytes()))
        if mode == "p2p":
            mine = torch.frombuffer(bytearray(f.p2pExport()), dtype=torch.uint8).cuda()
            hs = [torch.zeros(640, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(hs, mine)
            f.p2pInit(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in hs))
        f.seed(3)
        of = orc.OraclePf(num_particles=N, **q)
        of.noise_philox(3)
        res = {"weights": 0.0, "poses": 0.0, "ancestors_equal": True, "resampled": 0, "map_equal": True, "migrated": 0,
               "best_pose": 0.0, "best_map_equal": True, "best_unsupported_without_peer_memory": False}
        for i in range(scans):
            scan = orc.room_scan(poses[i + 1], rng=rng)
            f.SLAM(scan, pkg.Twist2D(*twists[i]), pkg.Pose(*poses[i + 1]), pkg.Pose(*poses[i]))
            of.slam(scan, twists[i], poses[i + 1], poses[i])
            st = of.state()
            neff, rs, anc = f.resampleInfo()
            oneff, ors, oanc = of.resample_info()
            res["resampled"] += rs
            res["ancestors_equal"] &= bool(rs == ors and neff == oneff and np.array_equal(anc, oanc))
            sl = slice(rank * Nl, (rank + 1) * Nl)
            w = f.weights()
            res["weights"] = max(res["weights"], float(np.max(np.abs(w - st["weights"][sl]) / np.maximum(np.abs(st["weights"][sl]), 1e-300))))
            p, _ = f.poses()
            res["poses"] = max(res["poses"], float(np.max(np.abs(p - st["poses"][sl]))))
            for j in (0, Nl // 2, Nl - 1):                          # full maps, bit for bit
                gg, go = f.grid(j), of.grid(rank * Nl + j)
                res["map_equal"] &= bool(np.array_equal(gg["log_odds"], go["log_odds"]) and np.array_equal(gg["occ_dist"], go["occ_dist"]))
                res["map_equal"] &= bool(np.array_equal(f.occOrder(j), of.occ_order(rank * Nl + j)))
            res["migrated"] += f.migration()[0]
            # the filter's answer = the best particle of ALL ranks (particle_filter.cpp:255-291), wherever it lives
            if mode == "p2p":
                res["best_pose"] = max(res["best_pose"], float(np.max(np.abs(np.subtract(f.getRobotState().displacement(), of.robot_state())))))
                res["best_map_equal"] &= bool(np.array_equal(f.newMap(), of.new_map()))
            elif i == 0:
                try:
                    f.getRobotState()
                except pkg.B2NError as e:
                    res["best_unsupported_without_peer_memory"] = e.code == -4
        out["rbpf_" + mode] = res
        dist.barrier()
        f.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(gathered), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
