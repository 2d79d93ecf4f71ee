"""Worker of tests/test_distributed_cpu.py: one process per rank, gloo backend on 127.0.0.1 (no GPU).

    python _dist_worker.py <case> <rank> <world> <port> <outfile>

case "mppi":  SURVEY.md 8(e) for MPPI - every rank simulates its slice of the rollouts with the CPU oracle (noise
    keyed by GLOBAL rollout id), forms the [T][6] partial (min J, sum e, sum e*duL, sum e*duR, sum duL, sum duR),
    the partials travel with ONE all_gather, every rank applies the identical update in rank order.
case "rbpf":  SURVEY.md 8(e) for RBPF - weights all_gather, identical sequential walk on every rank, particles
    whose ancestor lives on the other rank are exchanged following b2n_pf_plan_migration (the product's host code).
"""
import ctypes as C
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


def mppi_partial(J, du, lam):
    """[T][6] of one shard; J [T][K], du [K][T][2]"""
    m = J.min(axis=1)
    e = np.exp(-(J - m[:, None]) / lam)
    return np.stack([m, e.sum(axis=1), (e * du[:, :, 0].T).sum(axis=1), (e * du[:, :, 1].T).sum(axis=1),
                     du[:, :, 0].sum(axis=0), du[:, :, 1].sum(axis=0)], axis=1)


def mppi_merge_update(parts, u, lam, k_total, umax, uinit):
    """what mppi_update_kernel does with the gathered partials (mppi.cpp:112-137), in rank order"""
    m = parts[:, :, 0].min(axis=0)
    f = np.exp((m[None, :] - parts[:, :, 0]) / lam)
    S = (parts[:, :, 1] * f).sum(axis=0) + k_total * 1e-8
    A = (parts[:, :, 2] * f).sum(axis=0) + 1e-8 * parts[:, :, 4].sum(axis=0)
    B = (parts[:, :, 3] * f).sum(axis=0) + 1e-8 * parts[:, :, 5].sum(axis=0)
    new = np.clip(u + np.stack([A / S, B / S]), -umax, umax)
    out = (new[0, 0], new[1, 0])
    nxt = np.concatenate([new[:, 1:], np.array(uinit).reshape(2, 1)], axis=1)
    return out, nxt


def case_mppi(rank, world):
    K, hor, dt, calls = 512, 0.5, 0.02, 4
    prm = orc.MILD
    Kl = K // world
    o = orc.OracleMppi(hor, dt, Kl, **prm)
    o.noise_philox(7)
    o.set_shard(rank * Kl)
    o.setWaypoint(1.0, 0.0, 1.5707)
    T = o.T
    u = np.zeros((2, T))
    pose = (0.0, 0.0, 0.0)
    res = []
    for c in range(calls):
        o.set_plan(u)
        o.newControls(*pose)        # the shard's own (local) update is discarded; J and du are what is used
        g = o.get()
        part = torch.from_numpy(mppi_partial(g["J"], g["du"], prm["lambda_"]))
        gathered = [torch.zeros_like(part) for _ in range(world)]
        dist.all_gather(gathered, part)                      # the ONE exchange
        ctrl, u = mppi_merge_update(np.stack([t.numpy() for t in gathered]), u, prm["lambda_"], K, prm["max_wheel_vel"], (0.0, 0.0))
        res.append([float(ctrl[0]), float(ctrl[1])])
        pose = orc.unicycle_step(pose, ctrl[0], ctrl[1], dt)
    return {"controls": res, "plan": u.tolist()}


def case_rbpf(rank, world):
    pkg = _pkg.load()
    lib = pkg.load_library()
    Nl, P = 24, 5                      # particles per rank, payload words per particle
    N = Nl * world
    rng = np.random.default_rng(100 + rank)
    w_local = torch.from_numpy(rng.random(Nl) ** 6 + 1e-3)          # uneven weights: forces resampling and migration
    payload = (np.arange(Nl)[:, None] + rank * Nl) * 1000 + np.arange(P)[None, :]
    gathered = [torch.zeros_like(w_local) for _ in range(world)]
    dist.all_gather(gathered, w_local)                                # weights in global particle order
    w = np.concatenate([t.numpy() for t in gathered])
    # the identical sequential walk on every rank (particle_filter.cpp:442-500): the CPU oracle's
    o = orc.OraclePf(num_particles=N)
    o.noise_external(np.concatenate([np.zeros(3 * N), [0.3]]), 3)
    o.set_weights(w)
    resampled, anc = o.normalize_resample()
    assert resampled == 1
    anc = np.ascontiguousarray(anc, dtype=np.int32)
    copy1, copy2 = np.zeros(Nl, np.int32), np.zeros(Nl, np.int32)
    recv, send = np.zeros(3 * Nl, np.int32), np.zeros(2 * Nl * world, np.int32)
    n_recv, n_send = C.c_int(), C.c_int()
    pkg._capi.check(lib.b2n_pf_plan_migration(pkg._capi.as_ptr(anc), N, rank, world, pkg._capi.as_ptr(copy1), pkg._capi.as_ptr(copy2),
                                              pkg._capi.as_ptr(recv), recv.size, C.byref(n_recv), pkg._capi.as_ptr(send), send.size,
                                              C.byref(n_send)))
    new = np.full((Nl, P), -1, dtype=np.int64)
    for m in range(Nl):
        if copy1[m] >= 0:
            new[m] = payload[copy1[m]]
    reqs, bufs = [], []
    for i in range(n_send.value):
        reqs.append(dist.isend(torch.from_numpy(payload[send[2 * i]].copy()), int(send[2 * i + 1])))
    for i in range(n_recv.value):
        b = torch.zeros(P, dtype=torch.int64)
        bufs.append((int(recv[3 * i]), b))
        reqs.append(dist.irecv(b, int(recv[3 * i + 2])))
    for r in reqs:
        r.wait()
    for slot, b in bufs:
        new[slot] = b.numpy()
    for m in range(Nl):
        if copy2[m] >= 0:
            new[m] = new[copy2[m]]
    want = anc[rank * Nl:(rank + 1) * Nl].astype(np.int64)[:, None] * 1000 + np.arange(P)[None, :]
    return {"ok": bool(np.array_equal(new, want)), "n_recv": n_recv.value, "n_send": n_send.value, "ancestors": anc.tolist(),
            "unique_remote": int(len({int(a) for a in anc[rank * Nl:(rank + 1) * Nl] if a // Nl != rank}))}


def main():
    case, rank, world, port, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    res = {"mppi": case_mppi, "rbpf": case_rbpf}[case](rank, world)
    dist.barrier()
    dist.destroy_process_group()
    json.dump(res, open(out, "w"))


if __name__ == "__main__":
    main()
