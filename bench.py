#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 MPPI path (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" is one controller::MPPI::newControls(): K rollouts x T time steps (noise, RK4 diff-drive
integration, loss, cost-to-go, T softmaxes over K, control update).  At N GPUs every rank simulates
its own K rollouts of an N*K-rollout job (weak scaling) and the per-step partial sums are exchanged
over NVLink peer memory inside the call's kernel (--exchange nccl: one ncclAllGather).  Rank 0 prints ONE JSON line.
Extra keys: parity_check (a fresh handle against the CPU oracle), rbpf (configs[2]), c4 (configs[3], strong scaling),
c5 (configs[4], closed loop).

  value     whole-job trajectory-steps/s with everything resident on the device: `steps` calls are
            queued back to back on one stream and timed with CUDA events (max over ranks).
  e2e       the same metric through the public synchronous call (host pose in, host controls out,
            one stream synchronisation per step).
  roofline  the rollout phase of the call's kernel alone (launched without its merger CTAs): algorithmic
            bytes (12 B per trajectory-step, the fp32 state tensor) over its mean duration, `steps`
            launches back to back between two CUDA events on the launching stream in a second pass;
            peak from MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle port (oracle/liboracle_nav.so), one thread, on a bounded sample.
  --impl reference  times the UNMODIFIED reference controller::MPPI compiled at oracle/_ref
            (single thread: its RNG is one process-global engine) on bounded samples of the same
            workload.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "mppi_trajectory_steps_per_sec"
UNIT = "trajectory-steps/s"
K_ROLLOUTS = 16384
HORIZON, DT = 0.64, 0.01          # T = 64
STATE_RING = 16                   # 16 x 12.6 MB of state tensors = 201 MB > 126 MB of L2
RAMP_CALLS = 20480                # untimed pipelined calls (about 0.5 s) before the warm-up steps: clock / power-state ramp of a fresh box
ALGO_BYTES_PER_TRAJ_STEP = 12     # fp32 (x, y, theta) written once
WAYPOINT = (1.0, 0.0, 1.5707)


def workload_config(n_gpus, exchange="p2p"):
    return {
        "workload": "MPPI K=16384 T=64 diff-drive, quadratic waypoint cost (BASELINE configs[1]), shipped cost params",
        "rollouts_per_gpu": K_ROLLOUTS, "rollouts_total": K_ROLLOUTS * n_gpus, "horizon_steps": 64,
        "noise": "Philox4x32-10 counter-based + binary32 Box-Muller, seed 42",
        "l2": "state tensor written round-robin into %d buffers (%.0f MB) > 126 MB L2" % (STATE_RING, STATE_RING * K_ROLLOUTS * 64 * 12 / 1e6),
        "clock_ramp": "%d untimed pipelined calls (about 0.5 s) before the W warm-up steps (--no-ramp: none)" % RAMP_CALLS,
        "sharding": "rollouts; every step's [6] sums exchanged by the call's merger CTAs over NVLink peer memory (--exchange nccl: ncclAllGather)" if n_gpus > 1 else "none",
        "exchange": exchange if n_gpus > 1 else "none",
    }


# ------------------------------------------------------------------------------- clocks ---------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ------------------------------------------------------------------------ CPU arms ----------------
def cpu_oracle_port(budget_s=10.0):
    """The CPU oracle port, one thread, same workload; calls until `budget_s` seconds are spent."""
    import _oracle as orc
    o = orc.OracleMppi(HORIZON, DT, K_ROLLOUTS)
    o.noise_philox(42)
    o.setWaypoint(*WAYPOINT)
    o.newControls(0.0, 0.0, 0.0)            # warm-up
    calls, t0 = 0, time.perf_counter()
    while True:
        o.newControls(0.0, 0.0, 0.0)
        calls += 1
        el = time.perf_counter() - t0
        if el >= budget_s and calls >= 3:
            break
    return {
        "value": calls * K_ROLLOUTS * o.T / el, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": "%d newControls() calls at K=%d T=%d in %.1f s, oracle/liboracle_nav.so, Philox noise" % (calls, K_ROLLOUTS, o.T, el),
    }


# ------------------------------------------------------------------------------ RBPF leg --------
RBPF_N, RBPF_BEAMS, RBPF_CELLS = 4096, 360, 200 * 200      # BASELINE configs[2]
RBPF_MOTION_NOISE = (2e-3, 1e-3, 1e-3)                      # raised so that weights diverge and resampling fires
# per particle-update, distance field = read 8 B occupancy + write 4 B distance per cell
RBPF_DF_BYTES_PER_PARTICLE = 12 * RBPF_CELLS


def rbpf_inputs(n_scans, seed=4):
    import numpy as np
    import _pkg
    syn = _pkg.load().synthetic
    poses, twists = syn.circle_path(n_scans)
    rng = np.random.default_rng(seed)
    scans = [syn.room_scan(poses[i + 1], rng=rng) for i in range(n_scans)]
    return poses, twists, scans


def rbpf_cpu(kind, budget_s=12.0, n_particles=64):
    """bmapping::ParticleFilter::SLAM() on the host, one thread: the compiled reference (kind 'reference') or the
    oracle port, on `n_particles` of the 4096 particles (its cost is linear in the particle count)."""
    import _oracle as orc
    poses, twists, scans = rbpf_inputs(40)
    q = dict(num_particles=n_particles, init_pose=tuple(poses[0]), motion_noise=RBPF_MOTION_NOISE)
    if kind == "reference":
        f = orc.RefPf(**q)
        f.seed(1)
    else:
        f = orc.OraclePf(**q)
        f.noise_mt19937(1)
    done, t0 = 0, time.perf_counter()
    for i in range(len(scans)):
        f.slam(scans[i], twists[i], poses[i + 1], poses[i])
        done += 1
        el = time.perf_counter() - t0
        if el > budget_s:
            break
    return {"value": done * n_particles / el, "unit": "particle-updates/s", "cores": 1, "kind": kind,
            "sample": "%d SLAM() calls on %d of the %d particles (360 beams, 200x200 map) in %.1f s; cost is linear in the particle count"
                      % (done, n_particles, RBPF_N, el)}


def rbpf_gpu_leg(pkg, torch, n_scans, warmup, rank=0, world=1, local=0, dist=None, exchange="p2p"):
    """BASELINE configs[2]: RBPF 4096 particles (per GPU: weak scaling), 360-beam synthetic lidar, 200x200 map,
    motion-model branch: sample + beam weighting + ray integration + distance field + normalise + resample, every
    scan.  At N > 1 GPUs the weights travel with one allgather per scan, every rank runs the identical walk and
    particles whose ancestor lives on another GPU migrate (ncclSend/ncclRecv)."""
    poses, twists, scans = rbpf_inputs(n_scans + warmup)
    q = pkg.synthetic.pf_params(num_particles=RBPF_N, init_pose=tuple(poses[0]), motion_noise=RBPF_MOTION_NOISE)
    f, exchange = make_filter(pkg, torch, dist, q, RBPF_N, rank, world, local, exchange)
    f.seed(1)
    f.setKernelTiming(True)
    ms = [0.0, 0.0, 0.0]
    resampled, wall, migrated = 0, 0.0, 0
    dup = []
    n0 = f.launchCount()
    for i in range(n_scans + warmup):
        if i == warmup:
            n0 = f.launchCount()
            if world > 1:
                dist.barrier()
        t0 = time.perf_counter()
        f.SLAM(scans[i], pkg.Twist2D(*twists[i]), pkg.Pose(*poses[i + 1]), pkg.Pose(*poses[i]))
        dt_wall = time.perf_counter() - t0
        if i >= warmup:
            wall += dt_wall
            k = f.kernelTimes()
            for j in range(3):
                ms[j] += k[j]
            info = f.resampleInfo()
            resampled += info[1]
            if info[1]:
                import numpy as np
                dup.append(1.0 - len(np.unique(info[2])) / float(len(info[2])))
            migrated += f.migration()[0]
    launches = f.launchCount() - n0
    if world > 1:
        t = torch.tensor(ms + [wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall = [float(v) for v in t[:3]], float(t[3])
        mg = torch.tensor([migrated], dtype=torch.int64, device="cuda")
        dist.all_reduce(mg)
        migrated = int(mg.item())
    dev_ms = sum(ms)
    df_ms = ms[1] / n_scans
    n_all = RBPF_N * world
    achieved = RBPF_N * RBPF_DF_BYTES_PER_PARTICLE / (df_ms * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("rbpf_distance_field_kernel_dram_bytes_per_launch")
    out = {
        "metric": "rbpf_particle_updates_per_sec", "unit": "particle-updates/s", "n_gpus": world, "scaling": "weak",
        "value": n_all * n_scans / (dev_ms * 1e-3),
        "config": {"workload": "RBPF 4096 particles per GPU, 360-beam synthetic lidar, 200x200 occupancy grid (BASELINE configs[2]), motion-model branch",
                   "particles_per_gpu": RBPF_N, "particles_total": n_all, "beams": RBPF_BEAMS, "cells": RBPF_CELLS, "scans": n_scans,
                   "warmup_scans": warmup, "resampled_scans": int(resampled), "particles_migrated_between_gpus": int(migrated),
                   "duplicate_ancestor_fraction_per_resampling_scan": (sum(dup) / len(dup)) if dup else None,
                   "sharding": ("particles; weights allgather + identical walk + migration by %s" % ("peer-memory copy kernel over NVLink" if exchange == "p2p" else "ncclSend/ncclRecv")) if world > 1 else "none",
                   "l2": "per-particle planes total %.1f GB per GPU >> 126 MB L2" % (RBPF_N * RBPF_CELLS * 12 / 1e9)},
        "ms_per_scan": dev_ms / n_scans,
        "kernel_ms_per_scan": {"update(sample+weight+rays)": ms[0] / n_scans, "distance_field": df_ms,
                               "normalise+resample+copy": ms[2] / n_scans},
        "e2e": {"value": n_all * n_scans / wall, "unit": "particle-updates/s", "ms_per_scan": 1e3 * wall / n_scans,
                "h2d_bytes_per_step": 4 * RBPF_BEAMS + 72, "d2h_bytes_per_step": 16,
                "note": "synchronous SLAM(): host scan + twist + odometry in, status / N_eff / resample flag out"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "rbpf_distance_field_groups_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "kernel_ms": df_ms,
                     "algorithmic_bytes_per_launch": RBPF_N * RBPF_DF_BYTES_PER_PARTICLE,
                     "note": "the reference's brushfire is order-dependent (heap ties, seed order): serial per particle by definition, "
                             "parallel over particles only - latency bound, not HBM bound (DESIGN.md 4.3)"},
    }
    f.close()
    return out


def run_reference_arm(args):
    """--impl reference: the unmodified reference controller::MPPI (oracle/_ref), single thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import _oracle as orc
    kind = "reference" if orc.have_ref() else "port"
    Mk = (lambda K: orc.RefMppi(HORIZON, DT, K)) if kind == "reference" else (lambda K: orc.OracleMppi(HORIZON, DT, K))
    # probe the per-trajectory-step cost, then size each step's sample so the run takes ~2 minutes at most
    probe = Mk(256)
    if kind == "reference":
        probe.seed(42)
    else:
        probe.noise_mt19937(42)
    probe.setWaypoint(*WAYPOINT)
    probe.newControls(0.0, 0.0, 0.0)
    t0 = time.perf_counter()
    probe.newControls(0.0, 0.0, 0.0)
    per_ts = (time.perf_counter() - t0) / (256 * probe.T)
    total_calls = args.steps + args.warmup
    k_sample = int(min(K_ROLLOUTS, max(64, 120.0 / (per_ts * probe.T * total_calls))))
    m = Mk(k_sample)
    if kind == "reference":
        m.seed(42)
    else:
        m.noise_mt19937(42)
    m.setWaypoint(*WAYPOINT)
    for _ in range(args.warmup):
        m.newControls(0.0, 0.0, 0.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.newControls(0.0, 0.0, 0.0)
    el = time.perf_counter() - t0
    value = args.steps * k_sample * m.T / el
    sample = "each step = newControls() on %d of the %d rollouts (T=%d), mt19937_64 noise; throughput is linear in K on the CPU" % (k_sample, K_ROLLOUTS, m.T)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus, args.exchange),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle/_ref = unmodified reference sources compiled with a mini-Eigen stand-in (Eigen is not installed); "
                "single thread because the reference draws from one process-global mt19937_64",
    }
    if not args.no_rbpf:
        line["rbpf"] = rbpf_cpu(kind)
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------- handle factories ------
def make_mppi(pkg, torch, dist, horizon, dt, k_per_rank, rank, world, local, exchange):
    """controller::MPPI on this rank's slice of the rollouts; at world > 1 wired for the exchange (peer memory, or NCCL
    when a rank cannot map its peers).  Returns (handle, exchange actually used)."""
    prm = pkg.synthetic.SHIPPED          # inputs only; the CPU checkers are imported by the cpu_baseline / parity legs alone

    def create():
        return pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
                        prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], horizon, dt, k_per_rank,
                        rollout_offset=rank * k_per_rank, rollouts_total=world * k_per_rank, device=local)

    def nccl_wire(m):
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(pkg.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        m.commInit(rank, world, bytes(uid.cpu().numpy().tobytes()))

    m = create()
    if world == 1:
        return m, "none"
    if exchange == "p2p":
        # exchange inside the call's kernel over NVLink peer memory: gather every rank's CUDA IPC handle; if any rank cannot
        # map its peers (no peer access, GPUs hidden from each other) every rank falls back to ncclAllGather
        mine = torch.frombuffer(bytearray(m.p2pExport(world)), dtype=torch.uint8).cuda()
        hs = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(hs, mine)
        ok = torch.ones(1, dtype=torch.int32, device="cuda")
        try:
            m.p2pInit(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in hs))
        except pkg.B2NError:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            return m, "p2p"
        m.close()
        m = create()
    nccl_wire(m)
    return m, "nccl"


def make_filter(pkg, torch, dist, q, n_per_rank, rank, world, local, exchange):
    """bmapping::ParticleFilter on this rank's slice of the particles (weights allgather + migration by peer-memory copies,
    or ncclSend / ncclRecv when a rank cannot map its peers).  Returns (filter, exchange actually used)."""
    if world == 1:
        return pkg.bmapping.make_filter(q, device=local), "none"

    def create():
        f = pkg.bmapping.make_filter(q, particle_offset=rank * n_per_rank, particles_total=world * n_per_rank, device=local)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(pkg.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        f.commInit(rank, world, bytes(uid.cpu().numpy().tobytes()))
        return f

    f = create()
    if exchange != "p2p":
        return f, "nccl"
    mine = torch.frombuffer(bytearray(f.p2pExport()), dtype=torch.uint8).cuda()
    hs = [torch.zeros(640, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(hs, mine)
    ok = torch.ones(1, dtype=torch.int32, device="cuda")
    try:
        f.p2pInit(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in hs))
    except pkg.B2NError:
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 1:
        return f, "p2p"
    f.close()                                       # some rank cannot map its peers: everyone migrates with ncclSend/ncclRecv
    return create(), "nccl"


# --------------------------------------------------------------------------- parity check --------
def mppi_parity_check(pkg, torch, dist, rank, world, local, exchange, calls=3):
    """A fresh handle of the benched configuration (sharded exactly like the timed one), `calls` receding-horizon calls, against
    the UNSHARDED CPU oracle on all world x K rollouts (rank 0; the oracle is the checker here, as in the tests): the bench line
    carries its own evidence that the kernels it timed compute the reference's controls."""
    syn = pkg.synthetic
    m, _ = make_mppi(pkg, torch, dist, HORIZON, DT, K_ROLLOUTS, rank, world, local, exchange)
    m.seed(42)
    m.setWaypoint(pkg.Pose(theta=WAYPOINT[2], x=WAYPOINT[0], y=WAYPOINT[1]))
    o = None
    if rank == 0:
        import _oracle as orc
        o = orc.OracleMppi(HORIZON, DT, K_ROLLOUTS * world)
        o.noise_philox(42)
        o.setWaypoint(*WAYPOINT)
    pose, worst_c, worst_p, worst_s = (0.0, 0.0, 0.0), 0.0, 0.0, 0.0
    import numpy as np
    for _ in range(calls):
        v = m.newControls(pkg.Pose(theta=pose[2], x=pose[0], y=pose[1]))
        variant = m.lastVariant()
        if rank == 0:
            c = o.newControls(*pose)
            g = o.get()
            worst_c = max(worst_c, max(abs(v.ul - c[0]), abs(v.ur - c[1])) / max(1e-3, abs(c[0]), abs(c[1])))
            worst_p = max(worst_p, float(np.max(np.abs(m.plan() - g["plan"]) / np.maximum(np.abs(g["plan"]), 1e-3))))
            so = g["states"][:K_ROLLOUTS]
            worst_s = max(worst_s, float(np.max(np.abs(m.states() - so) / np.maximum(np.abs(so), 1e-2))))
        pose = syn.unicycle_step(pose, v.ul, v.ur, DT)      # the GPU's controls are identical on every rank: so are the poses
    m.close()
    return {"controls_rel_err": worst_c, "plan_rel_err": worst_p, "states_rel_err": worst_s, "calls": calls, "kernel_variant": variant,
            "against": "unsharded CPU oracle on %d rollouts, same Philox streams" % (K_ROLLOUTS * world), "tolerance": 1e-5}


# ------------------------------------------------------------------------------- C4 leg ----------
C4_K, C4_HORIZON, C4_DT = 65536, 1.28, 0.01                 # BASELINE configs[3]: K = 65536, T = 128, obstacle term on


def c4_obstacle_field():
    """distance (m) to the nearest obstacle of the synthetic room's box, 200 x 200 cells of 5 cm over [-5, 5]^2"""
    import numpy as np
    c = -5.0 + 0.05 * (np.arange(200) + 0.5)
    X, Y = np.meshgrid(c, c, indexing="ij")
    dx = np.maximum(np.maximum(0.45 - X, X - 0.85), 0.0)       # a box ahead of the start pose, inside the horizon's reach
    dy = np.maximum(np.maximum(-0.25 - Y, Y - 0.25), 0.0)
    return np.hypot(dx, dy).astype(np.float32)


def c4_leg(pkg, torch, dist, rank, world, local, exchange, steps):
    """BASELINE configs[3]: MPPI K = 65536, T = 128 with the occupancy-grid obstacle term, the rollouts SPLIT over the GPUs
    (strong scaling).  Per-call time pipelined (device) and synchronous (host pose in, host controls out); at N > 1 rank 0
    also times the whole job on its own GPU in the same run, so the speed-up is measured on one box."""
    def timed(m, n):
        stream = torch.cuda.Stream(device=local)
        m.setStream(stream.cuda_stream)
        m.setStateRing(2)
        m.seed(42)
        m.setObstacleField(c4_obstacle_field(), -5.0, -5.0, 0.05, 5e4, 0.4, 1e6)
        m.setWaypoint(pkg.Pose(theta=0.0, x=1.5, y=0.0))
        pose = pkg.Pose(theta=0.0, x=0.0, y=0.0)
        for _ in range(20):
            m.newControls(pose)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        m.enqueueMany(pose, n)
        e1.record(stream)
        m.wait()
        torch.cuda.synchronize()
        dev_us = 1e3 * e0.elapsed_time(e1) / n
        t0 = time.perf_counter()
        for _ in range(n):
            v = m.newControls(pose)
        sync_us = 1e6 * (time.perf_counter() - t0) / n
        return dev_us, sync_us, m.lastVariant(), (v.ul, v.ur)

    k_rank = C4_K // world
    m, used = make_mppi(pkg, torch, dist, C4_HORIZON, C4_DT, k_rank, rank, world, local, exchange)
    if world > 1:
        dist.barrier()
    dev_us, sync_us, variant, ctl = timed(m, steps)
    m.close()
    if world > 1:
        t = torch.tensor([dev_us, sync_us], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_us, sync_us = float(t[0]), float(t[1])
    out = {"workload": "MPPI K=65536 T=128 with occupancy-grid obstacle term (BASELINE configs[3]), rollouts split over the GPUs (strong scaling)",
           "n_gpus": world, "rollouts_per_gpu": k_rank, "exchange": used, "kernel_variant": variant,
           "us_per_call": dev_us, "us_per_call_synchronous": sync_us,
           "trajectory_steps_per_sec": C4_K * 128 / (dev_us * 1e-6), "last_controls": list(ctl)}
    if world > 1:
        if rank == 0:
            m1, _ = make_mppi(pkg, torch, None, C4_HORIZON, C4_DT, C4_K, 0, 1, local, "none")
            d1, s1, _, _ = timed(m1, max(20, steps // 4))
            m1.close()
            out["one_gpu_same_box"] = {"us_per_call": d1, "us_per_call_synchronous": s1}
            out["speedup_vs_one_gpu"] = d1 / dev_us
            out["speedup_vs_one_gpu_synchronous"] = s1 / sync_us
        dist.barrier()
    return out


# ------------------------------------------------------------------------------- C5 leg ----------
def c5_leg(pkg, torch, dist, rank, world, local, exchange, ticks):
    """BASELINE configs[4]: RBPF (4096 particles per GPU: 32768 on eight) + MPPI K = 16384 (split over the GPUs) in closed loop
    on a synthetic trajectory: control at 50 Hz, lidar at 5 Hz.  The loop is the reference's nodes fused
    (mppi_waypoints_node.cpp:231-282, turtle_mapping_node.cpp:451-494): waypoint bookkeeping, newControls() on the pose estimate,
    wheelsToTwist, the simulated robot (DiffDrive::feedforward as fake_diff_encoders drives it), odometers on its encoders,
    SLAM() + getRobotState() on every scan.  Every rank runs the same host loop on the same seeds, so scans and poses agree
    without communication; the filter's answer is the best particle of all ranks."""
    import numpy as np
    syn = pkg.synthetic
    prm = syn.SHIPPED
    dt, hor, scan_every = 0.02, 1.28, 10              # 50 Hz control, T = 64 at dt = 0.02, a scan every 10th tick
    start = (0.0, 0.0, 0.0)
    q = syn.pf_params(num_particles=RBPF_N, init_pose=start, motion_noise=RBPF_MOTION_NOISE)
    f, used_f = make_filter(pkg, torch, dist, q, RBPF_N, rank, world, local, exchange)
    f.seed(1)
    m, used_m = make_mppi(pkg, torch, dist, hor, dt, K_ROLLOUTS // world, rank, world, local, exchange)
    m.seed(2)
    sw = syn.WaypointSwitch([0.8, 1.05, 0.4, -0.25, 0.0], [0.0, 0.76, 1.23, 0.76, 0.0], [0.0, 1.5707, 2.3562, -2.3562, -1.5707], 0.08)
    wx, wy, wth = sw.current()
    m.setWaypoint(pkg.Pose(theta=wth, x=wx, y=wy))
    robot, odo, pf_drive = (syn.DiffDrive(start, prm["wheel_base"], prm["wheel_radius"]) for _ in range(3))
    rng = np.random.default_rng(0)
    # the first scan builds the map before the robot moves
    f.SLAM(syn.room_scan(start, rng=rng), pkg.Twist2D(0.0, 0.0, 0.0), pkg.Pose(*start), pkg.Pose(*start))
    est, prev_odom = start, start
    t_mppi, t_slam, reached, resampled = [], [], 0, 0
    for k in range(ticks):
        nw = sw.update(est[1], est[2])
        if nw is not None:
            reached += 1
            m.setWaypoint(pkg.Pose(theta=nw[2], x=nw[0], y=nw[1]))
        t0 = time.perf_counter()
        v = m.newControls(pkg.Pose(theta=est[0], x=est[1], y=est[2]))
        t_mppi.append(time.perf_counter() - t0)
        w, vx, _ = robot.wheelsToTwist(v.ul, v.ur)
        robot.feedforward(w * dt, vx * dt)
        left, right = robot.getEncoders()
        odo.updateOdometry(left, right)
        est = odo.pose()
        if (k + 1) % scan_every:
            continue
        pf_drive.updateOdometry(left, right)
        cur_odom = pf_drive.pose()
        twist = pf_drive.wheelsToTwist(*pf_drive.wheelVelocities())
        scan = syn.room_scan(robot.pose(), rng=rng)
        t0 = time.perf_counter()
        f.SLAM(scan, pkg.Twist2D(*twist), pkg.Pose(*cur_odom), pkg.Pose(*prev_odom))
        est = f.getRobotState().displacement()
        t_slam.append(time.perf_counter() - t0)
        resampled += f.resampleInfo()[1]
        prev_odom = cur_odom
        odo = syn.DiffDrive(est, prm["wheel_base"], prm["wheel_radius"])       # the odometer restarts from the filter's pose,
        odo.left_curr, odo.right_curr = left, right                            #   on the current encoder angles
    true = robot.pose()
    err = float(np.hypot(est[1] - true[1], est[2] - true[2]))
    stats = np.array([np.mean(t_mppi), np.percentile(t_mppi, 99), np.max(t_mppi), np.mean(t_slam), np.max(t_slam)])
    if world > 1:
        t = torch.from_numpy(stats).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stats = t.cpu().numpy()
    f.close()
    m.close()
    return {"workload": "closed loop RBPF + MPPI at 50 Hz control / 5 Hz lidar on a synthetic trajectory (BASELINE configs[4])",
            "n_gpus": world, "particles_total": RBPF_N * world, "particles_per_gpu": RBPF_N, "rollouts_total": K_ROLLOUTS,
            "rollouts_per_gpu": K_ROLLOUTS // world, "horizon_steps": m.steps, "ticks": ticks, "scans": len(t_slam),
            "exchange": {"mppi": used_m, "rbpf": used_f},
            "plant": "rigid2d::DiffDrive feedforward / updateOdometry restated in numpy (pinned against the compiled reference by tests/test_host_logic.py)",
            "mppi_ms_per_tick": {"mean": 1e3 * stats[0], "p99": 1e3 * stats[1], "max": 1e3 * stats[2], "budget": 1e3 * dt},
            "slam_ms_per_scan": {"mean": 1e3 * stats[3], "max": 1e3 * stats[4], "budget": 1e3 * dt * scan_every,
                                 "includes": "SLAM() + getRobotState() (at N > 1 the best particle of all ranks, read over peer memory)"},
            "meets_50hz_budget": bool(1e3 * stats[2] < 1e3 * dt and 1e3 * stats[4] < 1e3 * dt * scan_every),
            "waypoints_reached": reached, "scans_that_resampled": int(resampled), "final_position_error_m": err}


# ------------------------------------------------------------------------------ our arm ----------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import _pkg
    pkg = _pkg.load()
    lib = pkg.load_library()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    if lib.b2n_device_count() < 1:
        raise SystemExit("bench.py: libb2nav sees no CUDA device (there is no CPU path)")
    if torch.cuda.device_count() <= local:          # launcher restricted each rank to its own GPU
        local = 0
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    requested_exchange = args.exchange
    mppi, args.exchange = make_mppi(pkg, torch, dist, HORIZON, DT, K_ROLLOUTS, rank, world, local, args.exchange)
    T = mppi.steps
    # the handle launches on THIS stream and the timing events are recorded on it (torch's current stream is the
    # legacy default stream, handle 0, which b2n_mppi_set_stream reads as "use your own": events there would bracket
    # nothing but the host's enqueue loop)
    stream = torch.cuda.Stream(device=local)
    assert stream.cuda_stream != 0
    mppi.setStream(stream.cuda_stream)
    mppi.setStateRing(STATE_RING)
    mppi.seed(42)
    mppi.setWaypoint(pkg.Pose(theta=WAYPOINT[2], x=WAYPOINT[0], y=WAYPOINT[1]))
    pose = pkg.Pose(theta=0.0, x=0.0, y=0.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- clock ramp + warm-up -----------------------------------------------------------------------
    # A fresh box (GPU idle while the reference arm ran on the CPU) needs more than W x 24 us to reach its steady state:
    # the first bench of a box measured 26.3 us per call against 23.8 us for every later one.  So before the W warm-up
    # steps the pipelined path runs untimed for RAMP_CALLS calls (about half a second); the W warm-up steps and the K
    # timed steps follow.
    for _ in range(0 if args.no_ramp else RAMP_CALLS // 256):   # a fixed COUNT: sharded ranks must make the same number of calls
        mppi.enqueueMany(pose, 256)
        mppi.wait()
    for _ in range(max(args.warmup, 3)):
        mppi.newControls(pose)
    barrier()

    sampler = ClockSampler(local)
    sampler.start()

    # ---- value: device-resident, `steps` calls queued back to back ------------------------------
    n0 = mppi.launchCount()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    mppi.enqueueMany(pose, args.steps)          # one C loop over the C ABI: no interpreter (or clock-sampler thread) between two launches
    e1.record(stream)
    mppi.wait()
    barrier()
    ms_value = max_over_ranks(e0.elapsed_time(e1))
    launches = mppi.launchCount() - n0

    # ---- e2e: the public synchronous call, host pose in, host controls out ----------------------
    # b2n_mppi_new_controls() = what controller::MPPI::newControls() of the C++ drop-in calls, `steps` times in a row, every
    # call complete (controls on the host) before the next starts.  Timed twice: from a C loop inside the library (what a
    # C++ node pays; the headline) and from this Python loop through ctypes (adds the interpreter's per-call cost).
    barrier()
    ms_c, v = mppi.timeNewControls(pose, args.steps)
    barrier()
    ms_e2e = max_over_ranks(ms_c * args.steps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v = mppi.newControls(pose)
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    ms_e2e_py = max_over_ranks(wall_ms)

    # ---- roofline pass: the rollout kernel alone, launched back to back between two CUDA events on the ------
    # handle's stream (event pairs around single launches add ~7 us of launch latency to a 20 us kernel)
    barrier()
    n_k = min(max(args.steps, 256), 4096)       # at least 256 launches: a run of 20 would mostly time its own start
    k_ms = mppi.timeRollout(pose, n_k)
    k_n = n_k
    clocks = sampler.stop()
    barrier()

    if rank == 0:
        traj_steps = K_ROLLOUTS * T * world
        value = traj_steps * args.steps / (ms_value * 1e-3)
        e2e = traj_steps * args.steps / (ms_e2e * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        achieved = K_ROLLOUTS * T * ALGO_BYTES_PER_TRAJ_STEP / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("mppi_rollout_kernel_dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world, requested_exchange), "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": 24,
                    "d2h_bytes_per_step": 16, "caller": "C loop over b2n_mppi_new_controls (the C ABI the C++ drop-in class calls)",
                    "python_ctypes_ms_per_step": ms_e2e_py / args.steps,
                    "note": "input is the 24-byte pose (travels as kernel parameters), output the 16-byte wheel command written by the call's kernel into mapped pinned memory (four 8-byte words, each tagged with the call's sequence number)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "mppi_rollout_kernel (rollout phase: loop + CTA partials, launched without the merger CTAs)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": k_ms, "kernel_samples": k_n,
                         "algorithmic_bytes_per_launch": K_ROLLOUTS * T * ALGO_BYTES_PER_TRAJ_STEP,
                         "traffic_source": "profiles/roofline_traffic.json: ncu range replay over 16 consecutive launches (tools/measure_traffic.sh)",
                         "note": "the fp64 parity kernel is bound by instruction issue and dependent-instruction latency (ncu: 41 % issue-active at 20 warps per SM, FP64 pipe 25 %), not by HBM (DESIGN.md 3.3)"},
            "last_controls": [v.ul, v.ur],
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_oracle_port()
    mppi.close()
    parity = None
    if not args.no_parity:
        parity = mppi_parity_check(pkg, torch, dist if world > 1 else None, rank, world, local, args.exchange)
    rbpf = c4 = c5 = None
    if not args.no_rbpf:
        rbpf = rbpf_gpu_leg(pkg, torch, args.rbpf_scans, 3, rank, world, local, dist if world > 1 else None, args.exchange)
    if not args.no_c4:
        c4 = c4_leg(pkg, torch, dist if world > 1 else None, rank, world, local, args.exchange, args.c4_steps)
    if not args.no_c5:
        c5 = c5_leg(pkg, torch, dist if world > 1 else None, rank, world, local, args.exchange, args.c5_ticks)
    if rank == 0:
        if parity is not None:
            line["parity_check"] = parity
        if rbpf is not None:
            line["rbpf"] = rbpf
            if world == 1 and not args.no_cpu:
                line["rbpf"]["cpu_baseline"] = rbpf_cpu("port")
        if c4 is not None:
            line["c4"] = c4
        if c5 is not None:
            line["c5"] = c5
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-rbpf", action="store_true", help="skip the RBPF leg (BASELINE configs[2])")
    ap.add_argument("--rbpf-scans", type=int, default=20)
    ap.add_argument("--no-parity", action="store_true", help="skip the parity_check leg (3 calls of a fresh handle against the CPU oracle)")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 leg (BASELINE configs[3]: K=65536 T=128 obstacle term, strong scaling)")
    ap.add_argument("--c4-steps", type=int, default=200)
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 leg (BASELINE configs[4]: RBPF + MPPI closed loop at 50 Hz)")
    ap.add_argument("--c5-ticks", type=int, default=300)
    ap.add_argument("--no-ramp", action="store_true", help="skip the untimed clock-ramp calls before the warm-up steps (ncu launch lists)")
    ap.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p",
                    help="N > 1: how the [T][6] softmax partial travels - inside the update kernel over NVLink peer memory, or ncclAllGather")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 2000:
            args.steps, args.warmup = 20, 3
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
