#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 MPPI path (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" is one controller::MPPI::newControls(): K rollouts x T time steps (noise, RK4 diff-drive
integration, loss, cost-to-go, T softmaxes over K, control update).  At N GPUs every rank simulates
its own K rollouts of an N*K-rollout job (weak scaling) and the per-step partial sums are exchanged
with one ncclAllGather.  Rank 0 prints ONE JSON line.

  value     whole-job trajectory-steps/s with everything resident on the device: `steps` calls are
            queued back to back on one stream and timed with CUDA events (max over ranks).
  e2e       the same metric through the public synchronous call (host pose in, host controls out,
            one stream synchronisation per step).
  roofline  the rollout kernel alone: algorithmic bytes (12 B per trajectory-step, the fp32 state
            tensor) over its mean duration, `steps` launches back to back between two CUDA events
            on the launching stream in a second pass; peak from MEASURED_PEAKS.json.
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


def workload_config(n_gpus):
    return {
        "workload": "MPPI K=16384 T=64 diff-drive, quadratic waypoint cost (BASELINE configs[1]), shipped cost params",
        "rollouts_per_gpu": K_ROLLOUTS, "rollouts_total": K_ROLLOUTS * n_gpus, "horizon_steps": 64,
        "noise": "Philox4x32-10 counter-based + binary32 Box-Muller, seed 42",
        "l2": "state tensor written round-robin into %d buffers (%.0f MB) > 126 MB L2" % (STATE_RING, STATE_RING * K_ROLLOUTS * 64 * 12 / 1e6),
        "clock_ramp": "%d untimed pipelined calls (about 0.5 s) before the W warm-up steps (--no-ramp: none)" % RAMP_CALLS,
        "sharding": "rollouts; [T][6] partial exchanged inside the update kernel over NVLink peer memory (--exchange nccl: ncclAllGather)" if n_gpus > 1 else "none",
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
    if world > 1:
        f = pkg.bmapping.make_filter(q, particle_offset=rank * RBPF_N, particles_total=world * RBPF_N, device=local)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(pkg.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        f.commInit(rank, world, bytes(uid.cpu().numpy().tobytes()))
        if exchange == "p2p":
            mine = torch.frombuffer(bytearray(f.p2pExport()), dtype=torch.uint8).cuda()
            hs = [torch.zeros(640, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(hs, mine)
            ok = torch.ones(1, dtype=torch.int32, device="cuda")
            try:
                f.p2pInit(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in hs))
            except pkg.B2NError:
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:                 # some rank cannot map its peers: everyone migrates with ncclSend/ncclRecv
                exchange = "nccl"
                f.close()
                f = pkg.bmapping.make_filter(q, particle_offset=rank * RBPF_N, particles_total=world * RBPF_N, device=local)
                if rank == 0:
                    uid.copy_(torch.frombuffer(bytearray(pkg.comm_unique_id()), dtype=torch.uint8))
                dist.broadcast(uid, 0)
                f.commInit(rank, world, bytes(uid.cpu().numpy().tobytes()))
    else:
        f = pkg.bmapping.make_filter(q, device=local)
    f.seed(1)
    f.setKernelTiming(True)
    ms = [0.0, 0.0, 0.0]
    resampled, wall, migrated = 0, 0.0, 0
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
            resampled += f.resampleInfo()[1]
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
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle/_ref = unmodified reference sources compiled with a mini-Eigen stand-in (Eigen is not installed); "
                "single thread because the reference draws from one process-global mt19937_64",
    }
    if not args.no_rbpf:
        line["rbpf"] = rbpf_cpu(kind)
    print(json.dumps(line), flush=True)


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

    prm = pkg.synthetic.SHIPPED          # inputs only; the CPU checkers are imported by the cpu_baseline leg alone
    mppi = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
                    prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], HORIZON, DT, K_ROLLOUTS,
                    rollout_offset=rank * K_ROLLOUTS, rollouts_total=world * K_ROLLOUTS, device=local)
    T = mppi.steps
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(pkg.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        mppi.commInit(rank, world, bytes(uid.cpu().numpy().tobytes()))
        if args.exchange == "p2p":
            # exchange inside the update kernel over NVLink peer memory: gather every rank's CUDA IPC handle; if any
            # rank cannot map its peers (no peer access, GPUs hidden from each other) every rank stays on ncclAllGather
            mine = torch.frombuffer(bytearray(mppi.p2pExport(world)), dtype=torch.uint8).cuda()
            hs = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(hs, mine)
            ok = torch.ones(1, dtype=torch.int32, device="cuda")
            try:
                mppi.p2pInit(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in hs))
            except pkg.B2NError:
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                args.exchange = "nccl"
                mppi.close()
                mppi = pkg.MPPI(pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
                                prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], HORIZON, DT, K_ROLLOUTS,
                                rollout_offset=rank * K_ROLLOUTS, rollouts_total=world * K_ROLLOUTS, device=local)
                if rank == 0:
                    uid.copy_(torch.frombuffer(bytearray(pkg.comm_unique_id()), dtype=torch.uint8))
                dist.broadcast(uid, 0)
                mppi.commInit(rank, world, bytes(uid.cpu().numpy().tobytes()))
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
        for _ in range(256):
            mppi.enqueue(pose)
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
    for _ in range(args.steps):
        mppi.enqueue(pose)
    e1.record(stream)
    mppi.wait()
    barrier()
    ms_value = max_over_ranks(e0.elapsed_time(e1))
    launches = mppi.launchCount() - n0

    # ---- e2e: the public synchronous call, host pose in, host controls out ----------------------
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        v = mppi.newControls(pose)
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))

    # ---- roofline pass: the rollout kernel alone, launched back to back between two CUDA events on the ------
    # handle's stream (event pairs around single launches add ~7 us of launch latency to a 20 us kernel)
    barrier()
    n_k = min(args.steps, 4096)
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
            "dtype": "f64", "data": "synthetic", "config": dict(workload_config(world), exchange=(args.exchange if world > 1 else "none")), "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": 24,
                    "d2h_bytes_per_step": 16,
                    "note": "input is the 24-byte pose (travels as kernel parameters), output the 16-byte wheel command written by the update kernel into mapped pinned memory"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "mppi_rollout_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": k_ms, "kernel_samples": k_n,
                         "algorithmic_bytes_per_launch": K_ROLLOUTS * T * ALGO_BYTES_PER_TRAJ_STEP,
                         "note": "fp64 parity kernel is FP64-pipe bound, not HBM bound (DESIGN.md)"},
            "last_controls": [v.ul, v.ur],
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_oracle_port()
    rbpf = None
    if not args.no_rbpf:
        mppi.close()
        rbpf = rbpf_gpu_leg(pkg, torch, args.rbpf_scans, 3, rank, world, local, dist if world > 1 else None, args.exchange)
    if rank == 0:
        if rbpf is not None:
            line["rbpf"] = rbpf
            if world == 1 and not args.no_cpu:
                line["rbpf"]["cpu_baseline"] = rbpf_cpu("port")
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
