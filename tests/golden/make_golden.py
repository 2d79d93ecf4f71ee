"""Regenerates tests/golden/*.npz from the UNMODIFIED reference compiled at oracle/_ref
(run in the build container, where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures travel with the repo, so the GPU box can check the CUDA path against the
reference's own outputs without /root/reference being present.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as orc  # noqa: E402


def mppi_closed_loop(name, horizon, dt, K, params, calls, seed=42, wpt=(1.0, 0.0, 1.5707), pose0=(0.0, 0.0, 0.0)):
    ref = orc.RefMppi(horizon, dt, K, **params)
    ref.seed(seed)
    ref.setInitialControls(0.0, 0.0)
    ref.setWaypoint(*wpt)
    T = ref.T
    poses, ctrl, plans = [], [], []
    dus = np.zeros((calls, K, T, 2))
    Jsub = np.zeros((calls, T, K))
    pose = pose0
    for c in range(calls):
        poses.append(pose)
        ul, ur = ref.newControls(*pose)
        g = ref.get()
        ctrl.append((ul, ur))
        plans.append(g["plan"])
        dus[c, :, :, 0] = g["duL"].T
        dus[c, :, :, 1] = g["duR"].T
        Jsub[c] = g["Jsub"]
        pose = orc.unicycle_step(pose, ul, ur, dt)
    np.savez_compressed(os.path.join(HERE, name), horizon=horizon, dt=dt, K=K, T=T, seed=seed, wpt=np.array(wpt),
                        poses=np.array(poses), controls=np.array(ctrl), plans=np.array(plans), du=dus, Jsub=Jsub,
                        **{"p_" + k: np.array(v) for k, v in params.items()})
    print(name, "T", T, "controls[0]", ctrl[0])


def rng_vectors():
    L = orc.ref_lib()
    n = 256
    a = np.zeros(n)
    L.ref_rigid2d_seed(42)
    L.ref_rigid2d_normals(n, 0.0, np.sqrt(0.9), a)
    b = np.zeros(n)
    L.ref_bmapping_seed(7)
    L.ref_bmapping_std_normals(n, b)
    np.savez_compressed(os.path.join(HERE, "rng_ref.npz"), rigid2d_seed42_sigma_sqrt0p9=a, bmapping_seed7_std=b)
    print("rng", a[:2], b[:2])


def rollout_vectors():
    """Single rollouts through the reference RK4 + loss (rk4.cpp:49-69, mppi.hpp:87-105)."""
    ref = orc.RefMppi(0.64, 0.01, 1, **orc.SHIPPED)
    ref.setWaypoint(1.0, 0.5, 0.3)
    rng = np.random.default_rng(3)
    u = rng.normal(0.0, 2.0, size=(8, 2, ref.T))
    x0 = rng.normal(0.0, 0.5, size=(8, 3))
    traj = np.zeros((8, ref.T, 3))
    loss = np.zeros((8, ref.T))
    for i in range(8):
        traj[i], loss[i] = ref.rollout(x0[i], u[i])
    np.savez_compressed(os.path.join(HERE, "mppi_rollouts_ref.npz"), u=u, x0=x0, traj=traj, loss=loss, wpt=np.array([1.0, 0.5, 0.3]))
    print("rollouts", traj[0, -1], loss[0, -1])


# small map for fixtures: shipped launch geometry ([-2,2] @ 0.05 -> 80x80, bmapping/launch/slam.launch:40-42)
PF_SMALL = dict(xmin=-2.0, xmax=2.0, ymin=-2.0, ymax=2.0)
SMALL_ROOM = dict(half=1.5, boxes=((0.5, 0.9, -0.2, 0.3),))


def rbpf_grid_fixture(name="rbpf_grid_ref.npz", scans=4):
    """One stand-alone reference GridMapper fed `scans` scans along a circle: end points, free-cell lists,
    likelihoods at displaced poses, the whole map after every integrateScan, the occupied-set iteration order."""
    rng = np.random.default_rng(11)
    poses, _ = orc.circle_path(scans, radius=0.3, step=0.05)
    g = orc.RefGrid(**PF_SMALL)
    out = dict(poses=poses[:scans], scans=[], end_points=[], n_valid=[], lik_pose=[], lik=[], log_odds=[], prob=[],
               occ_dist=[], state=[], occ_order=[], n_occ=[], bucket_count=[], grid_map=[], free_pt=[], free_cells=[],
               free_n=[])
    for i in range(scans):
        scan = orc.room_scan(poses[i], rng=rng, **SMALL_ROOM)
        out["scans"].append(scan)
        ep = g.end_points(scan, poses[i])
        pad = np.zeros((scan.size, 2)); pad[:len(ep)] = ep
        out["end_points"].append(pad); out["n_valid"].append(len(ep))
        for b in (0, len(ep) // 3, len(ep) // 2, len(ep) - 1):
            cells = g.free_cells(ep[b], poses[i])
            padc = -np.ones(256, dtype=np.int32); padc[:len(cells)] = cells
            out["free_pt"].append(ep[b]); out["free_cells"].append(padc); out["free_n"].append(len(cells))
        off = poses[i] + np.array([0.03, -0.02, 0.04])
        rc, lik = g.likelihood(scan, off)
        assert rc == 0
        out["lik_pose"].append(off); out["lik"].append(lik)
        assert g.integrate(scan, poses[i]) == 0
        m = g.grid()
        for k in ("log_odds", "prob", "occ_dist", "state"):
            out[k].append(m[k])
        oo = g.occ_order()
        pado = -np.ones(g.G, dtype=np.int32); pado[:len(oo)] = oo
        out["occ_order"].append(pado); out["n_occ"].append(len(oo)); out["bucket_count"].append(g.bucket_count())
        out["grid_map"].append(g.grid_map())
    np.savez_compressed(os.path.join(HERE, name), **{k: np.array(v) for k, v in out.items()},
                        **{"p_" + k: np.array(v) for k, v in PF_SMALL.items()})
    print(name, "n_occ", out["n_occ"], "lik", out["lik"])


def rbpf_slam_fixture(name, N, scans, icp, seed=42, **kw):
    """The reference ParticleFilter::SLAM over `scans` scans; every standard normal it consumed is recorded per
    call in the layout of b2n_pf_set_noise (per particle draws, then the resampling draw)."""
    rng = np.random.default_rng(5)
    poses, twists = orc.circle_path(scans, radius=0.3, step=0.05)
    params = dict(PF_SMALL)
    params.update(kw)
    r = orc.RefPf(num_particles=N, init_pose=tuple(poses[0]), **params)
    r.seed(seed)
    k = r.q["k"]
    stream = np.zeros(scans * (N * 3 * (k + 1) + 1) + 16)
    L = orc.ref_lib()
    L.ref_bmapping_seed(seed)
    L.ref_bmapping_std_normals(stream.size, stream)      # the same engine, replayed from the same seed
    r.seed(seed)
    pos = 0
    out = dict(scans=[], twists=twists[:scans], odom=poses[:scans + 1], icp_ok=[], icp_pose=[], z=[], per_particle=[],
               weights=[], poses=[], prev_poses=[], resampled=[], robot_state=[], occ_dist0=[], log_odds0=[], new_map=[],
               occ_order0=[], n_occ0=[])
    zmax = N * 3 * (k + 1) + 1
    for i in range(scans):
        scan = orc.room_scan(poses[i + 1], rng=rng, **SMALL_ROOM)
        icp_ok = int(icp and i > 0)
        icp_pose = (twists[i][0], twists[i][1] * np.cos(twists[i][0] / 2), twists[i][1] * np.sin(twists[i][0] / 2))
        assert r.slam(scan, twists[i], poses[i + 1], poses[i], icp_ok, icp_pose) == 0
        per = 3 * (k + 1) if icp_ok else 3
        z = np.zeros(zmax)
        z[:N * per + 1] = stream[pos:pos + N * per + 1]
        pos += N * per + (1 if r.last_resampled else 0)
        st = r.state()
        g0 = r.grid(0)
        oo = r.occ_order(0)
        pado = -np.ones(r.G, dtype=np.int32); pado[:len(oo)] = oo
        out["scans"].append(scan); out["icp_ok"].append(icp_ok); out["icp_pose"].append(icp_pose); out["z"].append(z)
        out["per_particle"].append(per); out["weights"].append(st["weights"]); out["poses"].append(st["poses"])
        out["prev_poses"].append(st["prev_poses"]); out["resampled"].append(r.last_resampled)
        out["robot_state"].append(r.robot_state()); out["occ_dist0"].append(g0["occ_dist"]); out["log_odds0"].append(g0["log_odds"])
        out["new_map"].append(r.new_map()); out["occ_order0"].append(pado); out["n_occ0"].append(len(oo))
    np.savez_compressed(os.path.join(HERE, name), N=N, seed=seed, **{k_: np.array(v) for k_, v in out.items()},
                        **{"p_" + k_: np.array(v) for k_, v in params.items()})
    print(name, "resampled", out["resampled"], "w[-1] range", out["weights"][-1].min(), out["weights"][-1].max())


def rbpf_resample_fixture(name="rbpf_resample_ref.npz"):
    """The reference's normalise + N_eff + low-variance walk alone on hand-made weight vectors."""
    rng = np.random.default_rng(9)
    cases = []
    for N, kind in ((8, "peaked"), (64, "peaked"), (257, "lognormal"), (1024, "lognormal"), (64, "uniform"), (2, "peaked")):
        if kind == "peaked":
            w = np.full(N, 1e-3); w[rng.integers(0, N)] = 1.0; w[rng.integers(0, N)] = 0.5
        elif kind == "lognormal":
            w = np.exp(rng.normal(0.0, 3.0, N))
        else:
            w = np.full(N, 1.0 / N)
        r = orc.RefPf(num_particles=N, **PF_SMALL)
        seed = 100 + N
        r.seed(seed)
        r.set_weights(w)
        rs, anc = r.normalize_resample()
        z = np.zeros(1)
        L = orc.ref_lib(); L.ref_bmapping_seed(seed); L.ref_bmapping_std_normals(1, z)
        cases.append(dict(N=N, w=w, resampled=rs, anc=anc, w_after=r.state()["weights"], z=z[0]))
    np.savez_compressed(os.path.join(HERE, name), n_cases=len(cases),
                        **{"c%d_%s" % (i, k): np.array(v) for i, c in enumerate(cases) for k, v in c.items()})
    print(name, [(c["N"], c["resampled"]) for c in cases])


if __name__ == "__main__":
    if not orc.have_ref():
        raise SystemExit("oracle/_ref/libref_nav.so missing: run `make -C oracle ref` where /root/reference exists")
    rng_vectors()
    rollout_vectors()
    mppi_closed_loop("mppi_c1_shipped_ref.npz", 0.5, 0.02, 128, orc.SHIPPED, calls=6)
    mppi_closed_loop("mppi_c1_mild_ref.npz", 0.5, 0.02, 128, orc.MILD, calls=6)
    mppi_closed_loop("mppi_t64_shipped_ref.npz", 0.64, 0.01, 96, orc.SHIPPED, calls=3)
    mppi_closed_loop("mppi_t100_shipped_ref.npz", 1.0, 0.01, 40, orc.SHIPPED, calls=2)
    rbpf_grid_fixture()
    rbpf_slam_fixture("rbpf_slam_motion_ref.npz", N=8, scans=6, icp=False, motion_noise=(1e-3, 1e-3, 1e-3))
    rbpf_slam_fixture("rbpf_slam_icp_ref.npz", N=4, scans=4, icp=True, k=10, motion_noise=(1e-3, 1e-3, 1e-3),
                      sample_range=(1e-4, 1e-4, 1e-4))
    rbpf_resample_fixture()
