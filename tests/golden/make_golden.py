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


if __name__ == "__main__":
    if not orc.have_ref():
        raise SystemExit("oracle/_ref/libref_nav.so missing: run `make -C oracle ref` where /root/reference exists")
    rng_vectors()
    rollout_vectors()
    mppi_closed_loop("mppi_c1_shipped_ref.npz", 0.5, 0.02, 128, orc.SHIPPED, calls=6)
    mppi_closed_loop("mppi_c1_mild_ref.npz", 0.5, 0.02, 128, orc.MILD, calls=6)
    mppi_closed_loop("mppi_t64_shipped_ref.npz", 0.64, 0.01, 96, orc.SHIPPED, calls=3)
    mppi_closed_loop("mppi_t100_shipped_ref.npz", 1.0, 0.01, 40, orc.SHIPPED, calls=2)
