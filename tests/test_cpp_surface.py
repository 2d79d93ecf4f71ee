"""The C++ drop-in headers (include/controller/mppi.hpp, include/bmapping/particle_filter.hpp): programs
written like the reference's ROS nodes compile against them, link libb2nav.so and - on the GPU box -
produce the same numbers as the ctypes path."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "ros-turtlebot-navigation_b200", "lib")
REF_RIGID2D = "/root/reference/rigid2d/include"


def _compile(src, out, extra_inc=()):
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include")]
    cmd += ["-I" + i for i in extra_inc]
    cmd += [os.path.join(ROOT, "tests", "cpp", src), "-o", out, "-L" + LIBDIR, "-lb2nav", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True)
    return out


@pytest.fixture(scope="module")
def bindir(tmp_path_factory, pkg):
    pkg.load_library()   # fails loudly when libb2nav.so has not been built
    return tmp_path_factory.mktemp("cppbin")


def test_mppi_surface_compiles_and_links(bindir, pkg):
    exe = _compile("mppi_node_like.cpp", str(bindir / "mppi_node_like"))
    r = subprocess.run([exe], capture_output=True, text=True)
    if pkg.load_library().b2n_device_count() < 1:
        # no CPU path: the constructor must throw
        assert r.returncode == 3 and r.stdout.startswith("NO_DEVICE"), (r.returncode, r.stdout, r.stderr)
    else:
        assert r.returncode == 0, (r.stdout, r.stderr)


@pytest.mark.skipif(not os.path.isdir(REF_RIGID2D), reason="reference tree not present on this machine")
def test_mppi_surface_compiles_against_reference_rigid2d(bindir):
    """With the reference's own rigid2d headers first on the include path (the catkin situation) the
    drop-in header uses rigid2d::Pose / WheelVelocities from there."""
    _compile("mppi_node_like.cpp", str(bindir / "mppi_node_like_ref"), extra_inc=[REF_RIGID2D])


@pytest.mark.gpu
def test_mppi_cpp_equals_ctypes_path(bindir, gpu_pkg):
    import _oracle as orc
    exe = _compile("mppi_node_like.cpp", str(bindir / "mppi_node_like_gpu"))
    r = subprocess.run([exe, "128", "3", "42"], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    got = [tuple(float(v) for v in ln.split()) for ln in r.stdout.strip().splitlines()]
    prm = orc.SHIPPED
    m = gpu_pkg.MPPI(gpu_pkg.CartModel(prm["wheel_radius"], prm["wheel_base"]), gpu_pkg.LossFunc(prm["Q"], prm["R"], prm["P1"]),
                     prm["lambda_"], prm["max_wheel_vel"], prm["ul_var"], prm["ur_var"], 0.5, 0.02, 128)
    m.setInitialControls(0.0, 0.0)
    m.seed(42)
    m.setWaypoint(gpu_pkg.Pose(theta=1.5707, x=1.0, y=0.0))
    for c in range(3):
        v = m.newControls(gpu_pkg.Pose(theta=0.0, x=0.0, y=0.0))
        assert (v.ul, v.ur) == got[c]


# ------------------------------------------------------------------------------- bmapping surface ---
def test_slam_surface_compiles_and_links(bindir, pkg):
    exe = _compile("slam_node_like.cpp", str(bindir / "slam_node_like"))
    r = subprocess.run([exe], input="0 0\n", capture_output=True, text=True)
    if pkg.load_library().b2n_device_count() < 1:
        assert r.returncode == 3 and r.stdout.startswith("NO_DEVICE"), (r.returncode, r.stdout, r.stderr)
    else:
        assert r.returncode == 0, (r.stdout, r.stderr)


@pytest.mark.skipif(not os.path.isdir(REF_RIGID2D), reason="reference tree not present on this machine")
def test_slam_surface_compiles_against_reference_rigid2d(bindir):
    """The catkin situation: the reference's rigid2d headers (real Transform2D / Twist2D / Pose) come first."""
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I" + REF_RIGID2D, "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "slam_node_like.cpp"), "/root/reference/rigid2d/src/rigid2d/rigid2d.cpp",
           "-o", str(bindir / "slam_node_like_ref"), "-L" + LIBDIR, "-lb2nav", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True)


@pytest.mark.gpu
@pytest.mark.parametrize("gpu_matcher", [0, 1])
def test_slam_cpp_equals_ctypes_path(bindir, gpu_pkg, gpu_matcher):
    """bmapping::ParticleFilter through the C++ header and through the ctypes mirror: same robot state, same map -
    with the stand-in matcher (motion-model branch) and with GpuScanAlignment (improved-proposal branch)."""
    import numpy as np
    import _oracle as orc
    exe = _compile("slam_node_like.cpp", str(bindir / "slam_node_like_gpu"))
    n_scans, N = 4, 8
    poses, twists = orc.circle_path(n_scans)
    rng = np.random.default_rng(11)
    scans = [orc.room_scan(poses[i + 1], rng=rng) for i in range(n_scans)]
    lines = ["%d %d" % (n_scans, 360)]
    for i in range(n_scans):
        lines.append(" ".join(repr(float(v)) for v in (*twists[i], *poses[i + 1], *poses[i])))
        lines.append(" ".join(repr(float(v)) for v in scans[i]))
    th0, x0, y0 = poses[0]
    r = subprocess.run([exe, str(N), "1", repr(float(x0)), repr(float(y0)), repr(float(th0)), str(gpu_matcher)], input="\n".join(lines) + "\n",
                       capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    got = [ln.split() for ln in r.stdout.strip().splitlines()]
    f = gpu_pkg.bmapping.make_filter(orc.pf_params(num_particles=N, init_pose=tuple(poses[0]), motion_noise=(2e-3, 1e-3, 1e-3),
                                                   sample_range=(1e-3, 1e-3, 1e-3), srr=0.001, srt=0.001, str_=0.001, stt=0.001,
                                                   beam_max=6.28319, beam_delta=0.0174533))
    f.seed(1)
    if gpu_matcher:
        f.scan_matcher = gpu_pkg.bmapping.GpuScanAlignment(f.scan_matcher.props, None)
    for i in range(n_scans):
        f.SLAM(scans[i], gpu_pkg.Twist2D(*twists[i]), gpu_pkg.Pose(*poses[i + 1]), gpu_pkg.Pose(*poses[i]))
        m = f.newMap().astype(np.int64)
        th, x, y = f.getRobotState().displacement()
        assert (float(got[i][0]), float(got[i][1]), float(got[i][2])) == (th, x, y)
        assert int(got[i][3]) == int(np.sum((np.arange(m.size) % 977 + 1) * m)) and int(got[i][4]) == m.size
