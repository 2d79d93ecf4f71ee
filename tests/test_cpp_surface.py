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
