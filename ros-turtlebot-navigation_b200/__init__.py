"""ros-turtlebot-navigation_b200: the MPPI rollout and RBPF particle update of
bostoncleek/ROS-Turtlebot-Navigation as sm_100a CUDA kernels behind the reference's class surfaces.

The directory name carries hyphens (the project's name); import it through the repo-root helper
`_pkg.load()` or with importlib under the module name `ros_turtlebot_navigation_b200`.
"""
from . import _capi, rigid2d, controller, synthetic  # noqa: F401
from ._capi import B2NError, load_library  # noqa: F401
from .controller import CartModel, LossFunc, MPPI, comm_unique_id  # noqa: F401
from .rigid2d import Pose, WheelVelocities, Twist2D, Transform2D, Vector2D  # noqa: F401

try:  # the RBPF mirror appears with rbpf_api.cu
    from . import bmapping  # noqa: F401
    from .bmapping import LaserProperties, GridMapper, ScanAlignment, ParticleFilter  # noqa: F401
except ImportError:  # pragma: no cover
    pass
