"""The plain value types the two hot paths borrow from the reference's rigid2d package
(rigid2d/include/rigid2d/diff_drive.hpp:16-33, rigid2d/include/rigid2d/rigid2d.hpp:143-170).
Field names and order are the reference's (note Pose is theta, x, y)."""
from dataclasses import dataclass
import math


@dataclass
class Pose:
    theta: float = 0.0
    x: float = 0.0
    y: float = 0.0


@dataclass
class WheelVelocities:
    ul: float = 0.0
    ur: float = 0.0


@dataclass
class Twist2D:
    w: float = 0.0
    vx: float = 0.0
    vy: float = 0.0


@dataclass
class Vector2D:
    x: float = 0.0
    y: float = 0.0


class Transform2D:
    """Only what the call sites of ParticleFilter use: construction from (Vector2D, radians) and
    displacement() (rigid2d/src/rigid2d/rigid2d.cpp:151-158,233-241)."""

    def __init__(self, trans=None, radians=0.0):
        self.theta = float(radians)
        self.x = float(trans.x) if trans is not None else 0.0
        self.y = float(trans.y) if trans is not None else 0.0

    def displacement(self):
        return self.theta, self.x, self.y

    def __call__(self, v):
        c, s = math.cos(self.theta), math.sin(self.theta)
        return Vector2D(c * v.x - s * v.y + self.x, s * v.x + c * v.y + self.y)
