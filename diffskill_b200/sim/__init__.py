"""Drop-in mirrors of the reference's plb.engine classes over the CUDA engine."""
from .function import GradModel  # noqa: F401
from .losses import Loss  # noqa: F401
from .mlp import MLP  # noqa: F401
from .mpm_simulator import MPMSimulator  # noqa: F401
from .primitives import (Box, Capsule, Chopsticks, Cylinder, Gripper, Gripper2, Knife, Primitive, Primitives,  # noqa: F401
                         RollingPin, RollingPinExt, Sphere, Torus)
from .taichi_env import TaichiEnv  # noqa: F401
