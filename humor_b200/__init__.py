"""humor_b200 — B200-native (sm_100a) implementation of HuMoR's Stage-III test-time-optimisation
hot path behind the reference's own Python surfaces (BodyModel / HumorModel / MotionOptimizer).
The compute lives in csrc/*.cu behind the C-ABI of include/humor_b200.h; there is no CPU fallback."""
from . import _ext  # noqa: F401


def __getattr__(name):
    # lazy: importing the package must not require the built library (build() imports it first)
    if name == 'BodyModel':
        from .body_model import BodyModel
        return BodyModel
    if name == 'HumorModel':
        from .humor_model import HumorModel
        return HumorModel
    if name == 'MotionOptimizer':
        from .motion_optimizer import MotionOptimizer
        return MotionOptimizer
    if name == 'FittingLoss':
        from .fitting_loss import FittingLoss
        return FittingLoss
    raise AttributeError(name)
