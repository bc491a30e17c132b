"""routeformer_b200: B200-native (sm_100a) implementation of Routeformer's batched forward/backward hot path.

Python API mirrors the reference (`from routeformer import Routeformer`, `from routeformer.models import RouteformerConfig`);
everything underneath is hand-written CUDA behind the C ABI of include/routeformer_b200.h.
"""
from .backbone import PatchEmbedBackbone, VideoBackboneModule  # noqa: F401
from .config import (BaseConfig, GPSBackboneConfig, PatchBackboneConfig, RouteformerConfig,  # noqa: F401
                     VideoBackboneConfig)
from .experiment import ParallelTrainerSteps  # noqa: F401
from .informer import Informer  # noqa: F401
from .layers import PerceiveDecoder, PerceiveEncoder  # noqa: F401
from .metrics import FutureDiscountedLoss, ade, ade_fde_per_sample, fde  # noqa: F401
from .routeformer import Routeformer  # noqa: F401

__all__ = [
    "Routeformer", "RouteformerConfig", "GPSBackboneConfig", "VideoBackboneConfig", "PatchBackboneConfig", "BaseConfig",
    "VideoBackboneModule", "PatchEmbedBackbone", "Informer", "PerceiveEncoder", "PerceiveDecoder",
    "ade", "fde", "ade_fde_per_sample", "FutureDiscountedLoss", "ParallelTrainerSteps",
]
