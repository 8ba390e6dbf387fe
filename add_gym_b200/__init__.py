"""add_gym_b200 -- B200-native (sm_100a) implementation of rsamf/add-gym's rollout + update hot path.

Public names mirror the reference's plugin API.  Importing this package is cheap and works
without a GPU; constructing any hot-path object needs libaddk.so and a CUDA device.
"""
from .config import default_config  # noqa: F401
from .engine import (BaseEngine, BaseEntity, BaseJoint, BaseLink, BaseScene, SyntheticEngine)  # noqa: F401


def __getattr__(name):
    # heavy modules on demand
    import importlib
    table = {
        "ADDAgent": ".add_agent", "AgentMode": ".add_agent", "ADDModel": ".add_model", "ADDMotion": ".add_motion",
        "AdaptiveSegmentSampler": ".add_motion", "ADDObservation": ".add_observation", "ADDReward": ".add_observation",
        "ADDDone": ".add_observation", "DoneFlags": ".add_observation", "MotionLib": ".motion_lib",
        "ExperienceBuffer": ".experience_buffer", "Normalizer": ".normalizer", "DiffNormalizer": ".normalizer",
        "ImitationEnvironment": ".env", "Environment": ".env", "Manipulator": ".env", "KinCharModel": ".kinematics",
    }
    if name in table:
        return getattr(importlib.import_module(table[name], __name__), name)
    raise AttributeError(name)
