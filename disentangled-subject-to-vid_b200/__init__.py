"""disentangled-subject-to-vid_b200 — B200-native (sm_100a) implementation of the CogVideoX subject-to-video
denoising hot path of carpedkm/disentangled-subject-to-vid, behind the reference's own pipeline / module surface.

The directory name is not a Python identifier; import it as `s2v_b200` (the alias module at the repo root) or with
importlib.import_module("disentangled-subject-to-vid_b200").
"""
from . import _lib, ops  # noqa: F401

__all__ = ["_lib", "ops"]
