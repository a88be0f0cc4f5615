"""instance_nerf_b200: B200 (sm_100a) implementation of Instance-NeRF's
instance-field render / train hot path behind the reference's own API.

    from instance_nerf_b200 import raymarching            # raymarching/raymarching.py
    from instance_nerf_b200.gridencoder import GridEncoder
    from instance_nerf_b200.shencoder import SHEncoder
    from instance_nerf_b200.nerf.network_mask import NeRFNetwork
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
