"""Drop-in module name for nerf/mask_renderer.py of the reference."""
from .renderer import NeRFMaskRenderer, NeRFRenderer, sample_pdf  # noqa: F401
