from .network import MLP, NeRFNetwork  # noqa: F401
from .renderer import NeRFRenderer, default_opt, near_far_from_aabb  # noqa: F401
