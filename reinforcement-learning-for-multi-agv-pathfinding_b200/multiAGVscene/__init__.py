"""Mirror of the reference's src/multiAGVscene package (Layout, Explorer, Scene)."""
from .Layout import Layout
from .Explorer import Explorer
from .Scene import Scene

__all__ = ["Layout", "Explorer", "Scene"]
