"""Drop-in alias: ``import mantaray`` resolves to the B200 path.

The reference package exports exactly these two names
(python/mantaray/__init__.py:1-3).
"""

from mantaray_b200.core import ray_tracing, single_ray

__all__ = ["single_ray", "ray_tracing"]
