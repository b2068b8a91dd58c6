from .ba import BA, BA_host, neighbors, reproject, last_status, last_plan_hits  # noqa: F401  (same re-exports as cdvslam/fastba/__init__.py)
from .batched import BA_batched, linearize_debug  # noqa: F401
