from . import brute_force  # noqa: F401
from . import ivf_flat  # noqa: F401
from . import ivf_pq  # noqa: F401
from . import cagra  # noqa: F401
from . import filters  # noqa: F401
from .refine import refine  # noqa: F401
