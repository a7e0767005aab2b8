from . import brute_force  # noqa: F401
from . import filters  # noqa: F401
