from .resources import Resources, auto_sync_resources  # noqa: F401
from .._capi import CuvsError  # noqa: F401
