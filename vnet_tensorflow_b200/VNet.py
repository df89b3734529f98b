"""Name-compatible module for the reference's legacy `VNet.py`: `VNet.VNet(...).network_fn(x)`."""
from .networks import LegacyVNet as VNet  # noqa: F401
