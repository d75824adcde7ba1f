"""nyxus_b200 -- B200-native (sm_100a) implementation of Nyxus's per-ROI texture + intensity feature path."""
from .nyxus import Nyxus, Nyxus3D
from .engine import Engine, NyxusB200Error

__all__ = ["Nyxus", "Nyxus3D", "Engine", "NyxusB200Error"]
__version__ = "0.1.0"
