"""Agent mirrors (reference: agents/liftingDWT_agent.py, agents/compression_agent.py)."""
from .compression_agent import CompressionAgent  # noqa: F401
from .liftingDWT_agent import LiftingBasedDWTAgent  # noqa: F401
