"""mcpower_b200 — B200-native backend for MCPower's per-simulation hot loop (engine_core::run_batch)."""
from . import spec  # noqa: F401
