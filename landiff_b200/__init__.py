"""landiff_b200 — B200-native (sm_100a) drop-in for LanDiff's diffusion-stage DiT step.

Importing the package registers the C-ABI compute entry points as `torch.ops.landiff_b200.*` custom ops."""
from . import ops  # noqa: F401  (registers the torch custom ops)
