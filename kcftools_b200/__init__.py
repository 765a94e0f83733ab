"""kcftools_b200 — B200-native `getVariations` hot path of kcftools behind a C ABI (include/kcf_b200.h)."""
from .api import Context, KMC, Plan, KcfError, fixed_windows  # noqa: F401

__version__ = "0.1.0"
