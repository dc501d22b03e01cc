"""alp_b200 — B200-native ALP / ALP_RD column codec (host-side mirror of the reference's primitive API)."""
from . import _abi  # noqa: F401
