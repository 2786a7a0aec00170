"""Frontend name handling (the TF-free part of the reference `models/frontend.py:24-53`)."""

from __future__ import annotations

import warnings

VALID_FRONTENDS = ("librosa", "hybrid", "raw", "mfcc", "log_mel")
_ALIASES = {"precomputed": "librosa", "tf": "raw"}


def normalize_frontend_name(name: str) -> str:
    """Canonical frontend name; deprecated aliases warn, unknown names raise ValueError."""
    if name in VALID_FRONTENDS:
        return name
    if name in _ALIASES:
        warnings.warn(f"Frontend name '{name}' is deprecated, use '{_ALIASES[name]}' instead.", DeprecationWarning, stacklevel=2)
        return _ALIASES[name]
    raise ValueError(f"Invalid audio frontend: '{name}'. Valid options: {VALID_FRONTENDS}")
