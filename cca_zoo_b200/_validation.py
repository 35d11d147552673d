"""Input validation with the reference's semantics and error messages
(cca_zoo/_utils/_validation.py:14-75), extended to accept torch tensors (CPU or CUDA)."""
from __future__ import annotations

from typing import TypeVar

import numpy as np
import torch
from sklearn.utils.validation import check_array

_T = TypeVar("_T")


#: host arrays above this many elements skip sklearn's finiteness scan (a full extra pass over the data on one
#: core); BaseModel checks the moment buffer on the device instead -- NaN/inf anywhere poisons it
_HOST_SCAN_LIMIT = 1 << 22


def _check_host_array(v):
    size = int(np.prod(np.shape(v))) if hasattr(v, "shape") else 0
    if size <= _HOST_SCAN_LIMIT:
        return check_array(v, ensure_2d=True, allow_nd=False, dtype="numeric")
    try:
        return check_array(v, ensure_2d=True, allow_nd=False, dtype="numeric", ensure_all_finite=False)
    except TypeError:  # scikit-learn < 1.6 spells it force_all_finite
        return check_array(v, ensure_2d=True, allow_nd=False, dtype="numeric", force_all_finite=False)


def validate_views(views, min_views: int = 2):
    """Return a list of 2-D float arrays / tensors with equal row counts.

    numpy / array-like inputs go through ``sklearn.utils.validation.check_array`` exactly as in the
    reference (NaN/inf rejected, dtype kept); torch tensors are checked for shape and finiteness
    and kept where they are (no host round trip for CUDA tensors).
    """
    if len(views) < min_views:
        raise ValueError(f"At least {min_views} views are required, got {len(views)}.")
    processed = []
    for v in views:
        if isinstance(v, torch.Tensor):
            if v.dim() != 2:
                raise ValueError(f"Expected 2D array, got {v.dim()}D tensor instead.")
            if not v.dtype.is_floating_point:
                v = v.to(torch.float64)
            if v.numel() == 0:
                raise ValueError("Found array with 0 sample(s) or 0 feature(s).")
            # finiteness of torch inputs is checked on the device after the copy (see
            # BaseModel._fit_device): a host-side scan of a large pinned tensor would dominate fit()
            processed.append(v)
        else:
            processed.append(_check_host_array(v))
    n = processed[0].shape[0]
    if not all(v.shape[0] == n for v in processed):
        raise ValueError(
            "All views must have the same number of samples. "
            f"Got shapes: {[tuple(v.shape) for v in processed]}."
        )
    return processed


def perview_parameter(name: str, value, default, n_views: int):
    """One value per view from a scalar, a per-view list or ``None`` (same contract and error text as
    cca_zoo/_utils/_validation.py:45-75): ``None`` -> the default everywhere, a list must have exactly
    ``n_views`` entries and is returned as is, anything else is repeated."""
    if isinstance(value, list):
        if len(value) == n_views:
            return value
        raise ValueError(
            f"Parameter '{name}' must be a scalar or a list of length {n_views}, got length {len(value)}."
        )
    return [default if value is None else value for _ in range(n_views)]
