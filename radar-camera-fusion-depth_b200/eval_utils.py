"""Depth metrics with the reference's names and formulas (reference: src/eval_utils.py:17-85); NumPy, host side
(observability: not on the accelerated path)."""
import numpy as np


def root_mean_sq_err(src, tgt):
    """sqrt(mean((tgt - src)^2))  (reference :17-29)"""
    return np.sqrt(np.mean((tgt - src) ** 2))


def mean_abs_err(src, tgt):
    """mean(|tgt - src|)  (reference :31-43)"""
    return np.mean(np.abs(tgt - src))


def inv_root_mean_sq_err(src, tgt):
    """sqrt(mean((1/tgt - 1/src)^2))  (reference :45-57)"""
    return np.sqrt(np.mean(((1.0 / tgt) - (1.0 / src)) ** 2))


def inv_mean_abs_err(src, tgt):
    """mean(|1/tgt - 1/src|)  (reference :59-71)"""
    return np.mean(np.abs((1.0 / tgt) - (1.0 / src)))


def mean_abs_rel_err(src, tgt):
    """mean(|src - tgt| / tgt)  (reference :73-85)"""
    return np.mean(np.abs(src - tgt) / tgt)
