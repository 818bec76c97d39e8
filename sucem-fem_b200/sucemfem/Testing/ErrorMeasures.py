"""Error measures of the examples (mirror of /root/reference/sucemfem/Testing/ErrorMeasures.py:22-60).  Both return the
SUM of squared relative errors times 100, as the reference does (no mean, no root) -- its drivers print these numbers."""
import numpy as np


def _summed_squares(actual, desired, extra_error, scale):
    err = np.abs(np.asarray(actual) - np.asarray(desired))
    extra = np.zeros_like(err) if extra_error is None else np.abs(np.asarray(extra_error))
    return 100 * np.sum((err / scale) ** 2 + (extra / scale) ** 2)


def normalised_RMS(actual, desired, extra_error=None):
    """Errors normalised entry by entry with |actual|; extra_error is an optional orthogonal error component."""
    return _summed_squares(actual, desired, extra_error, np.abs(np.asarray(actual)))


def max_normalised_RMS(actual, desired, extra_error=None):
    """Errors normalised with max|actual|, so that entries close to zero do not dominate."""
    return _summed_squares(actual, desired, extra_error, np.abs(np.asarray(actual)).max())
