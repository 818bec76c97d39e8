"""Matrix converters (mirror of /root/reference/sucemfem/Utilities/Converters.py:27-52)."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Evan Lezar; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as np


def dolfin_ublassparse_to_scipy_csr(A, dtype=None, imagify=False):
    """Convert an assembled matrix (anything with .data() -> (rowptr, cols, values) and .size(0)) to
    scipy.sparse.csr_matrix; this is the CSR hand-off of the path (Converters.py:37-42)."""
    import scipy.sparse
    (row, col, data) = A.data()
    col = np.intc(col)
    row = np.intc(row) if row[-1] < 2 ** 31 else row
    n = A.size(0)
    if imagify:
        data = data * 1j
    return scipy.sparse.csr_matrix((data, col, row), shape=(n, n), dtype=dtype)
