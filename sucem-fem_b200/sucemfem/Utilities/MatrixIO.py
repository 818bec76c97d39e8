"""Save and load finite element matrices as .mat files (mirror of /root/reference/sucemfem/Utilities/MatrixIO.py:29-85)."""
import os

import numpy as np


def check_path(path, create=True):
    """Check to see if a path exists, and create it if desired"""
    if not os.path.exists(path):
        if not create:
            return True
        os.makedirs(path)
    return True


def save_scipy_matrix_as_mat(path, name, matrix):
    """Save a scipy sparse matrix (or anything with to_scipy(), e.g. the device matrices) as <path>/<name>.mat"""
    import scipy.io
    if not check_path(path):
        return False
    if hasattr(matrix, 'to_scipy'):
        matrix = matrix.to_scipy()
    scipy.io.savemat(os.path.join(path, name) + '.mat', {name: matrix}, oned_as='column')
    return True


def load_scipy_matrix_from_mat(path, name):
    """Load a scipy sparse matrix from <path>/<name>.mat (None if the file does not exist)"""
    import scipy.io
    filename = os.path.join(path, name)
    if not os.path.exists(filename + '.mat'):
        return None
    data = scipy.io.loadmat(filename + '.mat')[name]
    if type(data) is np.ndarray:
        matrix = data
    else:
        matrix = data.tocsr()
    return matrix
