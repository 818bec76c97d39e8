"""Files that live next to a module (mirror of /root/reference/sucemfem/Testing/Paths.py:22-41)."""
import os


def get_module_path(modfile):
    """Directory of the module whose __file__ is passed in."""
    return os.path.dirname(modfile)


def get_module_path_filename(filename, __file__):
    return os.path.join(get_module_path(__file__), filename)


def get_module_path_file(filename, __file__, mode='r'):
    return open(get_module_path_filename(filename, __file__), mode)
