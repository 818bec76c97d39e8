"""ctypes binding of libsucem_b200.so (C ABI: include/sucem_b200.h).

The library is the product: there is no CPU fallback.  Importing this module fails loudly when the shared
library has not been built (``make -C sucem-fem_b200`` or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SUCEM_B200_LIB') or os.path.join(os.path.dirname(_HERE), 'lib', 'libsucem_b200.so')

FORM_MASS, FORM_STIFFNESS, FORM_ABC = 0, 1, 2
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_CANONICAL, KERNEL_ROWS = 0, 1, 2, 3
MATS_LAST_IS_ABC_COMPACT = 0x100


class NativeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libsucem_b200.so not built: run `make -C sucem-fem_b200` (expected %s)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    P = C.POINTER
    sig = {
        'sfem_last_error': (C.c_char_p, []),
        'sfem_version': (C.c_char_p, []),
        'sfem_release_cached_memory': (i32, []),
        'sfem_reserve_cached_memory': (i32, [C.c_int64, vp]),
        'sfem_cached_memory_info': (i32, [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        'sfem_mesh_create': (i32, [i32, i64, vp, i64, vp, vp, P(vp)]),
        'sfem_mesh_destroy': (i32, [vp]),
        'sfem_fill_unit_mesh': (i32, [i32, i32, i32, P(dbl), P(dbl), vp, vp, vp]),
        'sfem_space_create': (i32, [vp, i32, vp, P(vp)]),
        'sfem_space_destroy': (i32, [vp]),
        'sfem_space_info': (i32, [vp, P(i64), P(i64), P(i64), P(i32)]),
        'sfem_space_cell_dofs': (i32, [vp, vp, vp]),
        'sfem_space_cell_entities': (i32, [vp, vp, vp, vp]),
        'sfem_space_num_facets': (i32, [vp, P(i64)]),
        'sfem_mark_exterior_facets': (i32, [vp, vp, C.c_uint32, vp]),
        'sfem_pec_row_flags': (i32, [vp, vp, C.c_uint32, vp, vp]),
        'sfem_expand_material': (i32, [i64, vp, i32, P(C.c_uint32), P(dbl), vp, P(C.c_uint32), vp]),
        'sfem_space_set_global_ids': (i32, [vp, vp, vp, i64, i64, i64, i64, i64, i64, vp]),
        'sfem_pattern_create': (i32, [vp, i64, i64, vp, P(vp)]),
        'sfem_pattern_destroy': (i32, [vp]),
        'sfem_pattern_info': (i32, [vp, P(i64), P(i64), P(i64)]),
        'sfem_pattern_export': (i32, [vp, vp, vp, vp]),
        'sfem_partition_rows': (i32, [vp, i32, P(i64), vp]),
        'sfem_partition_rows_by_class': (i32, [vp, i32, P(i64), P(i64), vp]),
        'sfem_assemble': (i32, [vp, i32, vp, vp, vp, vp, i32, vp]),
        'sfem_assemble_ST': (i32, [vp, vp, vp, vp, vp, vp, i32, vp]),
        'sfem_apply_ident': (i32, [vp, vp, vp, vp]),
        'sfem_launches_per_assemble_ST': (i32, [vp, i32]),
        'sfem_matvec': (i32, [vp, i32, P(vp), P(dbl), i32, vp, vp, vp]),
        'sfem_diagonal': (i32, [vp, i32, P(vp), P(dbl), vp, vp]),
        'sfem_bicgstab': (i32, [vp, i32, P(vp), P(dbl), vp, vp, i32, dbl, i32, P(i32), P(dbl), vp]),
        'sfem_host_job_create': (i32, [i32, i64, vp, i64, vp, i32, vp, vp, i32, vp, C.c_uint32, vp, P(i64), P(i64), P(vp)]),
        'sfem_host_job_run': (i32, [vp, vp, vp, vp, vp, vp, i32]),
        'sfem_host_job_destroy': (i32, [vp]),
        'sfem_host_job_space': (i32, [vp, P(vp)]),
        'sfem_host_job_run_rows': (i32, [vp, i64, i64, vp, vp, vp, vp, vp, i32]),
        'sfem_space_rows_nnz': (i32, [vp, i64, i64, vp, P(i64)]),
        'sfem_space_assemble_rows_to_host': (i32, [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
        'sfem_space_count_nnz': (i32, [vp, vp, P(i64)]),
        'sfem_space_assemble_to_host': (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
        'sfem_measure_fp64_peak': (i32, [P(dbl), vp]),
        'sfem_abc_info': (i32, [vp, vp, P(i64), P(i64)]),
        'sfem_abc_export': (i32, [vp, vp, vp, vp, vp]),
        'sfem_assemble_abc': (i32, [vp, vp, vp, vp, vp]),
        'sfem_host_alloc': (i32, [C.c_size_t, P(vp)]),
        'sfem_host_free': (i32, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)      # AttributeError here = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTED = _load()


def check(rc):
    if rc != 0:
        raise NativeError(lib.sfem_last_error().decode())


def version():
    return lib.sfem_version().decode()
