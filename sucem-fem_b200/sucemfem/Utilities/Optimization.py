"""Form-compiler switches (mirror of /root/reference/sucemfem/Utilities/Optimization.py:24-39).  The reference turns on
FFC's `optimize` / `cpp_optimize` for the generated element kernels; here the element kernels are the hand-written CUDA
ones and there is nothing to switch: the call is kept so that drivers run unchanged, and records what was asked for."""

parameters = {'form_compiler': {'optimize': True, 'cpp_optimize': True}}


def set_dolfin_optimisation(enable=True):
    parameters['form_compiler']['optimize'] = bool(enable)
    parameters['form_compiler']['cpp_optimize'] = bool(enable)
