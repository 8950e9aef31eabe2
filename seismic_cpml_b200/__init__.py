"""cpml-b200: B200-native implementation of the SEISMIC_CPML velocity-stress C-PML time loop.

    lib       ctypes binding of the C ABI (include/cpml_b200.h, libcpml_b200.so)
    programs  host-side mirror of the reference programs (parameters, set-up, driver loop)
    slab      z-slab decomposition across GPUs (torch.distributed point-to-point)
    build     nvcc build of libcpml_b200.so for sm_100a
"""
__version__ = "0.1.0"
