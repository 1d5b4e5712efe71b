"""Prints the resident CTAs per SM the driver grants the fp32 restraint kernel (development aid)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trx2dyn
from trx2dyn import capi
lib = capi.lib()
if hasattr(lib, "trx_debug_k1_occupancy"):
    n = C.c_int()
    capi.check(lib.trx_debug_k1_occupancy(C.byref(n)))
    print("restraints_kernel<float>: resident CTAs/SM =", n.value)
