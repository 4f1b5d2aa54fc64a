"""ctypes binding of libbas_probe.so (include/bas_probe.h): measurement kernels only.  The product
package never loads this library; bench.py and the scripts in tools/ do."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'binaural-audio-synthesis_b200', 'libbas_probe.so')


def load():
    lib = C.CDLL(LIB_PATH)
    vp, i = C.c_void_p, C.c_int
    for name, argtypes in {'bas_probe_clock': [i, i, i, i, vp, vp, vp], 'bas_probe_block': [i, i, i, vp, vp],
                           'bas_probe_fma': [i, i, i, i, vp, vp], 'bas_probe_stamp': [vp, vp],
                           'bas_probe_tc_render': [vp, C.c_longlong, vp, i, i, vp, C.c_longlong, C.c_longlong, i, i, vp, vp]}.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = argtypes, i
    return lib
