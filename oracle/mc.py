"""ctypes wrapper of oracle/mc_oracle.c (CPU marching-cubes checker; test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmc_oracle.so")


def _load():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "mc_oracle.c")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    lib = C.CDLL(_LIB)
    lib.mc_oracle.restype = C.c_int
    lib.mc_oracle.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    return lib


def marching_cubes(volume, iso, x_off=0.0):
    """-> (verts (V,3) f32 index coords, faces (F,3) i32, normals (V,3) f32)."""
    lib = _load()
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    nx, ny, nz = vol.shape
    nv, nt = C.c_int64(0), C.c_int64(0)
    assert lib.mc_oracle(vol.ctypes.data, nx, ny, nz, float(iso), float(x_off), None, None, None, C.byref(nv), C.byref(nt)) == 0
    verts = np.empty((nv.value, 3), np.float32)
    normals = np.empty((nv.value, 3), np.float32)
    faces = np.empty((nt.value, 3), np.int32)
    assert lib.mc_oracle(vol.ctypes.data, nx, ny, nz, float(iso), float(x_off), verts.ctypes.data, normals.ctypes.data,
                         faces.ctypes.data, C.byref(nv), C.byref(nt)) == 0
    return verts, faces, normals
