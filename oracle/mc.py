"""ctypes wrapper of oracle/mc_oracle.c (CPU marching-cubes checker; test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmc_oracle.so")
_lib = None

# stats slot = lewiner_case * 16 + (number of joined ambiguous faces) * 2 + tunnel  ->  Lewiner's sub-case name
SUBCASE = {(3, 0, 0): "3.1", (3, 1, 0): "3.2", (4, 0, 0): "4.1", (4, 0, 1): "4.2", (6, 0, 0): "6.1.1", (6, 0, 1): "6.1.2 (c)",
           (6, 1, 0): "6.2", (7, 0, 0): "7.1", (7, 1, 0): "7.2", (7, 2, 0): "7.3 (c)", (7, 3, 0): "7.4.1", (7, 3, 1): "7.4.2",
           (10, 0, 0): "10.1.1", (10, 0, 1): "10.1.2", (10, 1, 0): "10.2 (c)", (10, 2, 0): "10.1.1_",
           (12, 0, 0): "12.1.1", (12, 0, 1): "12.1.2", (12, 1, 0): "12.2 (c)", (12, 2, 0): "12.1.1_",
           (13, 0, 0): "13.1", (13, 1, 0): "13.2", (13, 2, 0): "13.3 (c) [or the impossible 2-matching]",
           (13, 3, 0): "13.4 (c) / 13.5.1", (13, 3, 1): "13.5.2", (13, 4, 0): "13.3_ (c)", (13, 5, 0): "13.2_", (13, 6, 0): "13.1_"}


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "mc_oracle.c")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    lib = C.CDLL(_LIB)
    lib.mc_oracle.restype = C.c_int
    lib.mc_oracle.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]
    lib.mc_cell_variant.restype = C.c_int
    lib.mc_cell_variant.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 9
    _lib = lib
    return lib


def marching_cubes(volume, iso, x_off=0, g_nx=None, own=None, v_base=0, stats=False):
    """-> (verts (V,3) f32 index coords, faces (F,3) i32, normals (V,3) f32[, stats dict]).
    volume: the buffer planes; x_off: global index of buffer plane 0; g_nx: planes of the global grid (default: the buffer is
    the whole grid); own = (p_lo, p_hi): buffer planes whose points this call owns (default: all)."""
    lib = _load()
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    nb, ny, nz = vol.shape
    if g_nx is None:                 # a stand-alone volume: x_off only shifts the axis-0 coordinates of the vertices
        g_x0, g_nx, x_shift = 0, nb, int(x_off)
    else:
        g_x0, g_nx, x_shift = int(x_off), int(g_nx), 0
    p_lo, p_hi = (0, nb) if own is None else own
    nv, nt = C.c_int64(0), C.c_int64(0)
    st = np.zeros(256, np.int64)
    args = (vol.ctypes.data, nb, ny, nz, float(iso), g_x0, g_nx, p_lo, p_hi, x_shift, int(v_base))
    rc = lib.mc_oracle(*args, None, None, None, C.byref(nv), C.byref(nt), st.ctypes.data)
    assert rc == 0, f"mc_oracle: error {rc}"
    verts = np.empty((nv.value, 3), np.float32)
    normals = np.empty((nv.value, 3), np.float32)
    faces = np.empty((nt.value, 3), np.int32)
    rc = lib.mc_oracle(*args, verts.ctypes.data, normals.ctypes.data, faces.ctypes.data, C.byref(nv), C.byref(nt), None)
    assert rc == 0, f"mc_oracle: error {rc}"
    if stats:
        named = {}
        for slot in np.nonzero(st)[0]:
            key = (int(slot) // 16, (int(slot) % 16) // 2, int(slot) % 2)
            named[SUBCASE.get(key, f"case {key[0]}")] = named.get(SUBCASE.get(key, f"case {key[0]}"), 0) + int(st[slot])
        return verts, faces, normals, named
    return verts, faces, normals


def cell_variant(m, J, tunnel):
    """The oracle's procedural triangulation of (mask, face decisions J, tunnel) and its run-time test spec (tests only)."""
    lib = _load()
    ints = [C.c_int(0) for _ in range(7)]
    tris = (C.c_ubyte * 36)()
    faces = (C.c_int * 6)()
    ntri, uses_c, itest, tif, nf, mu, case = ints
    lib.mc_cell_variant(m, J, tunnel, C.byref(ntri), C.byref(uses_c), tris, C.byref(itest), C.byref(tif), C.byref(nf), faces,
                        C.byref(mu), C.byref(case))
    return dict(ntri=ntri.value, uses_c=uses_c.value, tris=list(tris)[:3 * ntri.value], itest=itest.value, tunnel_if_I=tif.value,
                nf=nf.value, faces=list(faces)[:nf.value], mu_pos=mu.value, case=case.value)
