"""ctypes binding of oracle/_build/liboracle.so (the C++ restatement of the reference's CPU algorithm).

TEST INFRASTRUCTURE: used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
The struct definitions are shared with the product's ctypes mirror because the oracle consumes the
same flat LrSceneDesc the C ABI does (include/lumilly.h); no product code imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from lumillyrender_b200.capi import LrCamera, LrMaterial, LrRenderParams, LrSceneDesc, LrSky

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_PF = C.POINTER(C.c_float)
_PI = C.POINTER(C.c_int32)


class OrcStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("samples", C.c_uint64), ("nodes_visited", C.c_uint64), ("prims_tested", C.c_uint64),
                ("nonfinite_samples", C.c_uint64), ("build_seconds", C.c_double), ("render_seconds", C.c_double),
                ("threads", C.c_int32), ("bvh_nodes", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("oracle.cpp", "oracle.h", "Makefile")] + [os.path.join(HERE, "..", "include", "lumilly.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.run(["make", "-C", HERE, "-B", "_build/liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_scene_create.argtypes = [C.POINTER(LrSceneDesc), C.POINTER(C.c_void_p)]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_nodes.argtypes = [C.c_void_p]
        L.orc_render.argtypes = [C.c_void_p, C.POINTER(LrRenderParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _PF, _PF, C.POINTER(OrcStats)]
        L.orc_render_aov.argtypes = [C.c_void_p, C.POINTER(LrRenderParams), C.c_int, C.c_int, C.c_int, _PF]
        L.orc_trace_path.argtypes = [C.c_void_p, C.POINTER(LrRenderParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _PF, _PF, _PI, _PF, _PI]
        L.orc_set_math_mode.argtypes = [C.c_int]
        L.orc_spec_sincos.argtypes = [C.c_float, _PF, _PF]
        L.orc_trace_primary.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _PI, _PF]
        L.orc_trace_rays.argtypes = [C.c_void_p, C.c_int64, _PF, _PF, C.c_int, C.c_int, _PI, _PF, _PF]
        L.orc_camera_sample.argtypes = [C.POINTER(LrCamera), C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _PF]
        for n in ("orc_triangle_intersect_mt", "orc_triangle_intersect_3c"):
            getattr(L, n).argtypes = [_PF, _PF, _PF, _PF, _PF, _PF]
        L.orc_sphere_intersect.argtypes = [_PF, C.c_float, _PF, _PF, _PF, _PF, _PF]
        L.orc_aabb_is_intersect.argtypes = [_PF, _PF, _PF, _PF]
        L.orc_reflect.argtypes = [_PF, _PF, _PF]
        L.orc_refract.argtypes = [_PF, _PF, C.c_float, _PF]
        L.orc_orthonormal_basis.argtypes = [_PF, _PF, _PF]
        L.orc_checker.argtypes = [C.c_float, C.c_float]
        L.orc_checker.restype = C.c_float
        L.orc_material_brdf.argtypes = [C.POINTER(LrMaterial), _PF, _PF, _PF, _PF, _PF]
        L.orc_material_sample.argtypes = [C.POINTER(LrMaterial), _PF, _PF, C.c_float, C.c_float, _PF, _PF]
        L.orc_material_weight.argtypes = [C.POINTER(LrMaterial)]
        L.orc_material_weight.restype = C.c_float
        L.orc_material_coef.argtypes = [C.POINTER(LrMaterial), _PF, _PF, C.c_float, _PF]
        L.orc_fresnel.argtypes = [C.c_float, C.c_float, _PF, _PF, _PF]
        L.orc_fresnel.restype = C.c_float
        L.orc_ior_pair.argtypes = [C.POINTER(LrMaterial), _PF, _PF, _PF, _PF]
        L.orc_sky_radiance.argtypes = [C.POINTER(LrSky), _PF, _PF]
        L.orc_rng_float.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_rng_float.restype = C.c_float
        L.orc_matrix_unit.argtypes = [_PF]
        L.orc_matrix_translate.argtypes = [_PF, _PF]
        L.orc_matrix_scale.argtypes = [_PF, _PF]
        L.orc_matrix_axis_angle.argtypes = [_PF, C.c_float, _PF]
        L.orc_matrix_look_at.argtypes = [_PF, _PF, _PF, _PF]
        L.orc_matrix_mul.argtypes = [_PF, _PF, _PF]
        L.orc_matrix_apply.argtypes = [_PF, _PF, _PF]
        L.orc_camera_ideal_pinhole.argtypes = [_PF, C.c_float, C.c_int, C.c_int, C.POINTER(LrCamera)]
        L.orc_camera_thin_lens.argtypes = [_PF, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(LrCamera)]
        L.orc_camera_omnidirectional.argtypes = [_PF, C.c_int, C.c_int, C.POINTER(LrCamera)]
        L.orc_camera_pinhole.argtypes = [_PF, _PF, _PF, C.c_int, C.c_int, C.c_float, C.POINTER(LrCamera)]
        L.orc_triangle_area.argtypes = [_PF]
        L.orc_triangle_area.restype = C.c_float
        L.orc_sphere_area.argtypes = [C.c_float]
        L.orc_sphere_area.restype = C.c_float
        _lib = L
    return _lib


def f3(v):
    return (C.c_float * len(v))(*[float(x) for x in v])


def fp(a):
    return a.ctypes.data_as(_PF)


class OracleScene:
    """The reference-algorithm scene built from the same flat LrSceneDesc the GPU path consumes."""

    def __init__(self, desc_ptr, keepalive=None):
        self._L = lib()
        self._s = C.c_void_p()
        self._keep = keepalive
        rc = self._L.orc_scene_create(desc_ptr, C.byref(self._s))
        if rc != 0:
            raise RuntimeError("orc_scene_create failed: %d" % rc)
        cam = desc_ptr.contents.camera
        self.width, self.height = cam.width, cam.height

    def render(self, params, traversal=0, rng_mode=0, math_mode=1, threads=0, pixel_stride=1, sumsq=True):
        """Returns (per-pixel SUM image, sumsq, stats).  traversal 0 = faithful reference algorithm;
        math_mode 1 = the specified fp32 sincos shared with the device, 0 = libm like the reference."""
        h = params.crop_h if params.crop_w > 0 else self.height
        w = params.crop_w if params.crop_w > 0 else self.width
        out = np.zeros((h, w, 3), dtype=np.float32)
        sq = np.zeros((h, w, 3), dtype=np.float32) if sumsq else None
        st = OrcStats()
        rc = self._L.orc_render(self._s, C.byref(params), traversal, rng_mode, math_mode, threads, pixel_stride, fp(out),
                                fp(sq) if sumsq else None, C.byref(st))
        if rc != 0:
            raise RuntimeError("orc_render failed: %d" % rc)
        return out, sq, st.as_dict()

    def render_aov(self, params, kind, traversal=0, threads=0):
        """Scene::normal ("normal": HxWx3) / Scene::depth ("depth": HxW) averaged over the camera rays of the sample range."""
        h = params.crop_h if params.crop_w > 0 else self.height
        w = params.crop_w if params.crop_w > 0 else self.width
        k = {"normal": 0, "depth": 1}[kind]
        out = np.zeros((h, w, 3) if k == 0 else (h, w), dtype=np.float32)
        rc = self._L.orc_render_aov(self._s, C.byref(params), k, traversal, threads, fp(out))
        if rc != 0:
            raise RuntimeError("orc_render_aov failed: %d" % rc)
        return out

    def trace_path(self, params, x, y, sample, traversal=0, max_rays=4096):
        """(origins, directions, prim, t) of every closest-hit query one sample of pixel (x, y) issues, in order."""
        o = np.zeros((max_rays, 3), np.float32); d = np.zeros((max_rays, 3), np.float32)
        prim = np.zeros(max_rays, np.int32); t = np.zeros(max_rays, np.float32)
        n = C.c_int32()
        rc = self._L.orc_trace_path(self._s, C.byref(params), x, y, sample, traversal, max_rays, fp(o), fp(d), prim.ctypes.data_as(_PI), fp(t), C.byref(n))
        if rc != 0:
            raise RuntimeError("orc_trace_path failed: %d" % rc)
        k = min(n.value, max_rays)
        return o[:k], d[:k], prim[:k], t[:k]

    def trace_primary(self, u=0.5, v=0.5, ua=0.5, va=0.5, traversal=0, threads=0):
        prim = np.empty((self.height, self.width), dtype=np.int32)
        t = np.empty((self.height, self.width), dtype=np.float32)
        self._L.orc_trace_primary(self._s, u, v, ua, va, traversal, threads, prim.ctypes.data_as(_PI), fp(t))
        return prim, t

    def trace_rays(self, origins, directions, traversal=0, brute_force=False):
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        n = o.shape[0]
        prim = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float32)
        nrm = np.empty((n, 3), dtype=np.float32)
        self._L.orc_trace_rays(self._s, n, fp(o), fp(d), traversal, 1 if brute_force else 0, prim.ctypes.data_as(_PI), fp(t), fp(nrm))
        return prim, t, nrm

    @property
    def bvh_nodes(self):
        return self._L.orc_scene_nodes(self._s)

    def close(self):
        if self._s:
            self._L.orc_scene_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
