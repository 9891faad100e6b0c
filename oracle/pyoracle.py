"""ctypes front-ends of the TEST ORACLE.

* `Oracle`  -- oracle/liboracle.so, the C restatement (restir_oracle.c).  Travels to the GPU box.
* `RefLib`  -- oracle/_ref/libromis_ref.so, the reference's own translation units compiled with the
               harness shims (oracle/Makefile `ref`).  Buildable only where /root/reference exists; the
               built library travels with the snapshot.

Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product (romis_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from romis_b200 import abi
from romis_b200.scene import Camera, Features, RmisParams, Scene, LIGHT_DTYPE, VERTEX_DTYPE, Mesh

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libromis_ref.so")
DROPIN_SO = os.path.join(HERE, "_ref", "libromis_dropin.so")
REFERENCE_ROOT = "/root/reference"


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in ("restir_oracle.c", "tracer.c", "tracer.h", "Makefile")]
    srcs += [os.path.join(HERE, "..", "include", f) for f in ("romis_rng.h", "romis_detmath.h", "romis_gpu.h")]
    if force or not os.path.exists(ORACLE_SO) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    return ORACLE_SO


def build_ref() -> str | None:
    """Builds oracle/_ref when the reference tree is present; returns the path or None."""
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "-j8"])
        if os.path.exists(os.path.join(HERE, "..", "romis_b200", "libromis_gpu.so")):
            # the reference's translation units + integration/render_restir_gpu.cpp on top of libromis_gpu.so
            subprocess.check_call(["make", "-s", "-C", HERE, "dropin", "-j8"])
    return REF_SO if os.path.exists(REF_SO) else None


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class ReservoirState:
    """Per-stage reservoir arrays, [N, H, W(, 3)]."""

    def __init__(self, N, H, W, with_id=True):
        self.light_id = np.full((N, H, W), 0xFFFFFFFF, np.uint32) if with_id else None
        self.u = np.zeros((N, H, W), np.float32) if with_id else None
        self.v = np.zeros((N, H, W), np.float32) if with_id else None
        self.W = np.zeros((N, H, W), np.float32)
        self.M = np.zeros((N, H, W), np.uint32)
        self.position = np.zeros((N, H, W, 3), np.float32)
        self.color = np.zeros((N, H, W, 3), np.float32)
        self.wSum = np.zeros((N, H, W), np.float32)
        self.chosenW = np.zeros((N, H, W), np.float32)

    def as_romis_dump(self) -> abi.romis_reservoir_dump:
        d = abi.romis_reservoir_dump()
        d.light_id = _p(self.light_id, C.c_uint32); d.u = _p(self.u, C.c_float); d.v = _p(self.v, C.c_float)
        d.W = _p(self.W, C.c_float); d.M = _p(self.M, C.c_uint32)
        d.position = _p(self.position, C.c_float); d.color = _p(self.color, C.c_float)
        return d


class GBuffer:
    def __init__(self, H, W):
        self.t = np.zeros((H, W), np.float32)
        self.normal = np.zeros((H, W, 3), np.float32)
        self.texcoord = np.zeros((H, W, 2), np.float32)
        self.mesh = np.zeros((H, W), np.uint32)

    def as_romis_dump(self) -> abi.romis_gbuffer_dump:
        d = abi.romis_gbuffer_dump()
        d.t = _p(self.t, C.c_float); d.normal = _p(self.normal, C.c_float)
        d.texcoord = _p(self.texcoord, C.c_float); d.mesh = _p(self.mesh, C.c_uint32)
        return d


# ------------------------------------------------------------------------------------------------
class Oracle:
    """C restatement of renderReSTIR (oracle/restir_oracle.c)."""

    def __init__(self, tracer_mode: int = 1):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_create.restype = C.c_void_p
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_tracer_mode.argtypes = [C.c_void_p, C.c_int]
        L.orc_upload_scene.argtypes = [C.c_void_p, C.POINTER(abi.romis_mesh_desc), C.c_int, C.POINTER(abi.romis_texture), C.c_int]
        L.orc_upload_lights.argtypes = [C.c_void_p, C.POINTER(abi.romis_light), C.c_int]
        L.orc_reset_history.argtypes = [C.c_void_p]
        L.orc_render_frame.argtypes = [C.c_void_p, C.POINTER(abi.romis_features), C.POINTER(abi.romis_camera), C.c_int, C.c_int,
                                       C.c_int, C.POINTER(abi.romis_rng), C.c_void_p]
        L.orc_download_reservoirs.argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.romis_reservoir_dump), C.c_void_p, C.c_void_p]
        L.orc_download_gbuffer.argtypes = [C.c_void_p, C.POINTER(abi.romis_gbuffer_dump)]
        L.orc_trace_rays.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 5
        L.orc_ray_dirs.argtypes = [C.POINTER(abi.romis_camera), C.c_int, C.c_int, C.c_void_p]
        L.orc_powf.restype = C.c_float; L.orc_powf.argtypes = [C.c_float, C.c_float]
        L.orc_expf.restype = C.c_float; L.orc_expf.argtypes = [C.c_float]
        L.orc_rng_bits.restype = C.c_uint32
        L.orc_rng_bits.argtypes = [C.c_uint64] + [C.c_uint32] * 5
        self.ctx = C.c_void_p(L.orc_create())
        L.orc_set_tracer_mode(self.ctx, tracer_mode)
        self.W = self.H = self.N = 0

    def close(self):
        if self.ctx:
            self.lib.orc_destroy(self.ctx); self.ctx = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"oracle error {rc}: {self.lib.orc_last_error(self.ctx).decode()}")

    def upload_scene(self, scene: Scene):
        descs, nm, texs, nt, keep = scene.to_abi()
        self._check(self.lib.orc_upload_scene(self.ctx, descs, nm, texs, nt))
        self.upload_lights(scene.lights)

    def upload_lights(self, lights: np.ndarray):
        a = np.ascontiguousarray(lights, LIGHT_DTYPE)
        self._check(self.lib.orc_upload_lights(self.ctx, a.ctypes.data_as(C.POINTER(abi.romis_light)), len(a)))

    def reset_history(self):
        self.lib.orc_reset_history(self.ctx)

    def render_frame(self, features: Features, camera: abi.romis_camera, W: int, H: int, history_valid: bool,
                     seed: int, frame: int, want_image: bool = True):
        f = features.to_abi(); r = abi.romis_rng(seed, frame, 0)
        out = np.zeros((H, W, 3), np.float32) if want_image else None
        self._check(self.lib.orc_render_frame(self.ctx, C.byref(f), C.byref(camera), W, H, int(history_valid), C.byref(r),
                                              out.ctypes.data if want_image else None))
        self.W, self.H, self.N = W, H, features.numSamplesInReservoir
        return out

    def render_frame_rmis(self, features: Features, rmis: RmisParams, camera: abi.romis_camera, W: int, H: int, seed: int, frame: int):
        """renderRMIS restated (oracle/restir_oracle.c orc_render_frame_rmis).  Returns (image, neighbour xy, counts)."""
        f = features.to_abi(); rp = rmis.to_abi(); r = abi.romis_rng(seed, frame, 0)
        K1 = features.numNeighboursToSample + 1
        img = np.zeros((H, W, 3), np.float32); xy = np.full((H, W, K1, 2), -1, np.int32); cnt = np.zeros((H, W), np.uint32)
        self.lib.orc_render_frame_rmis.argtypes = [C.c_void_p, C.POINTER(abi.romis_features), C.POINTER(abi.romis_rmis_params),
                                                   C.POINTER(abi.romis_camera), C.c_int, C.c_int, C.POINTER(abi.romis_rng),
                                                   C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(self.lib.orc_render_frame_rmis(self.ctx, C.byref(f), C.byref(rp), C.byref(camera), W, H, C.byref(r),
                                                   img.ctypes.data, xy.ctypes.data, cnt.ctypes.data))
        self.W, self.H, self.N = W, H, features.numSamplesInReservoir
        return img, xy, cnt

    def render_frame_romis(self, features: Features, rmis: RmisParams, camera: abi.romis_camera, W: int, H: int, seed: int, frame: int):
        """renderROMIS restated (oracle/restir_oracle.c orc_render_frame_romis).  Returns (image, technique matrices
        [H, W, k+1, k+1], contribution vectors [H, W, 3, k+1]) -- the latter two as they stand after the last iteration."""
        f = features.to_abi(); rp = rmis.to_abi(); r = abi.romis_rng(seed, frame, 0)
        K1 = features.numNeighboursToSample + 1
        img = np.zeros((H, W, 3), np.float32); A = np.zeros((H, W, K1, K1), np.float32); B = np.zeros((H, W, 3, K1), np.float32)
        self.lib.orc_render_frame_romis.argtypes = [C.c_void_p, C.POINTER(abi.romis_features), C.POINTER(abi.romis_rmis_params),
                                                    C.POINTER(abi.romis_camera), C.c_int, C.c_int, C.POINTER(abi.romis_rng),
                                                    C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(self.lib.orc_render_frame_romis(self.ctx, C.byref(f), C.byref(rp), C.byref(camera), W, H, C.byref(r),
                                                    img.ctypes.data, A.ctypes.data, B.ctypes.data))
        self.W, self.H, self.N = W, H, features.numSamplesInReservoir
        return img, A, B

    def cod_solve(self, A: np.ndarray, b: np.ndarray):
        """One system through include/romis_cod.h: (x, rank)."""
        A = np.ascontiguousarray(A, np.float32); b = np.ascontiguousarray(b, np.float32); n = len(b)
        x = np.zeros(n, np.float32); rank = C.c_int()
        self.lib.orc_cod_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        self._check(self.lib.orc_cod_solve(A.ctypes.data, b.ctypes.data, n, x.ctypes.data, C.byref(rank)))
        return x, rank.value

    def reservoirs(self, pass_id: int) -> ReservoirState:
        st = ReservoirState(self.N, self.H, self.W)
        d = st.as_romis_dump()
        self._check(self.lib.orc_download_reservoirs(self.ctx, pass_id, C.byref(d), st.wSum.ctypes.data, st.chosenW.ctypes.data))
        return st

    def gbuffer(self) -> GBuffer:
        g = GBuffer(self.H, self.W); d = g.as_romis_dump()
        self._check(self.lib.orc_download_gbuffer(self.ctx, C.byref(d)))
        return g

    def ray_dirs(self, camera: abi.romis_camera, W, H):
        d = np.zeros((H, W, 3), np.float32)
        self.lib.orc_ray_dirs(C.byref(camera), W, H, d.ctypes.data)
        return d

    def trace_rays(self, origins, dirs, tfar, any_hit=False):
        n = len(tfar)
        o = np.ascontiguousarray(origins, np.float32); d = np.ascontiguousarray(dirs, np.float32); tf = np.ascontiguousarray(tfar, np.float32)
        hit = np.zeros(n, np.uint8); t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
        tri = np.full(n, 0xFFFFFFFF, np.uint32)
        self._check(self.lib.orc_trace_rays(self.ctx, o.ctypes.data, d.ctypes.data, tf.ctypes.data, n, int(any_hit),
                                            hit.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data, tri.ctypes.data))
        return hit, t, u, v, tri


# ------------------------------------------------------------------------------------------------
class _ref_camera_desc(C.Structure):
    _fields_ = [("fov_deg", C.c_float), ("distance", C.c_float), ("look_at", abi.f3), ("rotation_deg", abi.f3)]


class _ref_reservoir_dump(C.Structure):
    _fields_ = [("position", C.POINTER(C.c_float)), ("color", C.POINTER(C.c_float)), ("W", C.POINTER(C.c_float)),
                ("M", C.POINTER(C.c_uint64)), ("wSum", C.POINTER(C.c_float)), ("chosenW", C.POINTER(C.c_float))]


class _ref_frame_dump(C.Structure):
    _fields_ = [("gbuffer_t", C.POINTER(C.c_float)), ("gbuffer_normal", C.POINTER(C.c_float)),
                ("gbuffer_texcoord", C.POINTER(C.c_float)), ("gbuffer_mesh", C.POINTER(C.c_uint32)),
                ("ray_dir", C.POINTER(C.c_float)), ("ray_origin", C.POINTER(C.c_float)),
                ("initial", C.POINTER(_ref_reservoir_dump)), ("temporal", C.POINTER(_ref_reservoir_dump)),
                ("spatial", C.POINTER(_ref_reservoir_dump) * 8), ("final_", C.POINTER(_ref_reservoir_dump))]


class ref_timings(C.Structure):
    _fields_ = [(n, C.c_double) for n in "primary_ms initial_ms temporal_ms spatial_ms shade_ms total_ms grid_copy_ms".split()]


REF_FLAG_WHOLE_FRAME, REF_FLAG_TIMING_RNG, REF_FLAG_SPLIT_SPATIAL, REF_FLAG_ASIS_RNG = 1, 2, 4, 8
# SceneType of the reference (src/scene/scene.h:18-26)
SCENE_TYPES = {"SingleTriangle": 0, "Cube": 1, "CubeTextured": 2, "CornellBox": 3,
               "CornellBoxParallelogramLight": 4, "CornellNightClub": 5, "Monkey": 6}


class RefFrame:
    pass


class RefLib:
    """The compiled reference (oracle/_ref/libromis_ref.so).  Process-global state (one scene)."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built (run `make -C oracle ref dropin` where /root/reference exists)")
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_last_error.restype = C.c_char_p
        L.ref_render_frame.argtypes = [C.POINTER(abi.romis_features), C.POINTER(_ref_camera_desc), C.c_int, C.c_int, C.c_int,
                                       C.POINTER(abi.romis_rng), C.c_int, C.POINTER(_ref_frame_dump), C.c_void_p,
                                       C.POINTER(ref_timings)]
        L.ref_set_scene.argtypes = [C.POINTER(abi.romis_mesh_desc), C.c_int, C.POINTER(abi.romis_texture), C.c_int]
        L.ref_set_lights.argtypes = [C.POINTER(abi.romis_light), C.c_int]
        L.ref_make_camera.argtypes = [C.POINTER(_ref_camera_desc), C.c_int, C.c_int, C.POINTER(abi.romis_camera)]

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"reference harness error {rc}: {self.lib.ref_last_error().decode()}")

    @staticmethod
    def _cam(camera: Camera) -> _ref_camera_desc:
        return _ref_camera_desc(camera.fov_deg, camera.distance, abi.f3(*camera.look_at), abi.f3(*camera.rotation_deg))

    def num_threads(self) -> int:
        return int(self.lib.ref_num_threads())

    def set_tracer_mode(self, mode: int):
        self.lib.ref_set_tracer_mode(mode)

    def load_prebuilt(self, scene_type, data_dir=os.path.join(REFERENCE_ROOT, "data")):
        """loadScenePrebuilt (reference src/scene/scene.cpp:68-132)."""
        st = SCENE_TYPES[scene_type] if isinstance(scene_type, str) else int(scene_type)
        self._check(self.lib.ref_load_prebuilt(st, (data_dir.rstrip("/") + "/").encode()))

    def set_scene(self, scene: Scene):
        descs, nm, texs, nt, keep = scene.to_abi()
        self._check(self.lib.ref_set_scene(descs, nm, texs, nt))
        self.set_lights(scene.lights)

    def set_lights(self, lights: np.ndarray):
        a = np.ascontiguousarray(lights, LIGHT_DTYPE)
        self._check(self.lib.ref_set_lights(a.ctypes.data_as(C.POINTER(abi.romis_light)), len(a)))

    def export_scene(self, name="") -> Scene:
        L = self.lib
        nm, nt, nl = C.c_int(), C.c_int(), C.c_int()
        L.ref_scene_info(C.byref(nm), C.byref(nt), C.byref(nl))
        s = Scene(name=name)
        for i in range(nm.value):
            nv, ntri, mat = C.c_uint32(), C.c_uint32(), abi.romis_material()
            L.ref_mesh_info(i, C.byref(nv), C.byref(ntri), C.byref(mat))
            v = np.zeros(nv.value, VERTEX_DTYPE); t = np.zeros((ntri.value, 3), np.uint32)
            L.ref_mesh_data(i, v.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
            s.meshes.append(Mesh(v, t, tuple(mat.kd), tuple(mat.ks), mat.shininess, mat.transparency, mat.kd_texture))
        for i in range(nt.value):
            w, h = C.c_int(), C.c_int()
            L.ref_texture_info(i, C.byref(w), C.byref(h))
            px = np.zeros((h.value, w.value, 3), np.float32)
            L.ref_texture_data(i, px.ctypes.data_as(C.c_void_p))
            s.textures.append(px)
        lights = np.zeros(nl.value, LIGHT_DTYPE)
        if nl.value:
            L.ref_lights_data(lights.ctypes.data_as(C.c_void_p))
        s.lights = lights
        return s

    def make_camera(self, camera: Camera, W: int, H: int) -> abi.romis_camera:
        out = abi.romis_camera(); cd = self._cam(camera)
        self._check(self.lib.ref_make_camera(C.byref(cd), W, H, C.byref(out)))
        return out

    def reset_history(self):
        self.lib.ref_reset_history()

    def render_frame_rmis(self, features: Features, rmis: RmisParams, camera: Camera, W: int, H: int, seed: int, frame: int,
                          want_neighbours: bool = True):
        """renderRMIS (reference src/rendering/render.cpp:64-119), called whole.  Returns (image, neighbour xy, counts)."""
        f = features.to_abi(); rp = rmis.to_abi(); r = abi.romis_rng(seed, frame, 0); cd = self._cam(camera)
        K1 = features.numNeighboursToSample + 1
        img = np.zeros((H, W, 3), np.float32)
        xy = np.full((H, W, K1, 2), -1, np.int32) if want_neighbours else None
        cnt = np.zeros((H, W), np.uint32) if want_neighbours else None
        self.lib.ref_render_frame_rmis.argtypes = [C.POINTER(abi.romis_features), C.POINTER(abi.romis_rmis_params), C.POINTER(_ref_camera_desc),
                                                   C.c_int, C.c_int, C.POINTER(abi.romis_rng), C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(self.lib.ref_render_frame_rmis(C.byref(f), C.byref(rp), C.byref(cd), W, H, C.byref(r), img.ctypes.data,
                                                   xy.ctypes.data if want_neighbours else None, cnt.ctypes.data if want_neighbours else None))
        return img, xy, cnt

    def set_mis_timing(self, on: bool):
        """R-MIS / R-OMIS calls run for timing: thread-safe non-parity random stream, OpenMP on all host cores."""
        self.lib.ref_set_mis_timing(int(on))

    def render_frame_romis(self, features: Features, rmis: RmisParams, camera: Camera, W: int, H: int, seed: int, frame: int,
                           capture: bool = True):
        """renderROMIS (reference src/rendering/render.cpp:121-265), called whole.  Returns (image, technique matrices,
        contribution vectors); the latter two come through the visualiseAlphas hook of the harness (None without capture)."""
        f = features.to_abi(); rp = rmis.to_abi(); r = abi.romis_rng(seed, frame, 0); cd = self._cam(camera)
        K1 = features.numNeighboursToSample + 1
        img = np.zeros((H, W, 3), np.float32)
        A = np.zeros((H, W, K1, K1), np.float32) if capture else None
        B = np.zeros((H, W, 3, K1), np.float32) if capture else None
        self.lib.ref_render_frame_romis.argtypes = [C.POINTER(abi.romis_features), C.POINTER(abi.romis_rmis_params), C.POINTER(_ref_camera_desc),
                                                    C.c_int, C.c_int, C.POINTER(abi.romis_rng), C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(self.lib.ref_render_frame_romis(C.byref(f), C.byref(rp), C.byref(cd), W, H, C.byref(r), img.ctypes.data,
                                                    A.ctypes.data if capture else None, B.ctypes.data if capture else None))
        return img, A, B

    def render_frame(self, features: Features, camera: Camera, W: int, H: int, history_valid: bool, seed: int, frame: int,
                     flags: int = 0, dump: bool = True, want_image: bool = True) -> RefFrame:
        N = features.numSamplesInReservoir
        f = features.to_abi(); r = abi.romis_rng(seed, frame, 0); cd = self._cam(camera)
        res = RefFrame()
        res.image = np.zeros((H, W, 3), np.float32) if want_image else None
        res.timings = ref_timings()
        fd = None; keep = []
        if dump:
            fd = _ref_frame_dump()
            res.gbuffer = GBuffer(H, W)
            res.ray_dir = np.zeros((H, W, 3), np.float32); res.ray_origin = np.zeros((H, W, 3), np.float32)
            fd.gbuffer_t = _p(res.gbuffer.t, C.c_float); fd.gbuffer_normal = _p(res.gbuffer.normal, C.c_float)
            fd.gbuffer_texcoord = _p(res.gbuffer.texcoord, C.c_float); fd.gbuffer_mesh = _p(res.gbuffer.mesh, C.c_uint32)
            fd.ray_dir = _p(res.ray_dir, C.c_float); fd.ray_origin = _p(res.ray_origin, C.c_float)
            res.stages = {}

            def mk(name):
                st = ReservoirState(N, H, W, with_id=False)
                st.M64 = np.zeros((N, H, W), np.uint64)
                d = _ref_reservoir_dump(_p(st.position, C.c_float), _p(st.color, C.c_float), _p(st.W, C.c_float),
                                        _p(st.M64, C.c_uint64), _p(st.wSum, C.c_float), _p(st.chosenW, C.c_float))
                keep.append(d); res.stages[name] = st
                return C.pointer(d)
            fd.initial = mk(abi.ROMIS_PASS_INITIAL); fd.temporal = mk(abi.ROMIS_PASS_TEMPORAL); fd.final_ = mk(abi.ROMIS_PASS_FINAL)
            if flags & REF_FLAG_SPLIT_SPATIAL:
                for p in range(min(8, features.spatialResamplingPasses)):
                    fd.spatial[p] = mk(abi.ROMIS_PASS_SPATIAL0 + p)
        self._check(self.lib.ref_render_frame(C.byref(f), C.byref(cd), W, H, int(history_valid), C.byref(r), flags,
                                              C.byref(fd) if fd is not None else None,
                                              res.image.ctypes.data if want_image else None, C.byref(res.timings)))
        if dump:
            for st in res.stages.values():
                st.M = st.M64.astype(np.uint32)
        return res


class DropinLib(RefLib):
    """oracle/_ref/libromis_dropin.so: the reference's translation units plus integration/render_restir_gpu.cpp (the
    replacement body of renderReSTIR) linked against libromis_gpu.so.  render_frame_gpu drives the reference's own Scene /
    Trackball / Screen / Features objects through the GPU path."""

    def __init__(self):
        super().__init__(DROPIN_SO)
        self.lib.ref_render_frame_dropin.argtypes = [C.POINTER(abi.romis_features), C.POINTER(_ref_camera_desc), C.c_int, C.c_int,
                                                     C.c_int, C.POINTER(abi.romis_rng), C.c_void_p]

    def render_frame_gpu(self, features: Features, camera: Camera, W: int, H: int, history_valid: bool, seed: int, frame: int):
        f = features.to_abi(); r = abi.romis_rng(seed, frame, 0); cd = self._cam(camera)
        img = np.zeros((H, W, 3), np.float32)
        self._check(self.lib.ref_render_frame_dropin(C.byref(f), C.byref(cd), W, H, int(history_valid), C.byref(r), img.ctypes.data))
        return img

    def render_frame_mis_gpu(self, romis: bool, features: Features, rmis: RmisParams, camera: Camera, W: int, H: int, seed: int, frame: int):
        """renderRMIS_gpu / renderROMIS_gpu (integration/render_restir_gpu.cpp) on the reference's own objects."""
        f = features.to_abi(); rp = rmis.to_abi(); r = abi.romis_rng(seed, frame, 0); cd = self._cam(camera)
        img = np.zeros((H, W, 3), np.float32)
        self.lib.ref_render_frame_mis_dropin.argtypes = [C.c_int, C.POINTER(abi.romis_features), C.POINTER(abi.romis_rmis_params),
                                                         C.POINTER(_ref_camera_desc), C.c_int, C.c_int, C.POINTER(abi.romis_rng), C.c_void_p]
        self._check(self.lib.ref_render_frame_mis_dropin(int(romis), C.byref(f), C.byref(rp), C.byref(cd), W, H, C.byref(r), img.ctypes.data))
        return img
