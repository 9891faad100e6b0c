"""Host-side scene, light, camera and parameter types of the ReSTIR path.

These mirror what the reference hands to `renderReSTIR` (reference src/rendering/render.h:25-28):
`Scene{meshes, lights}` (src/scene/scene.h:28-33), `Trackball` camera (framework/include/framework/
trackball.h), `Features` (src/utils/common.h:89-136).  They only marshal data into the C-ABI structs of
include/romis_gpu.h; no rendering arithmetic lives here.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import abi

LIGHT_DTYPE = np.dtype([("type", "<u4"), ("p0", "<f4", 3), ("e1", "<f4", 3), ("e2", "<f4", 3),
                        ("c0", "<f4", 3), ("c1", "<f4", 3), ("c2", "<f4", 3), ("c3", "<f4", 3)])
assert LIGHT_DTYPE.itemsize == C.sizeof(abi.romis_light)
VERTEX_DTYPE = np.dtype([("position", "<f4", 3), ("normal", "<f4", 3), ("texcoord", "<f4", 2)])
assert VERTEX_DTYPE.itemsize == C.sizeof(abi.romis_vertex)


@dataclass
class Features:
    """Hot fields of the reference's `Features` with the reference's defaults (common.h:89-136).

    The reference's default `rayTraceMode` is ROMIS (common.h:103); this path is the ReSTIR mode.
    """
    enableShading: bool = True
    enableTextureMapping: bool = True
    initialSamplesVisibilityCheck: bool = False
    numSamplesInReservoir: int = 2
    initialLightSamples: int = 32
    numNeighboursToSample: int = 5
    spatialResampleRadius: int = 10
    unbiasedCombination: bool = False
    spatialReuse: bool = True
    spatialReuseVisibilityCheck: bool = False
    temporalReuse: bool = True
    spatialResamplingPasses: int = 2
    temporalClampM: int = 20
    enableToneMapping: bool = True
    gamma: float = 1.0
    exposure: float = 1.5

    def to_abi(self) -> abi.romis_features:
        f = abi.romis_features()
        for fld in dataclasses.fields(self):
            v = getattr(self, fld.name)
            setattr(f, fld.name, float(v) if fld.name in ("gamma", "exposure") else int(v))
        return f


@dataclass
class RmisParams:
    """The fields of `Features` only the R-MIS mode reads, with the reference's defaults (common.h:110-121)."""
    maxIterationsMIS: int = 5
    misWeightRMIS: int = abi.ROMIS_MIS_EQUAL
    neighbourSelectionStrategy: int = abi.ROMIS_NEIGHBOURS_SIMILAR
    neighbourSameGeometry: bool = True
    neighbourMaxDepthDifferenceFraction: float = 0.10
    neighbourMaxNormalAngleDifferenceRadians: float = 0.436332
    useProgressiveROMIS: bool = False           # R-OMIS only (common.h:119-120)
    progressiveUpdateMod: int = 1

    def to_abi(self) -> abi.romis_rmis_params:
        return abi.romis_rmis_params(int(self.maxIterationsMIS), int(self.misWeightRMIS), int(self.neighbourSelectionStrategy),
                                     int(self.neighbourSameGeometry), float(self.neighbourMaxDepthDifferenceFraction),
                                     float(self.neighbourMaxNormalAngleDifferenceRadians), int(self.useProgressiveROMIS),
                                     int(self.progressiveUpdateMod))


_CAMERA_CACHE = {}


@dataclass
class Camera:
    """`CameraConfig` of the reference (src/utils/config.h:21-26); defaults = the nightclub view."""
    fov_deg: float = 30.0
    distance: float = 25.0
    look_at: tuple = (2.57, 1.23, -1.35)
    rotation_deg: tuple = (10.3, 30.0, 0.0)

    def to_abi(self, width: int, height: int) -> abi.romis_camera:
        """origin = Trackball::position() (framework/src/trackball.cpp:75-78), quat = glm::quat(euler)
        (glm type_quat.inl:208-217), half extents per trackball.cpp:26-27; all in fp32.
        The numpy scalar arithmetic below costs ~60 us, which sits in front of a frame's first kernel launch (the reference's
        Trackball computes these once per camera change, not per frame): results are kept per (camera, resolution)."""
        key = (float(self.fov_deg), float(self.distance), tuple(float(t) for t in self.look_at),
               tuple(float(t) for t in self.rotation_deg), int(width), int(height))
        hit = _CAMERA_CACHE.get(key)
        if hit is None:
            if len(_CAMERA_CACHE) >= 4096:
                _CAMERA_CACHE.clear()
            hit = _CAMERA_CACHE[key] = self._to_abi(width, height)
        cam = abi.romis_camera()
        C.memmove(C.byref(cam), C.byref(hit), C.sizeof(cam))       # callers own (and may edit) what they get
        return cam

    def _to_abi(self, width: int, height: int) -> abi.romis_camera:
        f32 = np.float32
        rad = [f32(math.radians(1.0)) * f32(a) for a in self.rotation_deg]   # glm::radians: deg * 0.0174532925...
        c = [f32(np.cos(f32(a) * f32(0.5))) for a in rad]
        s = [f32(np.sin(f32(a) * f32(0.5))) for a in rad]
        w = f32(c[0] * c[1] * c[2] + s[0] * s[1] * s[2])
        x = f32(s[0] * c[1] * c[2] - c[0] * s[1] * s[2])
        y = f32(c[0] * s[1] * c[2] + s[0] * c[1] * s[2])
        z = f32(c[0] * c[1] * s[2] - s[0] * s[1] * c[2])
        q = np.array([x, y, z], f32)
        v = np.array([0, 0, -self.distance], f32)
        uv = np.cross(q, v).astype(f32)
        uuv = np.cross(q, uv).astype(f32)
        pos = np.array(self.look_at, f32) + (v + ((uv * w) + uuv) * f32(2)).astype(f32)
        cam = abi.romis_camera()
        cam.origin = abi.f3(*[float(t) for t in pos])
        cam.quat = abi.f4(float(w), float(x), float(y), float(z))
        hh = f32(np.tan(f32(math.radians(1.0)) * f32(self.fov_deg) / f32(2)))
        aspect = f32(width) / f32(height) if width and height else f32(1)
        cam.half_height = float(hh)
        cam.half_width = float(f32(aspect * hh))
        return cam


@dataclass
class Mesh:
    vertices: np.ndarray        # VERTEX_DTYPE [nv]
    triangles: np.ndarray       # uint32 [nt, 3]
    kd: tuple = (1.0, 1.0, 1.0)
    ks: tuple = (0.0, 0.0, 0.0)
    shininess: float = 1.0
    transparency: float = 1.0
    kd_texture: int = -1


@dataclass
class Scene:
    meshes: List[Mesh] = field(default_factory=list)
    textures: List[np.ndarray] = field(default_factory=list)    # float32 [h, w, 3]
    lights: np.ndarray = field(default_factory=lambda: np.zeros(0, LIGHT_DTYPE))
    name: str = ""

    @property
    def n_triangles(self) -> int:
        return int(sum(len(m.triangles) for m in self.meshes))

    # ---- fixture I/O (.npz written by tests/golden/gen_golden.py from the reference's own loader) ----
    def save(self, path: str) -> None:
        d = {"n_meshes": np.int32(len(self.meshes)), "n_textures": np.int32(len(self.textures)),
             "lights": self.lights, "name": np.array(self.name)}
        for i, m in enumerate(self.meshes):
            d[f"m{i}_v"] = m.vertices
            d[f"m{i}_t"] = m.triangles
            d[f"m{i}_mat"] = np.array([*m.kd, *m.ks, m.shininess, m.transparency], np.float32)
            d[f"m{i}_tex"] = np.int32(m.kd_texture)
        for i, t in enumerate(self.textures):
            d[f"tex{i}"] = t
        np.savez_compressed(path, **d)

    @staticmethod
    def load(path: str) -> "Scene":
        z = np.load(path)
        s = Scene(name=str(z["name"]))
        for i in range(int(z["n_meshes"])):
            mat = z[f"m{i}_mat"]
            s.meshes.append(Mesh(vertices=z[f"m{i}_v"].astype(VERTEX_DTYPE), triangles=z[f"m{i}_t"].astype(np.uint32),
                                 kd=tuple(mat[0:3]), ks=tuple(mat[3:6]), shininess=float(mat[6]),
                                 transparency=float(mat[7]), kd_texture=int(z[f"m{i}_tex"])))
        for i in range(int(z["n_textures"])):
            s.textures.append(z[f"tex{i}"].astype(np.float32))
        s.lights = z["lights"].astype(LIGHT_DTYPE)
        return s

    # ---- marshalling ----
    def to_abi(self):
        """Returns (mesh_desc_array, n_meshes, texture_array, n_textures, keepalive)."""
        keep = []
        descs = (abi.romis_mesh_desc * max(1, len(self.meshes)))()
        for i, m in enumerate(self.meshes):
            v = np.ascontiguousarray(m.vertices, VERTEX_DTYPE)
            t = np.ascontiguousarray(m.triangles, np.uint32)
            keep += [v, t]
            descs[i].vertices = v.ctypes.data_as(C.POINTER(abi.romis_vertex))
            descs[i].n_vertices = len(v)
            descs[i].triangles = t.ctypes.data_as(C.POINTER(C.c_uint32))
            descs[i].n_triangles = len(t)
            mat = descs[i].material
            mat.kd = abi.f3(*[float(x) for x in m.kd]); mat.ks = abi.f3(*[float(x) for x in m.ks])
            mat.shininess = float(m.shininess); mat.transparency = float(m.transparency)
            mat.kd_texture = int(m.kd_texture)
        texs = (abi.romis_texture * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            a = np.ascontiguousarray(t, np.float32)
            keep.append(a)
            texs[i].pixels = a.ctypes.data_as(C.POINTER(C.c_float))
            texs[i].height, texs[i].width = a.shape[0], a.shape[1]
        return descs, len(self.meshes), texs, len(self.textures), keep

    def lights_abi(self):
        a = np.ascontiguousarray(self.lights, LIGHT_DTYPE)
        return a.ctypes.data_as(C.POINTER(abi.romis_light)), len(a), a


def synthetic_lights(n_lights: int, seed: int = 1, intensity: float = 24.0) -> np.ndarray:
    """Seed-fixed many-light set of BASELINE config C3/C5: 50 % PointLight / 50 %
    ParallelogramLight, positions uniform in the shell 1.5 <= |p| <= 3 around the origin, edges uniform
    in [-0.05, 0.05]^3, colours uniform in [0.2, 1]^3 scaled by intensity / n_lights."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lights = np.zeros(n_lights, LIGHT_DTYPE)
    d = rng.normal(size=(n_lights, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rad = rng.uniform(1.5, 3.0, size=(n_lights, 1))
    lights["p0"] = (d * rad).astype(np.float32)
    kind = (np.arange(n_lights) % 2).astype(np.uint32)
    rng.shuffle(kind)
    lights["type"] = np.where(kind == 0, abi.ROMIS_LIGHT_POINT, abi.ROMIS_LIGHT_PARALLELOGRAM)
    par = kind == 1
    lights["e1"][par] = rng.uniform(-0.05, 0.05, size=(int(par.sum()), 3)).astype(np.float32)
    lights["e2"][par] = rng.uniform(-0.05, 0.05, size=(int(par.sum()), 3)).astype(np.float32)
    scale = np.float32(intensity / max(1, n_lights))
    for c in ("c0", "c1", "c2", "c3"):
        col = rng.uniform(0.2, 1.0, size=(n_lights, 3)).astype(np.float32) * scale
        if c != "c0":
            col[~par] = 0
        lights[c] = col
    return lights
