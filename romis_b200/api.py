"""Python host side of the B200 ReSTIR path: a thin ctypes layer over the C-ABI (include/romis_gpu.h).

`RestirRenderer.render_frame` is the drop-in for the reference's `renderReSTIR`
(reference src/rendering/render.h:25-28): same inputs (scene, camera, Features, previous-frame grid ->
`history_valid`), same outputs (Screen::pixels() layout image, the reservoir grid as device-resident
history).  There is no CPU path: if libromis_gpu.so is missing or CUDA is unusable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi
from .scene import Camera, Features, RmisParams, Scene, LIGHT_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ROMIS_GPU_LIB") or os.path.join(HERE, "libromis_gpu.so")      # override: tuning builds only

EXPORTS = [
    "romis_abi_version", "romis_create", "romis_destroy", "romis_last_error", "romis_upload_scene", "romis_upload_lights",
    "romis_render_frame", "romis_render_frame_device", "romis_reset_history", "romis_synchronize", "romis_set_band",
    "romis_frame_begin", "romis_frame_spatial_pass", "romis_frame_end", "romis_halo_region", "romis_stream",
    "romis_set_capture", "romis_download_reservoirs", "romis_download_gbuffer", "romis_trace_rays",
    "romis_set_stage_timing", "romis_last_frame_timings", "romis_host_alloc", "romis_host_free",
    "romis_row_hit_counts", "romis_band_prepare", "romis_peer_export", "romis_peer_attach", "romis_peer_detach", "romis_peer_error",
    "romis_render_frame_rmis", "romis_download_rmis_neighbours", "romis_specular_cutoff",
    "romis_render_frame_romis", "romis_download_romis_system",
    "romis_upload_lights_range", "romis_set_light_archive_auto", "romis_light_archive_marks", "romis_light_archive_release",
    "romis_light_archive_size", "romis_host_register", "romis_host_unregister", "romis_selftest_division",
]
PEER_BLOB_BYTES = 512


class RomisError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Loads libromis_gpu.so (built in-tree by `python -m romis_b200.build`).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RomisError(f"{LIB_PATH} is missing: build it with `python -m romis_b200.build` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, ci = C.c_void_p, C.c_int
    L.romis_last_error.restype = C.c_char_p; L.romis_last_error.argtypes = [vp]
    L.romis_create.argtypes = [C.POINTER(ci), ci, C.POINTER(vp)]
    L.romis_destroy.argtypes = [vp]; L.romis_destroy.restype = None
    L.romis_upload_scene.argtypes = [vp, C.POINTER(abi.romis_mesh_desc), ci, C.POINTER(abi.romis_texture), ci]
    L.romis_upload_lights.argtypes = [vp, C.POINTER(abi.romis_light), ci]
    L.romis_upload_lights_range.argtypes = [vp, C.POINTER(abi.romis_light), ci, ci, ci]
    L.romis_set_light_archive_auto.argtypes = [vp, ci]
    L.romis_light_archive_marks.argtypes = [vp, vp, ci, C.POINTER(ci)]
    L.romis_light_archive_release.argtypes = [vp, vp, ci]
    L.romis_light_archive_size.argtypes = [vp, C.POINTER(ci), C.POINTER(ci)]
    frame_args = [vp, C.POINTER(abi.romis_features), C.POINTER(abi.romis_camera), ci, ci, ci, C.POINTER(abi.romis_rng)]
    L.romis_render_frame.argtypes = frame_args + [vp]
    L.romis_render_frame_device.argtypes = frame_args + [C.POINTER(vp)]
    L.romis_frame_begin.argtypes = frame_args
    L.romis_render_frame_rmis.argtypes = [vp, C.POINTER(abi.romis_features), C.POINTER(abi.romis_rmis_params), C.POINTER(abi.romis_camera),
                                          ci, ci, C.POINTER(abi.romis_rng), vp]
    L.romis_download_rmis_neighbours.argtypes = [vp, vp, vp]
    L.romis_render_frame_romis.argtypes = L.romis_render_frame_rmis.argtypes
    L.romis_download_romis_system.argtypes = [vp, vp, vp]
    L.romis_frame_spatial_pass.argtypes = [vp, ci]
    L.romis_frame_end.argtypes = [vp, vp]
    L.romis_reset_history.argtypes = [vp]; L.romis_synchronize.argtypes = [vp]
    L.romis_set_band.argtypes = [vp, ci, ci]
    L.romis_halo_region.argtypes = [vp, ci, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.romis_stream.argtypes = [vp, C.POINTER(vp)]
    L.romis_set_capture.argtypes = [vp, ci]; L.romis_set_stage_timing.argtypes = [vp, ci]
    L.romis_download_reservoirs.argtypes = [vp, ci, C.POINTER(abi.romis_reservoir_dump)]
    L.romis_download_gbuffer.argtypes = [vp, C.POINTER(abi.romis_gbuffer_dump)]
    L.romis_trace_rays.argtypes = [vp, vp, vp, vp, ci, ci, vp, vp, vp, vp, vp]
    L.romis_selftest_division.argtypes = [vp, vp, vp, ci, vp, vp]
    L.romis_last_frame_timings.argtypes = [vp, C.POINTER(abi.romis_timings)]
    L.romis_row_hit_counts.argtypes = [vp, C.POINTER(abi.romis_camera), ci, ci, vp]
    L.romis_band_prepare.argtypes = [vp, C.POINTER(abi.romis_features), ci, ci]
    L.romis_peer_export.argtypes = [vp, vp]
    L.romis_peer_attach.argtypes = [vp, vp, vp]
    L.romis_peer_detach.argtypes = [vp]
    L.romis_peer_error.argtypes = [vp, C.POINTER(ci)]
    L.romis_specular_cutoff.restype = C.c_float; L.romis_specular_cutoff.argtypes = [C.c_float]
    L.romis_host_alloc.restype = vp; L.romis_host_alloc.argtypes = [C.c_size_t]
    L.romis_host_free.argtypes = [vp]; L.romis_host_free.restype = None
    L.romis_host_register.argtypes = [vp, C.c_size_t]; L.romis_host_unregister.argtypes = [vp]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class ReservoirState:
    """Per-stage reservoir arrays, [N, H, W(, 3)] (romis_reservoir_dump)."""

    def __init__(self, N, H, W):
        self.light_id = np.full((N, H, W), 0xFFFFFFFF, np.uint32)
        self.u = np.zeros((N, H, W), np.float32)
        self.v = np.zeros((N, H, W), np.float32)
        self.W = np.zeros((N, H, W), np.float32)
        self.M = np.zeros((N, H, W), np.uint32)
        self.position = np.zeros((N, H, W, 3), np.float32)
        self.color = np.zeros((N, H, W, 3), np.float32)

    def as_abi(self) -> abi.romis_reservoir_dump:
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

    def as_abi(self) -> abi.romis_gbuffer_dump:
        d = abi.romis_gbuffer_dump()
        d.t = _p(self.t, C.c_float); d.normal = _p(self.normal, C.c_float)
        d.texcoord = _p(self.texcoord, C.c_float); d.mesh = _p(self.mesh, C.c_uint32)
        return d


class PinnedImage:
    """Page-locked float RGB image in Screen::pixels() layout (romis_host_alloc), exposed as a numpy array."""

    def __init__(self, H, W):
        L = load_library()
        self.nbytes = H * W * 3 * 4
        self.ptr = L.romis_host_alloc(self.nbytes)
        if not self.ptr:
            raise RomisError("romis_host_alloc failed")
        buf = (C.c_float * (H * W * 3)).from_address(self.ptr)
        self.array = np.frombuffer(buf, np.float32).reshape(H, W, 3)

    def free(self):
        if self.ptr:
            self.array = None
            load_library().romis_host_free(self.ptr); self.ptr = None


class SharedImage:
    """ONE float RGB image in Screen::pixels() layout that several processes (one per GPU) fill together: POSIX shared memory,
    page-locked in every process that attaches (romis_host_register), so each band's rows arrive by asynchronous DMA over that
    GPU's own PCIe link and the caller -- the creating process -- ends up with the whole frame in one buffer."""

    def __init__(self, H, W, name=None):
        from multiprocessing import shared_memory
        self.nbytes = H * W * 3 * 4
        self.owner = name is None
        self.shm = shared_memory.SharedMemory(create=True, size=self.nbytes) if self.owner else shared_memory.SharedMemory(name=name)
        self.name = self.shm.name
        self.array = np.ndarray((H, W, 3), np.float32, buffer=self.shm.buf)
        self.ptr = self.array.ctypes.data
        if load_library().romis_host_register(self.ptr, self.nbytes) != 0:
            raise RomisError("romis_host_register failed on the shared image")

    def free(self):
        if self.shm is not None:
            load_library().romis_host_unregister(self.ptr)
            self.array = None
            self.shm.close()
            if self.owner:
                self.shm.unlink()
            self.shm = None


class RestirRenderer:
    """One context = one GPU (optionally one row band of the frame); with a list of devices: one context driving one row band
    per GPU from this single process (romis_create with n_devices > 1)."""

    def __init__(self, device=0):
        self.lib = load_library()
        ctx = C.c_void_p()
        devices = list(device) if isinstance(device, (list, tuple)) else [int(device)]
        dev = (C.c_int * len(devices))(*devices)
        rc = self.lib.romis_create(dev, len(devices), C.byref(ctx))
        if rc != 0:
            raise RomisError(f"romis_create failed ({rc}): {self.lib.romis_last_error(None).decode()}")
        self.ctx = ctx
        self.W = self.H = self.N = 0

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.romis_destroy(self.ctx); self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RomisError(f"romis error {rc}: {self.lib.romis_last_error(self.ctx).decode()}")

    # ---- scene ----
    def upload_scene(self, scene: Scene):
        descs, nm, texs, nt, keep = scene.to_abi()
        self._check(self.lib.romis_upload_scene(self.ctx, descs, nm, texs, nt))
        self.upload_lights(scene.lights)

    def upload_lights(self, lights: np.ndarray, dirty: tuple | None = None):
        """scene.lights as of this frame.  `dirty` = (first, count): only those lights are examined (romis_upload_lights_range)."""
        a = np.ascontiguousarray(lights, LIGHT_DTYPE)
        ptr = a.ctypes.data_as(C.POINTER(abi.romis_light))
        if dirty is None:
            self._check(self.lib.romis_upload_lights(self.ctx, ptr, len(a)))
        else:
            self._check(self.lib.romis_upload_lights_range(self.ctx, ptr, len(a), int(dirty[0]), int(dirty[1])))

    # ---- light archive (history samples of edited lights; see include/romis_gpu.h) ----
    def set_light_archive_auto(self, on: bool):
        self._check(self.lib.romis_set_light_archive_auto(self.ctx, int(on)))

    def light_archive_size(self):
        n, held = C.c_int(), C.c_int()
        self._check(self.lib.romis_light_archive_size(self.ctx, C.byref(n), C.byref(held)))
        return n.value, held.value

    def light_archive_marks(self) -> np.ndarray:
        n, _ = self.light_archive_size()
        m = np.zeros(max(n, 1), np.uint8); got = C.c_int()
        self._check(self.lib.romis_light_archive_marks(self.ctx, m.ctypes.data, len(m), C.byref(got)))
        return m[:got.value]

    def light_archive_release(self, keep: np.ndarray):
        k = np.ascontiguousarray(keep, np.uint8)
        self._check(self.lib.romis_light_archive_release(self.ctx, k.ctypes.data if len(k) else None, len(k)))

    # ---- frame ----
    @staticmethod
    def _cam(camera, W, H) -> abi.romis_camera:
        return camera.to_abi(W, H) if isinstance(camera, Camera) else camera

    def render_frame(self, features: Features, camera, W: int, H: int, history_valid: bool, seed: int, frame: int,
                     out: np.ndarray | None = None, want_image: bool = True):
        """renderReSTIR (reference src/rendering/render.cpp:28-62).  Returns the float RGB image [H, W, 3] in
        Screen::pixels() layout (row 0 = top), or None with want_image=False (image stays on the device)."""
        f = features.to_abi(); cam = self._cam(camera, W, H); r = abi.romis_rng(seed, frame, 0)
        if want_image and out is None:
            out = np.zeros((H, W, 3), np.float32)
        self._check(self.lib.romis_render_frame(self.ctx, C.byref(f), C.byref(cam), W, H, int(history_valid), C.byref(r),
                                                out.ctypes.data if want_image else None))
        self.W, self.H, self.N = W, H, features.numSamplesInReservoir
        return out if want_image else None

    def render_frame_device(self, features: Features, camera, W, H, history_valid, seed, frame) -> int:
        f = features.to_abi(); cam = self._cam(camera, W, H); r = abi.romis_rng(seed, frame, 0)
        dev = C.c_void_p()
        self._check(self.lib.romis_render_frame_device(self.ctx, C.byref(f), C.byref(cam), W, H, int(history_valid), C.byref(r), C.byref(dev)))
        self.W, self.H, self.N = W, H, features.numSamplesInReservoir
        return dev.value

    def render_frame_rmis(self, features: Features, rmis: RmisParams, camera, W: int, H: int, seed: int, frame: int,
                          want_image: bool = True, out=None):
        """renderRMIS (reference src/rendering/render.cpp:64-119): the R-MIS estimator, no temporal state.  Returns the
        float RGB image [H, W, 3] in Screen::pixels() layout, or None with want_image=False (image stays on the device)."""
        f = features.to_abi(); rp = rmis.to_abi(); cam = self._cam(camera, W, H); r = abi.romis_rng(seed, frame, 0)
        if out is None and want_image:
            out = np.zeros((H, W, 3), np.float32)       # with a row band set (set_band) only the band's rows are written
        self._check(self.lib.romis_render_frame_rmis(self.ctx, C.byref(f), C.byref(rp), C.byref(cam), W, H, C.byref(r),
                                                     out.ctypes.data if out is not None else None))
        self.W, self.H, self.N = W, H, features.numSamplesInReservoir
        self._rmis_k1 = features.numNeighboursToSample + 1
        return out

    def render_frame_romis(self, features: Features, rmis: RmisParams, camera, W: int, H: int, seed: int, frame: int,
                           want_image: bool = True, out=None):
        """renderROMIS (reference src/rendering/render.cpp:121-265), direct estimator.  Returns the float RGB image
        [H, W, 3] in Screen::pixels() layout, or None with want_image=False (image stays on the device)."""
        f = features.to_abi(); rp = rmis.to_abi(); cam = self._cam(camera, W, H); r = abi.romis_rng(seed, frame, 0)
        if out is None and want_image:
            out = np.zeros((H, W, 3), np.float32)       # with a row band set (set_band) only the band's rows are written
        self._check(self.lib.romis_render_frame_romis(self.ctx, C.byref(f), C.byref(rp), C.byref(cam), W, H, C.byref(r),
                                                      out.ctypes.data if out is not None else None))
        self.W, self.H, self.N = W, H, features.numSamplesInReservoir
        self._rmis_k1 = features.numNeighboursToSample + 1
        return out

    def romis_system(self):
        """Technique matrices [H, W, k+1, k+1] and contribution vectors [H, W, 3, k+1] of the last R-OMIS frame."""
        K1 = self._rmis_k1
        A = np.zeros((self.H, self.W, K1, K1), np.float32); B = np.zeros((self.H, self.W, 3, K1), np.float32)
        self._check(self.lib.romis_download_romis_system(self.ctx, A.ctypes.data, B.ctypes.data))
        return A, B

    def rmis_neighbours(self):
        """Neighbour grid of the last R-MIS frame: (xy [H, W, k+1, 2] with -1 for unused entries, count [H, W])."""
        xy = np.full((self.H, self.W, self._rmis_k1, 2), -1, np.int32); cnt = np.zeros((self.H, self.W), np.uint32)
        self._check(self.lib.romis_download_rmis_neighbours(self.ctx, xy.ctypes.data, cnt.ctypes.data))
        return xy, cnt

    def reset_history(self):
        self._check(self.lib.romis_reset_history(self.ctx))

    def synchronize(self):
        self._check(self.lib.romis_synchronize(self.ctx))

    # ---- row bands ----
    def set_band(self, y0: int, y1: int):
        self._check(self.lib.romis_set_band(self.ctx, y0, y1))

    def frame_begin(self, features: Features, camera, W, H, history_valid, seed, frame):
        f = features.to_abi(); cam = self._cam(camera, W, H); r = abi.romis_rng(seed, frame, 0)
        self._check(self.lib.romis_frame_begin(self.ctx, C.byref(f), C.byref(cam), W, H, int(history_valid), C.byref(r)))
        self.W, self.H, self.N = W, H, features.numSamplesInReservoir

    def frame_spatial_pass(self, p: int):
        self._check(self.lib.romis_frame_spatial_pass(self.ctx, p))

    def frame_end(self, out: np.ndarray | None):
        self._check(self.lib.romis_frame_end(self.ctx, out.ctypes.data if out is not None else None))

    def halo_region(self, which: int):
        ptr = C.c_void_p(); n = C.c_size_t()
        self._check(self.lib.romis_halo_region(self.ctx, which, C.byref(ptr), C.byref(n)))
        return ptr.value or 0, n.value

    def row_hit_counts(self, camera, W: int, H: int) -> np.ndarray:
        cam = self._cam(camera, W, H)
        out = np.zeros(H, np.uint32)
        self._check(self.lib.romis_row_hit_counts(self.ctx, C.byref(cam), W, H, out.ctypes.data))
        return out

    # ---- peer-mapped halos (romis_peer_*) ----
    def band_prepare(self, features: Features, W: int, H: int):
        f = features.to_abi()
        self._check(self.lib.romis_band_prepare(self.ctx, C.byref(f), W, H))

    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(PEER_BLOB_BYTES)
        self._check(self.lib.romis_peer_export(self.ctx, buf))
        return bytes(buf.raw)

    def peer_attach(self, low: bytes | None, high: bytes | None):
        lo = C.create_string_buffer(low, PEER_BLOB_BYTES) if low else None
        hi = C.create_string_buffer(high, PEER_BLOB_BYTES) if high else None
        self._check(self.lib.romis_peer_attach(self.ctx, lo, hi))

    def peer_detach(self):
        self._check(self.lib.romis_peer_detach(self.ctx))

    def peer_timed_out(self) -> bool:
        e = C.c_int()
        self._check(self.lib.romis_peer_error(self.ctx, C.byref(e)))
        return bool(e.value)

    def stream(self) -> int:
        s = C.c_void_p()
        self._check(self.lib.romis_stream(self.ctx, C.byref(s)))
        return s.value or 0

    # ---- parity / measurement ----
    def set_capture(self, on: bool):
        self._check(self.lib.romis_set_capture(self.ctx, int(on)))

    def set_stage_timing(self, on: bool):
        self._check(self.lib.romis_set_stage_timing(self.ctx, int(on)))

    def reservoirs(self, pass_id: int = abi.ROMIS_PASS_FINAL) -> ReservoirState:
        st = ReservoirState(self.N, self.H, self.W); d = st.as_abi()
        self._check(self.lib.romis_download_reservoirs(self.ctx, pass_id, C.byref(d)))
        return st

    def gbuffer(self) -> GBuffer:
        g = GBuffer(self.H, self.W); d = g.as_abi()
        self._check(self.lib.romis_download_gbuffer(self.ctx, C.byref(d)))
        return g

    def timings(self) -> abi.romis_timings:
        t = abi.romis_timings()
        self._check(self.lib.romis_last_frame_timings(self.ctx, C.byref(t)))
        return t

    def selftest_division(self, num, den):
        """(fast, plain) results of a / d per component on the device: div3_shared vs `/` (csrc/device_common.cuh)."""
        num = np.ascontiguousarray(num, np.float32).reshape(-1, 3); den = np.ascontiguousarray(den, np.float32).reshape(-1)
        fast = np.zeros_like(num); ref = np.zeros_like(num)
        self._check(self.lib.romis_selftest_division(self.ctx, num.ctypes.data, den.ctypes.data, len(den), fast.ctypes.data, ref.ctypes.data))
        return fast, ref

    def trace_rays(self, origins, dirs, tfar, any_hit=False):
        n = len(tfar)
        o = np.ascontiguousarray(origins, np.float32); d = np.ascontiguousarray(dirs, np.float32); tf = np.ascontiguousarray(tfar, np.float32)
        hit = np.zeros(n, np.uint8); t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
        tri = np.full(n, 0xFFFFFFFF, np.uint32)
        self._check(self.lib.romis_trace_rays(self.ctx, o.ctypes.data, d.ctypes.data, tf.ctypes.data, n, int(any_hit),
                                              hit.ctypes.data, t.ctypes.data, u.ctypes.data, v.ctypes.data, tri.ctypes.data))
        return hit, t, u, v, tri
