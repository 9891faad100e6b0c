"""ctypes mirror of include/romis_gpu.h (the C-ABI's POD types).

Field names follow the reference's own structs: `Features` (reference src/utils/common.h:89-136),
`Vertex`/`Material`/`Mesh` (framework/include/framework/mesh.h:14-43), the light structs
(src/utils/common.h:72-87).
"""
import ctypes as C

ROMIS_LIGHT_POINT, ROMIS_LIGHT_SEGMENT, ROMIS_LIGHT_PARALLELOGRAM = 0, 1, 2
ROMIS_PASS_INITIAL, ROMIS_PASS_TEMPORAL, ROMIS_PASS_SPATIAL0, ROMIS_PASS_FINAL = 0, 1, 2, 1000
ROMIS_HALO_SEND_LOW, ROMIS_HALO_SEND_HIGH, ROMIS_HALO_RECV_LOW, ROMIS_HALO_RECV_HIGH = 0, 1, 2, 3

f3 = C.c_float * 3
f4 = C.c_float * 4
f2 = C.c_float * 2


class romis_vertex(C.Structure):
    _fields_ = [("position", f3), ("normal", f3), ("texcoord", f2)]


class romis_material(C.Structure):
    _fields_ = [("kd", f3), ("ks", f3), ("shininess", C.c_float), ("transparency", C.c_float),
                ("kd_texture", C.c_int32)]


class romis_mesh_desc(C.Structure):
    _fields_ = [("vertices", C.POINTER(romis_vertex)), ("n_vertices", C.c_uint32),
                ("triangles", C.POINTER(C.c_uint32)), ("n_triangles", C.c_uint32),
                ("material", romis_material)]


class romis_texture(C.Structure):
    _fields_ = [("pixels", C.POINTER(C.c_float)), ("width", C.c_int32), ("height", C.c_int32)]


class romis_light(C.Structure):
    _fields_ = [("type", C.c_uint32), ("p0", f3), ("e1", f3), ("e2", f3),
                ("c0", f3), ("c1", f3), ("c2", f3), ("c3", f3)]


class romis_features(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "enableShading", "enableTextureMapping", "initialSamplesVisibilityCheck",
        "numSamplesInReservoir", "initialLightSamples", "numNeighboursToSample",
        "spatialResampleRadius", "unbiasedCombination", "spatialReuse",
        "spatialReuseVisibilityCheck", "temporalReuse", "spatialResamplingPasses",
        "temporalClampM", "enableToneMapping")] + [("gamma", C.c_float), ("exposure", C.c_float)]


class romis_rmis_params(C.Structure):
    _fields_ = [("maxIterationsMIS", C.c_uint32), ("misWeightRMIS", C.c_uint32), ("neighbourSelectionStrategy", C.c_uint32),
                ("neighbourSameGeometry", C.c_uint32), ("neighbourMaxDepthDifferenceFraction", C.c_float),
                ("neighbourMaxNormalAngleDifferenceRadians", C.c_float),
                ("useProgressiveROMIS", C.c_uint32), ("progressiveUpdateMod", C.c_uint32)]


ROMIS_MIS_EQUAL, ROMIS_MIS_BALANCE = 0, 1
ROMIS_NEIGHBOURS_RANDOM, ROMIS_NEIGHBOURS_SIMILAR, ROMIS_NEIGHBOURS_DISSIMILAR, ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR = 0, 1, 2, 3


class romis_camera(C.Structure):
    _fields_ = [("origin", f3), ("quat", f4), ("half_width", C.c_float), ("half_height", C.c_float)]


class romis_rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("frame", C.c_uint32), ("reserved", C.c_uint32)]


class romis_reservoir_dump(C.Structure):
    _fields_ = [("light_id", C.POINTER(C.c_uint32)), ("u", C.POINTER(C.c_float)), ("v", C.POINTER(C.c_float)),
                ("W", C.POINTER(C.c_float)), ("M", C.POINTER(C.c_uint32)),
                ("position", C.POINTER(C.c_float)), ("color", C.POINTER(C.c_float))]


class romis_gbuffer_dump(C.Structure):
    _fields_ = [("t", C.POINTER(C.c_float)), ("normal", C.POINTER(C.c_float)),
                ("texcoord", C.POINTER(C.c_float)), ("mesh", C.POINTER(C.c_uint32))]


class romis_timings(C.Structure):
    _fields_ = [("primary_ms", C.c_float), ("initial_ms", C.c_float), ("temporal_ms", C.c_float),
                ("spatial_ms", C.c_float * 8), ("shade_ms", C.c_float), ("total_ms", C.c_float),
                ("n_spatial", C.c_int32), ("n_launches", C.c_int32), ("exchange_ms", C.c_float * 8),
                ("neighbours_ms", C.c_float), ("gather_ms", C.c_float), ("resolve_ms", C.c_float)]
