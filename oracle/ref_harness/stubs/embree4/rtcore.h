/* Stand-in for <embree4/rtcore.h> (Intel Embree 4.3.1 is not in this image).
 * Types only: the reference's embree_interface.h needs RTCDevice / RTCScene / RTCRayHit as members
 * and utils.h needs `enum RTCError`.  No Embree function is declared: the harness supplies its own
 * EmbreeInterface member definitions (shims.cpp) backed by oracle/tracer.c. */
#pragma once
#include <cstdint>
typedef struct RTCDeviceTy* RTCDevice;
typedef struct RTCSceneTy* RTCScene;
typedef struct RTCGeometryTy* RTCGeometry;
enum RTCError { RTC_ERROR_NONE = 0, RTC_ERROR_UNKNOWN = 1 };
#define RTC_INVALID_GEOMETRY_ID ((unsigned int)-1)
struct RTCRay { float org_x, org_y, org_z, tnear, dir_x, dir_y, dir_z, time, tfar; unsigned int mask, id, flags; };
struct RTCHit { float Ng_x, Ng_y, Ng_z, u, v; unsigned int primID, geomID, instID[1]; };
struct RTCRayHit { RTCRay ray; RTCHit hit; };
