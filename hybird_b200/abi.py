"""ctypes binding of liblbgpu.so (include/lbgpu.h).  Fails loudly when the library or a CUDA
device is missing: there is no CPU path behind this module."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

Q = 19


class LbGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lbgpu error %d: %s" % (code, msg))
        self.code = code


class LbGpuParams(C.Structure):
    _fields_ = [("size", C.c_int32 * 3), ("boundary", C.c_int32 * 6), ("lbF", C.c_double * 3),
                ("initDynVisc", C.c_double), ("plasticVisc", C.c_double), ("yieldStress", C.c_double),
                ("turbConst", C.c_double), ("slipCoefficient", C.c_double),
                ("freeSurface", C.c_int32), ("forceField", C.c_int32), ("nonNewtonian", C.c_int32),
                ("turbulence", C.c_int32),
                ("unitLength", C.c_double), ("unitTime", C.c_double), ("unitDensity", C.c_double),
                ("nWalls", C.c_int32), ("device", C.c_int32),
                ("slabAxis", C.c_int32), ("nSlabs", C.c_int32), ("slabIndex", C.c_int32),
                ("nLocalSlabs", C.c_int32), ("reserved", C.c_int32 * 4)]


class LbGpuDemParams(C.Structure):
    _fields_ = [("contactModel", C.c_int32), ("multiStep", C.c_int32), ("knConst", C.c_double), ("ksConst", C.c_double),
                ("dampCoeff", C.c_double), ("viscTang", C.c_double), ("linearStiff", C.c_double), ("frictionCoefPart", C.c_double),
                ("frictionCoefWall", C.c_double), ("numVisc", C.c_double), ("demF", C.c_double * 3), ("deltat", C.c_double),
                ("nebrRange", C.c_double), ("maxDisp", C.c_double), ("nPbc", C.c_int32), ("pad", C.c_int32),
                ("pbcP", (C.c_double * 3) * 3), ("pbcV", (C.c_double * 3) * 3)]


DEM_ELEMENT_DTYPE = np.dtype([("x0", "<f8", 3), ("x1", "<f8", 3), ("w0", "<f8", 3), ("radius", "<f8"), ("m", "<f8"), ("I", "<f8", 3),
                              ("size", "<i4"), ("pad", "<i4")])
DEM_WALL_DTYPE = np.dtype([("n", "<f8", 3), ("p", "<f8", 3), ("vel", "<f8", 3), ("omega", "<f8", 3), ("rotCenter", "<f8", 3),
                           ("moving", "<i4"), ("pad", "<i4")])

# every symbol include/lbgpu.h declares
EXPORTS = ("lbGpuLastError", "lbGpuAbiVersion", "lbGpuDeviceCount", "lbGpuSlabRange", "lbGpuInit", "lbGpuInitBox", "lbGpuSetCurves", "lbGpuSetMassTarget", "lbGpuStep", "lbGpuCouple", "lbGpuRun",
           "lbGpuParticleForces", "lbGpuFetchFields", "lbGpuStateBytes", "lbGpuSaveState", "lbGpuLoadState", "lbGpuCounts", "lbGpuCountsLocal", "lbGpuSynchronize", "lbGpuLastStepMs",
           "lbGpuLastKernelMs", "lbGpuLaunchCount", "lbGpuSelfTest", "lbGpuFinalize", "lbGpuCommUniqueId",
           "lbGpuCommInit", "lbGpuCommInfo", "lbGpuCommFinalize", "lbGpuPeerHalo", "lbGpuFluidSummary", "lbGpuWriteVti",
           "lbGpuDemInit", "lbGpuDemStep", "lbGpuRunDem", "lbGpuDemState", "lbGpuDemContacts", "lbGpuDemParticles", "lbGpuGraphInfo", "lbGpuPhaseTrace", "lbGpuPhaseMs")

_lib = None


def load_library(build_if_missing=True):
    """dlopen liblbgpu.so (building it with nvcc if absent) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing and _build.needs_build():
        try:
            _build.build()
        except Exception:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise FileNotFoundError(path + " is missing: run `python -m hybird_b200.build`")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.lbGpuLastError.restype = C.c_char_p
    L.lbGpuAbiVersion.restype = C.c_int
    L.lbGpuDeviceCount.restype = C.c_int
    L.lbGpuSlabRange.restype = C.c_int
    L.lbGpuSlabRange.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.lbGpuInit.restype = C.c_int
    L.lbGpuInit.argtypes = [C.POINTER(LbGpuParams), vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.lbGpuInitBox.restype = C.c_int
    L.lbGpuInitBox.argtypes = [C.POINTER(LbGpuParams), vp, vp, vp, C.c_uint32, vp, C.c_uint32, C.POINTER(vp)]
    L.lbGpuSetCurves.restype = C.c_int
    L.lbGpuSetCurves.argtypes = [vp, C.c_uint32, vp, vp]
    L.lbGpuSetMassTarget.restype = C.c_int
    L.lbGpuSetMassTarget.argtypes = [vp, C.c_double]
    L.lbGpuStep.restype = C.c_int
    L.lbGpuStep.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32]
    L.lbGpuCouple.restype = C.c_int
    L.lbGpuCouple.argtypes = [vp, C.c_int, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32]
    L.lbGpuRun.restype = C.c_int
    L.lbGpuRun.argtypes = [vp, C.c_int, C.c_uint32]
    L.lbGpuParticleForces.restype = C.c_int
    L.lbGpuParticleForces.argtypes = [vp, vp, vp, vp, vp]
    L.lbGpuFetchFields.restype = C.c_int
    L.lbGpuFetchFields.argtypes = [vp] * 10
    L.lbGpuStateBytes.restype = C.c_int
    L.lbGpuStateBytes.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.lbGpuSaveState.restype = C.c_int
    L.lbGpuSaveState.argtypes = [vp, vp, C.c_uint64]
    L.lbGpuLoadState.restype = C.c_int
    L.lbGpuLoadState.argtypes = [vp, vp, C.c_uint64]
    L.lbGpuCounts.restype = C.c_int
    L.lbGpuCounts.argtypes = [vp, C.POINTER(C.c_uint64 * 4)]
    L.lbGpuCountsLocal.restype = C.c_int
    L.lbGpuCountsLocal.argtypes = [vp, C.POINTER(C.c_uint64 * 4)]
    L.lbGpuSynchronize.restype = C.c_int
    L.lbGpuSynchronize.argtypes = [vp]
    L.lbGpuLastStepMs.restype = C.c_int
    L.lbGpuLastStepMs.argtypes = [vp, C.POINTER(C.c_float)]
    L.lbGpuLastKernelMs.restype = C.c_int
    L.lbGpuLastKernelMs.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
    L.lbGpuLaunchCount.restype = C.c_int
    L.lbGpuLaunchCount.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.lbGpuSelfTest.restype = C.c_int
    L.lbGpuSelfTest.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64 * 3)]
    L.lbGpuCommUniqueId.restype = C.c_int
    L.lbGpuCommUniqueId.argtypes = [vp]
    L.lbGpuCommInit.restype = C.c_int
    L.lbGpuCommInit.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
    L.lbGpuCommInfo.restype = C.c_int
    L.lbGpuCommInfo.argtypes = [C.POINTER(C.c_int32)] * 3
    L.lbGpuFluidSummary.restype = C.c_int
    L.lbGpuFluidSummary.argtypes = [vp, C.POINTER(C.c_double * 4)]
    L.lbGpuWriteVti.restype = C.c_int
    L.lbGpuWriteVti.argtypes = [vp, C.c_char_p, C.c_int]
    L.lbGpuDemInit.restype = C.c_int
    L.lbGpuDemInit.argtypes = [vp, C.POINTER(LbGpuDemParams), vp, C.c_uint32, vp, C.c_uint32]
    L.lbGpuDemStep.restype = C.c_int
    L.lbGpuDemStep.argtypes = [vp, vp]
    L.lbGpuRunDem.restype = C.c_int
    L.lbGpuRunDem.argtypes = [vp, C.c_int, C.c_uint32]
    L.lbGpuDemState.restype = C.c_int
    L.lbGpuDemState.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_double * 3)]
    L.lbGpuDemParticles.restype = C.c_int
    L.lbGpuDemParticles.argtypes = [vp, C.POINTER(C.c_uint32), vp, vp, vp]
    L.lbGpuDemContacts.restype = C.c_int
    L.lbGpuDemContacts.argtypes = [vp, vp, vp, vp, vp]
    L.lbGpuPhaseTrace.restype = C.c_int
    L.lbGpuPhaseTrace.argtypes = [vp, C.c_int]
    L.lbGpuPhaseMs.restype = C.c_int
    L.lbGpuPhaseMs.argtypes = [vp, C.POINTER(C.c_float * 7), C.POINTER(C.c_uint32)]
    L.lbGpuGraphInfo.restype = C.c_int
    L.lbGpuGraphInfo.argtypes = [vp, C.POINTER(C.c_uint64 * 2)]
    L.lbGpuPeerHalo.restype = C.c_int
    L.lbGpuPeerHalo.argtypes = [vp, C.POINTER(C.c_int32)]
    L.lbGpuCommFinalize.restype = C.c_int
    L.lbGpuCommFinalize.argtypes = []
    L.lbGpuFinalize.restype = C.c_int
    L.lbGpuFinalize.argtypes = [vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise LbGpuError(rc, load_library().lbGpuLastError().decode())


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_params(p: dict, device=-1) -> LbGpuParams:
    P = LbGpuParams()
    P.size[:] = [int(v) for v in p["size"]]
    P.boundary[:] = [int(v) for v in p["boundary"]]
    P.lbF[:] = [float(v) for v in p["lbF"]]
    for k in ("initDynVisc", "plasticVisc", "yieldStress", "turbConst", "slipCoefficient", "unitLength", "unitTime",
              "unitDensity"):
        setattr(P, k, float(p[k]))
    for k in ("freeSurface", "forceField", "nonNewtonian", "turbulence"):
        setattr(P, k, int(p[k]))
    P.nWalls = int(p.get("nWalls", 0))
    P.device = int(device)
    P.slabAxis = 2
    P.nSlabs = int(p.get("nSlabs", 1))
    P.slabIndex = int(p.get("slabIndex", 0))
    P.nLocalSlabs = int(p.get("nLocalSlabs", P.nSlabs))
    return P
