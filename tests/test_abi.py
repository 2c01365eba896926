"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/lbgpu.h
declares; without a CUDA device the entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import common
from hybird_b200 import abi, build


def header_symbols():
    src = open(os.path.join(common.ROOT, "include", "lbgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbGpu[A-Za-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    lib_path = build.build()
    assert os.path.exists(lib_path)
    syms = header_symbols()
    assert "lbGpuInit" in syms and "lbGpuStep" in syms and "lbGpuParticleForces" in syms and "lbGpuFetchFields" in syms
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    missing = [s for s in syms if s not in exported]
    assert not missing, "declared in include/lbgpu.h but not exported: %s" % missing
    assert sorted(abi.EXPORTS) == syms, "abi.EXPORTS out of sync with include/lbgpu.h"


def test_sm100a_cubin_embedded():
    lib_path = build.build()
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert "sm_100a" in out, out[:500]


def test_struct_layout_matches_header():
    assert C.sizeof(abi.LbGpuParams) == 3 * 4 + 6 * 4 + 4 + 3 * 8 + 5 * 8 + 4 * 4 + 3 * 8 + 2 * 4 + 3 * 4 + 5 * 4 + 4 \
        or C.sizeof(abi.LbGpuParams) % 8 == 0
    from hybird_b200.lattice_init import ELEMENT_DTYPE, PARTICLE_DTYPE
    assert PARTICLE_DTYPE.itemsize == 64 and ELEMENT_DTYPE.itemsize == 56


def test_no_cpu_fallback_without_device():
    L = abi.load_library()
    assert L.lbGpuAbiVersion() >= 1
    if L.lbGpuDeviceCount() > 0:
        pytest.skip("a CUDA device is present")
    prm = dict(size=[6, 6, 6], boundary=[7] * 6, lbF=[0, 0, 0], initDynVisc=1 / 6, plasticVisc=1 / 6, yieldStress=0,
               turbConst=0, slipCoefficient=0, freeSurface=0, forceField=0, nonNewtonian=0, turbulence=0, unitLength=1,
               unitTime=1, unitDensity=1)
    from hybird_b200 import LB, LbGpuError
    N = 216
    with pytest.raises(LbGpuError) as ei:
        LB(prm).latticeBolzmannInit(np.full(N, 7, np.uint8), np.zeros(N, np.uint32), np.ones(N), np.zeros((N, 3)),
                                    np.ones(N), np.full(N, 1 / 6))
    assert ei.value.code == -2  # LBGPU_ENODEVICE


def test_product_does_not_import_oracle():
    """The product package must not reference oracle/ (import, link or dlopen)."""
    pkg = os.path.join(common.ROOT, "hybird_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "lbo" not in re.findall(r"^\s*(?:import|from)\s+(\w+)", txt, flags=re.M), f
                assert "liblboracle" not in txt and "lb_oracle" not in txt, f


def test_product_shim_links_no_reference_lb_step():
    """hybird_gpu (the drop-in driver) must not carry the reference's own LB time step: only its set-up
    (latticeBoltzmannGet / latticeBolzmannInit and what they call) stays linked; the verify binary has both."""
    shim = os.path.join(common.ROOT, "hybird_b200", "shim", "_build")
    prod, ver = os.path.join(shim, "hybird_gpu"), os.path.join(shim, "hybird_gpu_verify")
    if not (os.path.exists(prod) and os.path.exists(ver)):
        pytest.skip("shim binaries not built (needs the reference sources)")
    step = re.compile(r" [Tt] (LB::(streaming|collision|reconstruction|computeHydroForces|updateMass|updateInterface|"
                      r"findInterfaceMutants|smoothenInterface|removeIsolated|findNewActive|findNewSolid)\(|LB_ref_latticeBolzmannStep|"
                      r"LB_ref_latticeBoltzmannCouplingStep|LB_ref_latticeBoltzmannFreeSurfaceStep)")
    def syms(p):
        return subprocess.run(["nm", "-C", p], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert not step.findall(syms(prod)), "the product binary links the reference's LB step"
    assert step.findall(syms(ver)), "the verify binary should step the reference alongside"
    # IO's whole-lattice walks bind to the shim's device-backed versions in both
    for p in (prod, ver):
        assert "lbGpuFluidSummary" in syms(p) and "lbGpuWriteVti" in syms(p)
