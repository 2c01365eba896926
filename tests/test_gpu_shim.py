"""GPU: the drop-in boundary.  hybird_b200/shim/_build/hybird_gpu is the UNMODIFIED reference driver (hybird.cpp,
DEM.cpp, IO.cpp, ...) linked against the LB shim + liblbgpu.so; hybird_ref is the same objects linked as the
reference is.  Both binaries are built in the build container (make -C hybird_b200/shim, from /root/reference) and
travel to the GPU box; nothing here reads /root/reference.

(1) the shipped configuration's physics (cfg1: demChute free surface + Smagorinsky) run through both
    drivers gives the same console/export.dat lines and the same ParaView fluid files -- the shim serves IO's
    whole-lattice walks from the device (lbGpuFluidSummary, lbGpuWriteVti: raw appended data instead of text), so the
    files are compared array by array;
(2) hybird_gpu_verify with LBGPU_VERIFY=1 steps the reference's own LB on the host next to the device every cycle,
    with the reference's DEM in the loop, and compares cell types, fields and particle forces (exit code 2 on a
    mismatch).  The product binary hybird_gpu links no reference LB step at all (tests/test_abi.py)."""
import os
import re
import subprocess

import numpy as np
import pytest

import common

import cases

pytestmark = pytest.mark.gpu
BUILD = os.path.join(common.ROOT, "hybird_b200", "shim", "_build")
GPU_BIN, REF_BIN = os.path.join(BUILD, "hybird_gpu"), os.path.join(BUILD, "hybird_ref")
VERIFY_BIN = os.path.join(BUILD, "hybird_gpu_verify")
NUM = re.compile(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?")


def _need_binaries():
    if not (os.path.exists(GPU_BIN) and os.path.exists(REF_BIN) and os.path.exists(VERIFY_BIN)):
        pytest.skip("shim binaries not built (make -C hybird_b200/shim needs the reference sources)")


def _run(binary, cfg, outdir, env=None, threads=None):
    os.makedirs(outdir, exist_ok=True)
    e = dict(os.environ, OMP_NUM_THREADS=str(threads or os.cpu_count() or 1))
    e.update(env or {})
    r = subprocess.run([binary, "-c", cfg, "-d", outdir, "-n", "run"], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=1200)
    return r.returncode, r.stdout


def _numbers(text):
    return np.array([float(x) for x in NUM.findall(text)])


VTK_DTYPES = {"Int8": np.int8, "Int16": np.int16, "Float64": np.float64}


def read_vti(path):
    """{name: array} + header attributes of an ImageData file with ascii (the reference's) or raw appended (lbGpuWriteVti)
    DataArrays."""
    raw = open(path, "rb").read()
    head_end = raw.find(b"<AppendedData")
    xml = raw[:head_end if head_end >= 0 else len(raw)].decode()
    meta = dict(extent=re.search(r'WholeExtent="([^"]*)"', xml).group(1), spacing=[float(v) for v in re.search(r'Spacing="([^"]*)"', xml).group(1).split()])
    out, order = {}, []
    base = raw.find(b"_", head_end) + 1 if head_end >= 0 else None
    for m in re.finditer(r'<DataArray ([^>]*?)(/>|>(.*?)</DataArray>)', xml, flags=re.S):
        attrs = dict(re.findall(r'(\w+)="([^"]*)"', m.group(1)))
        dt = VTK_DTYPES[attrs["type"]]
        order.append((attrs["Name"], attrs["type"], int(attrs.get("NumberOfComponents", 1))))
        if attrs.get("format") == "appended":
            off = base + int(attrs["offset"])
            nb = int(np.frombuffer(raw, "<u4", 1, off)[0])
            out[attrs["Name"]] = np.frombuffer(raw, dt, nb // np.dtype(dt).itemsize, off + 4).astype(np.float64)
        else:
            out[attrs["Name"]] = np.array(m.group(3).split(), dtype=np.float64)
    return meta, order, out


def test_shipped_config_through_both_drivers(tmp_path):
    _need_binaries()
    case = dict(cases.catalogue()["cfg1"])
    case.update(maximumTimeSteps=40, screenExpTime=1e-3, fluidExpTime=4e-3)
    cfg = cases.write_case_files(case, str(tmp_path))
    rc_r, out_r = _run(REF_BIN, cfg, str(tmp_path / "ref"))
    rc_g, out_g = _run(GPU_BIN, cfg, str(tmp_path / "gpu"))
    assert rc_r == 0, out_r[-2000:]
    assert rc_g == 0, out_g[-2000:]
    assert "cells uploaded to the GPU" in out_g
    # export.dat: step; time; MaxFSpeed; Volume; Mass per screenExpTime
    er = open(tmp_path / "ref" / "run" / "export.dat").read()
    eg = open(tmp_path / "gpu" / "run" / "export.dat").read()
    assert len(er.splitlines()) == len(eg.splitlines()) >= 8
    a, b = _numbers(er), _numbers(eg)
    assert a.shape == b.shape
    assert np.allclose(a, b, rtol=2e-6, atol=0), np.abs(a - b).max()
    # ParaView fluid files (ASCII, ~6 significant digits): same file names, same numbers
    fr = sorted(os.listdir(tmp_path / "ref" / "run" / "fluidData"))
    fg = sorted(os.listdir(tmp_path / "gpu" / "run" / "fluidData"))
    assert fr == fg and len(fr) >= 2
    for name in fr:
        mr, orr, ar = read_vti(tmp_path / "ref" / "run" / "fluidData" / name)
        mg, org, ag = read_vti(tmp_path / "gpu" / "run" / "fluidData" / name)
        assert mr["extent"] == mg["extent"] and np.allclose(mr["spacing"], mg["spacing"], rtol=1e-6)
        assert orr == org, (orr, org)  # same arrays, types, components, order
        for k in ar:
            a, b = ar[k], ag[k]
            assert a.shape == b.shape, (name, k)
            # the reference prints 6 digits: a last-digit flip of a rounded value is the most that can differ
            assert np.allclose(a, b, rtol=2e-5, atol=1e-12), (name, k, np.abs(a - b).max())
    # no export step fetched the whole state: the screen lines came from device reductions
    assert os.path.getsize(tmp_path / "gpu" / "run" / "maxVel.dat") > 0


@pytest.mark.parametrize("name,steps", [("cfg3_mini", 60), ("cluster_dem", 40), ("cfg1_mini", 60), ("drum_mini", 150)])
def test_verify_mode_reference_steps_alongside(name, steps, tmp_path):
    _need_binaries()
    case = dict(cases.catalogue()[name])
    case.update(maximumTimeSteps=steps)
    cfg = cases.write_case_files(case, str(tmp_path))
    rc, out = _run(VERIFY_BIN, cfg, str(tmp_path / "gpu"), env={"LBGPU_VERIFY": "1"}, threads=1)
    assert rc == 0, out[-3000:]
    lines = [l for l in out.splitlines() if l.startswith("lbgpu verify: step")]
    assert len(lines) == steps, out[-2000:]
    assert "FAILED" not in out


@pytest.mark.parametrize("name,steps", [("spheres_hertz", 80), ("bed_dem", 120)])
def test_dem_on_device_through_the_reference_driver(name, steps, tmp_path):
    """LBGPU_DEM=1: the shim's DEM::discreteElementStep keeps only the DEM clock on the host; the sub-steps (contacts, tables,
    Gear integration) run on the device inside the cycle's lbGpuRunDem.  The unmodified driver's screen lines, force file and
    particle files (positions, velocities, spins, contact / wall / hydrodynamic forces per particle) must be the reference's."""
    _need_binaries()
    case = dict(cases.catalogue()[name])
    case.update(maximumTimeSteps=steps, screenExpTime=5.0, partExpTime=20.0)
    cfg = cases.write_case_files(case, str(tmp_path))
    rc_r, out_r = _run(REF_BIN, cfg, str(tmp_path / "ref"), threads=1)
    rc_g, out_g = _run(GPU_BIN, cfg, str(tmp_path / "gpu"), env={"LBGPU_DEM": "1"}, threads=1)
    assert rc_r == 0, out_r[-2000:]
    assert rc_g == 0, out_g[-2000:]
    assert "DEM sub-steps on the GPU" in out_g
    for f in ("export.dat", "force.dat"):
        tr, tg = open(tmp_path / "ref" / "run" / f).read(), open(tmp_path / "gpu" / "run" / f).read()
        assert len(tr.splitlines()) == len(tg.splitlines()) >= 5, f
        a, b = _numbers(tr), _numbers(tg)
        assert a.shape == b.shape, f
        assert np.allclose(a, b, rtol=2e-5, atol=1e-9), (f, np.abs(a - b).max())
    fr = sorted(os.listdir(tmp_path / "ref" / "run" / "particleData"))
    fg = sorted(os.listdir(tmp_path / "gpu" / "run" / "particleData"))
    assert fr == fg and len(fr) >= 3
    for f in fr:
        a = _numbers(open(tmp_path / "ref" / "run" / "particleData" / f).read())
        b = _numbers(open(tmp_path / "gpu" / "run" / "particleData" / f).read())
        assert a.shape == b.shape, f
        # six printed digits; forces that cancel to ~0 are compared against the file's force scale
        assert np.allclose(a, b, rtol=2e-5, atol=1e-7), (f, np.abs(a - b).max())
