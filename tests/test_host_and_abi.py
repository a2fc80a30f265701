"""CPU-only checks of the product's host side and of the C-ABI library (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    import swb200 as S

    lib = S._lib.load()
    header = open(os.path.join(ROOT, "include", "swb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(swb_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    assert sorted(S._lib.EXPORTS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.swb_abi_version() == 2  # SWB_ABI_VERSION of include/swb200.h (2: swb_abi_layout, swb_sim_gradient_l2_ex)


def test_no_cpu_fallback_without_device():
    import swb200 as S

    if S.device_count() > 0:
        pytest.skip("a device is present")
    bc = S.CPMLBoundaryConditionParameters(halo=5)
    params = S.InputParametersAcoustic(50, 1e-3, (40, 40), (10.0, 10.0), bc)
    with pytest.raises(RuntimeError):
        S.build_wavesim(params, S.VpAcousticCDMaterialProperties(2000.0 * np.ones((40, 40))), runparams=S.RunParameters())
    with pytest.raises(ValueError):
        S.build_wavesim(params, S.VpAcousticCDMaterialProperties(2000.0 * np.ones((40, 40))), runparams=S.RunParameters(parall="threads"))
    # a raw engine call must fail with a CUDA error, not silently succeed
    lib = S._lib.load()
    d = S._lib.swb_sim_desc(kind=1, dtype=1, ndim=2, device=0, dt=1e-3, nt=10, halo=2, freetop=1, gradient=0, check_freq=1)
    d.n[0], d.n[1] = 16, 16
    d.spacing[0], d.spacing[1] = 1.0, 1.0
    h = C.c_void_p()
    assert lib.swb_sim_create(C.byref(d), C.byref(h)) == 2  # SWB_ERR_CUDA
    assert b"no CPU fallback" in lib.swb_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "seismicwaves.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "import oracle" not in src and "from oracle" not in src and "libswref" not in src, f


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("freetop", [True, False])
def test_hostprep_cpml_matches_oracle(dtype, freetop):
    import swb200 as S

    T = np.dtype(dtype).type
    ours = S.hostprep.init_bdc(T(3456.7), T(1.1e-3), 20, T(1e-4), (T(10.0), T(7.5)), freetop, T(12.0), dtype)
    ref = O.init_bdc(T(3456.7), T(1.1e-3), 20, T(1e-4), (T(10.0), T(7.5)), freetop, T(12.0), dtype)
    for (a, a_h, b, b_h), r in zip(ours, ref):
        assert a.dtype == np.dtype(dtype) and len(a) == 42 and len(a_h) == 40
        assert np.array_equal(a, r.a) and np.array_equal(a_h, r.a_h) and np.array_equal(b, r.b) and np.array_equal(b_h, r.b_h)
    a, a_h, b, b_h = ours[-1]
    if freetop:
        assert np.all(a[:21] == 0) and np.all(b[:21] == 1) and np.all(a_h[:20] == 0) and np.all(b_h[:20] == 1)
    assert np.all(b[21:] <= 1) and np.all(a[22:] < 0)
    # halo = 0: no NaNs (cpmlcoeffs.jl:30-32)
    z = S.hostprep.init_bdc(T(2000), T(1e-3), 0, T(1.0), (T(5.0),), False, T(5.0), dtype)[0]
    assert len(z[0]) == 2 and len(z[1]) == 0 and np.all(np.isfinite(z[0])) and np.all(z[0] == 0)


def test_hostprep_positions_and_scaling_match_oracle():
    import swb200 as S

    rng = np.random.default_rng(0)
    for dtype in (np.float32, np.float64):
        T = np.dtype(dtype).type
        pos = rng.uniform(0, 500, size=(50, 2)).astype(dtype)
        pos[0] = [15.0, 25.0]  # exact ties (x.5 in grid units) round up
        a = S.hostprep.find_nearest_grid_points(pos, (T(10.0), T(10.0)), dtype)
        b = O.find_nearest_grid_points(pos, (T(10.0), T(10.0)), dtype)
        assert np.array_equal(a, b) and a.dtype == np.int64
        assert list(a[0]) == [3, 4]
    assert [list(r) for r in S.distribsrcs(10, 4)] == [list(r) for r in O.distribsrcs(10, 4)]


def test_types_mirror_reference_checks():
    import swb200 as S

    with pytest.raises(AssertionError):
        S.ScalarSources(np.zeros((2, 2)), np.zeros((10, 3)), 5.0)
    with pytest.raises(AssertionError):
        S.InputParametersAcoustic(0, 1e-3, (10, 10), (1.0, 1.0), S.CPMLBoundaryConditionParameters())
    with pytest.raises(ValueError):
        S.L2Misfit(observed=np.zeros(5))
    with pytest.raises(AssertionError):
        S.L2Misfit(observed=np.zeros((5, 2)), windows=[(0, 3)])
    with pytest.raises(AssertionError):
        S.GradParameters(check_freq=0)
    m = S.L2Misfit(observed=np.zeros((5, 2)))
    recs = S.ScalarReceivers(np.zeros((2, 2)), 5)
    recs.seismograms[:] = 2.0
    assert m.calcmisfit(recs) == 20.0
    assert np.array_equal(m.dchi_du(recs), 2.0 * np.ones((5, 2)))


def test_slab_partition_and_routing():
    """host logic of the z-slab decomposition: contiguous cover, ghost planes, point routing"""
    from swb200.multigpu import slab_local_planes, slab_range, slab_route_points

    for nz, world in [(1024, 8), (61, 2), (90, 4), (75, 7)]:
        owned = [slab_range(nz, world, r) for r in range(world)]
        assert owned[0].start == 0 and owned[-1].stop == nz
        assert all(owned[r].stop == owned[r + 1].start for r in range(world - 1))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
        for r in range(world):
            loc = slab_local_planes(nz, world, r)
            assert loc.start == owned[r].start - (r > 0) and loc.stop == owned[r].stop + (r < world - 1)
        idx = np.stack([np.zeros(nz, dtype=np.int64), np.zeros(nz, dtype=np.int64), np.arange(nz)], axis=1)
        routed = np.concatenate([slab_route_points(idx, nz, world, r) for r in range(world)])
        assert sorted(routed.tolist()) == list(range(nz))  # every point has exactly one owner


def test_julia_extension_binds_only_exported_symbols():
    """ext/SeismicWaves_B200BackendExt.jl is the reference-side binding of the C ABI: every symbol it
    ccalls must be declared in include/swb200.h and exported by libswb200.so, so the boundary cannot drift unnoticed."""
    import swb200 as S

    lib = S._lib.load()
    src = open(os.path.join(ROOT, "ext", "SeismicWaves_B200BackendExt.jl")).read()
    bound = sorted(set(re.findall(r":(swb_[a-z0-9_]+)", src)))
    assert len(bound) >= 20, bound
    missing = [s for s in bound if s not in S._lib.EXPORTS or not hasattr(lib, s)]
    assert not missing, missing


def test_header_is_valid_c_and_layout_matches_a_c_compiler(tmp_path):
    """include/swb200.h is a C header (the Julia ccall / cgo-style consumer sees C, not C++): it must compile as C11, and a C translation unit
    must see the same struct sizes the library (compiled as C++ by nvcc) reports through swb_abi_layout()."""
    import json
    import shutil
    import subprocess

    import swb200 as S

    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if cc is None:
        pytest.skip("no C compiler")
    layout = json.loads(S._lib.load().swb_abi_layout().decode())
    src = tmp_path / "abi.c"
    lines = ['#include "swb200.h"', "#include <stdio.h>", "int main(void) {"]
    for name in layout:
        lines.append(f'    printf("{name} %zu\\n", sizeof({name}));')
    lines += ["    return 0;", "}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.check_call([cc, "-std=c11", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    sizes = dict((ln.split()[0], int(ln.split()[1])) for ln in out.splitlines())
    assert sizes == {k: v["size"] for k, v in layout.items()}


@pytest.mark.parametrize("esize", [4, 8])
def test_fused_cd_kernels_partition_the_grid(esize):
    """The fused CD step (csrc/acou_cd_fused.cu) runs a grid as up to three launches that write disjoint cells -- the bulk march,
    the z-marched x / y strip boxes, the per-vector strips / faces -- and addresses sources / receivers as (kernel, CTA, slot).
    Host-only check through the C ABI (no device): every cell of a grid is owned exactly once and every slot lies inside its CTA,
    for 2D and 3D shapes, odd extents, every strip combination of the z ends (free surface, z-slab interior faces), with and
    without the march."""
    import swb200 as S

    lib = S._lib.load()
    counts = (C.c_int64 * 3)()
    v = 16 // esize
    shapes = [(37, 1, 29, 4), (300, 1, 280, 20), (131, 1, 64, 0),  # 2D (ny = 1)
              (40, 36, 44, 6), (70, 52, 90, 8), (129, 33, 47, 5), (64, 40, 61, 10), (48, 48, 48, 0), (260, 30, 33, 12), (33, 31, 30, 14)]
    for nx, ny, nz, halo in shapes:
        for zlo, zhi in [(1, 1), (0, 1), (1, 0), (0, 0)]:
            for rim_zc in (-1, 0, 5):
                for zc in (7, 64):
                    rc = lib.swb_diag_cd_partition(esize, nx, ny, nz, halo, zc, zlo, zhi, rim_zc, counts)
                    assert rc == 0, (nx, ny, nz, halo, zlo, zhi, rim_zc, zc, lib.swb_last_error())
                    assert sum(counts) == nx * ny * nz
                    marched_possible = ny > 1 and rim_zc != 0 and nx >= 2 * halo + v
                    assert (counts[1] > 0) == (marched_possible and ny > 2), (nx, ny, nz, halo, zlo, zhi, rim_zc, list(counts))
                    if halo > 0 and nx > 2 * halo + 2 * v + 8 and nz > 2 * halo + 4 and (ny == 1 or ny > 2 * halo + 2):
                        assert counts[0] > 0  # a bulk exists


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the B200 arm) needs no GPU: it must print one JSON line with
    the contract's keys -- the B200 arm's metric / unit / config, `impl`, a `cpu_baseline` describing the run and an `e2e` block
    that repeats the value with zero host <-> device bytes."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"].startswith("Gcell-updates/s") and line["unit"] == "Gcell-updates/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["value"] > 0
    assert line["config"]["workload"].startswith("C2:") and line["dtype"] == "f32" and line["data"] == "synthetic"
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and "4096x4096" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
