"""The C-ABI library loads on a CPU-only box and exports every symbol include/mlo_b200.h declares
(no compute call is made here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    txt = (ROOT / "include" / "mlo_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mlo_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built):
    from mola_lidar_odometry_b200 import capi
    lib = ctypes.CDLL(str(capi.LIB_PATH))
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in mlo_b200.h but not exported: {missing}"


def test_binding_covers_header(built):
    from mola_lidar_odometry_b200 import capi
    assert set(_declared()) == set(capi.declared_symbols())
    lib = capi.load()
    assert lib.mlo_abi_version() == 1


def test_struct_sizes_match_header(built, tmp_path):
    """sizeof() of every POD struct as the C compiler sees it == the ctypes mirror."""
    import subprocess
    from mola_lidar_odometry_b200 import capi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "mlo_b200_host.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(mlo_map_params),sizeof(mlo_decimate_params),sizeof(mlo_filter1_params),'
                   'sizeof(mlo_icp_params),sizeof(mlo_icp_result),sizeof(mlo_profile),'
                   'sizeof(mlo_scan_job),sizeof(mlo_scan_info),sizeof(mlo_align_job),sizeof(mlo_insert_job),'
                   'sizeof(mlo_map_counts),sizeof(mlo_lo_scan_output),sizeof(mlo_icp_iteration_record));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    from mola_lidar_odometry_b200 import host_api
    want = [ctypes.sizeof(t) for t in (capi.MapParams, capi.DecimateParams, capi.Filter1Params, capi.IcpParams,
                                       capi.IcpResult, capi.Profile, capi.ScanJob, capi.ScanInfo, capi.AlignJob,
                                       capi.InsertJob, capi.MapCounts, host_api.ScanOutput, capi.IcpIterationRecord)]
    assert host_api.SCAN_OUTPUT_DTYPE.itemsize == ctypes.sizeof(host_api.ScanOutput)
    assert got == want


def test_no_device_fails_loudly(built):
    """Without a CUDA device the product path refuses to run (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mola_lidar_odometry_b200.api import Context, MloError
    with pytest.raises(MloError):
        Context(0)


def test_host_math_helpers(built):
    """mlo_voxel_index is a pure host function: same rounding as the device (truncation toward zero)."""
    from mola_lidar_odometry_b200 import capi
    lib = capi.load()
    assert lib.mlo_voxel_index(1.99, 1.0) == 1
    assert lib.mlo_voxel_index(-0.5, 1.0) == 0
    assert lib.mlo_voxel_index(-1.0, 1.0) == -1
    assert lib.mlo_voxel_index(-1.5, 0.5) == -3
