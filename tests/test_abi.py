"""The C-ABI library loads on a CPU-only box, exports every symbol include/fcapp.h declares,
and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "fcapp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from freecappuccino_b200 import lib
    h = lib.load()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/fcapp.h but not exported"
    assert set(lib.SYMBOLS) == set(names)


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors of the ABI structs have the sizes / field offsets the C compiler gives the header."""
    import subprocess
    from freecappuccino_b200 import lib
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "fcapp.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %d %zu %zu %zu %zu\\n", sizeof(fc_mesh_desc),'
        ' sizeof(fc_solver_opts),'
        ' sizeof(fc_solver_report), sizeof(fc_calcp_opts), sizeof(fc_calcp_report), sizeof(fc_timings),'
        ' offsetof(fc_mesh_desc, gloCells), offsetof(fc_calcp_opts, sol), offsetof(fc_timings, launches),'
        ' (int)FC_NUM_FIELDS, sizeof(fc_calcuvw_opts), sizeof(fc_calcuvw_report), offsetof(fc_calcuvw_opts, sol),'
        ' offsetof(fc_calcuvw_opts, viscos));return 0;}\n')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(lib.MeshDesc), C.sizeof(lib.SolverOpts), C.sizeof(lib.SolverReport), C.sizeof(lib.CalcpOpts),
            C.sizeof(lib.CalcpReport), C.sizeof(lib.Timings), lib.MeshDesc.gloCells.offset, lib.CalcpOpts.sol.offset,
            lib.Timings.launches.offset, len(lib.FIELDS), C.sizeof(lib.CalcuvwOpts), C.sizeof(lib.CalcuvwReport),
            lib.CalcuvwOpts.sol.offset, lib.CalcuvwOpts.viscos.offset]
    assert got == want


def test_enum_constants_match_header(tmp_path):
    """Tuning keys, gradient / limiter ids, solver ids and error codes of lib.py are the header's values."""
    import subprocess
    from freecappuccino_b200 import lib
    names = ["FC_TUNE_SPMV_KERNEL", "FC_TUNE_DPCG_PERSISTENT", "FC_TUNE_CTAS_PER_SM", "FC_TUNE_PIPE_GEOMETRY",
             "FC_TUNE_SWEEP_P2P", "FC_TUNE_SWEEP_TILED", "FC_TUNE_FUSED_GRAD", "FC_TUNE_TILE_CTAS", "FC_TUNE_SWEEP_CHECK", "FC_DPCG", "FC_ICCG", "FC_BICGSTAB",
             "FC_OK", "FC_ERR_ARG", "FC_ERR_CUDA", "FC_ERR_NCCL", "FC_ERR_UNSUPPORTED", "FC_ERR_NODEVICE",
             "FC_VIS", "FC_SP", "FC_FLMASS", "FC_USER3"]
    src = tmp_path / "enums.c"
    src.write_text('#include <stdio.h>\n#include "fcapp.h"\nint main(void){printf("' + " ".join(["%d"] * len(names)) +
                   '\\n", ' + ", ".join(f"(int){n}" for n in names) + ");return 0;}\n")
    exe = tmp_path / "enums"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(zip(names, (int(x) for x in subprocess.check_output([str(exe)], text=True).split())))
    want = {"FC_TUNE_SPMV_KERNEL": lib.TUNE_SPMV_KERNEL, "FC_TUNE_DPCG_PERSISTENT": lib.TUNE_DPCG_PERSISTENT,
            "FC_TUNE_CTAS_PER_SM": lib.TUNE_CTAS_PER_SM, "FC_TUNE_PIPE_GEOMETRY": lib.TUNE_PIPE_GEOMETRY,
            "FC_TUNE_SWEEP_P2P": lib.TUNE_SWEEP_P2P, "FC_TUNE_SWEEP_TILED": lib.TUNE_SWEEP_TILED,
            "FC_TUNE_FUSED_GRAD": lib.TUNE_FUSED_GRAD, "FC_TUNE_TILE_CTAS": lib.TUNE_TILE_CTAS, "FC_TUNE_SWEEP_CHECK": lib.TUNE_SWEEP_CHECK, "FC_DPCG": lib.DPCG, "FC_ICCG": lib.ICCG,
            "FC_BICGSTAB": lib.BICGSTAB, "FC_OK": lib.FC_OK, "FC_ERR_ARG": lib.FC_ERR_ARG, "FC_ERR_CUDA": lib.FC_ERR_CUDA,
            "FC_ERR_NCCL": lib.FC_ERR_NCCL, "FC_ERR_UNSUPPORTED": lib.FC_ERR_UNSUPPORTED,
            "FC_ERR_NODEVICE": lib.FC_ERR_NODEVICE, "FC_VIS": lib.F["VIS"], "FC_SP": lib.F["SP"],
            "FC_FLMASS": lib.F["FLMASS"], "FC_USER3": lib.F["USER3"]}
    assert got == want


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from freecappuccino_b200 import lib
    with pytest.raises(lib.FcError) as e:
        lib.Context(0)
    assert e.value.code == lib.FC_ERR_NODEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under freecappuccino_b200/ may reference it."""
    pkg = os.path.join(ROOT, "freecappuccino_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "fc_oracle" not in src, f
                # nor the host build of the kernel bodies the CPU suite uses (tests/kernel_bodies_host)
                assert "fcm_host" not in src and "libfcm_host" not in src, f
