"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/ndt_b200.h declares,
the POD layouts agree between the header and the ctypes/numpy views, and -- with no GPU -- the
product refuses to run instead of falling back."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from ndt_slam_b200 import build, capi

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    txt = (ROOT / "include" / "ndt_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ndt_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = capi.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ndt_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names
    assert b"sm_100a" in L.ndt_version()


def test_struct_layouts_match_header():
    assert C.sizeof(capi.NdtResult) == capi.RESULT_DTYPE.itemsize == 208
    for name, off in [("pose", 0), ("T", 24), ("score", 88), ("trans_prob", 96), ("fitness", 104), ("hess", 112),
                      ("converged", 184), ("iters", 188), ("evals", 192), ("point_evals", 200)]:
        assert getattr(capi.NdtResult, name).offset == off
        assert capi.RESULT_DTYPE.fields[name][1] == off
    assert C.sizeof(capi.NdtEvalOut) == 8 * 13 + 8
    assert C.sizeof(capi.NdtGridInfo) == 40


def test_pod_layouts_match_what_a_c_compiler_sees(tmp_path):
    """gcc compiles include/ndt_b200.h as plain C and prints sizeof / offsetof of every POD: the ctypes views must agree."""
    import subprocess
    fields = {"ndt_params": capi.NdtParams, "ndt_result": capi.NdtResult, "ndt_eval_out": capi.NdtEvalOut, "ndt_grid_info": capi.NdtGridInfo}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ndt_b200.h"', 'int main(void) {']
    for cname, cls in fields.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in fields.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(out[f"{cname}.{f}"]) == getattr(cls, f).offset, (cname, f)


def test_default_params_are_the_reference_defaults():
    """PoseEstimator.h:63-64 C++ defaults + PCL internals."""
    p = capi.default_params()
    assert (p.resolution, p.step_size, p.trans_eps, p.max_iter) == (1.0, 0.1, 0.01, 35)
    assert (p.outlier_ratio, p.min_points, p.eig_mult) == (0.55, 6, 0.01)
    assert p.quirks == capi.QUIRKS_PCL_1_10
    assert (p.align_skip_fitness, p.pairs_schedule, p.pairs_batch_points, p.align_team) == (0, capi.PAIRS_AUTO, 0, 0)


def test_sass_is_sm100a_only():
    out = build._run(["cuobjdump", "-lelf", str(build.LIB_CUDA)])
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(capi.NdtError) as ei:
        capi.Ndt(capi.default_params())
    assert "no CUDA device" in str(ei.value) or "-3" in str(ei.value)


def test_product_never_references_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    for p in (ROOT / "ndt_slam_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".h", ".cpp"} and "build.py" != p.name:
            txt = p.read_text(errors="ignore")
            assert "oracle_api" not in txt and "libndt_oracle" not in txt and "ndt_oracle" not in txt, p
