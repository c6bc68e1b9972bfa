"""C-ABI library: loads, exports every symbol include/quatro_b200.h declares, and fails loudly without a GPU
(no CPU fallback).  No compute calls here -- those are the -m gpu parity tests."""
import ctypes as C
import re
from pathlib import Path

import pytest

from quatro_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    txt = (ROOT / "include" / "quatro_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qb200_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_are_exported():
    lib = capi.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/quatro_b200.h but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == names, "capi.py binding list and header disagree"
    assert lib.qb200_version() == 100


def test_default_params_match_python_mirror():
    lib = capi.load_library()
    p = capi.Params()
    lib.qb200_default_params(C.byref(p))
    q = capi.default_params()
    assert bytes(p) == bytes(q)
    # config/params.yaml values
    assert (round(p.voxel_size, 6), round(p.normal_radius, 6), round(p.fpfh_radius, 6)) == (0.3, 0.5, 0.75)
    assert (p.noise_bound, p.cbar2, p.rotation_max_iterations, p.rotation_gnc_factor, p.rotation_cost_threshold) == (0.3, 1.0, 50, 1.4, 0.00011)
    assert p.inlier_selection_mode == capi.PMC_HEU and p.cote_mode == capi.COTE_MEDIAN
    c = capi.Config()
    lib.qb200_default_config(C.byref(c))
    assert bytes(c) == bytes(capi.default_config())


def test_struct_layouts():
    assert C.sizeof(capi.Result) == 56 + 8 + 128 + 0 or C.sizeof(capi.Result) == capi.RESULT_DTYPE.itemsize
    assert C.sizeof(capi.Pair) == 24
    assert C.sizeof(capi.Params) % 8 == 0


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the failure path is exercised on the CPU-only box")
    with pytest.raises(capi.QuatroB200Error) as e:
        capi.Handle()
    assert e.value.code == -2  # QB200_ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    """The product path must never import, link or call the test oracle."""
    for f in list((ROOT / "quatro_b200").rglob("*.py")) + list((ROOT / "quatro_b200" / "csrc").glob("*")) + list((ROOT / "include").rglob("*.h*")):
        txt = f.read_text(errors="ignore")
        assert "quatro_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt and "qo_" not in txt.replace("qo_math", ""), f


def test_pod_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of every POD of include/quatro_b200.h have the size and field offsets a C compiler gives them."""
    import subprocess
    from quatro_b200 import capi
    src = tmp_path / "pod.c"
    src.write_text('''
#include <stddef.h>
#include <stdio.h>
#include "quatro_b200.h"
#define S(t) printf("%s %zu\\n", #t, sizeof(t))
#define O(t, f) printf("%s.%s %zu\\n", #t, #f, offsetof(t, f))
int main(void) {
  S(qb200_params); S(qb200_config); S(qb200_result); S(qb200_pair); S(qb200_patchwork_params); S(qb200_segment_params);
  O(qb200_params, seed); O(qb200_params, noise_bound); O(qb200_params, max_clique_node_limit); O(qb200_params, RyRx);
  O(qb200_result, flags); O(qb200_result, n_edges); O(qb200_result, T);
  O(qb200_patchwork_params, min_ranges_each_zone); O(qb200_patchwork_params, num_iter); O(qb200_patchwork_params, num_rings_each_zone);
  O(qb200_segment_params, segment_theta); O(qb200_segment_params, segment_valid_line_num);
  return 0;
}
''')
    exe = tmp_path / "pod"
    r = subprocess.run(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    mirror = {"qb200_params": capi.Params, "qb200_config": capi.Config, "qb200_result": capi.Result, "qb200_pair": capi.Pair,
              "qb200_patchwork_params": capi.PatchworkParams, "qb200_segment_params": capi.SegmentParams}
    for name, val in got.items():
        if "." in name:
            t, f = name.split(".")
            assert getattr(mirror[t], f).offset == int(val), name
        else:
            assert C.sizeof(mirror[name]) == int(val), name
