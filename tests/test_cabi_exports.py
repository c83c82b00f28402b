"""CPU-side checks of the drop-in boundary: the shared library builds/loads and exports exactly the
entry points include/mdf_b200.h declares; the Python mirror keeps the reference's names and
state-dict keys; nothing in the product package touches the oracle.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

HEADER = os.path.join(ROOT, "include", "mdf_b200.h")              # the drop-in boundary
DEBUG_HEADER = os.path.join(ROOT, "include", "mdf_b200_debug.h")  # test / bench / tuning entry points


def declared_symbols(headers=(HEADER, DEBUG_HEADER)):
    names = set()
    for h in headers:
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"#ifdef MDF_TUNING.*?#endif", "", text, flags=re.S)      # tuning builds only
        names |= set(re.findall(r"MDF_API\s+[\w\s\*]*?\b(mdf_\w+)\s*\(", text))
    return sorted(names)


@pytest.fixture(scope="module")
def libpath():
    from mdf_net_b200 import build
    return build.build_library()


def test_header_declares_the_path():
    names = declared_symbols((HEADER,))
    assert not [n for n in names if n.endswith("_ex") or n.startswith("mdf_debug_")], "diagnostics belong in mdf_b200_debug.h"
    for required in ("mdf_cost_volume_fwd", "mdf_homo_warp_fwd", "mdf_variance_volume_fwd", "mdf_softmax_regress_fwd",
                     "mdf_depth_regression_fwd", "mdf_confidence_fwd", "mdf_cost_volume_workspace_bytes"):
        assert required in names


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/mdf_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("mdf_"))
    assert exported == declared_symbols(), "exported mdf_* symbols and the header must match one to one"


def test_python_signatures_cover_the_header(libpath):
    from mdf_net_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == declared_symbols()
    lib = _cabi.lib()
    assert lib.mdf_abi_version() == 1
    assert lib.mdf_status_string(0) == b"ok"
    # pure host queries (no device needed)
    assert lib.mdf_cost_volume_workspace_bytes(1, 5, 64, 32, 48, 144, 200) > 4 * 144 * 200 * 32 * 4
    assert lib.mdf_cost_volume_workspace_bytes(1, 1, 64, 32, 48, 144, 200) == 0     # N < 2
    assert lib.mdf_homo_warp_workspace_bytes(2) % 256 == 0


def test_argument_validation_without_a_device(libpath):
    """Shape / null checks run before any CUDA call, so they can be exercised on a CPU box."""
    from mdf_net_b200 import _cabi
    lib = _cabi.lib()
    assert lib.mdf_depth_regression_fwd(None, None, 0, 1, 8, 4, 4, None, None) == -3          # null output
    assert lib.mdf_confidence_fwd(None, 1, 8, 4, 4, 0, 1, 2, 1, ctypes.c_void_p(256), None) == -1   # n = 0
    assert lib.mdf_cost_volume_fwd(None, 1, None, None, None, 0, None, None, None, None, None, 1e-5, None, None,
                                   1, 16, 8, 8, 4, 4, None, None, 0, None) == -1            # N < 2
    assert lib.mdf_cost_volume_fwd(None, 3, None, None, None, 0, None, None, None, None, None, 1e-5, None, None,
                                   1, 15, 8, 8, 4, 4, None, None, 0, None) == -1            # C % G != 0
    assert lib.mdf_cost_volume_fwd(None, 3, None, None, None, 0, None, None, None, None, None, 1e-5, None, None,
                                   0, 16, 8, 8, 4, 4, None, None, 0, None) == 0             # empty batch: nothing to do
    assert lib.mdf_cost_volume_fwd(None, 40, None, None, None, 0, None, None, None, None, None, 1e-5, None, None,
                                   1, 16, 8, 8, 4, 4, None, None, 0, None) == -2            # > MDF_MAX_VIEWS
    P = ctypes.c_void_p(256)
    # fused tails: curve / s consistency, supported depths, empty inputs
    assert lib.mdf_softmax_regress_fit_fwd(P, P, 0, 1, 8, 4, 4, None, P, None, 4, 1, 2, 2, 3, P, None) == -2     # unknown curve
    assert lib.mdf_softmax_regress_fit_fwd(P, P, 0, 1, 8, 4, 4, None, P, None, 4, 1, 2, 2, 1, None, None) == -3  # curve without s
    assert lib.mdf_softmax_regress_fit_fwd(P, P, 0, 0, 8, 4, 4, None, None, None, 4, 1, 2, 2, 0, None, None) == 0  # empty
    assert lib.mdf_prob_head_fwd(P, P, P, 0, 1, 8, 10, 4, 4, None, None, P, None, 4, 1, 2, 2, 0, None, None) == -2   # D = 10
    assert lib.mdf_prob_head_fwd(P, P, P, 0, 1, 65, 8, 4, 4, None, None, P, None, 4, 1, 2, 2, 0, None, None) == -2   # C > 64
    assert lib.mdf_prob_head_fwd(P, P, P, 0, 1, 8, 8, 4, 4, None, None, None, None, 4, 1, 2, 2, 0, None, None) == -3  # no output
    assert lib.mdf_prob_head_fwd(P, P, P, 0, 1, 8, 8, 4, 4, None, None, P, None, 4, 1, 2, 2, 2, None, None) == -3     # curve without s
    assert lib.mdf_prob_head_fwd(None, None, None, 0, 0, 8, 8, 4, 4, None, None, None, None, 4, 1, 2, 2, 0, None, None) == 0
    # FPN hand-off (optional fast entry)
    assert lib.mdf_fpn_out_prepped_fwd(P, P, 1, 64, 12, 4, 4, None, P, None, None, None) == -2              # G not in {8,16,32}
    assert lib.mdf_fpn_out_prepped_fwd(P, P, 1, 62, 8, 4, 4, None, P, None, None, None) == -2               # Cin % 4
    assert lib.mdf_fpn_out_prepped_fwd(P, P, 1, 64, 8, 4, 4, None, None, P, None, None) == -3               # reference view without cq4 / conv weight
    assert lib.mdf_fpn_out_prepped_fwd(P, P, 1, 64, 8, 4, 4, None, P, P, None, None) == -3                  # source view with q4
    assert lib.mdf_fpn_out_prepped_fwd(None, None, 0, 64, 8, 4, 4, None, None, None, None, None) == 0       # empty batch
    assert lib.mdf_cost_volume_prepped_workspace_bytes(1, 5) % 256 == 0 and lib.mdf_cost_volume_prepped_workspace_bytes(1, 1) == 0
    assert lib.mdf_cost_volume_fwd_prepped(P, P, P, 1, None, None, None, 0, None, None, None, None, None, 1e-5, None, None,
                                           1, 8, 8, 4, 4, None, None, 0, None) == -1                        # N < 2
    assert lib.mdf_cost_volume_fwd_prepped(P, P, P, 3, None, None, None, 0, None, None, None, None, None, 1e-5, None, None,
                                           1, 12, 8, 4, 4, None, None, 0, None) == -2                       # G not in {8,16,32}
    assert lib.mdf_cost_volume_fwd_prepped(P, P, P, 3, None, None, None, 0, None, None, None, None, None, 1e-5, None, None,
                                           1, 8, 8, 4, 4, None, None, 0, None) == -3                        # null pointers
    # the debug entry: both timing events or none
    assert lib.mdf_cost_volume_fwd_ex(None, 1, None, None, None, 0, None, None, None, None, None, 1e-5, None, None,
                                      1, 16, 8, 8, 4, 4, None, None, 0, 0, None, None, None) == -1          # N < 2
    assert lib.mdf_cost_volume_fwd_ex(None, 3, None, None, None, 0, None, None, None, None, None, 1e-5, None, None,
                                      1, 16, 8, 8, 4, 4, None, None, 0, 16, None, None, None) in (-2, -3)   # tuning variants: not in the product build
    # running statistics of the train-mode BatchNorm
    assert lib.mdf_bn_running_update(None, 0, 0.1, None, None, None, None) == 0                 # no source view: nothing to do
    assert lib.mdf_bn_running_update(None, 3, 0.1, P, P, None, None) == -3                      # no batch statistics
    assert lib.mdf_bn_running_update(P, 99, 0.1, P, P, None, None) == -1                        # more views than the path supports
    # geometric filter
    assert lib.mdf_geo_filter_workspace_bytes(4) >= (20 + 4 * 64) * 4
    assert lib.mdf_geo_filter_fwd(P, P, P, None, None, None, 33, 4, 4, None, 0.8, 5, 4.0, 1300.0, None, None, P, None, None, None,
                                  P, 1 << 20, None) == -2                                   # > MDF_MAX_FILTER_VIEWS
    assert lib.mdf_geo_filter_fwd(P, P, P, None, None, None, 0, 4, 4, None, 0.8, 5, 4.0, 1300.0, None, None, None, None, None, None,
                                  P, 1 << 20, None) == -3                                   # no output requested
    assert lib.mdf_geo_filter_fwd(P, P, P, None, None, None, 0, 4, 4, None, 0.8, 5, 4.0, 1300.0, None, None, P, None, None, None,
                                  P, 16, None) == -4                                        # workspace too small
    assert lib.mdf_geo_filter_fwd(P, P, P, None, None, None, 0, 4, 4, None, 0.8, 5, 0.0, 1300.0, None, None, P, None, None, None,
                                  P, 1 << 20, None) == -2                                   # thresholds must be positive
    assert lib.mdf_geo_filter_fwd(None, None, None, None, None, None, 0, 0, 4, None, 0.8, 5, 4.0, 1300.0, None, None, None, None,
                                  None, None, None, 0, None) == 0                           # empty map


def test_cpu_tensors_are_a_hard_error(libpath):
    import mdf_net_b200 as mdf
    with pytest.raises((NotImplementedError, RuntimeError)):
        mdf.depth_regression(torch.zeros(1, 2, 3, 4), torch.zeros(1, 2, 1, 1))
    m = mdf.VectorAggregate(8).eval()
    with torch.no_grad(), pytest.raises((NotImplementedError, RuntimeError)):
        m([torch.zeros(1, 16, 4, 4)] * 2, torch.eye(4)[None], [torch.eye(4)[None]], torch.ones(1, 2, 1, 1))


def test_state_dict_keys_match_the_reference():
    """Keys captured from the reference's config.model.Homoaggre (tests/golden/make_golden.py)."""
    import mdf_net_b200 as mdf
    ref_keys = [str(k) for k in load_golden("corenet_64x64_n3")["homoaggre_state_keys"]]
    mods = torch.nn.ModuleList([mdf.VectorAggregate(g) for g in (32, 16, 8)])
    assert sorted(mods.state_dict().keys()) == sorted(ref_keys)
    assert mods[0].depth_weight[0].conv.weight.shape == (1, 32, 1, 1, 1)
    assert sum(p.numel() for p in mods[0].parameters()) == 36       # 36 / 20 / 12 parameters
    assert sum(p.numel() for p in mods[2].parameters()) == 12


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "mdf_net_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b|c_oracle|mdf_oracle_", text, re.M), f
