"""CPU checks of bench.py's bookkeeping: the roofline numerator is SURVEY 8d's formula (BASELINE.md totals), the
workload is BASELINE.json's config, the reference arm prints the contract's JSON line."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mdf_net_b200 import synthetic as syn  # noqa: E402


def test_algorithmic_bytes_match_baseline_md():
    cv, head = bench.algorithmic_bytes(1152, 1600, 5, 1)
    assert [round(b / 1e6, 1) for b in cv] == [213.8, 261.7, 280.2]
    assert round(sum(cv) / 1e6, 1) == 755.7
    cv1, _ = bench.algorithmic_bytes(512, 640, 3, 1)
    assert round(sum(cv1) / 1e6, 1) == 116.0                                   # C1
    cv4, _ = bench.algorithmic_bytes(1056, 1920, 7, 1)
    assert round(sum(cv4) / 1e6, 1) == 944.8                                   # C4
    ob = bench.onchip_bound(1152, 1600, 5, 1, 1965.0, 0.55)
    assert ob["group_evaluations_per_step"] == 471859200                       # 471.9 M


def test_default_workload_is_the_metric_config():
    cfg = bench.make_config("dtu_1600x1152_n5")
    assert cfg["views"] == 5 and cfg["batch"] == 1
    assert cfg["stages"] == ["144x200 C64 D48 G32", "288x400 C32 D24 G16", "576x800 C16 D8 G8"]
    assert bench.LAUNCHES_PER_STEP == 12


def test_scene_hypotheses_are_reference_shaped():
    h = syn.scene_hypos(1, 8, 144, 200, seed=3)
    assert h.shape == (1, 8, 144, 200) and h.dtype == np.float32
    assert (np.diff(h, axis=1) >= 0).all() and h.min() >= 425.0 and h.max() <= 935.0      # ascending, clamped
    rng = h[:, -1] - h[:, 0]
    assert rng.max() <= 30.0 + 1e-3                                            # the search range requested
    # built at half resolution and bilinearly upsampled x2, like depthhypos.py:49-52
    assert np.abs(np.diff(h[0, 0], axis=1)).mean() < 4.0        # smooth surfaces, a few silhouettes


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "dtu_640x512_n3",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "views/s" and line["higher_is_better"] is True
    assert line["config"] == bench.make_config("dtu_640x512_n3")
    from oracle import ref_install
    # the unmodified reference on torch CPU wherever its copy is present (build container, GPU box); the C port otherwise
    assert line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_install.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and line["value"] > 0
    assert line["cpu_baseline"]["value"] == line["value"]


def test_reference_copy_is_the_unmodified_reference():
    """oracle/_ref (git-ignored, travels with the gpurun snapshot) hashes to the committed manifest of /root/reference."""
    from oracle import ref_install
    if not ref_install.available():
        import pytest
        pytest.skip("oracle/_ref not installed here")
    ref_install.verify()
    import json as js
    manifest = js.load(open(os.path.join(ROOT, "oracle", "ref_manifest.json")))
    assert sorted(manifest) == sorted(ref_install.FILES)
    if os.path.isdir(ref_install.REF_SRC):        # build container: the manifest itself matches the checkout
        for rel in ref_install.FILES:
            assert ref_install._sha(os.path.join(ref_install.REF_SRC, rel)) == manifest[rel]


def test_chain_logits_make_scene_like_hypotheses():
    """bench.py's end-to-end leg forms the stage-1/2 hypotheses on the device from synthetic logits: the oracle chain on the
    same logits gives smooth, scene-like hypotheses (not the adversarial i.i.d. case)."""
    from oracle import c_oracle as co
    h0, w0 = 256, 320
    drange = np.array([[425.0, 935.0]], np.float32)
    hyp, depth, prob = syn.uniform_hypos(1, 48), None, None
    curves, th = (None, "gauss1", "laplace"), (0.0, 0.95, 1e-5)
    for s in range(3):
        H, W = syn.stage_shapes(h0, w0)[s]
        D = syn.STAGE_DEPTHS[s]
        if s > 0:
            hyp = co.hypos_generate(depth, co.hypos_fit(prob, hyp, depth, curves[s]), drange, curves[s], th[s], D, True)
            assert hyp.shape == (1, D, H, W) and (np.diff(hyp, axis=1) >= 0).all()
            rng = hyp[:, -1] - hyp[:, 0]
            assert 4.0 < np.median(rng) < 60.0
            assert np.median(np.abs(np.diff(hyp[0, D // 2], axis=1))) < 2.0       # neighbouring pixels see the same surface
        prob = co.softmax_depth(syn.scene_logits(1, s, H, W, seed=6))
        depth = co.depth_regression(prob, hyp)
