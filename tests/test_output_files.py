"""The output side (eval.py:36-50, tools/data_io.py:44-75): PFM / PNG writers byte-identical to the reference's own
(golden from its functions), and the asynchronous writer produces the same files as synchronous calls."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden


def test_writers_are_byte_identical_to_the_reference(tmp_path):
    from mdf_net_b200 import io
    z = load_golden("output_files")
    io.save_pfm(tmp_path / "d.pfm", torch.from_numpy(z["depth"]))
    io.write_depth_img(tmp_path / "d.png", z["depth"])
    io.save_pfm(tmp_path / "c.pfm", z["confidence"])
    rd = lambda n: np.frombuffer(open(tmp_path / n, "rb").read(), np.uint8)
    assert np.array_equal(rd("d.pfm"), z["depth_pfm"])
    assert np.array_equal(rd("d.png"), z["depth_png"])
    assert np.array_equal(rd("c.pfm"), z["confidence_pfm"])
    # the device-side quantisation of the depth image is PIL's "F" -> "L" conversion
    io.write_depth_img(tmp_path / "d2.png", io.depth_to_u8(torch.from_numpy(z["depth"])).numpy())
    assert np.array_equal(rd("d2.png"), z["depth_png"])
    with pytest.raises(Exception):
        io.save_pfm(tmp_path / "bad.pfm", z["depth"].astype(np.float64))


def _roundtrip(tmp_path, device):
    from mdf_net_b200 import io
    rng = np.random.default_rng(3)
    views = [(torch.from_numpy((700 + 100 * rng.standard_normal((2, 96, 128))).astype(np.float32)).to(device),
              torch.from_numpy(rng.random((2, 96, 128), dtype=np.float32)).to(device)) for _ in range(5)]
    with io.OutputWriter(workers=3, slots=2) as w:
        for i, (d, c) in enumerate(views):
            names = [str(tmp_path / f"a/scan{i}_{b}" ) for b in range(2)]
            w.submit([n + "_depth.pfm" for n in names], [n + "_depth.png" for n in names], [n + "_conf.pfm" for n in names], d, c)
    for i, (d, c) in enumerate(views):
        for b in range(2):
            io.save_pfm(tmp_path / "ref.pfm", d[b].cpu())
            assert open(tmp_path / f"a/scan{i}_{b}_depth.pfm", "rb").read() == open(tmp_path / "ref.pfm", "rb").read()
            io.save_pfm(tmp_path / "ref.pfm", c[b].cpu())
            assert open(tmp_path / f"a/scan{i}_{b}_conf.pfm", "rb").read() == open(tmp_path / "ref.pfm", "rb").read()
            io.write_depth_img(tmp_path / "ref.png", d[b].cpu().numpy())
            assert open(tmp_path / f"a/scan{i}_{b}_depth.png", "rb").read() == open(tmp_path / "ref.png", "rb").read()


def test_async_writer_on_host_tensors(tmp_path):
    _roundtrip(tmp_path, "cpu")


@pytest.mark.gpu
def test_async_writer_overlaps_device_work(tmp_path):
    _roundtrip(tmp_path, "cuda")
