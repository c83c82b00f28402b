"""Output side of the path (second half): the files eval.py:36-50 writes per reference view -- depth PFM,
depth PNG, confidence PFM (tools/data_io.py:44-75) -- written WHILE the GPU computes the next view.

`save_pfm` / `write_depth_img` keep the reference's names, arguments and bytes (golden: tests/golden/output_files.npz from
the reference's own functions).  `OutputWriter` is the overlap: the reference copies each map to the host with a blocking
`.cpu()` and writes three files before the next forward starts (eval.py:46-49); here the device-to-host copies go to pinned
buffers on a side stream behind an event, and a small thread pool does the flips / PNG compression / file writes, so the
stage loop never waits for the disk.  The 8-bit depth image is quantised on the device ((d - 500) / 2, clipped, truncated:
what PIL's "F" -> "L" conversion does), which shrinks that copy by 4x.
"""
from __future__ import annotations

import os
import queue
import sys
import threading
from typing import Sequence

import numpy as np
import torch

__all__ = ["save_pfm", "write_depth_img", "OutputWriter"]


def save_pfm(filename: str, image, scale: float = 1) -> None:
    """tools/data_io.py:44-71, byte for byte: 'Pf' / 'PF' header, 'W H', the scale (negative = little endian) as '%f',
    then the rows bottom-up as raw float32."""
    image = np.asarray(image.detach().cpu().numpy() if isinstance(image, torch.Tensor) else image)
    image = np.flipud(image)
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if len(image.shape) == 3 and image.shape[2] == 3:
        color = True
    elif len(image.shape) == 2 or len(image.shape) == 3 and image.shape[2] == 1:
        color = False
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    endian = image.dtype.byteorder
    if endian == "<" or endian == "=" and sys.byteorder == "little":
        scale = -scale
    with open(filename, "wb") as f:
        f.write(b"PF\n" if color else b"Pf\n")
        f.write("{} {}\n".format(image.shape[1], image.shape[0]).encode("utf-8"))
        f.write(("%f\n" % scale).encode("utf-8"))
        np.ascontiguousarray(image).tofile(f)


def depth_to_u8(depth):
    """(depth - 500) / 2 as PIL converts mode "F" to "L" (tools/data_io.py:72-73): clipped to [0, 255], truncated."""
    if isinstance(depth, torch.Tensor):
        return ((depth - 500) / 2).clamp_(0, 255).to(torch.uint8)
    return np.clip((np.asarray(depth, np.float32) - 500) / 2, 0, 255).astype(np.uint8)


def write_depth_img(filename: str, depth) -> int:
    """tools/data_io.py:72-75.  `depth` float32 (H, W) as in the reference, or the already quantised uint8 image."""
    from PIL import Image
    depth = np.asarray(depth.detach().cpu().numpy() if isinstance(depth, torch.Tensor) else depth)
    img = depth if depth.dtype == np.uint8 else depth_to_u8(depth)
    Image.fromarray(img, mode="L").save(filename)
    return 1


class OutputWriter:
    """Asynchronous version of eval.py:36-50.  submit(depth_files, png_files, confidence_files, depth, confidence) queues the
    three files of every batch item; device tensors are copied to pinned host buffers on a side stream (after the work
    queued so far on the current stream) and written by `workers` threads.  close() waits for everything."""

    def __init__(self, workers: int = 4, slots: int = 4):
        self.q: "queue.Queue" = queue.Queue(maxsize=slots)
        self.errors = []
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.threads = [threading.Thread(target=self._run, daemon=True) for _ in range(workers)]
        for t in self.threads:
            t.start()

    def _run(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            try:
                ev, items = job
                if ev is not None:
                    ev.synchronize()
                for kind, path, arr in items:
                    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
                    if kind == "pfm":
                        save_pfm(path, arr.numpy())
                    else:
                        write_depth_img(path, arr.numpy())
            except Exception as e:  # pragma: no cover - reported by close()
                self.errors.append(e)
            finally:
                self.q.task_done()

    def _to_host(self, t: torch.Tensor) -> torch.Tensor:
        if not t.is_cuda:
            return t.detach()
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        return host

    def submit(self, depth_files: Sequence[str], png_files: Sequence[str], confidence_files: Sequence[str],
               depth: torch.Tensor, confidence: torch.Tensor) -> None:
        ev = None
        if depth.is_cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                u8 = depth_to_u8(depth)
                for t in (depth, confidence, u8):
                    t.record_stream(self.stream)
                hd, hc, hu = self._to_host(depth), self._to_host(confidence), self._to_host(u8)
                ev = torch.cuda.Event()
                ev.record(self.stream)
        else:
            hd, hc, hu = depth.detach(), confidence.detach(), depth_to_u8(depth)
        for b in range(hd.shape[0]):
            self.q.put((ev, [("pfm", depth_files[b], hd[b]), ("png", png_files[b], hu[b]), ("pfm", confidence_files[b], hc[b])]))

    def close(self) -> None:
        self.q.join()
        for _ in self.threads:
            self.q.put(None)
        for t in self.threads:
            t.join()
        if self.errors:
            raise self.errors[0]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
