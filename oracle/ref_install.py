"""The reference's own model code on the GPU box -- TEST / BENCH INFRASTRUCTURE, never a product path.

`install()` copies the reference's model package (net/*.py, net/unit/*.py, config.py: ~1 200 lines of Python, no
native code) UNMODIFIED from /root/reference into the git-ignored `oracle/_ref/`.  Ignored files travel with the
`gpurun` snapshot, so the unmodified `VectorAggregate` / `homo_warping` / `regress.*` / `HyposByFit` / `CoreNet` can
run on the B200 box (ATen / cuDNN) and on its host cores (torch CPU):

  * `tests/test_gpu_reference.py`   parity of the CUDA path against the reference modules at full size,
  * `bench.py --impl reference`     the reference's hot path on torch CPU (cpu_baseline.kind = "reference"),
  * `bench.py`                      `aten_cuda_baseline` and `pipeline` (FPN / 3-D CNN "timed separately", north_star).

Nothing is committed: `oracle/ref_manifest.json` (committed) holds the sha256 of every file of the reference checkout
this repo was built against, and `verify()` checks the shipped copy against it -- "unmodified" is testable on the box.
Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module (tests/test_cabi_exports.py
greps the product package for violations).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(HERE, "ref_manifest.json")
FILES = ["config.py", "net/__init__.py", "net/core.py", "net/loss.py", "net/unit/__init__.py", "net/unit/backbone.py",
         "net/unit/base.py", "net/unit/depthhypos.py", "net/unit/homoaggregate.py", "net/unit/refine.py",
         "net/unit/regress.py", "net/unit/regular.py", "net/unit/scale.py"]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def install(force: bool = False) -> bool:
    """Copy the files (only where /root/reference exists: the build container).  Returns available()."""
    if os.path.isdir(REF_SRC) and (force or not available()):
        for rel in FILES:
            dst = os.path.join(REF_DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(REF_SRC, rel), dst)
    return available()


def write_manifest() -> None:
    """(build container, once) record the hashes of the reference checkout."""
    with open(MANIFEST, "w") as f:
        json.dump({rel: _sha(os.path.join(REF_SRC, rel)) for rel in FILES}, f, indent=1, sort_keys=True)
        f.write("\n")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DST, rel)) for rel in FILES)


def verify() -> None:
    """Raises if the shipped copy differs from the recorded reference checkout."""
    want = json.load(open(MANIFEST))
    for rel in FILES:
        got = _sha(os.path.join(REF_DST, rel))
        if got != want[rel]:
            raise RuntimeError(f"oracle/_ref/{rel} differs from the reference checkout recorded in ref_manifest.json")


def modules():
    """Import the reference's unit modules (no `import config`: that module has import-time side effects).
    Returns a namespace with core, base, homoaggregate, regress, regular, backbone, depthhypos, refine, scale."""
    if not available():
        raise ImportError("oracle/_ref is missing: run `python -m oracle.ref_install` in the build container "
                          "(needs /root/reference); the GPU box receives it with the gpurun snapshot")
    verify()
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    import types
    from net import core
    from net.unit import backbone, base, depthhypos, homoaggregate, refine, regress, regular, scale
    return types.SimpleNamespace(core=core, base=base, homoaggregate=homoaggregate, regress=regress, regular=regular,
                                 backbone=backbone, depthhypos=depthhypos, refine=refine, scale=scale)


def config_model():
    """`config.model` exactly as eval.py:12 obtains it (config.py:186-218 assembles it at import time).  The import
    overwrites CUDA_VISIBLE_DEVICES / CUDA_DEVICE_ORDER and silences warnings (config.py:2-8); the environment is put back."""
    if not available():
        raise ImportError("oracle/_ref is missing (see oracle/ref_install.py)")
    verify()
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    keep = {k: os.environ.get(k) for k in ("CUDA_VISIBLE_DEVICES", "CUDA_DEVICE_ORDER")}
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):       # config.py prints at import; bench.py's stdout carries one JSON line
        import config
    for k, v in keep.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    return config.model


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    if "--manifest" in sys.argv:
        write_manifest()
    if ok:
        verify()
    print(f"oracle/_ref: {'installed and verified' if ok else 'NOT available (no /root/reference here)'}")
