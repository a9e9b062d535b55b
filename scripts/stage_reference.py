#!/usr/bin/env python
"""scripts/stage_reference.py -- stage the UNMODIFIED reference detector where the GPU box can see it.

    python scripts/stage_reference.py            # /root/reference -> baseline/_ref  (git-ignored, NOT gpurun-ignored)

The GPU box has no /root/reference; `gpurun` ships the repo directory, including git-ignored files.  This copies the
Python sources the two-stream detector needs (models/, utils/, global_var.py, the hyper-parameter YAMLs) byte for byte
into baseline/_ref/ -- outside git history, exactly like the `pip install --target baseline/_ref` of the base contract
(the reference has no setup.py, so there is nothing for pip to install).  Nothing under baseline/_ref is ever imported
by the product package except through `mmidet_b200.harness`, which treats it as the external reference checkout it is
(SURVEY 7.2 step 1 / App. C).  A MANIFEST with sha256 sums is written so a test can prove the staged tree is unmodified."""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MMIDET_REF", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
KEEP_DIRS = ("models", "utils")
KEEP_FILES = ("global_var.py", "data/hyp.scratch.yaml", "data/hyp.finetune.yaml", "LICENSE.txt")


def main():
    if not os.path.isdir(SRC):
        print(f"{SRC} not found: nothing staged (the GPU box uses the copy shipped in baseline/_ref)")
        return 0 if os.path.isdir(DST) else 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "*.jpg", "*.png", "*.pt", "wandb_logging", "aws", "google_app_engine",
                                    "flask_rest_api")
    for d in KEEP_DIRS:
        shutil.copytree(os.path.join(SRC, d), os.path.join(DST, d), ignore=ignore)
    for f in KEEP_FILES:
        if os.path.exists(os.path.join(SRC, f)):
            os.makedirs(os.path.dirname(os.path.join(DST, f)), exist_ok=True)
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    lines = []
    for base, _, files in sorted(os.walk(DST)):
        for fn in sorted(files):
            p = os.path.join(base, fn)
            lines.append(f"{hashlib.sha256(open(p, 'rb').read()).hexdigest()}  {os.path.relpath(p, DST)}")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(lines) + "\n")
    size = sum(os.path.getsize(os.path.join(b, fn)) for b, _, fs in os.walk(DST) for fn in fs)
    print(f"staged {len(lines)} files, {size / 1024:.0f} KiB -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
