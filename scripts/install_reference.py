"""Install the UNMODIFIED reference into baseline/_ref so that the GPU box can time its own torch CPU path.

    python scripts/install_reference.py [--src /root/reference]

The reference is a flat directory of Python modules without setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` has nothing to build (tried first, outcome printed).  The fallback is
what a `--target` install of a pure-Python project does: the modules of the detector path are placed, byte for byte, under
baseline/_ref/.  baseline/_ref/ is git-ignored (it is not product source and never enters the history) but travels to the
GPU box with the working tree, where `bench.py --impl reference` and the `cpu_reference_torch` leg import it.  Nothing in the
package, in tests/ or in the GPU legs of bench.py reads it.
"""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
# the modules the detector path imports (bamp.py / scamp.py / vamp.py and what they pull in); drivers, plotting and the
# stored simulation results stay behind
MODULES = ["config.py", "channel.py", "data.py", "loss.py", "bamp.py", "scamp.py", "vamp.py", "shrink.py"]


def install(src="/root/reference", quiet=False):
    say = (lambda *a: None) if quiet else (lambda *a: print("[install_reference]", *a, file=sys.stderr))
    if not os.path.isdir(src):
        say(f"{src} is absent (GPU box): keeping whatever baseline/_ref already holds")
        return os.path.isdir(DEST)
    os.makedirs(DEST, exist_ok=True)
    if os.path.exists(os.path.join(src, "setup.py")) or os.path.exists(os.path.join(src, "pyproject.toml")):
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", DEST, src]
        rc = subprocess.run(cmd, capture_output=True, text=True)
        say("pip install --target:", "ok" if rc.returncode == 0 else rc.stderr.strip().splitlines()[-1:])
        if rc.returncode == 0:
            return True
    else:
        say("the reference has no setup.py / pyproject.toml: pip has nothing to install; placing its modules unmodified")
    sums = []
    for m in MODULES:
        shutil.copyfile(os.path.join(src, m), os.path.join(DEST, m))
        sums.append(f"{hashlib.sha256(open(os.path.join(DEST, m), 'rb').read()).hexdigest()}  {m}")
    with open(os.path.join(DEST, "SHA256SUMS"), "w") as f:
        f.write("\n".join(sums) + "\n")
    say(f"{len(MODULES)} modules -> {os.path.relpath(DEST, ROOT)}")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    install(ap.parse_args().src)
