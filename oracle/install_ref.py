"""Stage the UNMODIFIED reference where the GPU box can import it: baseline/_ref/ (git-ignored, travels with gpurun).

TEST / BENCH INFRASTRUCTURE ONLY (bench.py --impl reference, cpu_baseline kind "reference").  Nothing under
dl-dkd_b200/ reads it.  The recipe follows the driver contract: `pip install --target baseline/_ref /root/reference`
first; the reference has neither setup.py nor pyproject.toml, so pip refuses it and the fallback stages the two
pure-Python packages the eval path imports (method/, utils/: *.py only, byte-identical copies, never committed).
"""
import filecmp
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
SRC = os.environ.get("DKD_REFERENCE_SRC", "/root/reference")
PACKAGES = ("method", "utils")


def staged() -> bool:
    return os.path.isfile(os.path.join(DEST, "method", "eval.py"))


def install(verbose=False) -> str:
    """Returns a one-line outcome.  No-op when the reference source tree is absent (the GPU box)."""
    if not os.path.isdir(os.path.join(SRC, "method")):
        return "reference source absent: using the staged copy" if staged() else "reference source absent, nothing staged"
    os.makedirs(DEST, exist_ok=True)
    outcome = "pip: not attempted"
    if not staged():
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", DEST, SRC],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        outcome = "pip install ok" if r.returncode == 0 else "pip install refused the tree (no setup.py / pyproject.toml)"
        if verbose:
            print(r.stdout[-400:])
    if not staged():
        for pkg in PACKAGES:
            for d, _, files in os.walk(os.path.join(SRC, pkg)):
                if "__pycache__" in d:
                    continue
                out = os.path.join(DEST, os.path.relpath(d, SRC))
                os.makedirs(out, exist_ok=True)
                for f in files:
                    if f.endswith(".py"):
                        shutil.copy2(os.path.join(d, f), os.path.join(out, f))
        outcome += "; staged method/ + utils/ (*.py) by copy"
    same = all(filecmp.cmp(os.path.join(SRC, p, f), os.path.join(DEST, p, f), shallow=False)
               for p in PACKAGES for f in os.listdir(os.path.join(DEST, p)) if f.endswith(".py"))
    return outcome + ("; byte-identical to the source tree" if same else "; WARNING: staged files differ from the source tree")


if __name__ == "__main__":
    print(install(verbose=True))
