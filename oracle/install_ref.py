"""TEST INFRASTRUCTURE ONLY -- makes the UNMODIFIED reference modules of the path travel to the GPU box.

The reference is a pure-Python research tree without setup.py / pyproject: there is nothing to `pip install`.
The seven files the path consists of are copied verbatim from /root/reference into
``baseline/_ref/`` (git-ignored, NOT gpurun-ignored: it ships with the snapshot like the built .so), keeping their
relative layout so that ``oracle/ref_loader.py`` can be pointed at either root.  Nothing is copied into tracked
paths, and the product package never looks here.

    python oracle/install_ref.py          # dev container only (needs /root/reference)
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "infty-Video-LLaMA/InfVideoLLaMA/models/long_term_attention_gibbs.py",
    "infty-Video-LLaMA/InfVideoLLaMA/models/long_term_attention.py",
    "infty-Video-LLaMA/InfVideoLLaMA/models/basis_functions.py",
    "infty-Video-LLaMA/InfVideoLLaMA/models/Qformer.py",
    "infty-VideoChat2/models/blip2/long_term_attention_gibbs.py",
    "infty-VideoChat2/models/blip2/basis_functions.py",
    "infty-VideoChat2/models/blip2/Qformer.py",
]


def install(verbose=False):
    """Returns the number of files present under baseline/_ref afterwards (0 when /root/reference is absent and
    nothing was installed earlier)."""
    have = 0
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if os.path.isfile(s):
            if not (os.path.isfile(d) and filecmp.cmp(s, d, shallow=False)):
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                if verbose:
                    print("installed", rel)
        have += os.path.isfile(d)
    return have


if __name__ == "__main__":
    n = install(verbose=True)
    print(f"{n}/{len(FILES)} reference files under {DST}")
    sys.exit(0 if n == len(FILES) else 1)
