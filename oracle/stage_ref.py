"""TEST INFRASTRUCTURE (oracle): stage the UNMODIFIED reference where the GPU box can find it.

`/root/reference` exists only in the build container.  The reference is pure Python with nothing to compile, so "building"
it for the benchmark's CPU arm means copying the handful of modules `base.py` imports -- byte for byte, nothing edited --
into `oracle/_ref/`, which is git-ignored (the reference's sources never enter this repository's history) but travels to
the GPU box with the snapshot, like the built `.so` files.  `oracle/shim.py` falls back to that directory when
`/root/reference` is absent; `bench.py --impl reference` then times `base.RespiratoryMonitor` itself
(`cpu_baseline.kind = "reference"`) instead of the restatement in `oracle/cpu_path.py` (`"port"`).

    python -m oracle.stage_ref          (also run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil
import sys

SRC = os.environ.get("RESPMON_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# what `import base` pulls in (base.py:1-12, transforms.py:1-11): the hot-path modules and the two prototypes that
# transforms.py imports at module level
FILES = ["__init__.py", "base.py", "pyramid.py", "transforms.py", "tools.py",
         os.path.join("prototypes", "__init__.py"), os.path.join("prototypes", "parabolic.py"),
         os.path.join("prototypes", "wavelets.py")]


def stage(verbose=False):
    """Copy the reference's modules to oracle/_ref/.  Returns the directory, or None when there is no reference tree."""
    if not os.path.isfile(os.path.join(SRC, "base.py")):
        return DST if os.path.isfile(os.path.join(DST, "base.py")) else None
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            print("staged", rel)
    return DST


if __name__ == "__main__":
    out = stage(verbose=True)
    print(out or "no reference tree at %s" % SRC)
    sys.exit(0 if out else 1)
