#!/usr/bin/env python
"""baseline/_ref: the UNMODIFIED reference modules of the hot path, so that they can be timed on the GPU box.

    python baseline/make_ref.py

/root/reference exists only in the build container; `gpurun` ships /root/repo (git-ignored files included), so this
script -- run by __graft_entry__.build() whenever /root/reference is present -- copies, byte for byte,

    scripts/distance_based_kmer_recruitment.py  scripts/read_kmer_cloud.py  scripts/ncrf_parser.py  scripts/read.py
    scripts/utils/*.py

into the git-ignored baseline/_ref/ (never into the history: .gitignore lists it) together with oracle/bio_shim's
20-line stand-in for Biopython's SeqIO (the only missing import of that path).  baseline/t0.py imports
them from there.  The reference is a directory of scripts without setup.py / pyproject.toml, so there is nothing for
`pip install` to build; this copy is the install.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
FILES = ["scripts/distance_based_kmer_recruitment.py", "scripts/read_kmer_cloud.py", "scripts/ncrf_parser.py",
         "scripts/read.py"]


def main():
    if not os.path.isdir(REF):
        print("baseline/make_ref.py: /root/reference is not here; keeping whatever baseline/_ref holds")
        return 0 if os.path.isdir(OUT) else 1
    os.makedirs(os.path.join(OUT, "scripts", "utils"), exist_ok=True)
    pairs = [(os.path.join(REF, f), os.path.join(OUT, f)) for f in FILES]
    udir = os.path.join(REF, "scripts", "utils")
    pairs += [(os.path.join(udir, f), os.path.join(OUT, "scripts", "utils", f)) for f in sorted(os.listdir(udir))
              if f.endswith(".py")]
    for src, dst in pairs:
        shutil.copyfile(src, dst)
        assert filecmp.cmp(src, dst, shallow=False)
    shim = os.path.join(OUT, "bio_shim")
    if os.path.isdir(shim):
        shutil.rmtree(shim)
    shutil.copytree(os.path.join(ROOT, "oracle", "bio_shim"), shim)
    print(f"baseline/_ref: {len(pairs)} reference files copied unmodified")
    return 0


if __name__ == "__main__":
    sys.exit(main())
