#!/usr/bin/env python
"""Stand-in for the third-party `fastp` binary (absent from this image) in the paired-end CLI drop-in test: the reference runs
`fastp -i R1 -I R2 --out1 A --out2 B ... ` before process_paired_fastq even in --crispresso_merge mode (CRISPRessoCORE.py:3671-3690);
with trimming and filtering disabled (the reference's default options string) that is a pass-through, which is all this does.
TEST INFRASTRUCTURE."""
import gzip
import shutil
import sys


def main():
    a = sys.argv[1:]
    if "--version" in a:                                     # CRISPRessoCORE.check_program parses `fastp X.Y.Z`
        print("fastp 0.23.4")
        return 0
    val = lambda flag: a[a.index(flag) + 1]
    for src, dst in ((val("-i"), val("--out1")), (val("-I"), val("--out2"))):
        opener = gzip.open if src.endswith(".gz") else open
        with opener(src, "rb") as fi, (gzip.open(dst, "wb", compresslevel=1) if dst.endswith(".gz") else open(dst, "wb")) as fo:
            shutil.copyfileobj(fi, fo)
    for flag in ("--json", "--html"):
        if flag in a:
            open(val(flag), "w").close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
