"""Native quality filter (c2b_fastq_filter behind crispresso2_b200.filter_fastqs.filterFastqs) against the reference's own
filterFastqs.filterFastqs (CRISPResso2/filterFastqs.py, imported from /root/reference when present -- it needs only numpy --
else against a restatement of its single-end record loop): byte-identical output text for every combination of the three
thresholds, plain and gzip, CRLF input, truncated files, qualities below '!' (uint8 wrap-around)."""
import gzip
import importlib.util
import itertools
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

from crispresso2_b200 import filter_fastqs

REF_PY = "/root/reference/CRISPResso2/filterFastqs.py"


@pytest.fixture(scope="module")
def lib():
    import build_emu
    return build_emu.build()


def reference_filter(path_in, path_out, mbp, mrq, mbpn):
    # (min_bp_qual_in_read + min_bp_qual_or_N without the mean filter is broken in the reference itself: run_mBP_mBPN masks a
    #  read-only numpy view, filterFastqs.py:191-192, and raises on the first record it keeps -- use the restatement there)
    if os.path.exists(REF_PY) and not (mbp and mbpn and not mrq):
        spec = importlib.util.spec_from_file_location("_ref_filterFastqs", REF_PY)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.filterFastqs(fastq_r1=path_in, fastq_r1_out=path_out, min_bp_qual_in_read=mbp, min_av_read_qual=mrq, min_bp_qual_or_N=mbpn)
        return
    # restatement of filterFastqs.py:128-229 (single-end record loop)
    opener = (lambda p: gzip.open(p, "rb")) if path_in.endswith(".gz") else (lambda p: open(p, "rb"))
    out = gzip.open(path_out, "wt") if path_out.endswith(".gz") else open(path_out, "w")
    with opener(path_in) as f, out:
        idl = f.readline().rstrip().decode()
        while idl:
            seq, plus, qual = f.readline().rstrip(), f.readline().rstrip(), f.readline().rstrip()
            q = np.frombuffer(qual, dtype=np.uint8) - 33
            keep = True
            if mbp and not (np.min(q) >= mbp):
                keep = False
            if keep and mrq and not (np.mean(q) >= mrq):
                keep = False
            if keep:
                s = np.frombuffer(seq, "c").copy()
                if mbpn:
                    s[q < mbpn] = b"N"
                out.write("%s\n%s\n%s\n%s\n" % (idl, s.tobytes().decode(), plus.decode(), qual.decode()))
            idl = f.readline().rstrip().decode()


def content(path):
    with (gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")) as fh:
        return fh.read()


def make_fastq(rng, n, L=60, low=0.08, crlf=False, weird=False):
    recs = []
    for k in range(n):
        seq = "".join(rng.choice(list("ACGT"), L))
        q = rng.integers(20, 41, size=L)
        q[rng.random(L) < low] = rng.integers(0, 15)
        if weird and k % 17 == 0:
            q[0] = -1                                      # ' ' (32): wraps to 255 after the uint8 subtraction
        qual = "".join(chr(33 + int(v)) for v in q)
        recs.append("@read%d extra\n%s\n+\n%s\n" % (k, seq, qual))
    text = "".join(recs)
    if crlf:
        text = text.replace("\n", "\r\n")
    return text.encode()


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("thr", [t for t in itertools.product([None, 10], [None, 30], [None, 20]) if any(t)])
def test_all_threshold_combinations(lib, tmp_path, thr, gz):
    rng = np.random.default_rng(hash(thr) % 1000)
    data = make_fastq(rng, 700, weird=True)
    src = str(tmp_path / ("in.fastq.gz" if gz else "in.fastq"))
    with (gzip.open(src, "wb") if gz else open(src, "wb")) as fh:
        fh.write(data)
    want, got = str(tmp_path / ("want.fastq.gz" if gz else "want.fastq")), str(tmp_path / ("got.fastq.gz" if gz else "got.fastq"))
    reference_filter(src, want, *thr)
    n_in, n_out = filter_fastqs.filterFastqs(fastq_r1=src, fastq_r1_out=got, min_bp_qual_in_read=thr[0], min_av_read_qual=thr[1],
                                             min_bp_qual_or_N=thr[2], lib_path=lib)
    assert content(got) == content(want)
    assert n_in == 700 and n_out == content(got).count(b"\n") // 4


def reference_filter_pair(r1, r2, o1, o2, mbp, mrq, mbpn):
    """The reference's own paired filterFastqs (filterFastqs.py:230-407), imported from /root/reference."""
    spec = importlib.util.spec_from_file_location("_ref_filterFastqs", REF_PY)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.filterFastqs(fastq_r1=r1, fastq_r2=r2, fastq_r1_out=o1, fastq_r2_out=o2, min_bp_qual_in_read=mbp, min_av_read_qual=mrq,
                     min_bp_qual_or_N=mbpn)


@pytest.mark.skipif(not os.path.exists(REF_PY), reason="needs the reference's filterFastqs.py")
@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("thr", [t for t in itertools.product([None, 12], [None, 31], [None, 20]) if any(t)])
def test_paired_all_threshold_combinations(lib, tmp_path, thr, gz):
    """Paired input, every combination of the three thresholds (the seven run_*_pair variants, including the two whose mate-2
    comparison is strict), thresholds chosen so that reads sit exactly ON them; byte-identical output for both mates."""
    rng = np.random.default_rng(abs(hash(thr)) % 1000 + 7)
    ext = ".fastq.gz" if gz else ".fastq"
    paths = {}
    for mate in (1, 2):
        data = bytearray(make_fastq(rng, 900, L=50, low=0.03))
        recs = bytes(data).split(b"\n")
        for k in range(0, 900, 3):                            # every third record: constant quality ON a threshold
            v = 12 if k % 2 else 31
            recs[4 * k + 3] = bytes([33 + v]) * 50
        paths[mate] = str(tmp_path / ("in_r%d%s" % (mate, ext)))
        with (gzip.open(paths[mate], "wb") if gz else open(paths[mate], "wb")) as fh:
            fh.write(b"\n".join(recs))
    want = [str(tmp_path / ("want%d%s" % (m, ext))) for m in (1, 2)]
    got = [str(tmp_path / ("got%d%s" % (m, ext))) for m in (1, 2)]
    reference_filter_pair(paths[1], paths[2], want[0], want[1], *thr)
    n_in, n_out = filter_fastqs.filterFastqs(fastq_r1=paths[1], fastq_r2=paths[2], fastq_r1_out=got[0], fastq_r2_out=got[1],
                                             min_bp_qual_in_read=thr[0], min_av_read_qual=thr[1], min_bp_qual_or_N=thr[2], lib_path=lib)
    assert content(got[0]) == content(want[0]) and content(got[1]) == content(want[1])
    assert n_in == 900 and n_out == content(got[0]).count(b"\n") // 4 == content(got[1]).count(b"\n") // 4
    assert 0 < n_out


@pytest.mark.skipif(not os.path.exists(REF_PY), reason="needs the reference's filterFastqs.py")
def test_paired_shorter_mate_file_and_default_names(lib, tmp_path):
    """Mate 2 runs out first: its lines read as empty -- with the mean filter the pair is dropped (mean of nothing is nan), with the
    min filter numpy raises; default output names of filterFastqs.py:48-79."""
    rng = np.random.default_rng(21)
    r1, r2 = str(tmp_path / "a_R1.fastq"), str(tmp_path / "a_R2.fastq")
    open(r1, "wb").write(make_fastq(rng, 40))
    open(r2, "wb").write(make_fastq(rng, 25))
    w1, w2 = str(tmp_path / "w1.fastq"), str(tmp_path / "w2.fastq")
    reference_filter_pair(r1, r2, w1, w2, None, 25, None)
    filter_fastqs.filterFastqs(fastq_r1=r1, fastq_r2=r2, min_av_read_qual=25, lib_path=lib)
    assert content(str(tmp_path / "a_R1_filtered.fastq")) == content(w1) and content(str(tmp_path / "a_R2_filtered.fastq")) == content(w2)
    with pytest.raises(ValueError):
        reference_filter_pair(r1, r2, w1, w2, 5, None, None)
    with pytest.raises(ValueError):
        filter_fastqs.filterFastqs(fastq_r1=r1, fastq_r2=r2, min_bp_qual_in_read=5, lib_path=lib)
    # min + mean without masking: the mean is tested first, so the short file drops pairs instead of raising (:300)
    reference_filter_pair(r1, r2, w1, w2, 5, 25, None)
    filter_fastqs.filterFastqs(fastq_r1=r1, fastq_r2=r2, fastq_r1_out=str(tmp_path / "g1.fastq"), fastq_r2_out=str(tmp_path / "g2.fastq"),
                               min_bp_qual_in_read=5, min_av_read_qual=25, lib_path=lib)
    assert content(str(tmp_path / "g1.fastq")) == content(w1) and content(str(tmp_path / "g2.fastq")) == content(w2)


def test_crlf_truncated_and_blank_id(lib, tmp_path):
    rng = np.random.default_rng(3)
    cases = {
        "crlf": make_fastq(rng, 50, crlf=True),
        "truncated": make_fastq(rng, 20)[:-35],
        "blank_id_stops": make_fastq(rng, 10) + b"\n" + make_fastq(rng, 10),
        "no_final_newline": make_fastq(rng, 5).rstrip(b"\n"),
    }
    for name, data in cases.items():
        src = str(tmp_path / (name + ".fastq"))
        open(src, "wb").write(data)
        want, got = str(tmp_path / (name + "_want.fastq")), str(tmp_path / (name + "_got.fastq"))
        for thr in ((None, 25, 20), (None, None, 25), (None, 25, None)):
            if name == "truncated" and thr[2]:              # last record: quality shorter than the sequence -> IndexError in both
                with pytest.raises(IndexError):
                    reference_filter(src, want, *thr)
                with pytest.raises(IndexError):
                    filter_fastqs.filterFastqs(fastq_r1=src, fastq_r1_out=got, min_bp_qual_in_read=thr[0], min_av_read_qual=thr[1],
                                               min_bp_qual_or_N=thr[2], lib_path=lib)
                continue
            reference_filter(src, want, *thr)
            filter_fastqs.filterFastqs(fastq_r1=src, fastq_r1_out=got, min_bp_qual_in_read=thr[0], min_av_read_qual=thr[1],
                                       min_bp_qual_or_N=thr[2], lib_path=lib)
            assert content(got) == content(want), (name, thr)


def test_large_multithreaded_and_default_output_name(lib, tmp_path):
    rng = np.random.default_rng(9)
    data = make_fastq(rng, 20000, L=100)
    src = str(tmp_path / "big.fastq.gz")
    with gzip.open(src, "wb") as fh:
        fh.write(data)
    want = str(tmp_path / "want.fastq.gz")
    reference_filter(src, want, 5, 28, 15)
    n_in, n_out = filter_fastqs.filterFastqs(fastq_r1=src, min_bp_qual_in_read=5, min_av_read_qual=28, min_bp_qual_or_N=15,
                                             lib_path=lib, n_threads=6)
    got = str(tmp_path / "big_filtered.fastq.gz")            # filterFastqs.py:50: default output name
    assert os.path.exists(got) and content(got) == content(want) and 0 < n_out < n_in == 20000


def test_error_behaviour(lib, tmp_path):
    src = str(tmp_path / "x.fastq")
    open(src, "wb").write(b"@a\nACGT\n+\nII\n")
    with pytest.raises(IndexError):
        filter_fastqs.filterFastqs(fastq_r1=src, min_bp_qual_or_N=20, lib_path=lib)
    open(src, "wb").write(b"@a\nACGT\n+\n\n")
    with pytest.raises(ValueError):
        filter_fastqs.filterFastqs(fastq_r1=src, min_bp_qual_in_read=20, lib_path=lib)
    with pytest.raises(SystemExit):
        filter_fastqs.filterFastqs(fastq_r1=src, lib_path=lib)
    with pytest.raises(Exception):
        filter_fastqs.filterFastqs(fastq_r1=str(tmp_path / "missing.fastq"), min_av_read_qual=3, lib_path=lib)
