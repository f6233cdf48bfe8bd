"""GPU FASTQ front end (c2b_fastq_dedup_gpu, csrc/c2b_fastq_gpu.cu) against the reference's own loop
(CRISPRessoCORE.py:1820-1849, restated in tests/test_fastq_ingest.py) and against the host front end: same unique
sequences in first-seen order, same counts, same first-record indices, same packed layout -- on the edge-case files
(CRLF / lone CR / blank lines / truncated records / whitespace / empty), gzip, ragged lengths, a large file with heavy
duplication and an all-unique one."""
import gzip
import os

import numpy as np
import pytest

from crispresso2_b200 import fastq, synth
from test_fastq_ingest import CASES, reference_loop

pytestmark = pytest.mark.gpu


def check_gpu(path):
    want, n = reference_loop(path)
    got = fastq.dedup_file(path, device=0)
    host = fastq.dedup_file(path)
    assert got.n_reads == n == host.n_reads
    assert got.uniques == list(want.keys())
    assert got.counts.tolist() == list(want.values())
    assert np.array_equal(got.off, host.off) and np.array_equal(got.buf, host.buf)
    assert np.array_equal(got.first_index, host.first_index)
    return got


@pytest.mark.parametrize("name", sorted(CASES))
def test_edge_cases(name, tmp_path):
    p = tmp_path / (name + ".fastq")
    p.write_bytes(CASES[name])
    check_gpu(str(p))
    g = tmp_path / (name + ".fastq.gz")
    with gzip.open(g, "wb") as fh:
        fh.write(CASES[name])
    check_gpu(str(g))


def test_buffer_entry():
    got = fastq.dedup_bytes(CASES["clean"], device=0)
    assert got.uniques == ["ACGT", "TTTT"] and got.counts.tolist() == [2, 1] and got.n_reads == 3


def test_ragged_lengths_mixed_line_ends(tmp_path):
    rng = np.random.default_rng(5)
    lines = []
    for k in range(60000):
        L = int(rng.integers(0, 40))
        s = "".join(rng.choice(list("ACGT"), L))
        end = ["\n", "\r\n", "\r"][k % 3]
        lines.append("@r%d%s%s%s+%s%s%s" % (k, end, s, end, end, "I" * L, end))
    p = tmp_path / "ragged.fastq"
    p.write_bytes("".join(lines).encode())
    check_gpu(str(p))


@pytest.mark.parametrize("templates,sub", [(512, 0.002), (4096, 0.02)])
def test_large_files(tmp_path, templates, sub):
    rng = np.random.default_rng(11)
    amp = synth.random_amplicon(rng, 250)
    reads = synth.synth_reads_fast(rng, amp, 300000, 250, sub_rate=sub, cut=126, n_templates=templates)
    p = tmp_path / "big.fastq"
    synth.write_fastq_fast(str(p), reads)
    assert os.path.getsize(p) > (128 << 20)                 # more than one upload block
    got = check_gpu(str(p))
    assert got.off[-1] == len(got.buf) and (np.diff(got.off) == 250).all()
    assert got.first_index[0] == 0 and (np.diff(got.first_index) > 0).all()


def test_process_fastq_with_gpu_ingest_equals_host_ingest(tmp_path, monkeypatch):
    """core.process_fastq with C2B_GPU_INGEST=1: identical variantCache keys, statistics and count block."""
    import golden_util as G
    import parity_util as PU
    from crispresso2_b200 import core
    from oracle import oracle as O
    rec = G.load("synth_single")
    refs = G.refs_from(rec)
    args = PU.args_from(rec["params"])
    fq = tmp_path / "g.fastq"
    with open(fq, "w") as fh:
        for k, s in enumerate(rec["reads"]):
            fh.write("@r%d\n%s\n+\n%s\n" % (k, s, "I" * len(s)))
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("C2B_GPU_INGEST", flag)
        cache = {}
        st, lost = core.process_fastq(str(fq), cache, rec["ref_names"], refs, args, [], str(tmp_path), aln_matrix=O.make_matrix())
        blk = core.quantify(cache)
        out.append((st, list(cache.keys()), sorted(lost), {r: {k: v.tolist() for k, v in blk.vectors(r).items()} for r in rec["ref_names"]},
                    blk.class_counts()))
    assert out[0] == out[1]


def test_blocked_gzip_through_the_gpu_front_end(tmp_path):
    rng = np.random.default_rng(13)
    amp = synth.random_amplicon(rng, 250)
    reads = synth.synth_reads_fast(rng, amp, 40000, 250, sub_rate=0.002, cut=126, n_templates=512)
    plain = tmp_path / "b.fastq"
    synth.write_fastq_fast(str(plain), reads)
    bg = tmp_path / "b.fastq.gz"
    bg.write_bytes(synth.bgzf_bytes(plain.read_bytes()))
    got = check_gpu(str(bg))
    want = fastq.dedup_file(str(plain))
    assert np.array_equal(got.buf, want.buf) and np.array_equal(got.counts, want.counts)
