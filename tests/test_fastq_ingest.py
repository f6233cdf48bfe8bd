"""Native FASTQ ingest + dedup (c2b_fastq_dedup) against the reference's own loop (CRISPRessoCORE.py:1820-1849),
restated here verbatim in Python as the checker: same unique sequences, same first-seen order, same counts,
same number of records -- on clean files, gzip, CRLF / lone-CR line ends, blank lines, truncated records,
whitespace around sequences, and large multi-threaded inputs."""
import gzip
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

from crispresso2_b200 import fastq, synth


@pytest.fixture(scope="module")
def lib():
    import build_emu
    return build_emu.build()        # the FASTQ front end is host code: identical in the CUDA and the emulator build


def reference_loop(path):
    """CRISPRessoCORE.py:1820-1849."""
    opener = (lambda x: gzip.open(x, "rt")) if str(path).endswith(".gz") else open
    cache, n = {}, 0
    with opener(path) as fh:
        fastq_id = fh.readline()
        while fastq_id:
            seq = fh.readline().strip()
            fh.readline().strip()
            fh.readline()
            if seq in cache:
                cache[seq] += 1
            else:
                cache[seq] = 1
            fastq_id = fh.readline()
            n += 1
    return cache, n


def check(path, lib, threads=0):
    want, n = reference_loop(path)
    got = fastq.dedup_file(path, threads, lib_path=lib)
    assert got.n_reads == n
    assert got.uniques == list(want.keys())
    assert got.counts.tolist() == list(want.values())
    assert int(got.counts.sum()) == n
    return got


CASES = {
    "clean": b"@a\nACGT\n+\nIIII\n@b\nACGT\n+\nIIII\n@c\nTTTT\n+\nIIII\n",
    "no_final_newline": b"@a\nACGT\n+\nIIII\n@b\nGGGG\n+\nIIII",
    "crlf": b"@a\r\nACGT\r\n+\r\nIIII\r\n@b\r\nACGT\r\n+\r\nIIII\r\n",
    "lone_cr": b"@a\rACGT\r+\rIIII\r@b\rAAAA\r+\rIIII\r",
    "mixed_ends": b"@a\nACGT\r\n+\rIIII\n@b\r\nACGT\n+\nIIII\r",
    "blank_lines": b"@a\nACGT\n+\nIIII\n\n\n\n\n@b\nACGT\n+\nIIII\n\n",
    "truncated_record": b"@a\nACGT\n+\nIIII\n@b\nTTGA\n",
    "only_id": b"@a\n",
    "spaces": b"@a\n  ACGT \t\n+\nIIII\n@b\nACGT\n+\nIIII\n@c\n\x0bAC GT\x0c\n+\nIIII\n",
    "lowercase_and_n": b"@a\nacgtN\n+\nIIIII\n@b\nacgtN\n+\nIIIII\n@c\nACGTN\n+\nIIIII\n",
    "empty": b"",
    "empty_seq": b"@a\n\n+\n\n@b\n\n+\n\n",
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_edge_cases_match_the_reference_loop(name, lib, tmp_path):
    p = tmp_path / (name + ".fastq")
    p.write_bytes(CASES[name])
    check(str(p), lib)
    g = tmp_path / (name + ".fastq.gz")
    with gzip.open(g, "wb") as fh:
        fh.write(CASES[name])
    check(str(g), lib)


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_large_file_multithreaded(lib, tmp_path, threads):
    rng = np.random.default_rng(threads)
    amp = synth.random_amplicon(rng, 250)
    reads = synth.synth_reads_fast(rng, amp, 40000, 250, sub_rate=0.002, cut=126, n_templates=512)
    p = tmp_path / "big.fastq"
    synth.write_fastq(str(p), reads)
    assert os.path.getsize(p) > (1 << 20)                  # above the single-thread cut-off
    got = check(str(p), lib, threads)
    assert 1000 < len(got.uniques) < 40000
    # packed layout is what the engine takes
    assert got.off[-1] == len(got.buf) and (np.diff(got.off) == 250).all()
    assert got.first_index[0] == 0 and (np.diff(got.first_index) > 0).all()


def test_ragged_lengths_and_crlf_multithreaded(lib, tmp_path):
    rng = np.random.default_rng(5)
    lines = []
    for k in range(60000):
        L = int(rng.integers(0, 40))
        s = "".join(rng.choice(list("ACGT"), L))
        end = ["\n", "\r\n", "\r"][k % 3]
        lines.append("@r%d%s%s%s+%s%s%s" % (k, end, s, end, end, "I" * L, end))
    p = tmp_path / "ragged.fastq"
    p.write_bytes("".join(lines).encode())
    assert os.path.getsize(p) > (1 << 20)
    check(str(p), lib, 7)


bgzf_bytes = synth.bgzf_bytes


@pytest.mark.parametrize("name", ["clean", "crlf", "blank_lines", "truncated_record", "empty"])
def test_blocked_gzip_edge_cases(name, lib, tmp_path):
    p = tmp_path / (name + ".fastq.gz")
    p.write_bytes(bgzf_bytes(CASES[name], block=7))        # members end in the middle of lines and of "\r\n" pairs
    with gzip.open(p, "rb") as fh:
        assert fh.read() == CASES[name]
    check(str(p), lib)


def test_blocked_gzip_large_parallel_inflate_and_fallbacks(lib, tmp_path):
    rng = np.random.default_rng(12)
    amp = synth.random_amplicon(rng, 250)
    reads = synth.synth_reads_fast(rng, amp, 30000, 250, sub_rate=0.002, cut=126, n_templates=512)
    plain = tmp_path / "big.fastq"
    synth.write_fastq(str(plain), reads)
    data = plain.read_bytes()
    want = fastq.dedup_file(str(plain), lib_path=lib)
    bg = tmp_path / "big.fastq.gz"
    bg.write_bytes(bgzf_bytes(data))
    assert len(bgzf_bytes(data)) > 250 * 30                # several hundred members: the threaded path
    got = check(str(bg), lib)
    assert np.array_equal(got.buf, want.buf) and np.array_equal(got.counts, want.counts)
    # a corrupted member (payload byte flipped: CRC mismatch) is not accepted by the blocked reader; the serial zlib path
    # then reports the error like the reference's gzip module would
    bad = bytearray(bgzf_bytes(data))
    bad[len(bad) // 2] ^= 0x55
    bp = tmp_path / "bad.fastq.gz"
    bp.write_bytes(bytes(bad))
    with pytest.raises(fastq.FastqError):
        fastq.dedup_file(str(bp), lib_path=lib)
    # ordinary multi-member gzip without the size subfield: serial path, same result
    mm = tmp_path / "multi.fastq.gz"
    with open(mm, "wb") as fh:
        half = data.index(b"\n@", len(data) // 2) + 1
        fh.write(gzip.compress(data[:half]) + gzip.compress(data[half:]))
    got = check(str(mm), lib)
    assert np.array_equal(got.buf, want.buf)


def test_buffer_entry(lib):
    got = fastq.dedup_bytes(CASES["clean"], lib_path=lib)
    assert got.uniques == ["ACGT", "TTTT"] and got.counts.tolist() == [2, 1] and got.n_reads == 3


def test_front_end_selection(lib, tmp_path, monkeypatch):
    """process_fastq's choice of FASTQ front end: the emulator build has no device front end (host threads); C2B_GPU_INGEST
    forces either way; unreadable paths fall to the host front end, which reports them."""
    p = tmp_path / "x.fastq"
    p.write_bytes(CASES["clean"])
    monkeypatch.delenv("C2B_GPU_INGEST", raising=False)
    assert fastq.gpu_ingest_device(str(p), 3, lib) is None
    monkeypatch.setenv("C2B_GPU_INGEST", "1")
    assert fastq.gpu_ingest_device(str(p), 3, lib) == 3
    monkeypatch.setenv("C2B_GPU_INGEST", "0")
    assert fastq.gpu_ingest_device(str(p), 3, lib) is None
    got = fastq.dedup_for_process_fastq(str(p), 0, lib)
    assert got.uniques == ["ACGT", "TTTT"]
    monkeypatch.setenv("C2B_GPU_INGEST", "1")                # forced onto a build without it: a loud failure naming the switch
    with pytest.raises(fastq.FastqError, match="C2B_GPU_INGEST=0"):
        fastq.dedup_for_process_fastq(str(p), 0, lib)
