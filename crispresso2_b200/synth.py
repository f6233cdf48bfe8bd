"""Deterministic synthetic amplicon-sequencing data (SURVEY.md section 8d generator spec).

Used by bench.py and the tests; not part of the alignment path.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def random_amplicon(rng, length=250):
    return _ACGT[rng.integers(0, 4, size=length)].tobytes().decode()


def amplicon_setup(seq, guide_start=110, guide_len=20, window_center=-3, window_size=1,
                   exclude_left=15, exclude_right=15, seed_len=10, seed_count=5, gap_incentive_value=1,
                   min_aln_score=60):
    """Builds the refs[...] entry the hot path reads (CRISPRessoCORE.py:3205-3268 semantics, restated):
    cut point = guide_end + window_center; gap_incentive[cut+1] = 1; include_idxs = window around the cut
    minus the excluded ends; fw/rc seeds every `seed_count` bp starting at exclude_left."""
    L = len(seq)
    cut = guide_start + guide_len + window_center      # index of the base left of the cut
    gi = np.zeros(L + 1, dtype=np.int64)
    gi[cut + 1] = gap_incentive_value
    win = set(range(cut - window_size + 1, cut + window_size + 1))
    keep = set(range(exclude_left, L - exclude_right))
    include = np.array(sorted(win & keep), dtype=np.int64)
    fw, rc = [], []
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    for s in range(exclude_left, L - exclude_right - seed_len, seed_count):
        k = seq[s:s + seed_len]
        fw.append(k)
        rc.append("".join(comp[c] for c in reversed(k)))
    return {"sequence": seq, "sequence_length": L, "gap_incentive": gi, "include_idxs": include,
            "fw_seeds": fw, "rc_seeds": rc, "min_aln_score": min_aln_score, "cut_point": cut}


def synth_reads(rng, amplicon, n_reads, read_len=250, sub_rate=0.005, del_frac=0.25, ins_frac=0.10,
                rc_frac=0.0, n_rate=0.0, cut=None):
    """-> uint8 array [n_reads, read_len] of ASCII bases.  Edits overlap the cut; reads shortened by a
    deletion are padded with random bases (reading into adapter), reads lengthened are truncated."""
    L = len(amplicon)
    if cut is None:
        cut = L // 2
    amp = np.frombuffer(amplicon.encode(), dtype=np.uint8)
    out = np.empty((n_reads, read_len), dtype=np.uint8)
    kind = rng.random(n_reads)
    del_len = rng.integers(1, 21, size=n_reads)
    del_off = rng.integers(0, 21, size=n_reads)
    ins_len = rng.integers(1, 11, size=n_reads)
    for r in range(n_reads):
        if kind[r] < del_frac:
            d = int(del_len[r])
            a = max(0, cut + 1 - int(del_off[r]) % (d + 1))
            s = np.concatenate([amp[:a], amp[a + d:]])
        elif kind[r] < del_frac + ins_frac:
            k = int(ins_len[r])
            s = np.concatenate([amp[:cut + 1], _ACGT[rng.integers(0, 4, size=k)], amp[cut + 1:]])
        else:
            s = amp
        if len(s) < read_len:
            s = np.concatenate([s, _ACGT[rng.integers(0, 4, size=read_len - len(s))]])
        out[r] = s[:read_len]
    if sub_rate > 0:
        m = rng.random(out.shape) < sub_rate
        out[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()))]
    if n_rate > 0:
        m = rng.random(out.shape) < n_rate
        out[m] = ord("N")
    if rc_frac > 0:
        flip = np.nonzero(rng.random(n_reads) < rc_frac)[0]
        out[flip] = _COMP[out[flip][:, ::-1]]
    return out


def synth_reads_fast(rng, amplicon, n_reads, read_len=250, sub_rate=0.005, del_frac=0.25, ins_frac=0.10,
                     cut=None, n_templates=4096):
    """Vectorised variant for million-read benches: draws edit templates once, then applies per-read
    substitutions.  Same edit distribution as synth_reads."""
    base = synth_reads(rng, amplicon, n_templates, read_len, 0.0, del_frac, ins_frac, 0.0, 0.0, cut)
    out = base[rng.integers(0, n_templates, size=n_reads)]
    if sub_rate > 0:
        m = rng.random(out.shape) < sub_rate
        out[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()))]
    return out


def write_fastq(path, reads):
    with open(path, "w") as fh:
        for k, r in enumerate(reads):
            s = r.tobytes().decode() if not isinstance(r, str) else r
            fh.write("@r%d\n%s\n+\n%s\n" % (k, s, "I" * len(s)))


def write_fastq_fast(path, reads):
    """FASTQ of a uint8 [n, L] read matrix in one vectorised pass (fixed-width names '@r%09d', quality 'I')."""
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    n, L = reads.shape
    rec = np.empty((n, 11 + 1 + L + 1 + 2 + L + 1), dtype=np.uint8)
    rec[:, 0], rec[:, 1] = ord("@"), ord("r")
    idx = np.arange(n, dtype=np.int64)
    for d in range(9):
        rec[:, 10 - d] = ord("0") + (idx // 10 ** d) % 10
    rec[:, 11] = 10
    rec[:, 12:12 + L] = reads
    rec[:, 12 + L] = 10
    rec[:, 13 + L], rec[:, 14 + L] = ord("+"), 10
    rec[:, 15 + L:15 + 2 * L] = ord("I")
    rec[:, 15 + 2 * L] = 10
    with open(path, "wb") as fh:
        fh.write(rec.tobytes())


# ------------------------------------------------------------------------------------------ BASELINE.json configs[2..4]
def hdr_workload(arng, rng, n_reads):
    """configs[2]: WT + HDR (WT with a 3-bp substitution + 6-bp insertion near the cut) + a third allele with
    5 SNPs; reads drawn 60/30/10 %.  -> (refs, ref_names, uint8 [n_reads, 250])"""
    amp = random_amplicon(arng, 250)
    hdr = amp[:120] + "TGA" + amp[123:127] + "ACGTAC" + amp[127:]
    snp = list(amp)
    for p in (30, 80, 140, 190, 230):
        snp[p] = "A" if snp[p] != "A" else "C"
    snp = "".join(snp)
    refs = {"WT": amplicon_setup(amp), "HDR": amplicon_setup(hdr), "SNP": amplicon_setup(snp)}
    parts = [synth_reads_fast(rng, a, int(n_reads * f), 250, cut=126) for a, f in ((amp, 0.6), (hdr, 0.3), (snp, 0.1))]
    reads = np.concatenate(parts)
    if len(reads) < n_reads:
        reads = np.concatenate([reads, synth_reads_fast(rng, amp, n_reads - len(reads), 250, cut=126)])
    return refs, ["WT", "HDR", "SNP"], reads[rng.permutation(n_reads)]


def pooled_workload(arng, rng, n_reads, n_amplicons=96):
    """configs[3]: n_amplicons independent random amplicons (length U[180,280]), reads of 250 bp carrying the index of their
    amplicon (post-demultiplex Pooled).  -> (refs, ref_names, packed uint8, int64 offsets, int32 ref_id)"""
    refs, names, parts, rid = {}, [], [], []
    per = -(-n_reads // n_amplicons)
    for k in range(n_amplicons):
        Lk = int(arng.integers(180, 281))
        a = random_amplicon(arng, Lk)
        nm = "amp%d" % k
        refs[nm] = amplicon_setup(a, guide_start=Lk // 2 - 10)
        names.append(nm)
        parts.append(synth_reads_fast(rng, a, per, 250, cut=refs[nm]["cut_point"], n_templates=1024))
        rid += [k] * per
    reads = np.concatenate(parts)
    rid = np.asarray(rid, dtype=np.int32)
    order = rng.permutation(len(reads))[:n_reads]
    reads, rid = reads[order], rid[order]
    return refs, names, reads.reshape(-1), np.arange(n_reads + 1, dtype=np.int64) * 250, np.ascontiguousarray(rid)


def mixed_length_reads(rng, amplicon, n_reads, lo=50, hi=300, cut=None):
    """configs[4]: read lengths U[lo, hi], truncated copies of edited amplicon reads.  -> (packed uint8, int64 offsets)"""
    base = synth_reads_fast(rng, amplicon, n_reads, hi, cut=cut)
    lens = rng.integers(lo, hi + 1, size=n_reads)
    off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return base[np.arange(hi)[None, :] < lens[:, None]], off


def bgzf_bytes(data, block=60000, level=1):
    """`data` as blocked gzip (BGZF, what bgzip / Illumina's converters write): members of at most `block` input bytes whose
    header carries the member's size in a 'BC' extra subfield, then the empty end-of-file member."""
    import struct
    import zlib
    out = bytearray()
    chunks = [data[k:k + block] for k in range(0, len(data), block)] + [b""]
    for ch in chunks:
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        raw = c.compress(ch) + c.flush()
        bsize = 12 + 6 + len(raw) + 8
        out += b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += raw + struct.pack("<II", zlib.crc32(ch) & 0xffffffff, len(ch))
    return bytes(out)
