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
