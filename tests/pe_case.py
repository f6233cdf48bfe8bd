"""Prime-editing test input shared by tests/test_cli_dropin.py and tests/golden/gen_golden.py: FANC reads + prime-edited reads +
reads carrying 1..30 bases of the pegRNA scaffold right after the extension (both strands, with sequencing errors, one with a
deletion next to it, some duplicated).  Needs /root/reference/tests/FANC.Cas9.fastq (generation time only)."""
import random

SPACER = "GGAATCCCTTCTGCAGCACC"
SCAFFOLD = "GTTTTAGAGCTAGAAATAGCAAGTTAAAATAAGGCTAGTCCGTTATCAACTTGAAAAAGTGGCACCGAGTCGGTGC"
_COMP = str.maketrans("ACGT", "TGCA")


def rc(s):
    return s.translate(_COMP)[::-1]


def extension(amp):
    return rc(amp[79:92] + "ACCTCGATCGCTTTT")                  # PBS + amp[92:107] with the PAM's G -> C


def write_fastq(path, amp):
    """-> (extension RNA, scaffold RNA)"""
    new15 = "ACCTCGATCGCTTTT"
    sd = rc(SCAFFOLD)
    pe = amp[:92] + new15 + amp[107:]
    rnd = random.Random(3)
    mut = lambda s, p: "".join(rnd.choice("ACGT") if rnd.random() < p else c for c in s)
    with open("/root/reference/tests/FANC.Cas9.fastq") as fh:
        lines = fh.read().split("\n")
    reads = [lines[k + 1] for k in range(0, len(lines) - 3, 4)][:120]
    reads += [mut(pe, 0.004) for _ in range(40)]
    for m in (1, 2, 3, 5, 8, 12, 20, 30):
        reads += [mut(pe[:107] + sd[:m] + pe[107:], 0.003)[:223 + m] for _ in range(3)]
    reads += [rc(pe[:107] + sd[:m] + pe[107:]) for m in (4, 9)]
    reads.append(pe[:107] + sd[:6] + pe[115:])
    reads += reads[130:150]                                     # duplicates: counts > 1 on re-labelled reads too
    with open(path, "w") as fh:
        for k, s in enumerate(reads):
            fh.write("@r%d\n%s\n+\n%s\n" % (k, s, "I" * len(s)))
    return extension(amp), SCAFFOLD


def write_pairs(path1, path2, amp, n=260, seed=5):
    """Paired-end reads over `amp` for --crispresso_merge: 150 bp mates overlapping in the middle, edits at the cut, sequencing
    errors inside the overlap with unequal qualities (the consensus then picks by quality and the pair may not be cached),
    pairs sequenced from the other strand, duplicates."""
    rnd = random.Random(seed)
    L = 150
    pairs = []
    for k in range(n):
        t = amp
        u = rnd.random()
        if u < 0.25:
            d = rnd.randint(1, 12)
            t = t[:92 - d // 2] + t[92 - d // 2 + d:]
        elif u < 0.35:
            t = t[:92] + "".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 5))) + t[92:]
        elif u < 0.45:
            t = t[:90] + rnd.choice("ACGT") + t[91:]
        if rnd.random() < 0.15:
            t = rc(t)
        m1, m2 = t[:L], rc(t[-L:])
        q1 = "".join(rnd.choice("II5#") for _ in m1)
        q2 = "".join(rnd.choice("II5#") for _ in m2)
        if rnd.random() < 0.3:                                # an error in one mate, inside the overlap
            p = rnd.randint(len(t) - L + 2, L - 3) if len(t) - L + 2 < L - 3 else L // 2
            m1 = m1[:p] + rnd.choice("ACGT") + m1[p + 1:]
        pairs.append((m1, q1, m2, q2))
    pairs += pairs[10:60]
    for path, a, b in ((path1, 0, 1), (path2, 2, 3)):
        with open(path, "w") as fh:
            for k, pr in enumerate(pairs):
                fh.write("@p%d\n%s\n+\n%s\n" % (k, pr[a], pr[b]))
    return len(pairs)
