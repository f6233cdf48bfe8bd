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
