"""CRISPRessoCOREResources-compatible surface (reference: CRISPResso2/CRISPRessoCOREResources.pyx).

`ResultsSlotsDict` mirrors the reference container (:18-65).  Payloads are materialised from what the CUDA
kernel emitted -- scalar counts in c2b_aln_rec plus the edit list (c2b_edit) -- by pure expansion
(`range(start, end)`, pairing flanks); no classification decision is taken on the host.
"""
import numpy as np

from . import _lib

_SLOTS = (
    'all_insertion_positions', 'all_insertion_left_positions', 'insertion_positions', 'insertion_coordinates',
    'insertion_sizes', 'insertion_n', 'all_deletion_positions', 'all_deletion_coordinates', 'deletion_positions',
    'deletion_coordinates', 'deletion_sizes', 'deletion_n', 'all_substitution_positions', 'substitution_positions',
    'all_substitution_values', 'substitution_values', 'substitution_n', 'ref_positions', 'ref_name', 'aln_scores',
    'classification', 'aln_seq', 'aln_ref', 'aln_strand', 'irregular_ends', 'insertions_outside_window',
    'deletions_outside_window', 'substitutions_outside_window', 'total_mods', 'mods_in_window', 'mods_outside_window',
)


class ResultsSlotsDict:
    """Slots object with dict-style access (same field names and behaviour as COREResources.pyx:18-65)."""
    __slots__ = _SLOTS

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    @property
    def __dict__(self):
        return {k: getattr(self, k) for k in self.__slots__ if hasattr(self, k)}

    def __reduce__(self):                      # picklable (ranks exchange payloads in core.process_fastq_sharded)
        return (_rebuild_slots, (self.__dict__,))


def _rebuild_slots(state):
    return ResultsSlotsDict(**state)


def ref_positions_from(aln_ref):
    """ref_positions list (COREResources.pyx:109-133) from the aligned reference string."""
    g = np.frombuffer(aln_ref.encode(), dtype=np.uint8) == ord('-')
    idx = np.cumsum(~g) - (~g)          # reference index of each column / bases seen so far at gap columns
    out = np.where(g, np.where(idx == 0, -1, -idx), idx)
    return out.tolist()


def ref_positions_fast(aln_ref):
    """Same list as ref_positions_from, built from the gap runs of the aligned reference (usually none or one): ranges for
    the stretches of reference bases, -idx (or -1 before the first base) repeated over each run (COREResources.pyx:109-133)."""
    k = aln_ref.find("-")
    if k < 0:
        return list(range(len(aln_ref)))
    out, idx, pos, n = [], 0, 0, len(aln_ref)
    while k >= 0:
        out.extend(range(idx, idx + k - pos))
        idx += k - pos
        e = k
        while e < n and aln_ref[e] == "-":
            e += 1
        out.extend([-idx if idx else -1] * (e - k))
        pos = e
        k = aln_ref.find("-", e)
    out.extend(range(idx, idx + n - pos))
    return out


def payload_from_lists(insertion_n, deletion_n, substitution_n, edits, aln_ref, legacy=False):
    """Batch fast path of payload_from_device: `edits` is a list of plain tuples (a, b, type, in_window, base, pad) --
    EDIT_DTYPE rows after one .tolist() per batch -- already cut to the alignment's n_edits.  Same lists, built with
    plain Python on the (usually 0-3) entries instead of numpy masks."""
    all_sub_pos, all_sub_val, sub_pos, sub_val = [], [], [], []
    all_ins_left, all_ins_pos, ins_coords, ins_pos, ins_sizes = [], [], [], [], []
    all_del_coords, all_del_pos, del_coords, del_pos, del_sizes = [], [], [], [], []
    for a, b, t, inw, base, _ in edits:
        if t == 1:
            all_sub_pos.append(a)
            all_sub_val.append(chr(base))
            if inw:
                sub_pos.append(a)
                sub_val.append(chr(base))
        elif t == 2:
            all_ins_left.append(a)
            all_ins_pos.append(a)
            all_ins_pos.append(a + 1)
            if inw:
                ins_coords.append((a, a + 1))
                ins_pos.append(a)
                ins_pos.append(a + 1)
                ins_sizes.append(b)
        else:
            all_del_coords.append((a, b))
            all_del_pos.extend(range(a, b))
            if inw:
                del_coords.append((a, b))
                del_pos.extend(range(a, b))
                del_sizes.append(b - a + base - 2 if legacy else b - a)     # legacy: the column count (c2b_edit.base = size - (b - a) + 2)
    if legacy:                                            # COREResources.pyx:311-312: numpy sums (0.0 for no window indel)
        insertion_n, deletion_n = np.sum(ins_sizes), np.sum(del_sizes)
    p = ResultsSlotsDict.__new__(ResultsSlotsDict)
    p.all_insertion_positions = all_ins_pos
    p.all_insertion_left_positions = all_ins_left
    p.insertion_positions = ins_pos
    p.insertion_coordinates = ins_coords
    p.insertion_sizes = ins_sizes
    p.insertion_n = insertion_n
    p.all_deletion_positions = all_del_pos
    p.all_deletion_coordinates = all_del_coords
    p.deletion_positions = del_pos
    p.deletion_coordinates = del_coords
    p.deletion_sizes = del_sizes
    p.deletion_n = deletion_n
    p.all_substitution_positions = all_sub_pos
    p.substitution_positions = sub_pos
    p.all_substitution_values = np.array(all_sub_val)
    p.substitution_values = np.array(sub_val)
    p.substitution_n = substitution_n
    p.ref_positions = ref_positions_fast(aln_ref)
    return p


def payload_from_device(aln, edits, aln_read, aln_ref, legacy=False):
    """aln: one ALN_DTYPE record; edits: EDIT_DTYPE array (at least aln['n_edits'] valid entries)."""
    n = int(aln["n_edits"])
    if aln["status"] & _lib.ST_EDIT_OVERFLOW or n > len(edits):
        raise OverflowError("edit list overflow")
    ed = edits[:n]
    sub, ins, dele = ed[ed["type"] == 1], ed[ed["type"] == 2], ed[ed["type"] == 3]
    all_sub_pos = sub["a"].astype(int).tolist()
    all_sub_val = [chr(b) for b in sub["base"]]
    sw = sub[sub["in_window"] != 0]
    sub_pos = sw["a"].astype(int).tolist()
    sub_val = [chr(b) for b in sw["base"]]
    all_ins_left = ins["a"].astype(int).tolist()
    all_ins_pos = [v for a in all_ins_left for v in (a, a + 1)]
    iw = ins[ins["in_window"] != 0]
    ins_coords = [(int(a), int(a) + 1) for a in iw["a"]]
    ins_pos = [v for c in ins_coords for v in c]
    ins_sizes = iw["b"].astype(int).tolist()
    all_del_coords = [(int(a), int(b)) for a, b in zip(dele["a"], dele["b"])]
    all_del_pos = [p for a, b in all_del_coords for p in range(a, b)]
    dw = dele[dele["in_window"] != 0]
    del_coords = [(int(a), int(b)) for a, b in zip(dw["a"], dw["b"])]
    del_pos = [p for a, b in del_coords for p in range(a, b)]
    del_sizes = [b - a for a, b in del_coords]
    ins_n, del_n = int(aln["insertion_n"]), int(aln["deletion_n"])
    if legacy:
        del_sizes = [int(b) - int(a) + int(x) - 2 for a, b, x in zip(dw["a"], dw["b"], dw["base"])]
        ins_n, del_n = np.sum(ins_sizes), np.sum(del_sizes)
    return ResultsSlotsDict(
        all_insertion_positions=all_ins_pos, all_insertion_left_positions=all_ins_left,
        insertion_positions=ins_pos, insertion_coordinates=ins_coords, insertion_sizes=ins_sizes,
        insertion_n=ins_n,
        all_deletion_positions=all_del_pos, all_deletion_coordinates=all_del_coords,
        deletion_positions=del_pos, deletion_coordinates=del_coords, deletion_sizes=del_sizes,
        deletion_n=del_n,
        all_substitution_positions=all_sub_pos, substitution_positions=sub_pos,
        all_substitution_values=np.array(all_sub_val), substitution_values=np.array(sub_val),
        substitution_n=int(aln["substitution_n"]),
        ref_positions=ref_positions_from(aln_ref),
    )


_pair_engine = None


def find_indels_substitutions(read_seq_al, ref_seq_al, _include_indx, _legacy=False):
    """Drop-in for CRISPRessoCOREResources.find_indels_substitutions (COREResources.pyx:71) for aligned
    pairs that obey the aligner's invariants (no column with two gaps, no insertion column next to a
    deletion column -- true of every global_align output, Align.pyx:394-413).  Runs the engine's
    row-classification kernel on the GPU; anything else raises."""
    from .engine import Engine, EngineError
    global _pair_engine
    if len(read_seq_al) != len(ref_seq_al) or len(ref_seq_al) == 0:
        raise ValueError("aligned strings must be non-empty and of equal length")
    if _pair_engine is None:
        _pair_engine = Engine()
    e = _pair_engine
    aln, edits = e.classify_pair(read_seq_al, ref_seq_al, [int(v) for v in _include_indx], legacy=_legacy)
    return payload_from_device(aln, edits, read_seq_al, ref_seq_al, legacy=_legacy)


def find_indels_substitutions_legacy(read_seq_al, ref_seq_al, _include_indx):
    """Drop-in for CRISPRessoCOREResources.find_indels_substitutions_legacy (COREResources.pyx:190-315,
    `--use_legacy_insertion_quantification`), same restrictions as find_indels_substitutions."""
    return find_indels_substitutions(read_seq_al, ref_seq_al, _include_indx, _legacy=True)
