"""View of the engine's int64 count block as the named per-position vectors / counters that the reference's
quantification loop builds (CRISPRessoCORE.py:3841-3907, :3964-4115).  Pure re-labelling plus the one
closed-form step the device leaves to the host: all_base_count = deviation + (ref base ? counts_total : 0).
"""
import numpy as np

from . import _lib


class CountBlock:
    def __init__(self, raw, ref_names, ref_seqs, alphabet, n_vec, stride, n_scal, n_hist=0, hstride=0, hist_zero=0,
                 flags=0):
        self.raw = raw
        self.ref_names, self.ref_seqs, self.alphabet = list(ref_names), list(ref_seqs), alphabet
        self.n_vec, self.stride, self.n_scal = n_vec, stride, n_scal
        self.n_hist, self.hstride, self.hist_zero, self.flags = n_hist, hstride, hist_zero, flags
        self.class_extra = {}         # joined class labels of --expand_ambiguous_alignments reads (core.process_fastq)
        self.class_whole = set()      # references whose reads carry the bare reference name as class ('Scaffold-incorporated')
        nv, nh = n_vec * stride, n_hist * hstride
        per = nv + nh + n_scal
        self._vec, self._scal, self._hist = {}, {}, {}
        for k, name in enumerate(ref_names):
            blk = raw[k * per:(k + 1) * per]
            self._vec[name] = blk[:nv].reshape(n_vec, stride)
            self._hist[name] = blk[nv:nv + nh].reshape(n_hist, hstride)
            self._scal[name] = blk[nv + nh:]

    def add_scaffold_reference(self, name, seq, pe_idx, raw_a, raw_b, weight):
        """Prime editing with a scaffold sequence: appends the reference the scaffold step of get_new_variant_object assigns
        reads to (CRISPRessoCORE.py:789-796, :3759-3764).  raw_a: block of the re-labelled reads bound to the prime-edited
        amplicon (segment pe_idx becomes the new segment); raw_b: block of the same reads bound to reference 0 (its all_* rows
        are the new segment's re-projection onto reference 0, :4226-4272); None, None = no such read."""
        nv, nh = self.n_vec * self.stride, self.n_hist * self.hstride
        per = nv + nh + self.n_scal
        seg = np.zeros(per, dtype=np.int64)
        if raw_a is not None:
            seg[:] = np.asarray(raw_a, dtype=np.int64)[pe_idx * per:(pe_idx + 1) * per]
            b = np.asarray(raw_b, dtype=np.int64)[:per]
            va, vb = seg[:nv].reshape(self.n_vec, self.stride), b[:nv].reshape(self.n_vec, self.stride)
            for dst, src in ((_lib.V_R1_ALL_INS, _lib.V_ALL_INS), (_lib.V_R1_ALL_INS_LEFT, _lib.V_ALL_INS_LEFT),
                             (_lib.V_R1_ALL_DEL, _lib.V_ALL_DEL), (_lib.V_R1_ALL_SUB, _lib.V_ALL_SUB)):
                va[dst] = vb[src]
            for q in range(len(self.alphabet) + 1):
                va[_lib.V_R1_BASEDEV0 + q] = vb[_lib.V_BASEDEV0 + q]
            seg[nv + nh + _lib.S["REF1_W"]] = b[nv + nh + _lib.S["TOTAL"]]
        self.raw = np.concatenate([np.asarray(self.raw, dtype=np.int64), seg])
        k = len(self.ref_names)
        self.ref_names.append(name)
        self.ref_seqs.append(seq)
        blk = self.raw[k * per:(k + 1) * per]
        self._vec[name] = blk[:nv].reshape(self.n_vec, self.stride)
        self._hist[name] = blk[nv:nv + nh].reshape(self.n_hist, self.hstride)
        self._scal[name] = blk[nv + nh:]
        self.class_whole.add(name)
        if weight:
            self.class_extra[name] = self.class_extra.get(name, 0) + int(weight)

    def scalar(self, ref, name):
        return int(self._scal[ref][_lib.S[name]])

    def scalars(self, ref):
        """counters of CRISPRessoCORE.py:3844-3863 under the reference's names"""
        g = lambda n: self.scalar(ref, n)
        return {"counts_total": g("TOTAL"), "counts_modified": g("MODIFIED"), "counts_unmodified": g("UNMODIFIED"),
                "counts_discarded": g("DISCARDED"), "counts_insertion": g("INS"), "counts_deletion": g("DEL"),
                "counts_substitution": g("SUB"), "counts_only_insertion": g("ONLY_INS"),
                "counts_only_deletion": g("ONLY_DEL"), "counts_only_substitution": g("ONLY_SUB"),
                "counts_insertion_and_deletion": g("INS_DEL"), "counts_insertion_and_substitution": g("INS_SUB"),
                "counts_deletion_and_substitution": g("DEL_SUB"),
                "counts_insertion_and_deletion_and_substitution": g("INS_DEL_SUB"),
                "counts_modified_frameshift": g("MOD_FRAMESHIFT"), "counts_modified_non_frameshift": g("MOD_NON_FRAMESHIFT"),
                "counts_non_modified_non_frameshift": g("NON_MOD_NON_FRAMESHIFT"),
                "counts_splicing_sites_modified": g("SPLICING_MODIFIED")}

    def class_counts(self):
        """class_counts of CRISPRessoCORE.py:3984-3986 (classes with a non-zero weight, like the reference's dict)."""
        out = {}
        amb = 0
        for r in self.ref_names:
            if r in self.class_whole:
                continue
            for lab, key, base in (("_MODIFIED", "CLASS_MODIFIED", "MODIFIED"), ("_UNMODIFIED", "CLASS_UNMODIFIED", "UNMODIFIED")):
                v = self.scalar(r, base) + self.scalar(r, key)          # counts_* plus the signed deviation (c2b200.h)
                if v:
                    out[r + lab] = v
            amb += self.scalar(r, "AMBIGUOUS_W")
        if amb:
            out["AMBIGUOUS"] = amb
        for k, v in self.class_extra.items():
            out[k] = out.get(k, 0) + v
        return out

    def size_histograms(self, ref):
        """inserted_n_dicts / deleted_n_dicts / substituted_n_dicts / effective_len_dicts of :4020-4043 as Counters with
        the reference's key sets.  The device does not store the commonest bucket (key 0, or len(ref) for the effective
        length): it is counts_total minus the stored ones."""
        from collections import Counter
        H = self._hist[ref]
        total = self.scalar(ref, "TOTAL")
        L = len(self.ref_seqs[self.ref_names.index(ref)])
        out = {}
        for name, row, ignored, hub in (("inserted_n", _lib.H_INS_N, self.flags & _lib.F_IGNORE_INSERTIONS, 0),
                                        ("deleted_n", _lib.H_DEL_N, self.flags & _lib.F_IGNORE_DELETIONS, 0),
                                        ("substituted_n", _lib.H_SUB_N, self.flags & _lib.F_IGNORE_SUBSTITUTIONS, 0),
                                        ("effective_len", _lib.H_EFF_LEN, 0, L)):
            c = Counter()
            if not ignored:
                nz = np.nonzero(H[row])[0]
                for k in nz:
                    c[int(k)] = int(H[row, k])
                rest = total - int(H[row].sum())
                if rest:
                    c[hub] = rest
            out[name] = c
        return out

    def frame_histograms(self, ref):
        """hists_inframe[ref], hists_frameshift[ref] (:3903-3906, :4134-4177): Counters, key 0 always present."""
        from collections import Counter
        H = self._hist[ref]
        out = []
        for row in (_lib.H_INFRAME, _lib.H_FRAMESHIFT):
            c = Counter()
            c[0] = 0
            for k in np.nonzero(H[row])[0]:
                c[int(k) - self.hist_zero] = int(H[row, k])
            out.append(c)
        return out[0], out[1]

    def vectors(self, ref):
        """float64 vectors under the names of oracle.VECTOR_NAMES (the reference keeps float64 too, :3865)"""
        k = self.ref_names.index(ref)
        L = len(self.ref_seqs[k])
        V = self._vec[ref]
        f = lambda row: V[row, :L].astype(np.float64)
        out = {"all_insertion_count": f(_lib.V_ALL_INS), "all_insertion_left_count": f(_lib.V_ALL_INS_LEFT),
               "all_deletion_count": f(_lib.V_ALL_DEL), "all_substitution_count": f(_lib.V_ALL_SUB),
               "insertion_count": f(_lib.V_INS), "deletion_count": f(_lib.V_DEL), "substitution_count": f(_lib.V_SUB),
               "insertion_length": f(_lib.V_INS_LEN), "deletion_length": f(_lib.V_DEL_LEN),
               "insertion_count_noncoding": f(_lib.V_INS_NONCODING), "deletion_count_noncoding": f(_lib.V_DEL_NONCODING),
               "substitution_count_noncoding": f(_lib.V_SUB_NONCODING)}
        total = self.scalar(ref, "TOTAL")
        seq = np.frombuffer(self.ref_seqs[k].encode(), dtype=np.uint8)
        for q, ch in enumerate(self.alphabet):
            out["all_substitution_base_" + ch] = f(_lib.V_SUBBASE0 + q)
            out["all_base_count_" + ch] = f(_lib.V_BASEDEV0 + q) + np.where(seq == ord(ch), float(total), 0.0)
        out["all_base_count_-"] = f(_lib.V_BASEDEV0 + len(self.alphabet))
        return out

    def aln_stats_partial(self):
        """The aln_stats sums the device accumulates (CRISPRessoCORE.py:1974-1979), summed over references."""
        keys = ["N_GLOBAL_SUBS", "N_SUBS_OUTSIDE_WINDOW", "N_MODS_IN_WINDOW", "N_MODS_OUTSIDE_WINDOW",
                "N_READS_IRREGULAR_ENDS"]
        return {k: sum(self.scalar(r, k) for r in self.ref_names) for k in keys}

    def vectors_ref1(self, ref):
        """HDR / prime-editing mode: ref1_all_*_count_vectors[ref] and ref1_all_base_count_vectors[ref + "_" + nuc]
        (CRISPRessoCORE.py:4195-4272).  For reference 0 they are copies of its own all_* vectors (:4217-4224)."""
        ref0 = self.ref_names[0]
        if ref == ref0:
            V = self.vectors(ref0)
            out = {"ref1_all_insertion_count": V["all_insertion_count"], "ref1_all_insertion_left_count": V["all_insertion_left_count"],
                   "ref1_all_deletion_count": V["all_deletion_count"], "ref1_all_substitution_count": V["all_substitution_count"]}
            for ch in self.alphabet + "-":
                out["ref1_all_base_count_" + ch] = V["all_base_count_" + ch]
        else:
            L = len(self.ref_seqs[0])
            V = self._vec[ref]
            f = lambda row: V[row, :L].astype(np.float64)
            out = {"ref1_all_insertion_count": f(_lib.V_R1_ALL_INS), "ref1_all_insertion_left_count": f(_lib.V_R1_ALL_INS_LEFT),
                   "ref1_all_deletion_count": f(_lib.V_R1_ALL_DEL), "ref1_all_substitution_count": f(_lib.V_R1_ALL_SUB)}
            w = self.scalar(ref, "REF1_W")
            seq0 = np.frombuffer(self.ref_seqs[0].encode(), dtype=np.uint8)
            for q, ch in enumerate(self.alphabet):
                out["ref1_all_base_count_" + ch] = f(_lib.V_R1_BASEDEV0 + q) + np.where(seq0 == ord(ch), float(w), 0.0)
            out["ref1_all_base_count_-"] = f(_lib.V_R1_BASEDEV0 + len(self.alphabet))
        out["ref1_all_indelsub_count"] = (out["ref1_all_insertion_count"] + out["ref1_all_deletion_count"]
                                          + out["ref1_all_substitution_count"])
        return out
