"""Allele-level consumers of the engine's unique-read table, without per-row Python.

Stands in for three pieces of the reference that walk every allele row in Python / pandas:
  * the allele rows of the quantification loop + `df_alleles` (CRISPRessoCORE.py:3909-3959, 3964-4060, 4298-4303):
        AlleleTable(variantCache).to_dataframe()
  * `Alleles_frequency_table.txt` / `.zip` (CRISPRessoCORE.py:4498-4535):
        AlleleTable.write_frequency_table(path) / .write_frequency_zip(zip_path)
  * `CRISPRessoShared.get_dataframe_around_cut_asymmetrical` (CRISPRessoShared.py:1513-1531; called per reference and guide by
    plots/data_prep.py:1537):   get_dataframe_around_cut_asymmetrical(df_alleles, cut_point, plot_left, plot_right)
The rows come from the compact results of the process_fastq call that filled `variantCache` (op streams, records, merged
weights); strings are spelled, sorted, grouped and formatted by host threads in csrc/c2b_alleles.cpp.  No alignment or
classification happens here -- those numbers are the kernels'.
"""
import ctypes as C
import os
import zipfile

import numpy as np

from . import _lib

try:
    from . import _c2b_pyext as _ext
except ImportError:
    _ext = None

CRISPRESSO2_COLS = ["Aligned_Sequence", "Reference_Sequence", "Reference_Name", "Read_Status", "n_deleted", "n_inserted",
                    "n_mutated", "#Reads", "%Reads"]


class RefPositions(list):
    """`ref_positions` of one allele row (CRISPRessoCOREResources.pyx:105-133: the reference index of every alignment column,
    -index for insertion columns), spelled on first use from the row's aligned reference string."""
    __slots__ = ("_t", "_r")

    def _fill(self):
        t = getattr(self, "_t", None)
        if t is not None:
            self._t = None
            list.extend(self, t._ref_positions(self._r))
        return self

    def index(self, *a):
        return list.index(self._fill(), *a)

    def __getitem__(self, k):
        return list.__getitem__(self._fill(), k)

    def __iter__(self):
        return list.__iter__(self._fill())

    def __len__(self):
        return list.__len__(self._fill())

    def __contains__(self, x):
        return list.__contains__(self._fill(), x)

    def __eq__(self, other):
        if isinstance(other, RefPositions):
            other._fill()
        return list.__eq__(self._fill(), other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return list.__repr__(self._fill())

    def __reduce__(self):
        return (list, (list(self._fill()),))

    def count(self, x):
        return list.count(self._fill(), x)

    def copy(self):
        return list(self._fill())

    def __add__(self, other):
        return list(self._fill()) + other

    def __reversed__(self):
        return list.__reversed__(self._fill())


def _cstr_array(strings):
    arr = (C.c_char_p * len(strings))()
    arr[:] = [s.encode() for s in strings]
    return arr


class AlleleTable:
    """Columnar allele rows of one run, in the order the reference's loop appends them (cache order, then winner order)."""

    def __init__(self, variantCache, lib_path=None):
        from . import core
        src = core.source_of(variantCache)
        scaffold = getattr(src, "scaffold", None)            # prime editing: reads re-labelled 'Scaffold-incorporated' (:789-796)
        self.L = _lib.load(lib_path or src.lib_path)
        self.ref_names = list(src.ref_names)
        self.ref_seqs = [src.refs[r]["sequence"] for r in self.ref_names]
        flags = src.flags
        parts = src.parts
        recs = np.concatenate([p[1].recs for p in parts]) if len(parts) > 1 else parts[0][1].recs
        alns = np.concatenate([p[1].alns for p in parts]) if len(parts) > 1 else parts[0][1].alns
        n, nr = alns.shape
        NW = parts[0][1].ops.shape[-1]
        ops = (np.concatenate([p[1].ops.reshape(-1, NW) for p in parts]) if len(parts) > 1 else parts[0][1].ops.reshape(-1, NW))
        meta = (np.concatenate([p[1].meta.reshape(-1) for p in parts]) if len(parts) > 1 else parts[0][1].meta.reshape(-1))
        if len(parts) > 1:
            bufs, offs, base = [], [np.zeros(1, dtype=np.int64)], 0
            for _, res, _fx in parts:
                bufs.append(res._buf)
                offs.append(res._off[1:] + base)
                base += int(res._off[-1])
            buf, off = np.concatenate(bufs), np.concatenate(offs)
        else:
            buf, off = parts[0][1]._buf, parts[0][1]._off
        ref_id = src.ref_id
        w = np.asarray(src.weights, dtype=np.int64)
        aligned = recs["best_score_milli"] > 0
        self.n_total = int(w[aligned].sum())                      # N_TOTAL of CRISPRessoCORE.py:3976
        sel = aligned & (w > 0)
        mask = recs["winner_mask"].astype(np.int64)
        first = np.zeros(n, dtype=np.int64)                       # first winner = aln_ref_names[0]
        if nr > 1:
            low = mask & -mask
            first = np.where(low > 0, np.log2(np.maximum(low, 1)).astype(np.int64), 0)
        amb = (recs["ambiguous"] != 0) if nr > 1 else np.zeros(n, dtype=bool)
        assign_first = bool(flags & _lib.F_ASSIGN_FIRST)
        hit, pe_idx = np.zeros(n, dtype=bool), -1
        if scaffold is not None:                              # one row per re-labelled read: its prime-edited alignment
            pe_idx = self.ref_names.index(core.PE_REF)
            hs = []
            for _lo, res, _fx in parts:
                h = getattr(res, "_scaffold_hits", None)
                hs.append(h if h is not None else core.scaffold_hits(res, pe_idx, scaffold[0], scaffold[1]))
            hit = np.concatenate(hs) if len(hs) > 1 else hs[0]
            amb = amb & ~hit
        rows_read, rows_r = [], []
        if nr == 1:
            idx = np.nonzero(sel)[0]
            rows_read.append(idx)
            rows_r.append(np.zeros(len(idx), dtype=np.int64))
        else:
            for r in range(nr):
                won = sel & (((mask >> r) & 1) != 0)
                if assign_first:
                    won &= (first == r) | hit
                won &= ~amb | (first == r)                          # AMBIGUOUS: one row, the first winner's payload (:3989-3993)
                won &= ~hit | (r == pe_idx)
                idx = np.nonzero(won)[0]
                rows_read.append(idx)
                rows_r.append(np.full(len(idx), r, dtype=np.int64))
        rr, rk = np.concatenate(rows_read), np.concatenate(rows_r)
        o = np.argsort(rr * nr + rk, kind="stable")
        self.row_read, self.row_r = np.ascontiguousarray(rr[o]), np.ascontiguousarray(rk[o])
        m = len(self.row_read)
        a = alns[self.row_read, self.row_r]
        self.n_deleted = np.ascontiguousarray(a["deletion_n"].astype(np.int32))
        self.n_inserted = np.ascontiguousarray(a["insertion_n"].astype(np.int32))
        self.n_mutated = np.ascontiguousarray(a["substitution_n"].astype(np.int32))
        self.status_id = np.ascontiguousarray((a["modified"] != 0).astype(np.int32))      # 0 UNMODIFIED, 1 MODIFIED
        self.count = np.ascontiguousarray(w[self.row_read])
        ref_of_row = (np.asarray(ref_id)[self.row_read].astype(np.int64) if ref_id is not None else self.row_r)
        # Reference_Name: the reference, or AMBIGUOUS_<first winner> (:3991), or DISCARDED_<first winner> (:3999)
        label_names = self.ref_names + ([core.SCAFFOLD_REF] if scaffold is not None else [])
        nn = len(label_names)
        hit_row = hit[self.row_read]
        kind = np.zeros(m, dtype=np.int64)
        discard = (a["deletion_n"] > 0) | (a["insertion_n"] > 0) if (flags & _lib.F_DISCARD_INDEL_READS) else np.zeros(m, dtype=bool)
        fw_row = first[self.row_read] if nr > 1 else ref_of_row
        kind[discard] = 2
        kind[amb[self.row_read]] = 1
        name_ref = np.where(hit_row, nn - 1, np.where(kind == 0, ref_of_row, fw_row))
        self.name_id = np.ascontiguousarray((kind * nn + name_ref).astype(np.int32))
        self.names = label_names + ["AMBIGUOUS_" + x for x in label_names] + ["DISCARDED_" + x for x in label_names]
        # Aligned_Reference_Names / _Scores: per unique read, spelled once per distinct value
        if ref_id is not None:
            self._names_id = ref_of_row.astype(np.int64)
            self._names_tab = list(self.ref_names)
        else:
            mk = np.where(assign_first & (mask > 0), mask & -mask, mask)[self.row_read]
            mk = np.where(hit_row, 0, mk)                           # aln_ref_names of a re-labelled read: the scaffold reference alone
            uq, inv = np.unique(mk, return_inverse=True)
            self._names_tab = ["&".join(self.ref_names[r] for r in range(nr) if (int(v) >> r) & 1) if v else core.SCAFFOLD_REF for v in uq]
            self._names_id = inv
        sc = alns["score_milli"][self.row_read]                                        # [m, nr]
        uq, inv = np.unique(sc, axis=0, return_inverse=True)
        self._scores_tab = ["&".join(str(int(x) / 1000.0) for x in row) for row in uq]
        self._scores_id = np.asarray(inv).reshape(-1)
        # native table: strings + order
        row_slot = np.ascontiguousarray(self.row_read * nr + (0 if ref_id is not None else self.row_r))
        row_ref = np.ascontiguousarray(ref_of_row.astype(np.int32))
        comp = np.arange(256, dtype=np.uint8)
        for x, y in zip("ACGTN", "TGCAN"):
            comp[ord(x)] = ord(y)
        seqs = _cstr_array(self.ref_seqs)
        lens = np.asarray([len(s) for s in self.ref_seqs], dtype=np.int32)
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        ops = np.ascontiguousarray(ops, dtype=np.uint64)
        meta = np.ascontiguousarray(meta, dtype=np.uint32)
        h = C.c_void_p()
        rc = self.L.c2b_alleles_build(buf.ctypes.data if len(buf) else None, off.ctypes.data, ops.ctypes.data, meta.ctypes.data, NW, m,
                                      self.row_read.ctypes.data, row_slot.ctypes.data, row_ref.ctypes.data, self.count.ctypes.data,
                                      len(self.ref_seqs), seqs, lens.ctypes.data, comp.ctypes.data, 0, C.byref(h))
        if rc != 0:
            raise RuntimeError("c2b_alleles_build failed (%d)" % rc)
        self.h = h
        self.m = m
        self.order = np.ctypeslib.as_array(C.cast(self.L.c2b_alleles_order(h), C.POINTER(C.c_int64)), shape=(max(m, 1),))[:m].copy()
        self._off = np.ctypeslib.as_array(C.cast(self.L.c2b_alleles_offsets(h), C.POINTER(C.c_int64)), shape=(m + 1,))
        self._len = np.ctypeslib.as_array(C.cast(self.L.c2b_alleles_lengths(h), C.POINTER(C.c_int32)), shape=(max(m, 1),))[:m]
        tot = int(self._off[m])
        self._arena = np.ctypeslib.as_array(C.cast(self.L.c2b_alleles_arena(h), C.POINTER(C.c_uint8)), shape=(max(tot, 1),))
        self.pct = self.count / self.n_total * 100 if self.n_total else np.zeros(m)

    def __deepcopy__(self, memo):                          # pandas deep-copies DataFrame.attrs on every derived frame
        return self

    def __copy__(self):
        return self

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self.L.c2b_alleles_free(h)

    # ------------------------------------------------------------------------------------------------ strings
    def _strings(self, rows, which):
        starts = np.ascontiguousarray(self._off[rows] + (self._len[rows] if which else 0), dtype=np.int64)
        lens = np.ascontiguousarray(self._len[rows], dtype=np.int32)
        if _ext is not None:
            return _ext.slices(self._arena, starts, lens)
        raw = self._arena.tobytes()
        return [raw[s:s + n].decode("latin-1") for s, n in zip(starts.tolist(), lens.tolist())]

    def _ref_positions(self, row):
        n, o = int(self._len[row]), int(self._off[row])
        ref = self._arena[o + n:o + 2 * n]
        out, idx = [], 0
        for c in ref.tolist():
            if c != 45:
                out.append(idx)
                idx += 1
            else:
                out.append(-1 if idx == 0 else -idx)
        return out

    # ------------------------------------------------------------------------------------------------ df_alleles
    def to_dataframe(self):
        """df_alleles as CRISPRessoCORE.py:4298-4303 leaves it: the allele_row columns, %Reads, integer n_* columns, sorted by
        (#Reads desc, Aligned_Sequence, Reference_Sequence); the index holds the rows' positions before the sort."""
        import pandas as pd
        o = self.order
        rp = []
        for r in o.tolist():
            x = RefPositions()
            x._t, x._r = self, r
            rp.append(x)
        status = np.asarray(["UNMODIFIED", "MODIFIED"], dtype=object)
        names = np.asarray(self.names, dtype=object)
        df = pd.DataFrame({
            "#Reads": self.count[o],
            "Aligned_Sequence": self._strings(o, 0),
            "Reference_Sequence": self._strings(o, 1),
            "n_inserted": self.n_inserted[o].astype(np.int64),
            "n_deleted": self.n_deleted[o].astype(np.int64),
            "n_mutated": self.n_mutated[o].astype(np.int64),
            "Reference_Name": names[self.name_id[o]].tolist(),
            "Read_Status": status[self.status_id[o]].tolist(),
            "Aligned_Reference_Names": np.asarray(self._names_tab, dtype=object)[self._names_id[o]].tolist(),
            "Aligned_Reference_Scores": np.asarray(self._scores_tab, dtype=object)[self._scores_id[o]].tolist(),
            "ref_positions": pd.Series(rp, dtype=object),
        })
        df["%Reads"] = self.pct[o]
        df.index = pd.Index(o)
        df.attrs["c2b_allele_table"] = self
        return df

    # ------------------------------------------------------------------------------------------------ frequency table
    def write_frequency_table(self, path):
        """Text of Alleles_frequency_table.txt: df_alleles.loc[:, crispresso2Cols].to_csv(sep='\\t', header=True, index=None)"""
        uq, inv = np.unique(self.count, return_inverse=True)
        pcts = [repr(float(v)) for v in (uq / self.n_total * 100 if self.n_total else np.zeros(len(uq))).tolist()]
        pct_id = np.ascontiguousarray(inv.astype(np.int32))
        names, statuses, pstr = _cstr_array(self.names), _cstr_array(["UNMODIFIED", "MODIFIED"]), _cstr_array(pcts)
        rc = self.L.c2b_alleles_write_tsv(self.h, os.fsencode(path), self.m, self.order.ctypes.data, self.name_id.ctypes.data, names,
                                          self.status_id.ctypes.data, statuses, self.n_deleted.ctypes.data, self.n_inserted.ctypes.data,
                                          self.n_mutated.ctypes.data, pct_id.ctypes.data, pstr, 0)
        if rc != 0:
            raise OSError("c2b_alleles_write_tsv failed (%d) for %s" % (rc, path))

    def write_frequency_zip(self, zip_path, member="Alleles_frequency_table.txt"):
        """Alleles_frequency_table.zip as CRISPRessoCORE.py:4529-4531 writes it (text file zipped, then removed)."""
        txt = os.path.join(os.path.dirname(os.path.abspath(zip_path)), member)
        self.write_frequency_table(txt)
        with zipfile.ZipFile(zip_path, "w", zipfile.ZIP_DEFLATED, allowZip64=True) as z:
            z.write(txt, member)
        os.remove(txt)

    # ------------------------------------------------------------------------------------------------ around the cut
    def around_cut(self, rows, cut_point, plot_left, plot_right):
        """get_dataframe_around_cut_asymmetrical over `rows` (row ids in DataFrame order)."""
        import pandas as pd
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        unedited = np.ascontiguousarray((self.status_id == 0).astype(np.uint8))
        pct = np.ascontiguousarray(self.pct, dtype=np.float64)
        g = self.L.c2b_alleles_around_cut(self.h, len(rows), rows.ctypes.data, int(cut_point), int(plot_left), int(plot_right),
                                          unedited.ctypes.data, self.n_deleted.ctypes.data, self.n_inserted.ctypes.data,
                                          self.n_mutated.ctypes.data, pct.ctypes.data)
        if g == -3:
            raise ValueError("%d is not in list" % cut_point)          # row['ref_positions'].index(cut_point)
        if g < 0:
            raise RuntimeError("c2b_alleles_around_cut failed (%d)" % g)
        Wd = max(1, self.L.c2b_alleles_cut_width(self.h))
        seq, ref = np.zeros((g, Wd), dtype=np.uint8), np.zeros((g, Wd), dtype=np.uint8)
        wl, un = np.zeros(g, dtype=np.int32), np.zeros(g, dtype=np.uint8)
        nd, ni, nm = np.zeros(g, dtype=np.int32), np.zeros(g, dtype=np.int32), np.zeros(g, dtype=np.int32)
        reads, pc = np.zeros(g, dtype=np.int64), np.zeros(g, dtype=np.float64)
        rc = self.L.c2b_alleles_cut_fetch(self.h, seq.ctypes.data, ref.ctypes.data, wl.ctypes.data, un.ctypes.data, nd.ctypes.data,
                                          ni.ctypes.data, nm.ctypes.data, reads.ctypes.data, pc.ctypes.data)
        if rc != 0:
            raise RuntimeError("c2b_alleles_cut_fetch failed (%d)" % rc)
        starts = np.arange(g, dtype=np.int64) * Wd
        if _ext is not None:
            s1, s2 = _ext.slices(seq, starts, wl), _ext.slices(ref, starts, wl)
        else:
            s1 = [seq[k, :wl[k]].tobytes().decode("latin-1") for k in range(g)]
            s2 = [ref[k, :wl[k]].tobytes().decode("latin-1") for k in range(g)]
        df = pd.DataFrame({"Aligned_Sequence": s1, "Reference_Sequence": s2, "Unedited": un.astype(bool),
                           "n_deleted": nd.astype(np.int64), "n_inserted": ni.astype(np.int64), "n_mutated": nm.astype(np.int64),
                           "#Reads": reads, "%Reads": pc})
        return df.set_index("Aligned_Sequence")


def allele_table(variantCache):
    return AlleleTable(variantCache)


def get_dataframe_around_cut_asymmetrical(df_alleles, cut_point, plot_left, plot_right, collapse_by_sequence=True):
    """Drop-in for CRISPRessoShared.get_dataframe_around_cut_asymmetrical (CRISPRessoShared.py:1518-1531) on a DataFrame made by
    AlleleTable.to_dataframe() (whole, or filtered / re-ordered by the caller, e.g. `.loc[df['Reference_Name'] == ref_name]`)."""
    if df_alleles.shape[0] == 0:
        return df_alleles
    t = df_alleles.attrs.get("c2b_allele_table")
    if t is None:
        raise TypeError("df_alleles was not built by crispresso2_b200.alleles.AlleleTable.to_dataframe()")
    return t.around_cut(df_alleles.index.values, cut_point, plot_left, plot_right)
