"""Native FASTQ ingest + de-duplication (c2b_fastq_dedup of include/c2b200.h) behind the variantCache semantics of
the reference's process_fastq loop (CRISPRessoCORE.py:1820-1849): unique sequences in first-seen order with their
multiplicities.  The packed arrays go straight to Engine.align_packed; Python strings are only made for the keys of
variantCache."""
import ctypes as C

import numpy as np

from . import _lib


class FastqError(RuntimeError):
    pass


class Dedup:
    """uniques: list[str] (first-seen order) | counts: int32[n_unique] | buf/off: packed bytes + int64 offsets |
    n_reads: records in the file | first_index: record index of each unique's first occurrence"""

    def __init__(self, buf, off, counts, first_index, n_reads):
        self.buf, self.off, self.counts, self.first_index, self.n_reads = buf, off, counts, first_index, n_reads
        self._uniques = None

    @property
    def uniques(self):
        if self._uniques is None:
            raw = self.buf.tobytes()
            o = self.off
            try:
                text = raw.decode("ascii")
                self._uniques = [text[o[k]:o[k + 1]] for k in range(len(o) - 1)]
            except UnicodeDecodeError:              # text-mode reading decodes UTF-8; keep the reference's key strings
                self._uniques = [raw[o[k]:o[k + 1]].decode("utf-8", errors="surrogateescape") for k in range(len(o) - 1)]
        return self._uniques


class _Handle:
    """Owns a c2b_fastq result: the arrays handed out are views of its memory (no copy of the packed reads) and keep it alive."""

    def __init__(self, L, h):
        self.L, self.h = L, h

    def __del__(self):
        h, self.h = self.h, None
        if h:
            self.L.c2b_fastq_free(h)

    def view(self, addr, nbytes, dtype):
        if not nbytes:
            return np.zeros(0, dtype=dtype)
        raw = (C.c_uint8 * nbytes).from_address(addr)
        raw._owner = self                                  # ndarray -> ctypes array -> this handle
        return np.frombuffer(raw, dtype=dtype)


def _collect(L, h):
    own = _Handle(L, h)
    nu, nr = int(L.c2b_fastq_n_unique(h)), int(L.c2b_fastq_n_reads(h))
    off = own.view(L.c2b_fastq_offsets(h), (nu + 1) * 8, np.int64)
    if not len(off):
        off = np.zeros(1, dtype=np.int64)
    tot = int(off[-1])
    buf = own.view(L.c2b_fastq_seqs(h), tot, np.uint8)
    counts = own.view(L.c2b_fastq_counts(h), nu * 4, np.int32)
    first = own.view(L.c2b_fastq_first_index(h), nu * 8, np.int64)
    return Dedup(buf, off, counts, first, nr)


def dedup_file(path, n_threads=0, lib_path=None):
    L = _lib.load(lib_path)
    h = C.c_void_p()
    rc = L.c2b_fastq_dedup(str(path).encode(), int(n_threads), C.byref(h))
    if rc != 0:
        raise FastqError("c2b_fastq_dedup failed (%d): %s" % (rc, L.c2b_fastq_last_error().decode()))
    return _collect(L, h)


def dedup_bytes(data, n_threads=0, lib_path=None):
    L = _lib.load(lib_path)
    h = C.c_void_p()
    arr = np.frombuffer(data, dtype=np.uint8)
    rc = L.c2b_fastq_dedup_buffer(arr.ctypes.data if len(arr) else None, len(arr), int(n_threads), C.byref(h))
    if rc != 0:
        raise FastqError("c2b_fastq_dedup_buffer failed (%d)" % rc)
    return _collect(L, h)
