"""filterFastqs-compatible surface (reference: CRISPResso2/filterFastqs.py) on the native reader of libc2b200.so.

  filterFastqs(fastq_r1, fastq_r2=None, fastq_r1_out=None, fastq_r2_out=None, min_bp_qual_in_read=None,
               min_av_read_qual=None, min_bp_qual_or_N=None, debug=False)                 filterFastqs.py:29-125

Single-end input (the form CRISPRessoCORE calls, :3716-3717) and paired input (filterFastqs.py:230-407, the command-line tool).
"""
import ctypes as C
import datetime
import os

from . import _lib


def filterFastqs(fastq_r1=None, fastq_r2=None, fastq_r1_out=None, fastq_r2_out=None, min_bp_qual_in_read=None,
                 min_av_read_qual=None, min_bp_qual_or_N=None, debug=False, lib_path=None, n_threads=0):
    if debug:
        print('--fastq_r1:' + str(fastq_r1))
        print('--fastq_r2:' + str(fastq_r2))
        print('--min_bp_qual_in_read:' + str(min_bp_qual_in_read))
        print('--min_av_read_qual:' + str(min_av_read_qual))
        print('--min_bp_qual_or_N:' + str(min_bp_qual_or_N))
        print('--fastq_r1_out:' + str(fastq_r1_out))
        print('--fastq_r2_out:' + str(fastq_r2_out))
    start = datetime.datetime.now()
    if not os.path.exists(fastq_r1):
        raise Exception("fastq_r1 file '" + fastq_r1 + "' does not exist.")
    if fastq_r2 is not None and not os.path.exists(fastq_r2):
        raise Exception("fastq_r2 file '" + fastq_r2 + "' does not exist.")

    def out_name(src, given):                                                               # filterFastqs.py:48-79
        if given:
            return given
        if src.endswith('.gz'):
            return src.replace('.fastq', '').replace('.gz', '') + '_filtered.fastq.gz'
        return src.replace('.fastq', '') + '_filtered.fastq'

    out = out_name(fastq_r1, fastq_r1_out)
    out2 = out_name(fastq_r2, fastq_r2_out) if fastq_r2 else None
    if not (min_bp_qual_in_read or min_av_read_qual or min_bp_qual_or_N):
        import gzip
        for o in (out, out2):
            if o:
                (gzip.open(o, 'wb') if o.endswith('.gz') else open(o, 'wb')).close()        # the reference has opened (created) the outputs by now
        exit('Finished -- No modifications requested')
    L = _lib.load(lib_path)
    n_in, n_out = C.c_int64(0), C.c_int64(0)
    if fastq_r2:
        rc = L.c2b_fastq_filter_pair(fastq_r1.encode(), fastq_r2.encode(), out.encode(), out2.encode(), int(min_bp_qual_in_read or 0),
                                     int(min_av_read_qual or 0), int(min_bp_qual_or_N or 0), int(n_threads), C.byref(n_in), C.byref(n_out))
    else:
        rc = L.c2b_fastq_filter(fastq_r1.encode(), out.encode(), int(min_bp_qual_in_read or 0), int(min_av_read_qual or 0),
                                int(min_bp_qual_or_N or 0), int(n_threads), C.byref(n_in), C.byref(n_out))
    if rc == _lib_E_LIMIT:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    if rc == _lib_E_ARG and b"lengths differ" in L.c2b_fastq_last_error():
        raise IndexError("boolean index did not match indexed array along axis 0")
    if rc != 0:
        raise RuntimeError("c2b_fastq_filter failed (%d): %s" % (rc, L.c2b_fastq_last_error().decode()))
    print("Completed in %d seconds\n" % (datetime.datetime.now() - start).total_seconds())
    return n_in.value, n_out.value


_lib_E_ARG, _lib_E_LIMIT = -2, -3
