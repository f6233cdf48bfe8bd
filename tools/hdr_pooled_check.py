# BASELINE configs[2] (3 amplicons, HDR mode) and configs[3] (Pooled: 96 amplicons, per-read amplicon id) on one GPU,
# device-resident, with a per-read parity sample against the oracle.  usage: python tools/hdr_pooled_check.py
import sys, numpy as np, torch
sys.path.insert(0, '.')
from crispresso2_b200 import synth, _lib
from crispresso2_b200.engine import Engine
from oracle import oracle as O

dev = torch.device('cuda', 0)
m = O.make_matrix()


def run(name, refs, names, reads, ref_id, flags, sample_check):
    n = len(reads)
    eng = Engine(0)
    eng.configure(refs, names, m, -20, -2, 5, 2, flags, 'ACGTN', 8)
    L = eng.L
    nr = 1 if ref_id is not None else len(names)        # Pooled: compact outputs [read][0]
    J = reads.shape[1]
    W = eng.string_width(J)
    d_reads = torch.from_numpy(reads.reshape(-1)).to(dev)
    d_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * J
    d_rid = torch.from_numpy(ref_id).to(dev) if ref_id is not None else None
    d_ord = None
    if ref_id is not None:                               # pairing order: reads of one amplicon adjacent (c2b_align_batch does this itself)
        d_ord = torch.from_numpy(np.argsort(ref_id, kind='stable').astype(np.int32)).to(dev)
        L.c2b_set_pair_order(eng.h, d_ord.data_ptr())
    d_recs = torch.empty(n * 16, dtype=torch.uint8, device=dev)
    d_alns = torch.empty(n * nr * 32, dtype=torch.uint8, device=dev)
    d_str = torch.empty(n * nr * 2 * W, dtype=torch.uint8, device=dev)
    d_ed = torch.empty(n * nr * 8 * 8, dtype=torch.uint8, device=dev)
    ms = []
    for it in range(4):
        eng.counts_reset()
        rc = L.c2b_align_batch_device(eng.h, d_reads.data_ptr(), d_off.data_ptr(), n, J, None, None,
                                      d_rid.data_ptr() if d_rid is not None else None, d_recs.data_ptr(), d_alns.data_ptr(),
                                      d_str.data_ptr(), d_ed.data_ptr())
        assert rc == 0, L.c2b_last_error(eng.h)
        eng.sync()
        ms.append(eng.last_kernel_ms())
    k = float(np.median(ms[1:]))
    recs = np.frombuffer(d_recs.cpu().numpy().tobytes(), dtype=_lib.REC_DTYPE)
    alns = np.frombuffer(d_alns.cpu().numpy().tobytes(), dtype=_lib.ALN_DTYPE).reshape(n, nr)
    bad = sample_check(recs, alns)
    dps = n * nr
    print('%-44s %8d reads x %2d refs: kernel %.1f ms -> %.2f M reads/s (%.2f M alignments/s), paths %s ring %s, aligned %.3f, parity sample %s'
          % (name, n, len(names), k, n / k / 1e3, dps / k / 1e3, eng.path_counts(), eng.ring_counts(), (recs['best_score_milli'] > 0).mean(),
             'OK' if not bad else 'MISMATCH %s' % bad[:3]))


# ---- configs[2]: WT + HDR (3-bp substitution + 6-bp insertion near the cut) + a third allele with 5 SNPs; reads 60/30/10 %
rng = np.random.default_rng(7)
amp = synth.random_amplicon(np.random.default_rng(42), 250)
hdr = amp[:120] + 'TGA' + amp[123:127] + 'ACGTAC' + amp[127:]
snp = list(amp)
for p in (30, 80, 140, 190, 230):
    snp[p] = 'A' if snp[p] != 'A' else 'C'
snp = ''.join(snp)
refs = {'WT': synth.amplicon_setup(amp), 'HDR': synth.amplicon_setup(hdr), 'SNP': synth.amplicon_setup(snp)}
names = ['WT', 'HDR', 'SNP']
n = 1 << 20
parts = [synth.synth_reads_fast(rng, a, int(n * f), 250, cut=126) for a, f in ((amp, 0.6), (hdr, 0.3), (snp, 0.1))]
reads = np.concatenate(parts)
reads = np.concatenate([reads, synth.synth_reads_fast(rng, amp, n - len(reads), 250, cut=126)])[rng.permutation(n)]
params = O.Params(expected_hdr_amplicon_seq=hdr)


def check_hdr(recs, alns):
    bad = []
    for i in range(0, 300):
        v = O.new_variant(params, reads[i].tobytes().decode(), refs, names, m)
        want = [int(round(s * 1000)) for s in v['aln_scores']]
        if [int(x) for x in alns[i]['score_milli']] != want:
            bad.append((i, want))
        if v['best_match_score'] > 0 and int(recs[i]['best_score_milli']) != int(round(v['best_match_score'] * 1000)):
            bad.append((i, 'best'))
    return bad


run('configs[2] HDR mode, 3 amplicons', refs, names, reads, None, _lib.F_HDR_REF1, check_hdr)
if len(sys.argv) > 1 and sys.argv[1] == 'hdr':
    sys.exit(0)

# ---- configs[3]: 96 amplicons (len U[180,280]), reads carry the index of their amplicon (post-demultiplex Pooled)
rng = np.random.default_rng(11)
refs, names, parts, rid = {}, [], [], []
per = 10923                                                    # 96 x 10 923 = 1 048 608 reads
for k in range(96):
    Lk = int(rng.integers(180, 281))
    a = synth.random_amplicon(rng, Lk)
    nm = 'amp%d' % k
    refs[nm] = synth.amplicon_setup(a, guide_start=Lk // 2 - 10)
    names.append(nm)
    r = synth.synth_reads_fast(rng, a, per, 250, cut=refs[nm]['cut_point'], n_templates=1024)
    parts.append(r)
    rid += [k] * per
reads = np.concatenate(parts)
rid = np.asarray(rid, dtype=np.int32)
order = rng.permutation(len(reads))
reads, rid = reads[order], rid[order]


def chk_pooled(recs, alns):
    bad = []
    p1 = O.Params()
    for i in range(0, 300):
        nm = names[rid[i]]
        v = O.new_variant(p1, reads[i].tobytes().decode(), {nm: refs[nm]}, [nm], m)
        if int(alns[i, 0]['score_milli']) != int(round(v['aln_scores'][0] * 1000)) or int(recs[i]['best_ref']) != (rid[i] if v['best_match_score'] > 0 else -1):
            bad.append((i, nm))
    return bad


run('configs[3] Pooled, 96 amplicons, one configuration', refs, names, reads, rid, 0, chk_pooled)
