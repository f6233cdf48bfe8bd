# BASELINE configs[4] shape (load-balance stress): mixed read lengths 50..300, one 250-bp amplicon, device-resident.
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from crispresso2_b200 import synth, _lib
from crispresso2_b200.engine import Engine
from oracle import oracle as O
rng = np.random.default_rng(5)
amp = synth.random_amplicon(np.random.default_rng(42), 250); ref = synth.amplicon_setup(amp)
n = 1 << 20
base = synth.synth_reads_fast(rng, amp, n, 300, cut=ref['cut_point'])
lens = rng.integers(50, 301, size=n)
off = np.zeros(n + 1, dtype=np.int64); np.cumsum(lens, out=off[1:])
buf = base[np.arange(300)[None, :] < lens[:, None]]
eng = Engine(0); eng.configure({'Reference': ref}, ['Reference'], O.make_matrix(), -20, -2, 5, 2, 0, 'ACGTN', 8)
W = eng.string_width(300); L = eng.L; dev = torch.device('cuda', 0)
d_reads = torch.from_numpy(buf).to(dev); d_off = torch.from_numpy(off).to(dev)
d_recs = torch.empty(n * 16, dtype=torch.uint8, device=dev); d_alns = torch.empty(n * 32, dtype=torch.uint8, device=dev)
d_str = torch.empty(n * 2 * W, dtype=torch.uint8, device=dev); d_ed = torch.empty(n * 8 * 8, dtype=torch.uint8, device=dev)
for name, order in (("input order", None), ("pair order (sorted by length)", np.argsort(lens, kind='stable').astype(np.int32))):
    d_ord = None
    if order is not None:
        d_ord = torch.from_numpy(order).to(dev); L.c2b_set_pair_order(eng.h, d_ord.data_ptr())
    else:
        L.c2b_set_pair_order(eng.h, None)
    for it in range(3):
        eng.counts_reset()
        rc = L.c2b_align_batch_device(eng.h, d_reads.data_ptr(), d_off.data_ptr(), n, 300, None, None, None, d_recs.data_ptr(), d_alns.data_ptr(), d_str.data_ptr(), d_ed.data_ptr())
        assert rc == 0
        eng.sync()
    ms = eng.last_kernel_ms()
    cells = float((lens * 250).sum())
    print('%-32s kernel %.1f ms -> %.2f M reads/s, %.1f G cells/s, paths %s' % (name, ms, n / ms / 1e3, cells / ms / 1e6, eng.path_counts()))
recs = np.frombuffer(d_recs.cpu().numpy().tobytes(), dtype=_lib.REC_DTYPE)
print('aligned fraction %.3f (reads shorter than ~0.6 x amplicon score below the default 60)' % (recs['best_score_milli'] > 0).mean())
