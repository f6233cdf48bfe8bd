# config-5 style check: mixed lengths 50..300, 1 amplicon: throughput + pairing stats
import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from crispresso2_b200 import synth
from crispresso2_b200.engine import Engine
from oracle import oracle as O
rng=np.random.default_rng(5)
amp=synth.random_amplicon(np.random.default_rng(42),250); ref=synth.amplicon_setup(amp)
n=1<<20
base=synth.synth_reads_fast(rng, amp, n, 300, cut=ref['cut_point'])
lens=rng.integers(50,301,size=n)
off=np.zeros(n+1,dtype=np.int64); np.cumsum(lens,out=off[1:])
mask=np.arange(300)[None,:]<lens[:,None]
buf=base[mask]
eng=Engine(0); eng.configure({'Reference':ref},['Reference'],O.make_matrix(),-20,-2,5,2,0,'ACGTN',8)
for it in range(3):
    eng.counts_reset(); t0=time.time(); res=eng.align_packed(buf,off); dt=time.time()-t0
    print('mixed 50-300: %.1f ms -> %.2f M reads/s, paths %s, aligned %.3f'%(dt*1e3, n/dt/1e6, eng.path_counts(), (res.recs['best_score_milli']>0).mean()))
