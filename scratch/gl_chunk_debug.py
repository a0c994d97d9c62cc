import sys; sys.path.insert(0,'.')
import numpy as np, torch
from oracle import audio_lib_oracle as oracle
from speech_cloner_b200 import synth, _lib
from speech_cloner_b200.audio_lib import DspPlan, _GlLayout, griffin_lim_device
lib=_lib.load(); plan=DspPlan.get(n_fft=400, win_length=400, hop_length=80)
T=301
P=oracle.calc_MFCC_input(synth.utterance(3500,2.0), **synth.HP_ENC)[2][:T]
amp=torch.from_numpy(np.ascontiguousarray(np.sqrt(np.power(np.float32(10.0), np.float32(0.1)*(P/np.float32(0.01)-np.float32(80.0)))))).cuda()
np.random.seed(9); ph=torch.from_numpy((np.pi*np.random.rand(T,201)).astype(np.float32)).cuda()
lay=_GlLayout([T],80); Lw=80*(T-1)
st=torch.cuda.current_stream().cuda_stream
for iters in (1,2,3):
    whole=griffin_lim_device(plan,amp,ph,lay,iters)[:Lw].clone()
    cuts=[0,80*140,Lw]
    state=torch.zeros(Lw,dtype=torch.float32,device='cuda'); nxt=torch.zeros_like(state)
    for it in range(iters):
        for r in range(2):
            lo,hi=cuts[r],cuts[r+1]
            f_lo=max(0,lo//80-4); f_hi=min(T,hi//80+6); w_lo=max(0,lo-1000); w_hi=min(Lw,hi+1000)
            rc=lib.sc_griffinlim_chunk_step(plan._h, amp[f_lo:f_hi].contiguous().data_ptr(), ph[f_lo:f_hi].contiguous().data_ptr() if it==0 else None,
                f_lo,f_hi-f_lo,T, state[w_lo:w_hi].data_ptr() if it else None, w_lo, w_hi-w_lo, nxt[lo:hi].data_ptr(), lo, hi-lo, st)
            _lib.check(rc,'x'); torch.cuda.synchronize()
        state,nxt=nxt,state
    d=(state!=whole).cpu().numpy(); idx=np.nonzero(d)[0]
    print('iters',iters,'mismatch',d.sum(), 'range', (idx.min(),idx.max()) if len(idx) else None, 'hist per 2240', np.bincount(idx//2240, minlength=11) if len(idx) else None)
# whole-run determinism
a=griffin_lim_device(plan,amp,ph,lay,3)[:Lw].clone(); b=griffin_lim_device(plan,amp,ph,lay,3)[:Lw].clone()
print('rerun equal', bool((a==b).all()))
