import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import audio_lib_oracle as o
from speech_cloner_b200 import synth
hp=dict(synth.HP_ENC)
y=synth.batch(1,32,3.0,ds_norm=(0,10.))[2]
want=o.calc_MFCC_input(y,**hp)
g=np.float32(np.float64(0.003)/np.float64(np.abs(y).mean()))
for d in (-1,1):
    g2=np.nextafter(g,np.float32(np.inf*d))
    ys=(y*g2).astype(np.float32)
    kw=dict(hp); kw['mean_abs_amp_norm']=1.0
    got=o.calc_MFCC_input(ys,**kw)
    for a,b,n in zip(got,want,("MFCC","M","P")):
        err=np.abs(a.astype(np.float64)-b); tol=1e-5+1e-4*np.abs(b)
        print(d,n,"max err %.2e"%err.max(),"viol",(err>tol).sum())
# exact f64 mean vs numpy f32 pairwise mean
m32=np.abs(y).mean(); m64=np.float32(np.abs(y).astype(np.float64).mean())
print("mean32",repr(m32),"mean from f64",repr(m64), "gains", repr(np.float32(0.003/np.float64(m32))), repr(np.float32(0.003/np.float64(m64))))
