import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import audio_lib_oracle as o
from speech_cloner_b200 import synth

def radix4(a, dt):
    a0,a1,a2,a3=a
    s0=a0+a2; d0=a0-a2; s1=a1+a3; d1=a1-a3
    X0=s0+s1; X2=s0-s1
    X1=(d0.real+d1.imag)+1j*(d0.imag-d1.real)
    X3=(d0.real-d1.imag)+1j*(d0.imag+d1.real)
    return [X0.astype(dt),X1.astype(dt),X2.astype(dt),X3.astype(dt)]
def radix5(a, dt, rt):
    C1=rt(0.30901699437494745);C2=rt(-0.8090169943749473);S1=rt(0.9510565162951535);S2=rt(0.5877852522924731)
    a0,a1,a2,a3,a4=a
    t1=a1+a4;t3=a1-a4;t2=a2+a3;t4=a2-a3
    m1=a0+C1*t1+C2*t2; m2=a0+C2*t1+C1*t2
    q1=S1*t3+S2*t4; q2=S2*t3-S1*t4
    X0=a0+t1+t2
    X1=(m1.real+q1.imag)+1j*(m1.imag-q1.real); X4=(m1.real-q1.imag)+1j*(m1.imag+q1.real)
    X2=(m2.real+q2.imag)+1j*(m2.imag-q2.real); X3=(m2.real-q2.imag)+1j*(m2.imag+q2.real)
    return [x.astype(dt) for x in (X0,X1,X2,X3,X4)]
def dft20(v, dt, rt):   # v: list of 20 arrays
    t=[[None]*5 for _ in range(4)]
    for n2 in range(5):
        r=radix4([v[(5*n1+4*n2)%20] for n1 in range(4)], dt)
        for k1 in range(4): t[k1][n2]=r[k1]
    out=[None]*20
    for k1 in range(4):
        r=radix5(t[k1], dt, rt)
        for k2 in range(5): out[(5*k1+16*k2)%20]=r[k2]
    return out

def fft400(frames, cdt1, rdt1, cdt2, rdt2, in_dt):
    """frames: (F,400) float64 windowed. Full complex 400-pt via 20x20 (no real packing; precision study)."""
    F=frames.shape[0]
    x=frames.astype(in_dt).astype(cdt1)
    # step1: for each n2, DFT over n1 of x[20 n1+n2]
    v=[x[:, 20*n1:20*n1+20] for n1 in range(20)]   # each (F,20[n2])
    Y=dft20(v, cdt1, rdt1)                         # Y[k1] (F,20[n2])
    n2=np.arange(20)
    out=np.zeros((F,400),dtype=np.complex128)
    cols=[]
    for k1 in range(20):
        tw=np.exp(-2j*np.pi*n2*k1/400).astype(cdt1)
        cols.append((Y[k1]*tw).astype(cdt2))       # slot rounding to cdt2
    for k1 in range(20):
        u=[cols[k1][:, n] for n in range(20)]
        V=dft20(u, cdt2, rdt2)
        for k2 in range(20): out[:, k1+20*k2]=V[k2]
    return out

def run(y, name):
    hp=dict(synth.HP_ENC)
    want=o.calc_MFCC_input(y, **hp)[2]
    g=np.float64(0.003)/np.float64(np.abs(y).mean()); ys=y*np.float32(g)
    pe=o.calc_preemphasis(ys,0.97)
    n=len(y); T=1+n//80
    pos=(np.arange(T)[:,None]*80+np.arange(400)[None,:]-200)
    w=o.padded_window('hann',400,400)
    fr64=pe[o.reflect_index(pos,n)]*w
    fr32in=(pe.astype(np.float32)[o.reflect_index(pos,n)]*w.astype(np.float32)).astype(np.float64)  # f32 input roundings
    def pdb(X):
        P=(np.abs(X[:, :201])**2).astype(np.float32)
        d=o.power_to_db(P); d=np.float32(0.01)*(d-d.min()); return np.clip(d,-1,1)
    def rep(tag, X):
        got=pdb(X); err=np.abs(got-want); tol=1e-5+1e-4*np.abs(want)
        print(f"{name:18s} {tag:34s} max err {err.max():.2e} viol {(err>tol).sum():5d} / {err.size}")
    c64,c128,f32,f64=np.complex64,np.complex128,np.float32,np.float64
    rep("all f64 (sanity)", fft400(fr64,c128,f64,c128,f64,f64))
    rep("f32 inputs only", fft400(fr32in,c128,f64,c128,f64,f64))
    rep("f64 in, round frame to f32", fft400(fr64,c128,f64,c128,f64,f32))
    rep("step1 f32, step2 f64", fft400(fr64,c64,f32,c128,f64,f32))
    rep("step1 f64, step2 f32", fft400(fr64,c128,f64,c64,f32,f64))
    rep("all f32 (exact-in)", fft400(fr64,c64,f32,c64,f32,f32))
    rep("all f32 (f32 inputs)", fft400(fr32in,c64,f32,c64,f32,f32))

rng=np.random.default_rng(7); t=np.arange(16000)/16000.
run(synth.utterance(1000,3.0,ds_norm=(0,10.)),"hp3s")
run((0.1*np.sin(2*np.pi*1000*t)+1e-4*rng.standard_normal(16000)).astype(np.float32),"sine_on_bin")
run((0.1*np.sin(2*np.pi*1020*t)+1e-4*rng.standard_normal(16000)).astype(np.float32),"sine_between")
