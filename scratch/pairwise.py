import numpy as np
f32=np.float32
def pw(a):
    n=len(a)
    if n<8:
        r=f32(0.)
        for x in a: r=f32(r+x)
        return r
    if n<=128:
        r=[f32(a[i]) for i in range(8)]
        i=8
        while i < n-(n%8):
            for j in range(8): r[j]=f32(r[j]+a[i+j])
            i+=8
        res=f32(f32(f32(r[0]+r[1])+f32(r[2]+r[3]))+f32(f32(r[4]+r[5])+f32(r[6]+r[7])))
        while i<n:
            res=f32(res+a[i]); i+=1
        return res
    n2=n//2; n2-=n2%8
    return f32(pw(a[:n2])+pw(a[n2:]))
rng=np.random.default_rng(0)
ok_first=ok_zero=0; tot=0
for n in list(range(1,300))+[1000,4097,8191,8192,8193,16000,48000,48001,64000,100003]:
    a=np.abs(rng.standard_normal(n)).astype(f32)
    ref=a.sum()
    v0=pw(a)                         # identity start: 0 + pw(all)
    v1=f32(a[0]+pw(a[1:])) if n>1 else a[0]
    tot+=1; ok_zero+= (v0==ref); ok_first+=(v1==ref)
print("n cases",tot,"match 0+pw(all):",ok_zero,"match a0+pw(rest):",ok_first)
# mean: sum / n in f32?
a=np.abs(rng.standard_normal(48000)).astype(f32)
print(a.mean()==f32(a.sum()/f32(48000)), a.mean()==f32(np.float64(a.sum())/48000))
