"""torchrun -n 2 debug: where does the chunked run differ from the unchunked one?"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_cloner_b200 import audio_lib as al, distributed as D, synth, _lib

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hp = dict(synth.HP_ENC)
gl_wavs = synth.batch(3, 2, 5.0)
feats = al.calc_MFCC_input_batch(gl_wavs, return_device=True, **hp)
p1000 = feats[0][2][:1000].contiguous()
np.random.seed(3000)
ph1000 = torch.from_numpy((np.pi * np.random.rand(201, 1000)).T.astype(np.float32)).cuda()
plan = al.DspPlan.get(n_fft=400, win_length=400, hop_length=80)
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream

def report(tag, a, b):
    ne = (a != b)
    n = int(ne.sum().item())
    if n:
        idx = torch.nonzero(ne).flatten()
        print(f"[rank {rank}] {tag}: {n} of {a.numel()} differ, first {int(idx[0])} last {int(idx[-1])} max |d| {float((a - b).abs().max()):.3e}", flush=True)
    else:
        print(f"[rank {rank}] {tag}: identical ({a.numel()})", flush=True)

for T, n_iter, k in ((2001, 25, 4), (2001, 25, 20), (2001, 3, 1), (20001, 10, 4)):
    ii = torch.arange(T, device="cuda") % 1000
    P = p1000[ii].contiguous(); PH = ph1000[ii].contiguous()
    lay = al._GlLayout([T], 80)
    amp = torch.empty_like(P)
    _lib.check(lib.sc_power_to_amp_batch(plan._h, P.data_ptr(), lay.c_frame_offsets, lay.c_frame_counts, 1, 0.01, 1.0, amp.data_ptr(), st), "p2a")
    whole = al.griffin_lim_device(plan, amp, PH, lay, n_iter)[: 80 * (T - 1)].clone()
    c = D.ChunkedGriffinLim(T, 80, 400, steps_per_exchange=k)
    f_lo, f_hi = c.frame_range(n_iters=n_iter)
    chunk = c.run(amp[f_lo:f_hi].contiguous(), PH[f_lo:f_hi].contiguous(), n_iter)
    report(f"T={T} it={n_iter} k={k} bounds={c.bounds} GL chunk", chunk, whole[c.lo:c.hi])
    y = c.deemph_renorm(chunk, 0.97, 0.045)
    out = torch.empty(80 * (T - 1), dtype=torch.float64, device="cuda")
    _lib.check(lib.sc_deemph_renorm_batch(plan._h, whole.data_ptr(), _lib.i64_array([0, 80 * (T - 1)]), _lib.i64_array([80 * (T - 1)]), 1, 0.97, 0.045, out.data_ptr(), st), "de")
    report(f"T={T} epilogue", y, out[c.lo:c.hi])
    # p1000 differs per rank in this script (seed 3 + 10 rank)? it must not: print a checksum
    print(f"[rank {rank}] checksum P {float(P.double().sum()):.6f} PH {float(PH.double().sum()):.6f}", flush=True)
dist.destroy_process_group()
