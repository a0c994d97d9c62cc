"""Window samplers on a device-resident cache: time per batch and copy rate of
sc_window_gather against the reference's host loop (oracle.spec_window_sampler over the npz cache) on the same cache.

    python scripts/sampler_bench.py [n_utts] [seconds]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import audio_lib_oracle as oracle                     # noqa: E402  (the CPU arm of this script only)
from speech_cloner_b200 import dataset_cache as dc                # noqa: E402

n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
T, n_t, B = 1 + int(seconds * 16000) // 80, 400, 32
rng = np.random.default_rng(0)
rows = n_utts * T
groups = {"mfcc": torch.rand((rows, 80), device="cuda"), "mel_dB": torch.rand((rows, 80), device="cuda"),
          "power_dB": torch.rand((rows, 201), device="cuda")}
cache = dc.DeviceSpecCache(groups, np.arange(n_utts) * T, np.full(n_utts, T))
ids = np.arange(n_utts)


def run(n_epochs):
    np.random.seed(1)
    n = 0
    for _ in dc.spec_window_sampler(cache, ids, n_t, batch_size=B, n_epochs=n_epochs, prop_val=0.0, verbose=False):
        n += 1
    return n


run(1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
n_batches = run(4)
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = e0.elapsed_time(e1)
bytes_batch = B * n_t * 361 * 4
# kernel alone: one batch plan replayed
first = (np.arange(B) * T + 7).astype(np.int64)
valid = np.full(B, n_t, np.int32)
for _ in range(5):
    cache.gather(dc.DeviceSpecCache.FEATURES, first, valid, n_t)
torch.cuda.synchronize()
big_first = (rng.integers(0, n_utts, size=2048) * T + rng.integers(0, T - n_t, size=2048)).astype(np.int64)
big_valid = np.full(2048, n_t, np.int32)
cache.gather(dc.DeviceSpecCache.FEATURES, big_first, big_valid, n_t)
torch.cuda.synchronize()
e0.record()
cache.gather(dc.DeviceSpecCache.FEATURES, big_first, big_valid, n_t)
e1.record()
torch.cuda.synchronize()
ms_big = e0.elapsed_time(e1)

# the reference's host loop on the same data (a slice of the cache, host arrays)
class _HostCache(dict):
    pass


n_cpu = min(n_utts, 256)
host = {g: {str(i): t[i * T:(i + 1) * T].cpu().numpy() for i in range(n_cpu)} for g, t in groups.items()}
np.random.seed(1)
t0 = time.perf_counter()
n_cpu_batches = sum(1 for _ in oracle.spec_window_sampler(host, np.arange(n_cpu), n_t, batch_size=B, n_epochs=4, prop_val=0.0))
cpu_s = time.perf_counter() - t0
print(json.dumps({
    "cache": {"utterances": n_utts, "rows": rows, "GB": round(rows * 361 * 4 / 1e9, 2)},
    "batch": {"windows": B, "n_timesteps": n_t, "MB": round(bytes_batch / 1e6, 2)},
    "device_sampler": {"batches": n_batches, "us_per_batch_stream": round(1e3 * ms / n_batches, 1),
                       "us_per_batch_wall": round(1e6 * wall / n_batches, 1)},
    "gather_2048_windows": {"ms": round(ms_big, 3), "GB/s_read_plus_write": round(2 * 2048 * n_t * 361 * 4 / ms_big / 1e6, 1)},
    "host_reference_loop": {"batches": n_cpu_batches, "us_per_batch": round(1e6 * cpu_s / max(n_cpu_batches, 1), 1),
                            "note": "in-memory arrays: the reference's h5py reads would add to this"},
}))
