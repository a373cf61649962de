#!/usr/bin/env python
"""How much of the step is launch overhead?  Captures ONE whole train step (forward + backward + clip + Adam) of
bench.py's mlm workload in a CUDA graph (fixed batch, frozen seeds: a timing experiment, not a training path) and
compares replay time with the eager step on the same batch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from frozenbilm_b200.optim import FusedClipAdam  # noqa: E402
from frozenbilm_b200.synth import synth_batch  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = bench.workload_cfg("mlm", int(os.environ.get("LAYERS", "0")))
model = bench.build_model("mlm", cfg, dev).train()
opt = FusedClipAdam(model, lr=3e-5, betas=(0.9, 0.95), max_norm=0.1)
b = {k: v.to(dev) for k, v in synth_batch(cfg, 32, 342, 1234, n_frames=10, ragged=False).items()}


def run():
    out = model(video=b["video"], video_mask=b["video_mask"], input_ids=b["input_ids"],
                attention_mask=b["attention_mask"], labels=b["labels"])
    out["loss"].backward()
    opt.step()
    return out["loss"]


def timed(fn, n):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n, 1e3 * (t1 - t0) / n


for _ in range(3):
    run()
ms, host = timed(run, 6)
print(f"eager: {ms:.2f} ms/step (host loop {host:.2f} ms/step incl. the per-step sync)")

# capture: the labelled-row count is precomputed (reading it back is the forward's only host wait)
n_lab = int((b["labels"] != -100).sum())


class _Done:
    def synchronize(self):
        pass


model._count_host = torch.tensor([n_lab])
model._count_async = lambda flags: _Done()
ms, host = timed(run, 6)
print(f"eager, count known on the host: {ms:.2f} ms/step (host enqueue {host:.2f} ms/step)")
g = torch.cuda.CUDAGraph()
cs = torch.cuda.Stream()
cs.wait_stream(torch.cuda.current_stream())
with torch.cuda.graph(g, stream=cs, capture_error_mode="relaxed"):
    run()
torch.cuda.synchronize()
for _ in range(3):
    g.replay()
ms, host = timed(g.replay, 10)
print(f"graph replay: {ms:.2f} ms/step (host {host:.3f} ms/step)")
