#!/usr/bin/env python
"""bench.py -- FrozenBiLM masked-LM training step (BASELINE.json configs[1]) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = forward + selective backward + gradient all-reduce + clip + Adam of DeBERTa-V2-XL (frozen)
+ adapters (ds 8) + 10-frame visual prefix on a WebVid-shaped synthetic batch (B = 32 per GPU, 10 frames +
342 tokens, 15 % MLM labels, train-mode dropout 0.1), bf16 tensor-core GEMMs with fp32 accumulation.
Prints ONE JSON line (see the task's bench contract).  `--impl reference` times the CPU oracle port of the
reference's own path (oracle/frozenbilm_oracle.py; /root/reference does not exist on the GPU box) on the
host cores with a bounded sample (B = 1 per step) of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from frozenbilm_b200.config import CONFIG_XL  # noqa: E402
from frozenbilm_b200.synth import synth_batch, synth_params, is_trainable  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum of one FFN up-projection launch at B = 32 (ncu --set full capture,
# profiles/r1_ncu_full_summary.md: 66.5 MB read + 227.0 MB written)
NCU_FFN1_DRAM_BYTES = 293.5e6
METRIC = "samples/sec DeBERTa-XL+adapters seq352 bf16 train step (fwd+bwd+allreduce+clip+Adam)"
B_PER_GPU, N_FRAMES, N_TEXT = 32, 10, 342


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path, timed on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_step_fn(cfg, seed=1234, batch=1):
    from oracle import frozenbilm_oracle as orc

    torch.set_num_threads(os.cpu_count() or 1)
    P = synth_params(cfg, seed)
    names_t = [n for n in P if is_trainable(n, cfg)]
    for n in names_t:
        P[n].requires_grad_(True)
    b = synth_batch(cfg, batch, N_TEXT, seed + 1, n_frames=N_FRAMES, ragged=False)
    opt = torch.optim.Adam([P[n] for n in names_t], lr=3e-5, betas=(0.9, 0.95))

    def step():
        opt.zero_grad()
        out = orc.forward(P, cfg, video=b["video"], video_mask=b["video_mask"], input_ids=b["input_ids"],
                          attention_mask=b["attention_mask"], labels=b["labels"])
        out["loss"].backward()
        torch.nn.utils.clip_grad_norm_([P[n] for n in names_t], 0.1)
        opt.step()
        return float(out["loss"].detach())

    return step, batch


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = CONFIG_XL
    if args.tiny:   # contract self-test only (tests/test_bench_contract_cpu.py): same code path, toy dimensions
        from frozenbilm_b200.config import CONFIG_TINY
        base = CONFIG_TINY
    cfg = base.replace(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, adapter_dropout=0.0)
    step, bsz = cpu_step_fn(cfg, batch=1)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = bsz * args.steps / dt
    cores = torch.get_num_threads()
    sample = f"B=1 per step ({N_FRAMES} frames + {N_TEXT} tokens), fp32, fwd+bwd+clip+Adam, dropout off"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("TOY MODEL (contract self-test), " if args.tiny else "") +
                                   "DeBERTa-V2-XL frozen + adapters ds8, 10 frames + 342 tokens (WebVid MLM shape)",
                       "global_batch": bsz, "seq_len": N_FRAMES + N_TEXT, "parallelism": "cpu"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", type=int, default=0, help="debug: override the layer count")
    ap.add_argument("--tiny", action="store_true", help="reference arm only: toy model (JSON-contract self-test)")
    ap.add_argument("--profile-step", action="store_true",
                    help="profiling aid (ncu --profile-from-start off): warm up, bracket ONE resident step with "
                         "cudaProfilerStart/Stop, print nothing, exit")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from frozenbilm_b200 import lib
    from frozenbilm_b200.model import DebertaV2ForMaskedLM
    from frozenbilm_b200.optim import FusedClipAdam, broadcast_trainable

    cfg = CONFIG_XL if not args.layers else CONFIG_XL.replace(num_hidden_layers=args.layers)
    torch.manual_seed(1234)
    # weights: seed 1234 on every rank (random init of the XL architecture; there are no checkpoints offline)
    model = DebertaV2ForMaskedLM(cfg.to_hf(), max_feats=cfg.max_feats, features_dim=cfg.features_dim,
                                 ds_factor_attn=cfg.ds_factor_attn, ds_factor_ff=cfg.ds_factor_ff,
                                 dropout=cfg.adapter_dropout, n_ans=0)
    model.lm_predictions.lm_head.decoder.weight.data = torch.empty(0)   # unused by forward (deberta.py:1535-1542)
    model.to(dev)
    model.train()
    opt = FusedClipAdam(model, lr=3e-5, betas=(0.9, 0.95), max_norm=0.1)
    broadcast_trainable(model)

    B = args.batch
    # per-rank data shard: seed 1234 + rank; a ring of distinct batches so inputs are not reused back to back
    n_batches = 4
    host = []
    for i in range(n_batches):
        b = synth_batch(cfg, B, N_TEXT, 1234 + rank + 1000 * i, n_frames=N_FRAMES, ragged=False)
        host.append({k: v.pin_memory() for k, v in b.items()})
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def step_resident(i):
        b = devb[i % n_batches]
        out = model(video=b["video"], video_mask=b["video_mask"], input_ids=b["input_ids"],
                    attention_mask=b["attention_mask"], labels=b["labels"])
        out["loss"].backward()
        opt.step()
        return out["loss"]

    def step_e2e(i):
        hb = host[i % n_batches]
        b = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
        out = model(video=b["video"], video_mask=b["video_mask"], input_ids=b["input_ids"],
                    attention_mask=b["attention_mask"], labels=b["labels"])
        out["loss"].backward()
        opt.step()
        return float(out["loss"].item())   # D2H read of the step's result (4 bytes)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- warm-up (also sizes every workspace buffer)
    for i in range(args.warmup):
        step_resident(i)
    barrier()

    if args.profile_step:
        torch.cuda.profiler.start()
        step_resident(args.warmup)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    # ---- per-kernel roofline probes, timed live with CUDA events around each launch inside the timed region:
    #      the FFN up-projection GEMM (M = B*L tokens, N = 4H, K = H; fb::gemm_kernel is the dominant kernel family,
    #      ~46 % of the step) and the three attention kernels (the rest of the tensor-core work)
    eng = model._engine_obj
    T = B * (N_FRAMES + N_TEXT)
    H, I = cfg.hidden_size, cfg.intermediate_size
    import frozenbilm_b200.engine as engmod
    probes = {k: {"events": [], "ms": 0.0, "n": 0} for k in ("ffn1", "attn_fwd", "attn_bwd", "attn_bwd_rel")}
    probe_on = {"on": False}
    originals = {}

    def wrap(name, key, select=None):
        orig = getattr(lib, name)
        originals[name] = orig

        def f(*a, **kw):
            if not probe_on["on"] or (select is not None and not select(a, kw)):
                return orig(*a, **kw)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig(*a, **kw)
            e.record()
            probes[key]["events"].append((s, e))
            return r

        setattr(engmod.lib, name, f)

    wrap("gemm", "ffn1", lambda a, kw: kw.get("act") == lib.FB_ACT_GELU and a[1].shape[0] == I and a[0].shape[0] == T)
    wrap("attn_fwd", "attn_fwd")
    wrap("attn_bwd", "attn_bwd")
    wrap("attn_bwd_rel", "attn_bwd_rel")

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    probe_on["on"] = True
    l0 = lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = lib.launch_count() - l0
    probe_on["on"] = False
    clk = clocks.stop() if rank == 0 else None
    for pr in probes.values():
        for s, e in pr["events"]:
            pr["ms"] += s.elapsed_time(e)
            pr["n"] += 1
    for name, orig in originals.items():
        setattr(engmod.lib, name, orig)
    probe = probes["ffn1"]

    # ---- end-to-end arm through the public API with host buffers
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)

    global_batch = B * world
    value = global_batch * args.steps / (ms / 1e3)
    e2e_value = global_batch * args.steps / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_src = load_peaks()
    ffn_flops = 2.0 * T * I * H
    ach = (ffn_flops / (probe["ms"] / probe["n"] / 1e3) / 1e12) if probe["n"] else None
    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    L = N_FRAMES + N_TEXT
    # algorithmic fwd+bwd FLOPs per sample (SURVEY 8d / BASELINE.md section 3)
    S = cfg.position_buckets
    n_pass = cfg.num_hidden_layers + 1
    per_pass = L * H * (51 * H + 12 * L + 24 * S)
    n_lab = max(1, round(0.15 * N_TEXT))
    flops_sample = n_pass * per_pass + 12 * L * H * H + 4 * L * H * H + 4 * n_lab * H * cfg.vocab_size
    # attention share of the tensor-core work (L*H*(4L + 8S) forward, x3 with the backward), per launch
    attn_fwd_flops = float(B) * L * H * (4 * L + 8 * S)
    attn_roof = {"bound": "tensor", "unit": "TFLOP/s", "peak": peak_tf,
                 "note": "algorithmic FLOPs (SURVEY 8d) / live CUDA-event time per launch; these kernels are bound by "
                         "the softmax / relative-position regrouping on the CUDA cores, not by the tensor pipe"}
    # backward = 2x forward: dP, dV, dQ, dK content terms (8L) + positional terms of dQ / dK (8S) in fb_attn_bwd,
    # gradients of the two positional projections (8S) in fb_attn_bwd_rel
    for key, fl in (("attn_fwd", attn_fwd_flops), ("attn_bwd", float(B) * L * H * (8 * L + 8 * S)),
                    ("attn_bwd_rel", float(B) * L * H * 8 * S)):
        pr = probes[key]
        if pr["n"]:
            us = 1e3 * pr["ms"] / pr["n"]
            attn_roof[key] = {"us_per_launch": us, "achieved": fl / us / 1e6, "frac": fl / us / 1e6 / peak_tf,
                              "launches_timed": pr["n"]}
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "DeBERTa-V2-XL frozen + adapters ds8, 10 frames + 342 tokens (WebVid MLM shape), "
                               "train step", "global_batch": global_batch, "seq_len": L,
                   "parallelism": f"dp{world}", "layers": cfg.num_hidden_layers,
                   "l2": "per-step working set (~13 GB of saved activations + 2.2 GB of weights) >> 126 MB L2; "
                         "4 distinct input batches cycled",
                   "model_tflops_per_gpu": flops_sample * B * args.steps / (ms / 1e3) / 1e12,
                   "algorithmic_tflop_per_sample": flops_sample / 1e12},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "fb::gemm_kernel<256> FFN up-projection 11264x6144x1536 (+bias+GELU)",
                     "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": (ach / peak_tf) if ach else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel (ncu --set full, per launch)
                     "traffic": NCU_FFN1_DRAM_BYTES if B == B_PER_GPU else None,
                     "traffic_unit": "bytes/launch", "traffic_source": "profiles/r1_ncu_full_summary.md",
                     "algorithmic_bytes": 2.0 * (T * H + I * H + 2 * T * I),
                     "peak_source": peak_src + " (bf16_tflops_sustained)",
                     "launches_timed": probe["n"]},
        "roofline_attention": attn_roof,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            cfg_cpu = cfg.replace(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, adapter_dropout=0.0)
            # bounded sample of the same workload: B = 2 per step, 1 untimed + 3 timed steps (about 10-15 s of CPU work)
            step, bsz = cpu_step_fn(cfg_cpu, batch=2)
            step()
            t0 = time.perf_counter()
            for _ in range(3):
                step()
            dt = (time.perf_counter() - t0) / 3
            line["cpu_baseline"] = {"value": bsz / dt, "unit": "samples/s", "cores": torch.get_num_threads(),
                                    "kind": "port", "sample": "3 timed steps (after 1 warm-up) of B=2 (10 frames + 342 "
                                    "tokens), fp32 oracle port of the reference path, fwd+bwd+clip+Adam, dropout off"}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "error": repr(ex)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
