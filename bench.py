#!/usr/bin/env python
"""bench.py -- the FrozenBiLM hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mlm|mc|bert] [--impl reference]

--workload mlm (default; BASELINE.json configs[1] / [2], the configuration the metric is quoted on):
    one "step" = forward + selective backward + gradient all-reduce + clip + Adam of DeBERTa-V2-XL (frozen) +
    adapters (ds 8) + 10-frame visual prefix on a WebVid-shaped synthetic batch (B = 32 per GPU, 10 frames + 342
    tokens, 15 % MLM labels, train-mode dropout 0.1), bf16 tensor-core GEMMs with fp32 accumulation.
--workload mc (configs[3]): multiple-choice inference (mc.py:150-172): B = 16 questions x 5 candidates of 128 tokens
    + 10 frames, DeBERTa-V2-XL, n_ans = 2; one step = all 80 rows through the encoder as one batch (visual prefix
    projected once per clip), head at the [MASK] rows, softmax; questions/s.
--workload bert (configs[4]): BERT-large (model/bert.py), no adapters, B = 64 per GPU, 10 frames + 256 tokens, train step.

Prints ONE JSON line (see the task's bench contract).  `--impl reference` times the reference's own implementation
of the same workload on the host cores: the UNMODIFIED reference model files staged under baseline/_ref
(baseline/install_ref.py; kind "reference"), or the oracle port when they are absent (kind "port"), on a bounded
sample of the workload per step.
"""
import argparse
import csv
import glob
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

from frozenbilm_b200.config import CONFIG_BERT_LARGE, CONFIG_XL  # noqa: E402
from frozenbilm_b200.synth import is_trainable, synth_batch, synth_params  # noqa: E402

N_FRAMES = 10
WORKLOADS = {
    # name: (metric, unit, per-GPU batch, text tokens, description)
    "mlm": ("samples/sec DeBERTa-XL+adapters seq352 bf16 train step (fwd+bwd+allreduce+clip+Adam)", "samples/s", 32, 342,
            "DeBERTa-V2-XL frozen + adapters ds8, 10 frames + 342 tokens (WebVid MLM shape), train step"),
    "mc": ("questions/sec DeBERTa-XL multiple-choice inference (5 candidates x 128 tokens + 10 frames)", "questions/s",
           16, 128, "mc.py path: DeBERTa-V2-XL n_ans=2, 16 questions x 5 candidates (10 frames + 128 tokens), "
                    "inference, candidates batched as [80, 138]"),
    "bert": ("samples/sec BERT-large seq266 bf16 train step (fwd+bwd+allreduce+clip+Adam)", "samples/s", 64, 256,
             "BERT-large frozen (LayerNorms + linear_video trainable), 10 frames + 256 tokens, train step"),
}
MC_CANDIDATES = 5
MASK_ID = 128000


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_dram_bytes(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the newest committed
    `ncu --set full` raw-page export under profiles/ (profiles/r*_ncu_raw_*.csv); None when there is none."""
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_raw_*.csv"))):
        try:
            rows = list(csv.reader(open(path, newline="")))
        except OSError:
            continue
        if len(rows) < 3:
            continue
        hdr = rows[0]
        try:
            kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        except ValueError:
            continue
        units = rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = []
        for r in rows[2:]:
            if len(r) > max(kn, rd, wr) and kernel_substr in r[kn]:
                try:
                    vals.append(float(r[rd].replace(",", "")) * scale.get(units[rd], 1.0) +
                                float(r[wr].replace(",", "")) * scale.get(units[wr], 1.0))
                except ValueError:
                    pass
        if vals:
            best = (sum(vals) / len(vals), os.path.relpath(path, ROOT))
    return best


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
# workload definitions shared by both arms
# ------------------------------------------------------------------------------------------------------
def workload_cfg(name, layers=0, tiny=False):
    if tiny:   # contract self-test only (tests/test_bench_contract_cpu.py): same code path, toy dimensions
        from frozenbilm_b200.config import CONFIG_BERT_TINY, CONFIG_TINY
        cfg = CONFIG_BERT_TINY if name == "bert" else CONFIG_TINY
    else:
        cfg = CONFIG_BERT_LARGE if name == "bert" else CONFIG_XL
    if name == "mc":
        cfg = cfg.replace(n_ans=2)
    if layers:
        cfg = cfg.replace(num_hidden_layers=layers)
    return cfg


def mc_batch(cfg, n_questions, n_text, seed):
    """[B, 10, 768] clips + [B*C, n_text] candidate rows (one [MASK] each), question-major."""
    b = synth_batch(cfg, n_questions * MC_CANDIDATES, n_text, seed, n_frames=N_FRAMES, ragged=False)
    g = torch.Generator().manual_seed(seed + 7)
    ids = b["input_ids"]
    mask_id = min(MASK_ID, cfg.vocab_size - 1)
    ids[ids == mask_id] = 5
    pos = torch.randint(1, n_text - 1, (ids.shape[0],), generator=g)
    ids[torch.arange(ids.shape[0]), pos] = mask_id
    return dict(video=b["video"][::MC_CANDIDATES].contiguous(), video_mask=b["video_mask"][::MC_CANDIDATES].contiguous(),
                input_ids=ids, attention_mask=b["attention_mask"]), mask_id


def algorithmic_flops_per_unit(name, cfg, n_text):
    """algorithmic FLOPs per sample (train step: fwd + selective bwd) or per question (mc, fwd only)."""
    L, H, S, NL = N_FRAMES + n_text, cfg.hidden_size, cfg.position_buckets, cfg.num_hidden_layers
    if name == "mlm":
        per_pass = L * H * (51 * H + 12 * L + 24 * S)
        n_lab = max(1, round(0.15 * n_text))
        return (NL + 1) * per_pass + 12 * L * H * H + 4 * L * H * H + 4 * n_lab * H * cfg.vocab_size
    if name == "mc":
        per_pass = L * H * (25 * H + 4 * L + 8 * S)
        per_row = (NL + 1) * per_pass + 6 * L * H * H + 2 * H * H
        return MC_CANDIDATES * per_row
    # bert: no adapters / positional terms; frozen linears dX only (x2), attention x3; head on the labelled rows
    n_lab = max(1, round(0.15 * n_text))
    return NL * L * H * (48 * H + 12 * L) + 4 * n_lab * H * H + 4 * n_lab * H * cfg.vocab_size


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (baseline/_ref) or the oracle port, timed on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_step_fn(name, cfg, batch, seed=1234, prefer_reference=True):
    """Returns (step(), units per step, kind, description).  Train workloads: fwd + loss.backward() +
    clip_grad_norm_(0.1) + Adam (main.py:75-84) in train mode with the workload's dropout; mc: the 5 per-candidate
    forwards of mc.py:150-172 under no_grad."""
    torch.set_num_threads(os.cpu_count() or 1)
    n_text = WORKLOADS[name][3]
    kind = "port"
    model = None
    if prefer_reference:
        try:
            from oracle import ref_loader
            if ref_loader.reference_model_files_available():
                kw = dict(max_feats=cfg.max_feats, features_dim=cfg.features_dim, n_ans=cfg.n_ans)
                if cfg.arch == "bert":
                    model = ref_loader.build_reference_bert(cfg.to_hf(), **kw)
                else:
                    model = ref_loader.build_reference_deberta(cfg.to_hf(), ds_factor_attn=cfg.ds_factor_attn,
                                                               ds_factor_ff=cfg.ds_factor_ff,
                                                               dropout=cfg.adapter_dropout, **kw)
                kind = "reference"
        except Exception as ex:  # noqa: BLE001
            print(f"[bench] reference model files unusable ({ex!r}); timing the oracle port", file=sys.stderr)
            model = None
    if name == "mc":
        b, mask_id = mc_batch(cfg, batch, n_text, seed + 1)
        ids = b["input_ids"].view(batch, MC_CANDIDATES, n_text)
        am = b["attention_mask"].view(batch, MC_CANDIDATES, n_text)
        if model is not None:
            model.eval()

            def step():
                with torch.no_grad():
                    sc = []
                    for c in range(MC_CANDIDATES):
                        lg = model(video=b["video"], video_mask=b["video_mask"], input_ids=ids[:, c].contiguous(),
                                   attention_mask=am[:, c].contiguous())["logits"]
                        sc.append(lg[:, N_FRAMES:][ids[:, c] == mask_id].softmax(-1)[:, 0])
                    return float(torch.stack(sc, 1).max(1).indices.sum())
        else:
            from oracle import frozenbilm_oracle as orc
            P = synth_params(cfg, seed)

            def step():
                with torch.no_grad():
                    sc = []
                    for c in range(MC_CANDIDATES):
                        lg = orc.forward(P, cfg, video=b["video"], video_mask=b["video_mask"],
                                         input_ids=ids[:, c].contiguous(), attention_mask=am[:, c].contiguous())["logits"]
                        sc.append(lg[:, N_FRAMES:][ids[:, c] == mask_id].softmax(-1)[:, 0])
                    return float(torch.stack(sc, 1).max(1).indices.sum())
        desc = f"B={batch} questions x {MC_CANDIDATES} per-candidate forwards (10 frames + {n_text} tokens), fp32, no_grad"
        return step, batch, kind, desc
    b = synth_batch(cfg, batch, n_text, seed + 1, n_frames=N_FRAMES, ragged=False)
    if model is not None:
        model.train()
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=3e-5, betas=(0.9, 0.95))

        def step():
            opt.zero_grad()
            out = model(video=b["video"], video_mask=b["video_mask"], input_ids=b["input_ids"],
                        attention_mask=b["attention_mask"], labels=b["labels"])
            out["loss"].backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
            opt.step()
            return float(out["loss"].detach())
        drop = f"dropout {cfg.hidden_dropout_prob}"
    else:
        from oracle import frozenbilm_oracle as orc
        cfg0 = cfg.replace(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, adapter_dropout=0.0)
        P = synth_params(cfg0, seed)
        names_t = [n for n in P if is_trainable(n, cfg0)]
        for n in names_t:
            P[n].requires_grad_(True)
        opt = torch.optim.Adam([P[n] for n in names_t], lr=3e-5, betas=(0.9, 0.95))
        fwd = orc.bert_forward if cfg.arch == "bert" else orc.forward

        def step():
            opt.zero_grad()
            out = fwd(P, cfg0, video=b["video"], video_mask=b["video_mask"], input_ids=b["input_ids"],
                      attention_mask=b["attention_mask"], labels=b["labels"])
            out["loss"].backward()
            torch.nn.utils.clip_grad_norm_([P[n] for n in names_t], 0.1)
            opt.step()
            return float(out["loss"].detach())
        drop = "dropout off (the port has no dropout)"
    desc = f"B={batch} per step (10 frames + {n_text} tokens), fp32, fwd+bwd+clip+Adam, train mode, {drop}"
    return step, batch, kind, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    metric, unit, _, n_text, wdesc = WORKLOADS[name]
    cfg = workload_cfg(name, args.layers, args.tiny)
    # bounded sample of the workload per step: B_cpu = 2 (the reference materialises ~2-3 GB of saved attention
    # tensors per layer pass at B = 32; one B = 2 XL train step of the unmodified reference takes ~10 s on 8 cores, so
    # the driver's 20 + 5 steps end within a few minutes -- SURVEY 8d's B_cpu = 4 would double that for the same
    # samples/s)
    bsz = args.cpu_batch or 2
    step, units, kind, sample = cpu_step_fn(name, cfg, bsz)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = units * args.steps / dt
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("TOY MODEL (contract self-test), " if args.tiny else "") + wdesc,
                       "global_batch": units, "seq_len": N_FRAMES + n_text, "parallelism": "cpu",
                       "layers": cfg.num_hidden_layers},
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def build_model(name, cfg, dev):
    from frozenbilm_b200.model import BertForMaskedLM, DebertaV2ForMaskedLM
    torch.manual_seed(1234)
    # weights: seed 1234 on every rank (random init of the architecture; there are no checkpoints offline)
    if cfg.arch == "bert":
        model = BertForMaskedLM(cfg.to_hf(), max_feats=cfg.max_feats, features_dim=cfg.features_dim, n_ans=cfg.n_ans)
    else:
        model = DebertaV2ForMaskedLM(cfg.to_hf(), max_feats=cfg.max_feats, features_dim=cfg.features_dim,
                                     ds_factor_attn=cfg.ds_factor_attn, ds_factor_ff=cfg.ds_factor_ff,
                                     dropout=cfg.adapter_dropout, n_ans=cfg.n_ans)
        model.lm_predictions.lm_head.decoder.weight.data = torch.empty(0)   # unused by forward (deberta.py:1535-1542)
    return model.to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="mlm", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--cpu-batch", type=int, default=0, help="CPU legs: units per step of the bounded sample")
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
    name = args.workload
    metric, unit, B_default, n_text, wdesc = WORKLOADS[name]

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from frozenbilm_b200 import lib
    from frozenbilm_b200.optim import FusedClipAdam, broadcast_trainable

    cfg = workload_cfg(name, args.layers)
    model = build_model(name, cfg, dev)
    train = name != "mc"
    B = args.batch or B_default
    L = N_FRAMES + n_text
    # per-rank data shard: seed 1234 + rank; a ring of distinct batches so inputs are not reused back to back
    n_batches = 4
    host = []
    mask_id = None
    for i in range(n_batches):
        if name == "mc":
            b, mask_id = mc_batch(cfg, B, n_text, 1234 + rank + 1000 * i)
        else:
            b = synth_batch(cfg, B, n_text, 1234 + rank + 1000 * i, n_frames=N_FRAMES, ragged=False)
        host.append({k: v.pin_memory() for k, v in b.items()})
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    if train:
        model.train()
        opt = FusedClipAdam(model, lr=3e-5, betas=(0.9, 0.95), max_norm=0.1)
        broadcast_trainable(model)

        def run(b):
            out = model(video=b["video"], video_mask=b["video_mask"], input_ids=b["input_ids"],
                        attention_mask=b["attention_mask"], labels=b["labels"])
            out["loss"].backward()
            opt.step()
            return out["loss"]
        d2h_bytes = 4
    else:
        model.eval()

        def run(b):
            sc = model.answer_scores(b["input_ids"], b["attention_mask"], video=b["video"], video_mask=b["video_mask"],
                                     mask_token_id=mask_id, candidates=MC_CANDIDATES, return_probs=True)
            return sc["probs"][:, 0].view(B, MC_CANDIDATES).max(1).indices     # mc.py:172-177: the predicted candidate
        d2h_bytes = 8 * B

    def step_resident(i):
        return run(devb[i % n_batches])

    def step_e2e(i):
        hb = host[i % n_batches]
        r = run({k: v.to(dev, non_blocking=True) for k, v in hb.items()})
        return r.cpu()   # D2H read of the step's result (the loss / the predicted candidates)

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

    # ---- per-kernel roofline probes, timed live with CUDA events around each launch inside the timed region, on the
    #      stream the kernel is launched on: the FFN up-projection GEMM (the GEMM family's biggest member) and the
    #      attention kernels.  The DOMINANT kernel (largest measured share of the step) goes into "roofline".
    T = B * L * (MC_CANDIDATES if name == "mc" else 1)
    H, I = cfg.hidden_size, cfg.intermediate_size
    S = cfg.position_buckets
    rel = cfg.arch != "bert"
    import frozenbilm_b200.engine as engmod
    probes = {k: {"events": [], "ms": 0.0, "n": 0} for k in ("ffn1", "attn_fwd", "attn_bwd", "attn_bwd_rel")}
    probe_on = {"on": False}
    originals = {}

    def wrap(fname, key, select=None):
        orig = getattr(lib, fname)
        originals[fname] = orig

        def f(*a, **kw):
            if not probe_on["on"] or (select is not None and not select(a, kw)):
                return orig(*a, **kw)
            st = torch.cuda.current_stream()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(st)
            r = orig(*a, **kw)
            e.record(st)
            probes[key]["events"].append((s, e))
            return r

        setattr(engmod.lib, fname, f)

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
    for fname, orig in originals.items():
        setattr(engmod.lib, fname, orig)

    # ---- end-to-end arm through the public API with host buffers
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)

    units = B * world
    value = units * args.steps / (ms / 1e3)
    e2e_value = units * args.steps / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_src = load_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    flops_unit = algorithmic_flops_per_unit(name, cfg, n_text)
    rows = T / L                                     # sequences per launch
    # algorithmic FLOPs per launch: attention share L*H*(4L + 8S) forward; backward = dP, dV, dQ, dK
    # content terms (8L) + positional terms of dQ / dK (8S) in fb_attn_bwd; the gradients of the two positional
    # projections (8S) in fb_attn_bwd_rel
    kern = {
        "ffn1": ("fb::gemm_kernel<256> FFN up-projection %dx%dx%d (+bias+GELU)" % (T, I, H), 2.0 * T * I * H,
                 2.0 * (T * H + I * H + (2 if train else 1) * T * I), "gemm_kernel<256, 0, 0, 0, 2"),
        "attn_fwd": ("fb::attn_fwd_kernel (QK^T + c2p + p2c + softmax + PV)", rows * L * H * (4 * L + (8 * S if rel else 0)),
                     2.0 * (4 * T * H) + (2.0 * rows * cfg.num_attention_heads * L * ((L + 127) // 128 * 128) if train else 0),
                     "attn_fwd"),
        "attn_bwd": ("fb::attn_bwd_kernel (dP, dS, dV, dQ, dK incl. positional terms)",
                     rows * L * H * (8 * L + (8 * S if rel else 0)),
                     2.0 * (8 * T * H) + 2.0 * rows * cfg.num_attention_heads * L * ((L + 127) // 128 * 128) * (2 if rel else 1),
                     "attn_bwd_kernel"),
        "attn_bwd_rel": ("fb::attn_bwd_rel_kernel (gradients of the positional projections)", rows * L * H * 8 * S,
                         2.0 * (2 * T * H) + 2.0 * rows * cfg.num_attention_heads * L * ((L + 127) // 128 * 128),
                         "attn_bwd_rel_kernel"),
    }
    roofs = {}
    for key, (label, fl, alg_bytes, ncu_name) in kern.items():
        pr = probes[key]
        if not pr["n"]:
            continue
        us = 1e3 * pr["ms"] / pr["n"]
        tr = ncu_dram_bytes(ncu_name) if (name == "mlm" and B == B_default and not args.layers) else None
        roofs[key] = {"bound": "tensor", "kernel": label, "achieved": fl / us / 1e6, "peak": peak_tf, "unit": "TFLOP/s",
                      "frac": fl / us / 1e6 / peak_tf, "us_per_launch": us, "launches_timed": pr["n"],
                      "share_of_step": pr["ms"] / ms, "algorithmic_flop_per_launch": fl,
                      "algorithmic_bytes": alg_bytes, "traffic": tr[0] if tr else None,
                      "traffic_unit": "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)",
                      "traffic_source": tr[1] if tr else None,
                      "peak_source": peak_src + " (bf16_tflops_sustained: kernel timed inside a long step)"}
    dominant = max(roofs, key=lambda k: roofs[k]["share_of_step"]) if roofs else None
    tflops_step = flops_unit * B * args.steps / (ms / 1e3) / 1e12
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wdesc, "global_batch": units, "seq_len": L, "parallelism": f"dp{world}",
                   "layers": cfg.num_hidden_layers,
                   "l2": "per-step working set (GBs of saved activations + the bf16 weights) >> 126 MB L2; "
                         "4 distinct input batches cycled",
                   "model_tflops_per_gpu": tflops_step,
                   "algorithmic_tflop_per_unit": flops_unit / 1e12},
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clk,
        # the kernel with the largest measured share of the step
        "roofline": dict(roofs[dominant], dominant_of=sorted(roofs)) if dominant else None,
        "roofline_gemm": roofs.get("ffn1"),
        "roofline_attention": {k: v for k, v in roofs.items() if k.startswith("attn")},
        # whole step: algorithmic FLOPs of the workload / step time, against the same peak
        "roofline_step": {"bound": "tensor", "achieved": tflops_step, "peak": peak_tf, "unit": "TFLOP/s",
                          "frac": tflops_step / peak_tf},
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            # bounded sample of the same workload on the host cores: 1 untimed + 2 timed steps
            step, n_units, kind, sample = cpu_step_fn(name, cfg, args.cpu_batch or 2)
            step()
            t0 = time.perf_counter()
            for _ in range(2):
                step()
            dt = (time.perf_counter() - t0) / 2
            line["cpu_baseline"] = {"value": n_units / dt, "unit": unit, "cores": torch.get_num_threads(), "kind": kind,
                                    "sample": "2 timed steps (after 1 warm-up) of " + sample}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "error": repr(ex)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
