#!/usr/bin/env python
"""Benchmark of the NACF inference hot path (BASELINE.json config 2) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16x3|bf16|fp32]

A *step* = one pass of the hot path over one batch of synthetic input: ``model.encode`` +
``Translator.translate_batch`` (mask-predict T=5 + coarse-grained templates = 6 decoder passes,
length beam 6) for B=128 videos of MSRVTT shape (2 x 60 x 2048 features, max_len 30, 6-layer
d_model 512, vocab 10547).  Metric: captions/s (one caption = one video's final hypothesis).

  value        kernel-side throughput, inputs already resident in HBM, CUDA-event timed.
  e2e          same metric through the public API with HOST inputs: per step the pinned-host ->
               device copy of the features/category and the device -> host read of the token ids
               are inside the timed region (inputs of step i+1 prefetched on a copy stream, ids of
               step i-1 awaited after step i has been launched: a one-step software pipeline).
  roofline     every launch class of the decoder layer (qkv, self, so, cq, cross, co, f1, f2, vocab, ...)
               timed live with CUDA events on the launching stream: the decode graph is re-captured with
               event-record nodes around every launch class and replayed (device time, no host latency);
               `classes` holds {us, share, flops / bytes, frac_useful, frac_issued} per class and the
               top-level fields describe the class with the largest time share.
  parity       token ids of the GPU path (fp32 mode and the timed mode, through the replayed CUDA graph)
               against the CPU reference's ids on the same 16 videos, video by video, with the oracle's
               decision margins (N=1 only).
  cpu_baseline the UNMODIFIED reference (baseline/_ref or /root/reference; else the oracle port) on the
               host cores, on a bounded sample of the same workload (N=1, rank 0 only).
  extra.train  data-parallel training (BASELINE configs 3 and 5): samples/s with the single NCCL gradient
               all-reduce inside the timed step; config 5 = NACF global batch 1024 split over the N ranks.
  --impl reference   times the reference's CPU implementation alone (rank 0 only under torchrun).
Multi-GPU inference: videos are independent -> each rank decodes its own B=128 batch (weak scaling), no
data-path collective; time = max over ranks.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import cases  # noqa: E402

METRIC = "captions/sec (NACF, max_len=30, n_frames=60)"
UNIT = "captions/s"
WORKLOAD = "NACF 6-layer d512 h8, feats 2x60x2048, vocab 10547, B=128/GPU, mask-predict T=5 + CT (6 passes), lbs=6"
CONFIG = {"workload": WORKLOAD}   # identical in both arms; run details live under "details"
PARITY_B = 16                     # videos of the parity / cpu_baseline sample
MARGIN = {"fp32": 2e-5, "bf16x3": 1e-4, "tf32": 5e-3, "bf16": 2e-2}   # log-unit error bound of each mode (tests/test_gpu_parity.py)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clocks / throttle reasons sampled during the timed regions: an NVML thread in this process (a few
    cheap driver queries every 50 ms); `nvidia-smi -lms` as a separate process when NVML is not importable or
    $NAVC_SAMPLER=smi (heavier: every poll re-enters the driver through a fresh query of all fields)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.proc, self.thread, self.stop_flag = index, [], None, None, False

    def _nvml_loop(self, nv, handle):
        R = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
             "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
             "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
             "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        mx = nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                watts = nv.nvmlDeviceGetPowerUsage(handle) / 1000.0
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                bits = int(get(handle))
                self.samples.append([str(mhz), str(mx), "%.2f" % watts] +
                                    ["Active" if bits & R[n] else "Not Active"
                                     for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        if os.environ.get("NAVC_NO_SAMPLER"):
            return
        if os.environ.get("NAVC_SAMPLER", "nvml") != "smi":
            try:
                import threading
                import pynvml as nv
                nv.nvmlInit()
                # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [int(x) for x in vis.split(",")] if vis and all(x.strip().isdigit() for x in vis.split(",")) else None
                phys = ids[self.index] if ids and self.index < len(ids) else self.index
                handle = nv.nvmlDeviceGetHandleByIndex(phys)
                self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
                self.thread.start()
                return
            except Exception:
                self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 7 and f[0].replace(".", "", 1).isdigit():
                self.samples.append(f)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        busy = [s for s in self.samples if float(s[2]) > 250.0] or self.samples  # samples taken under load
        mhz = sorted(int(float(s[0])) for s in busy)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons,
                "samples": len(self.samples), "samples_under_load": len(busy),
                "power_w_max": max(float(s[2]) for s in self.samples),
                "sampler": "nvml thread, 50 ms" if self.thread is not None else "nvidia-smi -lms 200"}


# ------------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (preferred) or the oracle port
# ------------------------------------------------------------------------------------------------------
def reference_kind():
    """'reference' when the unmodified reference is importable (from /root/reference, or from baseline/_ref with every
    file matching the digests recorded at install time), else 'port' (the oracle restatement)."""
    import refutil
    if not refutil.reference_available():
        return "port"
    if os.path.realpath(refutil.REF_ROOT) == os.path.realpath(refutil.INSTALLED):
        import install_reference
        return "reference" if install_reference.verify() else "port"
    return "reference"


def cpu_run(opt, batch, steps, warmup, kind):
    """encode + Translator.translate_batch on the host cores (all threads), perf_counter timed
    (misc/run.py:130-141 drives the reference exactly so).  Returns (captions/s, s/step, ids [B, Smax])."""
    torch.set_num_threads(os.cpu_count() or 1)
    feats, category = cases.synth_inputs(opt, batch)
    if kind == "reference":
        import refutil
        model = refutil.ref_get_model(opt, seed=0)   # models.get_model(opt) under torch.manual_seed(0), .eval()
        vocab = {i: "w%d" % i for i in range(opt["vocab_size"])}
        with refutil.reference_on_path():
            from models.Translator import Translator
            tr = Translator(model, dict(opt), device=torch.device("cpu"), teacher_model=None)

            def once():
                with torch.no_grad():
                    enc = model.encode(feats=list(feats))
                    hyp, _ = tr.translate_batch(enc, category, None, vocab)
                return hyp

            for _ in range(warmup):
                once()
            t0 = time.perf_counter()
            for _ in range(steps):
                hyp = once()
            dt = time.perf_counter() - t0
    else:
        from oracle import navc_oracle as O
        import navc_b200
        torch.manual_seed(0)
        model = navc_b200.get_model(opt)  # parameter container only (same seeded init as the reference factory)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        for _ in range(warmup):
            O.translate(sd, opt, feats, category)
        t0 = time.perf_counter()
        for _ in range(steps):
            hyp = O.translate(sd, opt, feats, category)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, hyp


def run_reference(args, opt, rank):
    if rank != 0:
        return
    kind = reference_kind()
    # bounded sample: probe B=4 once, then the largest batch for which the whole run ends within ~4 min
    _, t4, _ = cpu_run(opt, 4, 1, 0, kind)
    budget = 240.0 / max(1, args.steps + args.warmup)
    batch = 4
    cap = int(os.environ.get("NAVC_BENCH_REF_MAXB", "128"))   # (tests bound the sample)
    for cand in (128, 64, 32, 16, 8):
        if cand <= cap and t4 * cand / 4.0 * 1.15 <= budget:
            batch = cand
            break
    value, sec, _ = cpu_run(opt, batch, args.steps, args.warmup, kind)
    cores = torch.get_num_threads()
    sample = "B=%d videos per step of the B=128 workload (same model, opts, seeds), %d steps + %d warm-up, %.2f s/step" % (
        batch, args.steps, args.warmup, sec)
    what = "unmodified reference (models.get_model + Seq2Seq.encode + Translator.translate_batch, fp32, %d threads)" % cores \
        if kind == "reference" else "oracle port (CPU restatement; the reference is not installed on this box)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(CONFIG),
            "details": {"sample": sample, "code": what},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: point fd 1 at stderr for everything else (NCCL prints its
    version banner with printf) and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


# ------------------------------------------------------------------------------------------------------
# roofline table of the launch classes
# ------------------------------------------------------------------------------------------------------
def class_table(prof, samples, opt, precision, B, step_ms, pk):
    """prof: Engine.prof of the eager profiling steps; samples: per profiled step {"R", "sq", "N", "S"}."""
    D, I, L_, E = opt["dim_hidden"], opt["intermediate_size"], opt["num_hidden_layers_decoder"], 2 * opt["n_frames"]
    V, F_, din = opt["vocab_size"], opt["n_frames"], opt["dim_i"]
    n = len(samples)
    R = sum(s["R"] for s in samples) / n          # decoder rows per launch (packed: sum of candidate lengths)
    SQ = sum(s["sq"] for s in samples) / n        # sum over candidates of len^2
    b = {"bf16x3": 4, "bf16": 2, "fp32": 4, "tf32": 4}[precision]   # operand bytes per element (hi + lo pairs in bf16x3)
    mult = {"bf16x3": 3.0, "bf16": 1.0, "tf32": 2.0, "fp32": 0.0}[precision]   # bf16-MMA time units issued per useful product
    peak_tf, peak_gb = pk["bf16_tflops_sustained"], pk["hbm_gbs"]
    vocab_rows = prof.get("vocab_rows") or []
    rv = (sum(vocab_rows) / len(vocab_rows)) if vocab_rows else R
    def layer_tail(r, sfx, note):   # everything behind the self-attention core, on r rows
        return {
            "so" + sfx: (2.0 * r * D * D, 3 * r * D * b + D * D * b, "tensor", "M=%d N=%d K=%d +residual%s" % (r, D, D, note)),
            "cq" + sfx: (2.0 * r * D * D, 2 * r * D * b + D * D * b, "tensor", "M=%d N=%d K=%d%s" % (r, D, D, note)),
            "co" + sfx: (2.0 * r * D * D, 3 * r * D * b + D * D * b, "tensor", "M=%d N=%d K=%d +residual%s" % (r, D, D, note)),
            "f1" + sfx: (2.0 * r * D * I, r * D * b + D * I * b + r * I * b, "tensor", "M=%d N=%d K=%d +gelu%s" % (r, I, D, note)),
            "f2" + sfx: (2.0 * r * D * I, r * I * b + D * I * b + 2 * r * D * b, "tensor", "M=%d N=%d K=%d +residual%s" % (r, D, I, note)),
            "cross" + sfx: (4.0 * r * E * D, 2 * r * D * b + B * E * 2 * D * b, "hbm", "rows=%d x E=%d keys, 8 heads%s" % (r, E, note)),
        }
    # last layer of the passes that only merge re-masked positions: the compacted rows (the vocabulary GEMM's rows)
    per = max(len(vocab_rows) // max(n, 1), 1)
    kp = len(prof.get("so_p") or []) // max(n, 1)
    pr = [v for j, v in enumerate(vocab_rows) if (j % per) >= per - kp]
    rp = (sum(pr) / len(pr)) if pr else R
    spec = {   # class: (flops per launch, algorithmic bytes per launch, bound, shape string)
        "qkv": (2.0 * R * D * 3 * D, R * D * b + 3 * D * D * b + R * 3 * D * b, "tensor", "M=%d N=%d K=%d" % (R, 3 * D, D)),
        "self": (4.0 * SQ * D, R * 3 * D * b + R * D * b, "hbm", "sum(len^2)=%d, 8 heads, dk 64" % SQ),
        **layer_tail(R, "", ""),
        **layer_tail(rp, "_p", " (last layer, re-masked rows only: mean M)"),
        "gather": (0.0, 2 * 2 * rp * D * b, "hbm", "context + residual rows of the re-masked positions, %d rows (mean)" % rp),
        "vocab": (2.0 * rv * D * V, rv * D * b + V * D * b + rv * ((V + 127) // 128) * 12, "tensor", "M=%d (mean) N=%d K=%d, softmax statistics epilogue" % (rv, V, D)),
        "kv": (2.0 * B * E * D * L_ * 2 * D, B * E * D * b + L_ * 2 * D * D * b + B * E * L_ * 2 * D * b, "tensor", "M=%d N=%d K=%d (once per batch)" % (B * E, L_ * 2 * D, D)),
        "enc0": (2.0 * B * F_ * din * D, B * F_ * din * b + din * D * b + B * F_ * D * (4 + b), "tensor", "M=%d N=%d K=%d" % (B * F_, D, din)),
        "enc12": (2.0 * B * F_ * D * 2 * D, B * F_ * D * b + 2 * D * D * b + B * F_ * 2 * D * 4, "tensor", "M=%d N=%d K=%d" % (B * F_, 2 * D, D)),
    }
    ncu = None
    path = os.path.join(ROOT, "profiles", "r2_ncu_classes_%s.json" % precision)
    if os.path.isfile(path) and B == 128:
        ncu = json.load(open(path))
    total_us = 0.0
    rows = {}
    for tag, (flops, nbytes, bound, shape) in spec.items():
        ev = prof.get(tag)
        if not ev:
            continue
        durs = list(ev)   # microseconds per launch
        us = sum(durs) / len(durs)
        per_step = sum(durs) / n
        total_us += per_step
        tf = flops / us / 1e6
        gb = nbytes / us / 1e3
        row = {"us": round(us, 2), "launches_per_step": len(durs) // n, "us_per_step": round(per_step, 1), "bound": bound,
               "shape": shape, "flops": flops, "bytes": nbytes, "tflops": round(tf, 1), "gbs": round(gb, 1),
               "frac_useful": round(tf / peak_tf, 4), "frac_issued": round(tf * mult / peak_tf, 4) if mult else None,
               "frac_hbm": round(gb / peak_gb, 4)}
        if ncu and tag in ncu.get("classes", {}):
            row["traffic"] = ncu["classes"][tag].get("dram_bytes")
            row["tensor_pipe_pct"] = ncu["classes"][tag].get("tensor_pct")
        rows[tag] = row
    for tag, row in rows.items():
        row["share_of_profiled"] = round(row["us_per_step"] / total_us, 4)
        row["share_of_step"] = round(row["us_per_step"] / (step_ms * 1e3), 4)
    top = max(rows, key=lambda t: rows[t]["us_per_step"]) if rows else None
    return rows, top, total_us, (ncu or {}).get("source")


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="navc")
    ap.add_argument("--precision", default=os.environ.get("NAVC_PRECISION", "bf16x3"))
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the extra.train legs (configs 3 and 5)")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    opt = cases.config2()

    if args.impl == "reference":
        run_reference(args, opt, rank)
        return

    import torch.distributed as dist
    import navc_b200
    from navc_b200 import _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep NCCL's version / debug lines off stdout (one JSON line)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    torch.manual_seed(0)
    model = navc_b200.get_model(opt).to(dev).eval()
    model.set_precision(args.precision)
    tr = navc_b200.Translator(model, opt, device=dev)
    B = args.batch
    # distinct input batches rotated between iterations: 4 x 126 MB of features > 126 MB L2
    n_rot = 4
    host, devin = [], []
    for r in range(n_rot):
        feats, category = cases.synth_inputs(opt, B, seed=1234 + 17 * r + 1000 * rank)
        host.append(([f.pin_memory() for f in feats], category.pin_memory()))
        devin.append(([f.to(dev) for f in feats], category.to(dev)))
    h2d = sum(f.numel() * 4 for f in host[0][0]) + host[0][1].numel() * 8

    def step_resident(i):
        feats, category = devin[i % n_rot]
        enc = model.encode(feats=feats)
        hyp, _ = tr.translate_batch(enc, category, None, {})
        return hyp

    # e2e: host -> device copies of step i+1 run on a copy stream into a second device buffer while
    # step i computes (every step's H2D copy and D2H read stay inside the timed region)
    copy_stream = torch.cuda.Stream()
    slots = [([torch.empty_like(f) for f in devin[0][0]], torch.empty_like(devin[0][1])) for _ in range(2)]
    pending = {}

    def issue_copy(i):
        feats_h, cat_h = host[i % n_rot]
        feats_d, cat_d = slots[i % 2]
        if i in pending:
            return
        with torch.cuda.stream(copy_stream):
            for d, h in zip(feats_d, feats_h):
                d.copy_(h, non_blocking=True)
            cat_d.copy_(cat_h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending[i] = ev

    # results: every step's ids are copied to pinned host memory on the compute stream (inside the timed region); the host
    # waits for step i-1's copy only after it has launched step i, so the ~0.3 ms of host launch work per step (eager encoder
    # launches + graph inputs) overlaps the previous step instead of idling the GPU (a synchronous .cpu() per step: -3 %)
    out_host = [torch.empty((B * 32,), dtype=torch.int64).pin_memory() for _ in range(2)]
    out_done = {}

    def step_e2e(i):
        if i not in pending:
            issue_copy(i)
        torch.cuda.current_stream().wait_event(pending.pop(i))
        feats, category = slots[i % 2]
        enc = model.encode(feats=feats)
        hyp, _ = tr.translate_batch(enc, category, None, {})
        dst = out_host[i % 2][:hyp.numel()].view(hyp.shape)
        dst.copy_(hyp, non_blocking=True)          # device -> host read of this step's result
        ev = torch.cuda.Event()
        ev.record()
        prev = out_done.pop(i - 1, None)
        if prev is not None:
            prev.synchronize()                     # step i-1 (compute + its D2H) complete: its input slot is free again
        out_done[i] = ev
        issue_copy(i + 1)        # slot (i+1)%2 was last read by step i-1, which has completed
        return dst               # valid after the closing synchronize of the timed region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_ms = {}

    def timed(fn, steps, name=None):
        for i in range(3):  # untimed settle steps of this very loop (PCIe / copy engine / clocks after the idle gap)
            fn(-3 + i)
        pending.clear()
        gc.collect()
        gc.freeze()  # long-lived objects (modules, captured graphs) out of the young generations: no multi-ms collections mid-run
        barrier()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        marks[0].record()
        for i in range(steps):
            out = fn(i)
            marks[i + 1].record()
        barrier()
        ms = marks[0].elapsed_time(marks[steps])
        step_ms[name or fn.__name__] = [round(marks[i].elapsed_time(marks[i + 1]), 2) for i in range(steps)]
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, out

    with torch.no_grad():
        # every rotated batch once eagerly and once more (first call per (B, Smax) shape runs eagerly,
        # the second captures the CUDA graph of the refinement loop), then the counted warm-up
        for i in range(2 * n_rot + args.warmup):
            step_resident(i)
            step_e2e(i)
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        # ---- timed region 1: resident inputs (kernel-side throughput) ----
        launches0 = L.launches
        ms, hyp = timed(step_resident, args.steps, "resident")
        launches = L.launches - launches0
        stats = dict(navc_b200.generate.last_stats)
        # ---- timed region 2: end to end from host buffers ----
        pending.clear()
        torch.cuda.synchronize()
        ms_e2e, hyp_host = timed(step_e2e, args.steps, "e2e")
        pending.clear()
        out_done.clear()
        sampler.stop()
        # ---- region 3 (not part of `value`): per-class device times from the REPLAYED graph.  The decode graphs are
        # dropped and re-captured with CUDA events around every launch class recorded INTO the graph (external event
        # nodes); each replay re-records them, so elapsed_time() is device time with no host launch latency inside ----
        prof_mode = "graph replay with event-record nodes around every launch class"
        samples, acc = [], {}
        model.engine.graphs.clear()
        model.engine.prof = {}
        try:
            step_resident(0)                      # eager ("warm") call of this shape; its events are discarded
            model.engine.prof = {}
            step_resident(0)                      # capture: the events below are nodes of the graph
            prof = model.engine.prof
            if not navc_b200.generate.last_stats.get("graph"):
                raise RuntimeError("no graph")
            for i in range(6):
                step_resident(0)                  # same batch shape -> replay of the captured graph
                torch.cuda.synchronize()
                st = navc_b200.generate.last_stats
                packed = bool(st.get("packed"))
                samples.append({"R": st["rows_real"] if packed else st["N"] * st["S"],
                                "sq": st["rows_sq"] if packed else st["N"] * st["S"] * st["S"], "N": st["N"], "S": st["S"]})
                for tag, ev in prof.items():
                    if tag == "vocab_rows":
                        acc.setdefault(tag, []).extend(int(x) for x in torch.stack(ev).flatten().tolist())
                    elif not tag.startswith("enc"):
                        acc.setdefault(tag, []).extend(a.elapsed_time(b_) * 1e3 for a, b_ in ev)
            # the encoder runs outside the decode graph: its events are appended eagerly, once per call (7 calls here)
            for tag, ev in prof.items():
                if tag.startswith("enc"):
                    acc[tag] = [a.elapsed_time(b_) * 1e3 for a, b_ in ev][len(ev) // 7:]
        except Exception as exc:                  # e.g. a paradigm that does not replay graphs: eager timing instead
            prof_mode = "eager re-run (graph replay unavailable: %s)" % type(exc).__name__
            tr.opt = dict(opt, navc_graphs=False)
            model.engine.prof = {}
            samples, acc = [], {}
            for i in range(min(args.steps, 4)):
                step_resident(i)
                st = navc_b200.generate.last_stats
                packed = bool(st.get("packed"))
                samples.append({"R": st["rows_real"] if packed else st["N"] * st["S"],
                                "sq": st["rows_sq"] if packed else st["N"] * st["S"] * st["S"], "N": st["N"], "S": st["S"]})
            torch.cuda.synchronize()
            for tag, ev in model.engine.prof.items():
                if tag == "vocab_rows":
                    acc[tag] = [int(x) for x in torch.stack(ev).flatten().tolist()]
                else:
                    acc[tag] = [a.elapsed_time(b_) * 1e3 for a, b_ in ev]
            tr.opt = opt
        prof = acc
        model.engine.prof = None
        model.engine.graphs.clear()
    d2h = hyp_host.numel() * 8

    value = world * B * args.steps / (ms / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    pk, pk_src = peaks()
    roof = None
    if prof:
        rows, top, total_us, ncu_src = class_table(prof, samples, opt, args.precision, B, ms / args.steps, pk)
        t = rows[top]
        tensor = t["bound"] == "tensor"
        roof = {"bound": t["bound"], "achieved": t["tflops"] if tensor else t["gbs"],
                "peak": pk["bf16_tflops_sustained"] if tensor else pk["hbm_gbs"], "unit": "TFLOP/s" if tensor else "GB/s",
                "frac": t["frac_useful"] if tensor else t["frac_hbm"], "traffic": t.get("traffic"),
                "traffic_source": ncu_src, "kernel": "%s (%s)" % (top, t["shape"]), "mean_us": t["us"],
                "issued_mma_frac": t["frac_issued"], "share_of_step": t["share_of_step"],
                "peak_source": pk_src + (" (sustained cuBLAS bf16, MEASURED_PEAKS.json)" if tensor else " (copy bandwidth, MEASURED_PEAKS.json)"),
                "timed_in": "%s, %d steps" % (prof_mode, len(samples)),
                "profiled_us_per_step": round(total_us, 1),
                "note": "frac = useful (fp32-equivalent) FLOPs / measured sustained bf16 peak; %s issues %.0fx as many bf16-MMA time "
                        "units per product (frac_issued)" % (args.precision, {"bf16x3": 3, "bf16": 1, "tf32": 2}.get(args.precision, 0)),
                "classes": rows}

    # ---- CPU baseline + parity on the same PARITY_B videos (N = 1 only) ----
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind = reference_kind()
        v16, sec16, hyp_cpu = cpu_run(opt, PARITY_B, 2, 1, kind)        # the parity sample (also warms the code paths)
        # the baseline itself: ONE full step of the workload (B = 128, as timed on the GPU) when that stays near 20 s
        if sec16 * (B / float(PARITY_B)) <= 30.0:
            v, sec, _ = cpu_run(opt, B, 1, 0, kind)
            sample = "one full step of the workload (B=%d videos, same model / opts / seeds): 1 timed repetition, %.1f s, after " \
                     "warm-up on a B=%d sample (%.1f captions/s there)" % (B, sec, PARITY_B, v16)
        else:
            v, sample = v16, "B=%d videos of the same workload (same model, opts, seeds), 2 timed repetitions + 1 warm-up " \
                             "(%.1f s each)" % (PARITY_B, sec16)
        cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample}
        if not args.no_parity:
            parity = parity_block(opt, model, dev, args.precision, hyp_cpu, kind)

    # ---- data-parallel training (BASELINE configs 3 and 5): the single NCCL all-reduce inside the timed step ----
    train = None
    if not args.no_train:
        del tr
        model.engine.graphs.clear()
        gc.unfreeze()
        gc.collect()
        torch.cuda.empty_cache()
        train = train_block(args, world, rank, dev)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"bf16x3": "bf16x3 (split-bf16 tensor-core products, fp32 accumulate; fp32-equivalent)",
                          "bf16": "bf16", "fp32": "f32", "tf32": "tf32"}[args.precision],
                "data": "synthetic", "config": dict(CONFIG),
                "details": {"batch_per_gpu": B, "precision": args.precision, "passes": stats.get("passes"),
                            "S": stats.get("S"), "rows": stats.get("N"), "cuda_graph": bool(stats.get("graph")),
                            "packed_rows": bool(stats.get("packed")), "positions_real": stats.get("rows_real"),
                            "positions_padded": (stats.get("N") or 0) * (stats.get("S") or 0),
                            "l2": "inputs rotate over %d distinct batches (%d MB of features > 126 MB L2)" % (n_rot, n_rot * h2d >> 20)},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu, "parity": parity,
                "extra": {"train": train},
                "per_step_ms": {"resident": step_ms.get("resident"), "e2e": step_ms.get("e2e")}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def parity_block(opt, model, dev, precision, hyp_cpu, kind):
    """GPU ids (fp32 mode and the timed mode; third call of the shape = replayed CUDA graph) against the CPU
    reference's ids on the same PARITY_B videos, video by video.  A video counts as a real mismatch only when
    every decision that reaches its ids clears the mode's arithmetic error (the oracle's `video_margin`, log units)."""
    import navc_b200
    from oracle import navc_oracle as O
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    feats, category = cases.synth_inputs(opt, PARITY_B)
    hyp_o, det = O.translate(sd, opt, feats, category, return_details=True)
    vm = det["video_margin"]
    out = {"against": "unmodified reference on the host" if kind == "reference" else "oracle port on the host",
           "sample": "the %d videos of the cpu_baseline leg (config 2 model, opts, seeds)" % PARITY_B,
           "reference_equals_oracle": bool(hyp_cpu.shape == hyp_o.shape and torch.equal(hyp_cpu, hyp_o)),
           "video_margin_log_units": {"min": float(vm.min()), "median": float(vm.median()), "max": float(vm.max())},
           "modes": {}}
    feats_d, cat_d = [f.to(dev) for f in feats], category.to(dev)
    ref = hyp_cpu if hyp_cpu.shape == hyp_o.shape else hyp_o
    for mode in dict.fromkeys(("fp32", precision)):
        model.set_precision(mode)
        tr = navc_b200.Translator(model, opt, device=dev)
        with torch.no_grad():
            for _ in range(3):  # eager, capture, replay
                hyp, _ = tr.translate_batch(model.encode(feats=feats_d), cat_d, None, {})
        st = navc_b200.generate.last_stats
        hyp = hyp.cpu()
        if hyp.shape != ref.shape:
            out["modes"][mode] = {"error": "shape %s vs %s" % (tuple(hyp.shape), tuple(ref.shape))}
            continue
        eq = (hyp == ref).all(1)
        above = vm > MARGIN[mode]
        out["modes"][mode] = {
            "videos": PARITY_B, "videos_equal": int(eq.sum()), "tokens_equal": float((hyp == ref).float().mean()),
            "margin_threshold": MARGIN[mode], "videos_above_margin": int(above.sum()),
            "videos_above_margin_equal": int((eq & above).sum()), "mismatch_above_margin": int((~eq & above).sum()),
            "differing_videos": [{"video": int(b), "margin": float(vm[b])} for b in (~eq).nonzero().flatten().tolist()],
            "passes_equal": st["passes"] == det["passes"],
            "path": "%s, %s" % ("CUDA graph replay" if st.get("graph") else "eager", "packed rows" if st.get("packed") else "padded rows")}
    model.set_precision(precision)
    return out


def train_block(args, world, rank, dev):
    """extra.train: config 3 (ARB / NAB, B=256 per GPU; N=1 only) and config 5 (NACF, global batch 1024 split over the
    ranks: strong scaling) through tools/train_bench.measure -- forward + loss + backward + ONE gradient all-reduce +
    fused clip/Adam per step; plus the unmodified reference's train step on the host cores at B=32 (N=1)."""
    import train_bench as tb
    out = {"config5_nacf_global1024": None, "config3": None, "cpu_reference": None}
    steps = max(4, min(args.steps, 10))
    r = tb.measure("NACF", 1024 // world, steps, 6, args.precision, dev=dev)   # 6 warm-up steps: every rotated batch twice (allocator steady state)
    r.pop("_step"), r.pop("_model")
    r["scaling"] = "strong (global batch 1024 fixed, %d per GPU)" % (1024 // world)
    out["config5_nacf_global1024"] = r
    if world == 1:
        gc.collect()
        torch.cuda.empty_cache()
        c3 = {}
        for m in ("ARB", "NAB"):
            x = tb.measure(m, 256, steps, 3, args.precision, dev=dev)
            x.pop("_step"), x.pop("_model")
            c3[m] = x
            gc.collect()
            torch.cuda.empty_cache()
        out["config3"] = c3
        if rank == 0 and not args.no_cpu_baseline:
            out["cpu_reference"] = {m: tb.reference_train_cpu(m, 32, 1, 1) for m in ("NACF", "NAB", "ARB")}
    return out if rank == 0 else None


if __name__ == "__main__":
    main()
