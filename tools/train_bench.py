#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs 3 and 5) -- one JSON line on stdout.

    python tools/train_bench.py --method NACF|NAB|ARB [--batch B] [--global-batch G] [--steps K] [--warmup W]
    torchrun --nproc-per-node N tools/train_bench.py --method NACF --global-batch 1024     (config 5)

A step = forward (train mode, dropout 0.5, BatchNorm batch statistics) + the reference's loss
(masked NLL sum / batch + KLDiv on the length head: misc/crit.py, caller code in PyTorch) + backward
(navc kernels) + ONE gradient all-reduce (N > 1) + clip_grad_value_(5) + Adam(lr 5e-4, wd 5e-4)
(misc/run.py:254-261, misc/optim.py:61-62; caller code).  Metric: samples/s, whole job.
Inputs are resident in HBM and rotate over distinct batches."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import cases  # noqa: E402


def reference_loss(opt, results, labels, length_target):
    """misc/crit.py:62-84, 156-181, 223 (the caller's loss, plain PyTorch)."""
    lps = results["tgt_word_logprobs"]
    labs = labels if isinstance(labels, (list, tuple)) else [labels] * len(lps)
    weights = opt.get("nv_weights", [0.8, 1.0]) if opt.get("visual_word_generation", False) else [1.0] * len(lps)
    bsz = lps[0].shape[0]
    loss = 0.0
    for w, lp, lab in zip(weights, lps, labs):
        nll = F.nll_loss(lp.reshape(-1, lp.shape[-1]), lab.reshape(-1), reduction="none")
        loss = loss + w * (nll * lab.reshape(-1).ne(0).float()).sum() / bsz
    if length_target is not None and "pred_length" in results:
        loss = loss + F.kl_div(results["pred_length"], length_target, reduction="mean")
    return loss


def make_batch(opt, B, seed, dev):
    nar = opt["decoding_type"] == "NARFormer"
    feats, category = cases.synth_inputs(opt, B, seed=seed)
    toks = cases.synth_tokens(opt, B, seed=seed + 1, kind="nar" if nar else "ar")
    dis = opt["decoder"] == "BertDecoderDisentangled"
    if nar:
        tgt = [toks["tokens_1"], toks["tokens"]] if dis else toks["tokens"]
        labels = [toks["labels_1"], toks["labels"]] if dis else toks["labels"]
        lt = toks["length_target"]
    else:
        tgt, labels, lt = toks["tokens"], toks["labels"], None
    mv = lambda t: [x.to(dev) for x in t] if isinstance(t, (list, tuple)) else (None if t is None else t.to(dev))
    return dict(feats=mv(feats), category=category.to(dev), tgt=mv(tgt), labels=mv(labels), lt=mv(lt))


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


def train_opt(method, fused_ce=False):
    """BASELINE configs 3 / 5: the config-2 model shape in training."""
    kw = dict(dim_hidden=512, num_hidden_layers_decoder=6, intermediate_size=2048, dim_i=2048, dim_m=2048,
              n_frames=60, max_len=30, vocab_size=10547)
    opt = cases.make_opt(method, **kw)
    opt["navc_fused_ce"] = bool(fused_ce)
    nar = opt["decoding_type"] == "NARFormer"
    opt.update(crit_key=[("tgt_word_logprobs", "tgt_word_labels")] + ([("pred_length", "tgt_length")] if nar else []),
               crit_name=["Cap Loss"] + (["Length Loss"] if nar else []), crit_scale=[1.0] + ([1.0] if nar else []),
               optim="adam", learning_rate=5e-4, minimum_learning_rate=5e-5, decay=0.9, weight_decay=5e-4, grad_clip=5)
    return opt


def measure(method, B, steps, warmup, precision="bf16x3", fused_ce=False, torch_optim=False, dev=None, seed_rank=0):
    """Time `steps` training steps of `method` at B samples on THIS rank (torch.distributed, if initialised, supplies
    the ONE gradient all-reduce per step and the max-over-ranks time).  Returns a dict (every rank)."""
    import torch.distributed as dist
    import navc_b200
    from navc_b200 import _lib as L, parallel
    from navc_b200.misc import crit as ncrit
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    opt = train_opt(method, fused_ce)
    crit = ncrit.get_criterion(opt)
    torch.manual_seed(0)
    model = navc_b200.get_model(opt).to(dev)
    model.set_precision(precision)
    model.train()
    dp = parallel.GradientAllReduce(model)
    if torch_optim:
        optim = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=5e-4)
    else:
        from navc_b200 import optim as nopt
        optim = nopt.get_optimizer(opt, model, grads=dp)
    n_rot = 3
    batches = [make_batch(opt, B, 1234 + 31 * r + 1000 * (rank + seed_rank), dev) for r in range(n_rot)]

    def step(i):
        b = batches[i % n_rot]
        dp.zero_grad()
        res = model(feats=b["feats"], tgt_tokens=b["tgt"], category=b["category"])
        if not fused_ce:
            loss = reference_loss(opt, res, b["labels"], b["lt"])
        else:
            res["tgt_word_labels"], res["tgt_length"] = b["labels"], b["lt"]
            loss = crit.get_loss(res)
        loss.backward()
        dp.allreduce()
        if torch_optim:
            torch.nn.utils.clip_grad_value_(model.parameters(), 5)
        optim.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(warmup, 3)):
        loss = step(i)
    barrier()
    l0 = L.launches
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    marks[0].record()
    for i in range(steps):
        loss = step(i)
        marks[i + 1].record()
    barrier()
    ms = marks[0].elapsed_time(marks[-1])
    per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(steps))
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    out = {"method": method, "samples_per_s": world * B * steps / (ms / 1e3), "ms_per_step": ms / steps,
           "batch_per_gpu": B, "global_batch": B * world, "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
           "precision": precision, "gpu_launches": L.launches - l0, "final_loss": float(loss.item()),
           "allreduce_bytes": dp.nbytes if world > 1 else 0, "params": sum(p.numel() for p in model.parameters()),
           "per_step_ms": {"min": round(per_step[0], 3), "median": round(per_step[len(per_step) // 2], 3),
                           "max": round(per_step[-1], 3)},
           "rows": "%d of %d decoder rows (packed)" % model.engine.last_train_rows
                   if getattr(model.engine, "last_train_rows", None) else None,
           "step": "fwd (dropout 0.5, BN batch stats) + loss + bwd + %s + %s" % (
               "one NCCL all-reduce (AVG) of the flat fp32 gradient buffer" if world > 1 else "no collective (1 rank)",
               "torch clip_grad_value_ + Adam" if torch_optim else "fused clip + Adam (navc_clip_adam, one launch)")}
    out["_step"], out["_model"] = step, model
    return out


def reference_train_cpu(method, B, steps, warmup):
    """The reference's OWN training step (misc/run.py:254-261: model(**batch) -> Criterion.get_loss -> backward ->
    clip_grad_value_ -> Adam) on the host cores, unmodified code from /root/reference or baseline/_ref; None if the
    reference is not installed."""
    import contextlib
    import io
    import warnings
    import refutil
    if not refutil.reference_available():
        return None
    opt = train_opt(method)
    torch.set_num_threads(os.cpu_count() or 1)
    b = make_batch(opt, B, 1234, torch.device("cpu"))
    with refutil.reference_on_path(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models import get_model
        from misc.crit import get_criterion
        from misc.optim import get_optimizer
        from torch.nn.utils import clip_grad_value_
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model = get_model(dict(opt))
        model.train()
        crit = get_criterion(opt)
        crit.reset_loss_recorder()
        optimizer = get_optimizer(opt, model)

        def step():
            optimizer.zero_grad()
            res = model(feats=b["feats"], tgt_tokens=b["tgt"], category=b["category"], opt=opt, vocab=None)
            res["tgt_word_labels"] = b["labels"]
            if b["lt"] is not None:
                res["tgt_length"] = b["lt"]
            loss = crit.get_loss(res)
            loss.backward()
            clip_grad_value_(model.parameters(), opt["grad_clip"])
            optimizer.step()
            return loss

        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = step()
        dt = time.perf_counter() - t0
    return {"method": method, "samples_per_s": B * steps / dt, "ms_per_step": dt / steps * 1e3, "batch": B,
            "cores": torch.get_num_threads(), "kind": "reference", "final_loss": float(loss.item()),
            "step": "unmodified reference: model(**batch), misc.crit.Criterion, backward, clip_grad_value_(5), Adam (misc/run.py:254-261)"}


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--method", default="NACF")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--global-batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default=os.environ.get("NAVC_PRECISION", "bf16x3"))
    ap.add_argument("--fused-ce", action="store_true", help="navc_b200.misc.crit with opt['navc_fused_ce'] (fused projection + "
                    "log-softmax + masked NLL: no [B,S,V] tensors; measured 31.9 vs 30.7 ms at B=256 because the backward recomputes "
                    "the logits and the criterion's meters add host syncs) instead of the reference's loss on materialised log-probs")
    ap.add_argument("--torch-optim", action="store_true", help="caller-side clip_grad_value_ + torch.optim.Adam instead of the fused navc_clip_adam step")
    ap.add_argument("--profile", action="store_true", help="print a per-kernel device-time breakdown of one step")
    ap.add_argument("--cpu-reference", type=int, default=0, help="also time the unmodified reference's train step on the host cores at this batch")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep NCCL's version / debug lines off stdout (one JSON line)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B = args.global_batch // world if args.global_batch else args.batch
    r = measure(args.method, B, args.steps, args.warmup, args.precision, args.fused_ce, args.torch_optim, dev)
    step = r.pop("_step")
    r.pop("_model")
    if rank == 0:
        line = {"metric": "training samples/sec (%s 6-layer d512, fwd+bwd+allreduce+clip+Adam)" % args.method,
                "value": r["samples_per_s"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if args.global_batch else "weak", "dtype": args.precision, "data": "synthetic",
                "config": {"workload": "%s training step, feats 2x60x2048, max_len 30, vocab 10547, dropout 0.5" % args.method,
                           "batch_per_gpu": B, "global_batch": B * world, "params": r["params"], "step": r["step"],
                           "loss": "torch log-probs + NLL" if not args.fused_ce else "fused cross-entropy (navc_b200.misc.crit, incl. its accuracy / perplexity meters)",
                           "allreduce_bytes": r["allreduce_bytes"], "l2": "inputs rotate over 3 distinct batches"},
                "gpu_launches": r["gpu_launches"], "final_loss": r["final_loss"], "per_step_ms": r["per_step_ms"], "rows": r["rows"]}
        if args.cpu_reference:
            line["cpu_reference"] = reference_train_cpu(args.method, args.cpu_reference, 2, 1)
        emit(line)
    if args.profile and rank == 0:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(0)
            torch.cuda.synchronize()
        agg = {}
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                a = agg.setdefault(ev.name[:80], [0, 0.0]); a[0] += 1; a[1] += ev.device_time
        tot = sum(v[1] for v in agg.values())
        print("total device time: %.2f ms over %d kernels" % (tot / 1e3, sum(v[0] for v in agg.values())), file=sys.stderr)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            print("%9.1f us %5d x %8.1f us  %5.1f%%  %s" % (v[1], v[0], v[1] / v[0], 100 * v[1] / tot, k), file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
