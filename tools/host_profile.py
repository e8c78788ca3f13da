#!/usr/bin/env python
"""Host-side cost of one training step (cProfile over a few steps): where the Python time between
kernel launches goes.  python tools/host_profile.py --method NAB [--steps 10]"""
import argparse
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import cases  # noqa: E402
import train_bench as TB  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--method", default="NAB")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import navc_b200
    from navc_b200 import parallel, optim as nopt
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    opt = cases.make_opt(args.method, dim_hidden=512, num_hidden_layers_decoder=6, intermediate_size=2048, dim_i=2048,
                         dim_m=2048, n_frames=60, max_len=30, vocab_size=10547)
    torch.manual_seed(0)
    model = navc_b200.get_model(opt).to(dev)
    model.set_precision("bf16x3")
    model.train()
    dp = parallel.GradientAllReduce(model)
    opt.update(optim="adam", learning_rate=5e-4, minimum_learning_rate=5e-5, decay=0.9, weight_decay=5e-4, grad_clip=5)
    optim = nopt.get_optimizer(opt, model, grads=dp)
    batches = [TB.make_batch(opt, args.batch, 1234 + 31 * r, dev) for r in range(3)]

    def step(i):
        b = batches[i % 3]
        dp.zero_grad()
        res = model(feats=b["feats"], tgt_tokens=b["tgt"], category=b["category"])
        loss = TB.reference_loss(opt, res, b["labels"], b["lt"])
        loss.backward()
        dp.allreduce()
        optim.step()

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    print("allocator:", torch.cuda.get_allocator_backend(), os.environ.get("PYTORCH_CUDA_ALLOC_CONF"))
    for i in range(12):  # per-step host time and allocator activity (cudaMalloc calls, reserved bytes)
        ms0 = torch.cuda.memory_stats()
        t0 = time.perf_counter()
        step(i)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ms1 = torch.cuda.memory_stats()
        print("step %2d: host %.2f ms, drained %.2f ms, cudaMalloc +%d, cudaFree +%d, reserved %.2f GB, rows %s" % (
            i, (t1 - t0) * 1e3, (t2 - t0) * 1e3, ms1["num_device_alloc"] - ms0["num_device_alloc"],
            ms1["num_device_free"] - ms0["num_device_free"], ms1["reserved_bytes.all.current"] / 2**30,
            model.engine.last_train_rows))
    # pure host time: launch without waiting (the GPU queue is deep enough for a few steps)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("host launch time %.2f ms/step, with drain %.2f ms/step" % ((t1 - t0) * 1e3 / args.steps, (t2 - t0) * 1e3 / args.steps))
    pr = cProfile.Profile()
    pr.enable()
    for i in range(args.steps):
        step(i)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr).strip_dirs()
    st.sort_stats("tottime").print_stats(22)
    st.sort_stats("cumulative").print_stats(40)


if __name__ == "__main__":
    main()
