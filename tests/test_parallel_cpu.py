"""Host-side multi-process logic on CPU: world_size-2 gloo.
Sharding of videos across ranks, weight broadcast, and the single flat-buffer gradient all-reduce:
the averaged gradient equals the mean of the per-rank gradients, for every parameter, through views
of ONE buffer.  (The kernels themselves need a GPU; gradients here are synthetic.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import navc_b200
from navc_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        opt = cases.config1()
        torch.manual_seed(100 + rank)  # different initial weights per rank on purpose
        model = navc_b200.get_model(opt)
        dp = parallel.GradientAllReduce(model)  # broadcasts rank 0's weights
        w0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        gathered = [torch.empty_like(w0) for _ in range(world)]
        dist.all_gather(gathered, w0)
        same_weights = all(torch.equal(gathered[0], t) for t in gathered)
        # synthetic per-rank gradients, accumulated the way autograd does (in place into p.grad)
        dp.zero_grad()
        g = torch.Generator().manual_seed(7 + rank)
        local = []
        for p in dp.params:
            gr = torch.randn(p.shape, generator=g)
            p.grad.add_(gr)
            local.append(gr)
        is_view = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(dp.params, dp.views))
        # an optimizer that detaches the grads (set_to_none) must not break the flat buffer
        model.zero_grad(set_to_none=True)
        for p, gr in zip(dp.params, local):
            p.grad = gr.clone()
        flat = dp.allreduce()
        expect = []
        for i in range(len(dp.params)):
            parts = []
            for r in range(world):
                gg = torch.Generator().manual_seed(7 + r)
                for j, p in enumerate(dp.params):
                    t = torch.randn(p.shape, generator=gg)
                    if j == i:
                        parts.append(t)
                        break
            expect.append(sum(parts) / world)
        ok = all(torch.allclose(p.grad, e, atol=1e-6) for p, e in zip(dp.params, expect))
        ok_flat = all(torch.allclose(flat[o:o + e.numel()], e.reshape(-1), atol=1e-6) for o, e in zip(dp.offsets, expect)) \
            and all(o % parallel.ALIGN == 0 for o in dp.offsets)
        # video sharding
        feats, category = cases.synth_inputs(opt, 7)
        sh = parallel.shard({"feats": feats, "category": category, "video_ids": ["v%d" % i for i in range(7)]})
        lo, hi = parallel.shard_bounds(7)
        q.put((rank, same_weights, is_view, ok, ok_flat, dp.numel, sh["category"].shape[0], sh["feats"][0].shape[0],
               sh["video_ids"], lo, hi))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_and_sharding_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = 0
    for rank, same_weights, is_view, ok, ok_flat, numel, nb, nf, ids, lo, hi in res:
        assert same_weights and is_view and ok and ok_flat
        assert nb == nf == hi - lo == len(ids)
        total += nb
    assert total == 7 and res[0][8] + res[1][8] == ["v%d" % i for i in range(7)]
    n_params = sum(p.numel() for p in navc_b200.get_model(cases.config1()).parameters())
    assert n_params <= res[0][5] < n_params + 64 * 200  # flat buffer = parameters + alignment padding


def _worker_early(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        opt = cases.config1()
        torch.manual_seed(3)
        model = navc_b200.get_model(opt)
        dp = parallel.GradientAllReduce(model)
        assert model.engine.grads_final_hook == dp.begin_early
        named = dict(model.named_parameters())
        enc = [p for k, p in named.items() if not k.startswith(("decoder.", "tgt_word_prj."))]
        late = set(id(p) for p in enc)
        res = []
        for weight in (1.0, 0.5 + rank):
            dp.zero_grad()
            dp.weight = weight
            g = torch.Generator().manual_seed(11 + rank)
            grads = [torch.randn(p.shape, generator=g) for p in dp.params]
            for p, gr in zip(dp.params, grads):          # "decoder side" first, as the backward pass does
                if id(p) not in late:
                    p.grad.add_(gr)
            dp.begin_early(enc)                          # what EncodeFn.backward calls before producing its own gradients
            in_flight = dp._early is not None
            split = dp._early[0] if in_flight else -1
            for p, gr in zip(dp.params, grads):
                if id(p) in late:
                    p.grad.add_(gr)
            flat = dp.allreduce().clone()
            # the same step without overlap
            dp.zero_grad()
            dp.overlap = False
            for p, gr in zip(dp.params, grads):
                p.grad.add_(gr)
            dp.begin_early(enc)
            assert dp._early is None
            ref = dp.allreduce().clone()
            dp.overlap = True
            res.append((in_flight, split, torch.allclose(flat, ref, atol=1e-6), float(ref.abs().sum())))
        first_late = min(o for p, o in zip(dp.params, dp.offsets) if id(p) in late)
        last_late = max(o for p, o in zip(dp.params, dp.offsets) if id(p) in late)
        # passing a different weight at allreduce() time while a part is in flight must fail loudly
        dp.zero_grad(); dp.weight = 1.0
        dp.begin_early(enc)
        try:
            dp.allreduce(weight=2.0)
            raised = False
        except RuntimeError:
            raised = True
        q.put((rank, res, first_late, last_late, dp.numel, raised))
    finally:
        dist.destroy_process_group()


def test_early_allreduce_of_finished_gradients_world2():
    """parallel.GradientAllReduce.begin_early: the tail of the flat buffer (every parameter behind the encoder's) is
    reduced while the encoder's gradients are still being produced; the result equals the single collective."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_early, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res, first_late, last_late, numel, raised in out:
        assert raised
        for in_flight, split, same, norm in res:
            assert in_flight and same and norm > 0
            assert last_late < split < numel      # most of the buffer (the decoder) goes out early
        assert first_late == 0                    # encoder parameters come first in the module tree


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 128, 1000):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gradient_sink_follows_grad_ownership():
    """GradientAllReduce.sink (what the backward kernels accumulate into in place): the flat-buffer view while
    p.grad IS that view, nothing once an optimizer detached it, the view again after attach() / zero_grad();
    the engine of the model points at it and keeps doing so across set_precision()."""
    opt = cases.config1()
    torch.manual_seed(0)
    model = navc_b200.get_model(opt)
    dp = parallel.GradientAllReduce(model, broadcast=False)
    assert model.engine.grad_sink == dp.sink
    p0, v0 = dp.params[0], dp.views[0]
    assert dp.sink(p0) is v0
    assert dp.sink(torch.nn.Parameter(torch.zeros(3))) is None            # not one of this model's parameters
    model.zero_grad(set_to_none=True)
    assert dp.sink(p0) is None                                            # detached: kernels must not write behind autograd
    dp.attach()
    assert dp.sink(p0) is v0
    p0.grad = torch.ones_like(p0)
    assert dp.sink(p0) is None
    dp.zero_grad()
    assert dp.sink(p0) is v0 and float(v0.abs().sum()) == 0.0
    model.set_precision("fp32")
    assert model.engine.grad_sink == dp.sink
    # offsets are 256-byte aligned (vectorised / TMA operands) and the views tile the buffer without overlap
    assert all(o % parallel.ALIGN == 0 for o in dp.offsets)
    assert all(o2 >= o1 + p.numel() for o1, o2, p in zip(dp.offsets, dp.offsets[1:], dp.params))
