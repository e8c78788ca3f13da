import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, cases, navc_b200
from navc_b200 import _lib as L
dev = torch.device("cuda", 0)
for nl in (1, 2):
    opt = cases.small("NACF", dim_hidden=512, num_attention_heads=8, intermediate_size=1024, max_len=30, length_beam_size=5,
                      num_hidden_layers_decoder=max(nl, 1))
    torch.manual_seed(0)
    model = navc_b200.get_model(opt).to(dev).eval(); model.set_precision("bf16x3")
    eng = model.engine
    feats, category = cases.synth_inputs(opt, 9)
    with torch.no_grad():
        enc = model.encode(feats=[f.to(dev) for f in feats])
    mem = eng.memory(enc["enc_output"], enc.get("_navc"))
    N, S, lbs = 45, 27, 5
    g = torch.Generator().manual_seed(3)
    lens = torch.randint(4, S + 1, (N,), generator=g).int(); lens[0] = S
    toks = torch.randint(4, 300, (N, S), generator=g)
    toks[torch.arange(S).unsqueeze(0) >= lens.unsqueeze(1)] = 0
    toks, lens_d, cat = toks.to(dev), lens.to(dev), category.to(dev)
    if nl == 0:
        eng.P["layers"] = []
    packed = eng.pack_rows(lens_d, S)
    so = packed["seq_off"].cpu()
    ref = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.int32), lens]), 0).int()
    print("layers", nl, "seq_off ok", torch.equal(so, ref), "rowmap ok", all(int(packed["rowmap"][so[n] + s]) == n * S + s for n in range(N) for s in range(int(lens[n]))))
    h0, _ = eng.decoder_pass(toks, mem, lbs, cat, "NARFormer", want_f32=True)
    h1, _ = eng.decoder_pass(toks, mem, lbs, cat, "NARFormer", want_f32=True, packed=packed)
    a, b = h0.f32.view(N, S, -1), h1.f32
    worst = 0.0; bad = []
    for n in range(N):
        ln = int(lens[n]); d = (a[n, :ln] - b[so[n]:so[n] + ln]).abs().max().item()
        if d > 1e-3: bad.append((n, ln, round(d, 4)))
        worst = max(worst, d)
    print("  max diff %.3e  bad seqs %d %s" % (worst, len(bad), bad[:12]))
    pm0 = eng.vocab_partials(h0); pm1 = eng.vocab_partials(h1, m_dev=packed["count"])
    d = max((pm0[0].view(N, S, -1)[n, :int(lens[n])] - pm1[0][so[n]:so[n] + int(lens[n])]).abs().max().item() for n in range(N))
    print("  vocab partial max diff %.3e" % d)
