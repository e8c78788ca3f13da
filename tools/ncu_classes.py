#!/usr/bin/env python
"""Per launch class (qkv, self, so, cq, cross, co, f1, f2, vocab, kv) DRAM bytes and tensor-pipe activity from an
`ncu --set full` capture of ONE eager config-2 step (tools/profile_step.py <precision> 128 range), written as the JSON
bench.py cites in roofline.classes[*].traffic:   python tools/ncu_classes.py rep.ncu-rep precision out.json

Classes are assigned from the launch ORDER, which is fixed by Engine.decoder_pass: after every embed_ln kernel, each
layer is  gemm(qkv) attn(self) gemm(so) gemm(cq) attn(cross) gemm(co) gemm(f1) gemm(f2)."""
import csv, io, json, os, re, subprocess, sys

rep, precision, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def num(v, k):
    try:
        x = float(v[ix[k]].replace(",", ""))
    except Exception:
        return 0.0
    u = units[ix[k]]
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)


LAYER = ["qkv", "self", "so", "cq", "cross", "co", "f1", "f2"]
acc = {}
pos = None
seen_kv = after_beam = False
for v in vals:
    name = v[ix["Kernel Name"]]
    is_gemm, is_attn = "gemm_tc_kernel" in name or "gemm2_tc_kernel" in name, "attn_tc_kernel" in name or "attn2_tc_kernel" in name
    tag = None
    if "embed_ln" in name:
        pos = 0
        continue
    if "length_beam" in name:
        after_beam = True
        continue
    if pos is None:
        if is_gemm and after_beam and not seen_kv:   # K|V projection of the encoder memory (Engine.memory)
            tag, seen_kv = "kv", True
    elif is_gemm or is_attn:
        if pos < 0:
            continue
        tag = LAYER[pos % 8]
        assert (tag in ("self", "cross")) == is_attn, (tag, name)
        pos += 1
    elif "refine_step" in name or "gather_rows" in name:
        continue
    if is_gemm and re.search(r"<[^>]*, 1, 256>", name):   # vocabulary epilogue: ends the pass
        tag, pos = "vocab", -1
    if tag is None:
        continue
    a = acc.setdefault(tag, {"n": 0, "us": 0.0, "dram": 0.0, "tensor": 0.0})
    a["n"] += 1
    a["us"] += num(v, "gpu__time_duration.sum")
    a["dram"] += num(v, "dram__bytes_read.sum") + num(v, "dram__bytes_write.sum")
    a["tensor"] += num(v, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
res = {"source": "profiles/" + os.path.basename(out).replace(".json", "") + " (ncu --set full --clock-control none, one eager config-2 step, B=128, %s)" % precision,
       "precision": precision, "classes": {t: {"launches": a["n"], "us_under_ncu": round(a["us"] / a["n"], 2), "dram_bytes": round(a["dram"] / a["n"]),
                                                "tensor_pct": round(a["tensor"] / a["n"], 1)} for t, a in acc.items()}}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
