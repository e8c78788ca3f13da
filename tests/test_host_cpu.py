"""CPU-only checks of the host side: C-ABI exports, state_dict contract, loud failure without a GPU."""
import os
import re
import subprocess

import pytest
import torch

import cases
import refutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _navc():
    import navc_b200
    return navc_b200


def test_library_exports_every_declared_symbol():
    navc = _navc()
    header = open(os.path.join(ROOT, "include", "navc.h")).read()
    declared = set(re.findall(r"\b(navc_[a-z0-9_]+)\s*\(", header))
    declared -= {"navc_epilogue_t", "navc_step_t"}
    assert len(declared) >= 20
    lib = navc._lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", navc._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (navc_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    assert set(navc._lib.EXPORTS) == declared
    assert lib.navc_version() == 4


def test_ctypes_struct_layout_matches_header():
    import ctypes
    navc = _navc()
    # navc_epilogue_t: 3 pointers, 2 int32, 3 pointers, 4 int32, 3 pointers, 2 int32 ; navc_step_t per include/navc.h
    assert ctypes.sizeof(navc._lib.Epilogue) == 3 * 8 + 8 + 3 * 8 + 8 + 8 + 16 + 8 + 8
    assert ctypes.sizeof(navc._lib.Step) == 3 * 8 + 8 * 4 + 16 * 8


def test_product_path_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    navc = _navc()
    opt = cases.config1()
    torch.manual_seed(0)
    model = navc.get_model(opt)
    model.eval()
    feats, category = cases.synth_inputs(opt, 2)
    with pytest.raises(navc._lib.NavcError):
        model.encode(feats=feats)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "non-autoregressive-video-captioning_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


@pytest.mark.parametrize("method,kw", [("NACF", {}), ("NAB", {"with_layernorm": True}), ("ARB", {}),
                                       ("NAB", {"with_category": False, "no_encoder_bn": True}),
                                       ("NACF", {"tie_weights": True})])
def test_state_dict_contract_matches_golden_inventory(method, kw):
    """Keys/shapes must equal the reference's; checked against the inventory the
    golden fixtures recorded from the reference, and against the live reference when mounted."""
    navc = _navc()
    opt = cases.small(method, **kw)
    torch.manual_seed(0)
    model = navc.get_model(opt)
    mine = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    if refutil.reference_available():
        ref = refutil.ref_get_model(opt, seed=0)
        theirs = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        assert mine == theirs
        # same construction order => same seeded initial weights
        for k, v in ref.state_dict().items():
            assert torch.equal(v, model.state_dict()[k]), k
    if method == "NACF" and not kw:
        g = torch.load(os.path.join(ROOT, "tests", "golden", "fwd_small_nacf.pt"), weights_only=False)
        assert mine == {k: tuple(v) for k, v in g["shapes"].items()}


def test_golden_inventory_loads_into_product_model():
    navc = _navc()
    g = torch.load(os.path.join(ROOT, "tests", "golden", "dec_small_nacf_teacher.pt"), weights_only=False)
    model = navc.get_model(g["opt"])
    model.load_state_dict(cases.synth_state_dict(g["shapes"], g["wseed"]))
    teacher = navc.get_model(g["teacher_opt"])
    teacher.load_state_dict(cases.synth_state_dict(g["teacher_shapes"], g["wseed"] + 1))
    # teacher remap used by misc/run.py:275-279 of the reference: decoder.bert.* <- decoder.*
    tk = set(teacher.state_dict())
    assert all(k.replace("decoder.bert.", "decoder.") in tk for k in model.state_dict() if k.startswith("decoder.bert."))


def test_training_packing_plan_decisions():
    """training.plan_packing (host logic): packs only when PAD is a suffix of every row, no row is empty and enough
    rows are PAD; one plan per token tensor, keyed so that a sliced view (AR input tokens[:, :-1]) finds its own."""
    import torch
    from navc_b200 import training as T

    class FakeEngine:
        opt = {}

        def tc_attention_ok(self, S, E):
            return True

        def pack_rows(self, lens, S):
            N = lens.numel()
            rowmap = torch.tensor([n * S + s for n in range(N) for s in range(int(lens[n]))], dtype=torch.int32)
            return dict(seq_off=None, rowmap=torch.cat([rowmap, torch.zeros(N * S - rowmap.numel(), dtype=torch.int32)]), N=N, S=S)

    eng = FakeEngine()
    t = torch.tensor([[5, 6, 0, 0], [7, 0, 0, 0], [1, 2, 3, 0]])
    called = []
    plans = T.plan_packing(eng, [t, t[:, :-1]], 10, between=lambda: called.append(1))
    assert called == [1]
    pk = plans[T._plan_key(t)]
    assert pk["rows"] == 6 and pk["rowmap"].tolist() == [0, 1, 4, 8, 9, 10] and pk["pad_rows"].tolist() == [2, 3, 5, 6, 7, 11]
    pk2 = plans[T._plan_key(t[:, :-1])]
    assert pk2["rows"] == 6 and pk2["pad_rows"].tolist() == [2, 4, 5]
    # PAD in the middle of a row, an empty row, (almost) no PAD at all -> padded path
    for bad in (torch.tensor([[5, 0, 6, 0], [7, 0, 0, 0]]), torch.tensor([[5, 6, 0, 0], [0, 0, 0, 0]]),
                torch.tensor([[5, 6, 7, 8], [1, 2, 3, 4]])):
        assert list(T.plan_packing(eng, [bad], 10).values()) == [None]
    # switched off
    eng.opt = {"navc_train_packed": 0}
    assert list(T.plan_packing(eng, [t], 10).values()) == [None]
