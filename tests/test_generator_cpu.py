"""Host logic of chore_b200.Generator.gen_pc_batch (recon/generator.py:123-217) against the REFERENCE's own run of the
same loop (tests/golden/gen_pc_batch.npz, made by oracle/make_golden.py with the reference Generator on the CPU).
Both are driven by the closed-form field of oracle/analytic_field.py, so every stage is deterministic: the surface
filter, the resampling indices drawn from the CPU generator, the min-count truncation and the argmax / mean reductions
must agree BIT FOR BIT (the "query-point indices" parity set)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ref_shim
from oracle.analytic_field import AnalyticField


def run_ours(df_type, seed, init, num_points, num_steps):
    import chore_b200
    gen = chore_b200.Generator(AnalyticField(), threshold=2.0, filter_val=0.004, device="cpu")
    assert chore_b200.Generator(AnalyticField(), device="cpu").filter_val == 0.03      # the reference's defaults (generator.py:18-26)
    torch.manual_seed(seed)
    drawn = gen.init_samples(3000, batch_size=2)                 # consumes the generator like the reference's call did
    drawn[1] = drawn[0].flip(0)
    assert torch.equal(drawn, init)
    return gen.gen_pc_batch(gen.model, df_type, init, num_points, {"crop_center": torch.tensor([[1008., 995.]] * 2)},
                            num_steps=num_steps, sample_num=20000)


@pytest.mark.parametrize("df_type", ["human", "object"])
def test_gen_pc_batch_bit_exact_vs_reference_run(df_type):
    g = load_golden("gen_pc_batch.npz")
    init = torch.from_numpy(g[f"{df_type}_init"])
    out = run_ours(df_type, int(g[f"seed_{df_type}"]), init, int(g["num_points"]), int(g["num_steps"]))
    for k in ("points", "pca_axis", "parts", "centers"):
        want = torch.from_numpy(g[f"{df_type}_{k}"])
        assert out[k].shape == want.shape and out[k].dtype == want.dtype, k
        assert torch.equal(out[k], want), (k, (out[k].double() - want.double()).abs().max().item())
    # every returned point lies on the surface of its sphere (the filter value is 4 mm)
    f = AnalyticField()
    c, r = (f.c_h, f.r_h) if df_type == "human" else (f.c_o, f.r_o)
    assert ((out["points"] - c).norm(dim=-1) - r).abs().max() < 0.004 + 1e-6


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_gen_pc_batch_equals_live_reference():
    Generator = ref_shim.load_generator_class()
    ref = object.__new__(Generator)
    ref.threshold, ref.filter_val, ref.device = 2.0, 0.004, "cpu"
    torch.manual_seed(5)
    init = ref.init_samples(2000, batch_size=1)
    want = ref.gen_pc_batch(AnalyticField(), "object", init, 21000, {"crop_center": torch.tensor([[1008., 995.]])}, 6, mute=True)
    import chore_b200
    gen = chore_b200.Generator(AnalyticField(), threshold=2.0, filter_val=0.004, device="cpu")
    torch.manual_seed(5)
    init2 = gen.init_samples(2000, batch_size=1)
    assert torch.equal(init, init2)
    got = gen.gen_pc_batch(gen.model, "object", init2, 21000, {"crop_center": torch.tensor([[1008., 995.]])}, num_steps=6)
    for k in want:
        assert torch.equal(got[k], want[k]), k
