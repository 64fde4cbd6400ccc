"""GPU: the benchmarked tensor-core modes against the fp32 parity mode at FULL size (BASELINE configs[2]: base model,
64 x 500 frames) -- per-frame alignment agreement, token agreement on the rows the pick loop reads, exact-hypothesis
rate.  Reference contract: src/models/cassnat.py:367-389 (argmax -> alignment) and :611 (token pick).  The same report
is the `parity` key of the bench line."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def report():
    import bench
    from cassnat_asr_public_b200.cassnat import make_model
    w = bench.WORKLOADS["cfg3"]
    args = bench.model_args(w)
    torch.manual_seed(1234)
    model = make_model(80, args).to(DEV).eval()
    enc_h, mask, ratio, _ = bench.synth_encoder_output(w, w["batch"], 1234)
    inp = [t.to(DEV) for t in (enc_h, mask, ratio)]
    from cassnat_asr_public_b200 import ops
    try:
        rep = bench.parity_report(model, args, w, inp, "bf16")
    finally:
        ops.set_precision("fp32")                  # the precision switch is process-wide
    print(rep)
    return rep


def test_fp32_mode_is_the_oracle_at_full_size(report):
    """The fp32 parity mode against the CPU oracle port on the first utterances: same alignment, same lengths, and no
    token differs where the oracle's own top-2 margin is above fp32 round-off (1e-4)."""
    r = report["fp32_mode_vs_oracle_port"]
    assert r["alignment_identical"] and r["same_lengths"]
    assert r["differing_with_oracle_margin_gt_1e-4"] == 0
    assert r["tokens_differing"] <= 0.005 * r["tokens"]


def test_decision_safe_heads_keep_the_alignment(report):
    """With the CTC head as bf16x3 the best-path alignment equals the fp32 one on every frame whose fp32 top-2 margin
    exceeds 1e-4 (frames inside fp32 round-off of a tie are undecidable in any arithmetic)."""
    r = report["bf16 (decision-safe heads: K1/K9 as bf16x3)"]
    a = r["alignment_frames_where_fp32_margin_gt_1e-4"]
    assert a["equal"] == a["of"]
    assert r["alignment_frames"]["rate"] >= 0.9999
    plain = report["bf16 with plain bf16 heads"]
    assert plain["alignment_frames"]["equal"] <= r["alignment_frames"]["equal"]


def test_token_agreement(report):
    """Stated tolerance of the fast mode: tokens equal the fp32 pick wherever the fp32 top-2 margin exceeds 0.05 nat
    (bf16 activations move the decoder output by ~4e-3 relative); the bf16x3 mode tightens that."""
    r = report["bf16 (decision-safe heads: K1/K9 as bf16x3)"]
    t = r["tokens_where_fp32_margin_gt_0.05"]
    assert t["equal"] == t["of"] and t["of"] > 1000
    assert r["tokens"]["rate"] >= 0.98
    assert r["rel_err_dec_h"] < 1e-2
    x3 = report["bf16x3 (all GEMMs and the attention three-pass on tcgen05)"]
    assert x3["tokens"]["rate"] >= 0.999 and x3["tokens"]["rate"] >= r["tokens"]["rate"]      # the tensor-core parity mode
    t3 = x3["tokens_where_fp32_margin_gt_0.001"]
    assert t3["equal"] == t3["of"]
    assert x3["rel_err_dec_h"] < 1e-4
