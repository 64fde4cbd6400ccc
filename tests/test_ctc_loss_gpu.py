"""GPU: K10 CTC loss forward-backward against torch.nn.functional.ctc_loss (the arithmetic the
reference calls, cassnat.py:126,232) evaluated on CPU in float64, compared at the logits level.
Tolerance: rel 1e-3 (north_star) — observed ~1e-5."""
import pytest
import torch
import torch.nn.functional as F

from cassnat_asr_public_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def torch_reference(logits, targets, in_len, tgt_len):
    lg = logits.double().clone().requires_grad_(True)
    lp = torch.log_softmax(lg, -1)
    per = F.ctc_loss(lp.transpose(0, 1), targets, in_len, tgt_len, blank=0, reduction="none", zero_infinity=True)
    loss = F.ctc_loss(lp.transpose(0, 1), targets, in_len, tgt_len, blank=0, reduction="mean", zero_infinity=True)
    loss.backward()
    return loss.detach(), per.detach(), lg.grad


@pytest.mark.parametrize("B,T,V,L", [(4, 30, 20, 5), (3, 50, 97, 12), (8, 250, 5000, 40), (2, 120, 50, 60),
                                     (5, 64, 33, 1), (2, 400, 300, 150)])
def test_ctc_loss_vs_torch(B, T, V, L):
    g = torch.Generator().manual_seed(B * 100 + L)
    logits = torch.randn(B, T, V, generator=g) * 2
    targets = torch.randint(1, V, (B, L), generator=g)
    targets[:, 1::3] = targets[:, 0::3][:, :targets[:, 1::3].shape[1]]        # repeated labels
    tgt_len = torch.randint(max(L // 2, 1), L + 1, (B,), generator=g)
    in_len = torch.randint(min(T, 2 * L + 2), T + 1, (B,), generator=g)
    in_len[0] = T
    tgt_len[-1] = L
    loss_ref, per_ref, grad_ref = torch_reference(logits, targets, in_len, tgt_len)
    lp = torch.log_softmax(logits.to(DEV), -1)
    loss, nll, grad = ops.ctc_loss(lp, targets, in_len, tgt_len)
    assert abs(loss.item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
    assert torch.allclose(nll.cpu().double(), per_ref, rtol=1e-4, atol=1e-3)
    err = (grad.cpu().double() - grad_ref).abs().max().item() / grad_ref.abs().max().item()
    assert err < 1e-3, err


def test_ctc_loss_infeasible_and_empty_targets():
    """zero_infinity: a target longer than the input contributes 0 loss and 0 gradient; L_b = 0 works."""
    B, T, V, L = 3, 6, 11, 5
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(B, T, V, generator=g)
    targets = torch.randint(1, V, (B, L), generator=g)
    tgt_len = torch.tensor([5, 0, 2])
    in_len = torch.tensor([4, 6, 6])            # utterance 0: 5 labels in 4 frames -> infeasible
    loss_ref, per_ref, grad_ref = torch_reference(logits, targets, in_len, tgt_len)
    loss, nll, grad = ops.ctc_loss(torch.log_softmax(logits.to(DEV), -1), targets, in_len, tgt_len)
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    assert torch.isinf(nll[0]) and float(grad[0].abs().max()) == 0.0
    assert (grad.cpu().double() - grad_ref).abs().max().item() < 1e-6


def test_ctc_loss_golden(blocks_golden):
    """The reference model's own CTC loss and d loss / d logits (tests/golden/blocks_*.npz)."""
    name, z, sd, cfg = blocks_golden
    logits = torch.from_numpy(z["train::ctc_logits"]).to(DEV)
    loss, nll, grad = ops.ctc_loss(torch.log_softmax(logits, -1), torch.from_numpy(z["train::labels"]),
                                   torch.from_numpy(z["train::frames"]), torch.from_numpy(z["train::label_sizes"]))
    assert abs(loss.item() - float(z["train::ctc_loss"])) < 1e-4 * abs(float(z["train::ctc_loss"]))
    ref = torch.from_numpy(z["train::dctc_dlogits"])
    assert (grad.cpu() - ref).abs().max().item() / ref.abs().max().item() < 1e-3
