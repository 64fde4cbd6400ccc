"""Host-side logic of the decode call that needs no GPU: the greedy hypothesis assembly from the packed read-back
rows (reference: /root/reference/src/models/cassnat.py:574-637 at beam width 1), the deferred-result handle, and the
n-gram ranker's detokenisation (cassnat.py:525-536)."""
from types import SimpleNamespace

import numpy as np
import torch

from cassnat_asr_public_b200.cassnat import CassNAT, PendingHypotheses


def _reference_pick(ids, logp, ylen, ymax, sos):
    """The reference's loop for beam_width 1, lm_weight 0: position i is taken while i <= ylen[b] (cassnat.py:590),
    the score is the running python-float sum of the picked log-probabilities (:619-621)."""
    out = []
    for b in range(ids.shape[0]):
        hyp, score = [sos], 0.0
        for i in range(ymax):
            if i > ylen[b]:
                continue
            score = score + float(logp[b, i])
            hyp.append(int(ids[b, i]))
        out.append((hyp, score))
    return out


def test_greedy_assembly_matches_reference_loop():
    rng = np.random.default_rng(5)
    B, U = 6, 9
    ids = rng.integers(0, 50, (B, U)).astype(np.int32)
    logp = (-rng.random((B, U)) * 3).astype(np.float32)
    ylen = np.asarray([8, 0, 3, 9, 5, 1], dtype=np.int32)            # 9 == U: the `i <= ylen` rule runs off the end
    ymax = U
    packed = torch.from_numpy(np.concatenate([ids, logp.view(np.int32), ylen[:, None]], 1))
    args = SimpleNamespace(use_gpu=False)
    got = CassNAT._assemble_greedy(packed, U, B, ymax, 1, args, torch.device("cpu"))
    want = _reference_pick(ids, logp, ylen, ymax, 1)
    for g, (hyp, score) in zip(got, want):
        assert g[0]["hyp"] == hyp
        assert g[0]["score"] == score                                     # same float64 summation order
        assert g[0]["ys"].tolist() == [[1]]


def test_pending_hypotheses_assembles_once():
    calls = []
    p = PendingHypotheses(lambda: calls.append(1) or ["x"], None)
    assert p.result() == ["x"] and p.result() is p.result() and calls == [1]


def test_ngram_ranker_detokenisation():
    """rank_model='n-gram' on CPU tensors: eos pieces (id 2) are dropped, word marks become blanks, the score is
    divided by the target length, the first maximum wins."""
    seen = []

    class Scorer:
        def score(self, sentence):
            seen.append(sentence)
            return -float(len(sentence))

    vocab = SimpleNamespace(index2word={0: "<b>", 1: "<s>", 2: "</s>", 3: "▁a", 4: "b", 5: "▁cd"})
    B, S, U = 1, 3, 4
    ids = torch.tensor([[3, 4, 2, 0], [5, 2, 0, 0], [3, 4, 5, 4]], dtype=torch.int32)
    tgt_mask = torch.tensor([[1, 1, 1, 0], [1, 1, 0, 0], [1, 1, 1, 1]], dtype=torch.uint8).unsqueeze(1)
    logp = torch.zeros(B * S, U)
    ylen = torch.tensor([3, 2, 4])
    out_ids, _, out_ylen, ymax, best = CassNAT._rescore_ngram(None, ids, logp, ylen, tgt_mask, Scorer(), vocab, B, S)
    assert seen == ["ab", "cd", "ab cdb"]
    # scores / lengths: -2/3, -2/2, -6/4 -> the first sample wins
    assert best.tolist() == [0] and out_ylen.tolist() == [3] and ymax == 3
    assert out_ids.tolist() == [[3, 4, 2, 0]]
