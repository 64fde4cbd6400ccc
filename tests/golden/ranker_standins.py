"""Deterministic stand-ins for the caller-supplied rankers of the sampled-alignment decode (cassnat.py:499-536), shared
by make_golden_rankers.py (which runs the UNMODIFIED reference with them) and the tests (which run this package with
them): an n-gram scorer with kenlm's `score(sentence)` call, an autoregressive ASR model with the three calls the
'at_baseline' branch makes, and the piece table behind `vocab.index2word`."""
import torch


def index2word(V):
    """Sentencepiece-like pieces: every third id starts a word."""
    return {i: ("▁" if i % 3 == 0 else "") + "p%d" % i for i in range(V)}


class TinyNgram:
    """score(sentence) -> log10-probability-like number that depends on every character and on the word count."""

    def score(self, sentence):
        h = 0
        for ch in sentence:
            h = (h * 131 + ord(ch)) % 1000003
        return -0.37 * len(sentence.split()) - (h % 9973) / 997.0


class TinyAT(torch.nn.Module):
    """src_embed(src, x_mask) -> (x, mask); encoder(x, mask) -> h; forward_decoder(h, ys, mask, ys_mask) -> log-probs:
    the calls cassnat.py:515-519 makes on the AT baseline model."""

    def __init__(self, feat, V, d=16, seed=11):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        P = lambda *s: torch.nn.Parameter(torch.randn(*s, generator=g) * 0.5, requires_grad=False)
        self.w_in, self.w_enc, self.emb, self.out = P(feat, d), P(d, d), P(V, d), P(d, V)

    def src_embed(self, src, x_mask):
        return src[:, ::4] @ self.w_in, x_mask[:, :, ::4]

    def encoder(self, x, mask):
        return torch.tanh(x @ self.w_enc) * mask.transpose(1, 2).to(x.dtype)

    def forward_decoder(self, h, ys, mask, ys_mask):
        m = mask.to(h.dtype)                                                    # [N,1,T]
        ctx = (m @ h) / m.sum(-1, keepdim=True).clamp(min=1)                    # [N,1,d] utterance summary
        y = self.emb[ys]
        a = (ys_mask != 0).to(h.dtype)
        y = (a @ y) / a.sum(-1, keepdim=True).clamp(min=1)                       # causal running mean of the ids
        return torch.log_softmax((y + ctx) @ self.out, -1)
