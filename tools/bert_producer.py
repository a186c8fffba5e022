"""Gradient producer for the end-to-end numbers: a plain PyTorch BERT encoder + classifier head.

NOT part of the product path (SURVEY.md 3.2 "forward + backward of the model ... not our path"):
it only exists so bench.py can report micro-steps/s with a real forward/backward in the loop, on
synthetic CoLA-shaped batches (input_ids[8,128], input_mask, segment_ids, label_ids[8] in {0,1}).
Parameter creation order and TensorFlow-style names follow upstream google-research/bert
(modeling.py, referenced by the reference's README.md:14), so T, P and the weight-decay mask are
exactly those of SURVEY.md 8 (BERT-Small: T=73, P=28 764 674).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

CONFIGS = {"bert_small": (4, 512, 8), "bert_base": (12, 768, 12), "bert_large": (24, 1024, 16)}


class Layer(nn.Module):
    def __init__(self, h, heads):
        super().__init__()
        self.heads = heads
        self.q, self.k, self.v = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)
        self.ao = nn.Linear(h, h)
        self.aln = nn.LayerNorm(h, eps=1e-12)
        self.inter = nn.Linear(h, 4 * h)
        self.out = nn.Linear(4 * h, h)
        self.oln = nn.LayerNorm(h, eps=1e-12)

    def forward(self, x, mask):
        b, s, h = x.shape
        sp = lambda t: t.view(b, s, self.heads, h // self.heads).transpose(1, 2)
        a = F.scaled_dot_product_attention(sp(self.q(x)), sp(self.k(x)), sp(self.v(x)), attn_mask=mask)
        x = self.aln(x + self.ao(a.transpose(1, 2).reshape(b, s, h)))
        return self.oln(x + self.out(F.gelu(self.inter(x))))


class Bert(nn.Module):
    def __init__(self, name="bert_small", vocab=30522, max_pos=512, type_vocab=2, num_labels=2):
        super().__init__()
        L, h, heads = CONFIGS[name]
        self.word = nn.Embedding(vocab, h)
        self.ttype = nn.Embedding(type_vocab, h)
        self.pos = nn.Embedding(max_pos, h)
        self.eln = nn.LayerNorm(h, eps=1e-12)
        self.layers = nn.ModuleList([Layer(h, heads) for _ in range(L)])
        self.pool = nn.Linear(h, h)
        self.output_weights = nn.Parameter(torch.empty(num_labels, h))
        self.output_bias = nn.Parameter(torch.zeros(num_labels))
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.normal_(p, 0.0, 0.02)          # BERT initializer_range

    def forward(self, input_ids, input_mask, segment_ids, label_ids):
        s = input_ids.shape[1]
        x = self.word(input_ids) + self.ttype(segment_ids) + self.pos.weight[:s]
        x = self.eln(x)
        mask = input_mask[:, None, None, :].bool()
        for l in self.layers:
            x = l(x, mask)
        pooled = torch.tanh(self.pool(x[:, 0]))
        logits = pooled @ self.output_weights.t() + self.output_bias
        return F.cross_entropy(logits, label_ids)


def tf_names(model: Bert):
    """(tf_name, parameter) in upstream BERT's variable creation order (LayerNorm: beta before gamma)."""
    out = []
    e = "bert/embeddings/"
    out += [(e + "word_embeddings", model.word.weight), (e + "token_type_embeddings", model.ttype.weight),
            (e + "position_embeddings", model.pos.weight), (e + "LayerNorm/beta", model.eln.bias),
            (e + "LayerNorm/gamma", model.eln.weight)]
    for i, l in enumerate(model.layers):
        b = f"bert/encoder/layer_{i}/"
        for nm, lin in (("query", l.q), ("key", l.k), ("value", l.v)):
            out += [(b + f"attention/self/{nm}/kernel", lin.weight), (b + f"attention/self/{nm}/bias", lin.bias)]
        out += [(b + "attention/output/dense/kernel", l.ao.weight), (b + "attention/output/dense/bias", l.ao.bias),
                (b + "attention/output/LayerNorm/beta", l.aln.bias), (b + "attention/output/LayerNorm/gamma", l.aln.weight),
                (b + "intermediate/dense/kernel", l.inter.weight), (b + "intermediate/dense/bias", l.inter.bias),
                (b + "output/dense/kernel", l.out.weight), (b + "output/dense/bias", l.out.bias),
                (b + "output/LayerNorm/beta", l.oln.bias), (b + "output/LayerNorm/gamma", l.oln.weight)]
    out += [("bert/pooler/dense/kernel", model.pool.weight), ("bert/pooler/dense/bias", model.pool.bias),
            ("output_weights", model.output_weights), ("output_bias", model.output_bias)]
    assert len(out) == sum(1 for _ in model.parameters())
    return out


def synthetic_batch(micro_bs, seq_len, device, gen, vocab=30522):
    ids = torch.randint(0, vocab, (micro_bs, seq_len), device=device, generator=gen)
    return ids, torch.ones_like(ids), torch.zeros_like(ids), torch.randint(0, 2, (micro_bs,), device=device, generator=gen)
