"""Test scaffolding for SURVEY.md §8(f) F4: the host layers around the attention block of the reference's
`CachedTransformerModel` (ch02/cached_generation.py:101-205: RMSNorm, SwiGLU FFN, embedding, LM head) and its
`cached_generate` loop (:208-274), with the attention block injected.  Those layers are dense GEMMs and host control flow —
out of scope for the product — so they live here, not in the package; submodules are created in the
reference's order so that `torch.manual_seed(s)` reproduces the reference model's weights exactly (checked against the
weight checksum stored in tests/golden/r2_ch02_generate.npz)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class RMSNorm(nn.Module):
    def __init__(self, hidden_dim, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_dim))
        self.eps = eps

    def forward(self, x):
        return x / torch.sqrt(torch.mean(x ** 2, dim=-1, keepdim=True) + self.eps) * self.weight


class SwiGLUFFN(nn.Module):
    def __init__(self, hidden_dim, intermediate_dim):
        super().__init__()
        self.gate_proj = nn.Linear(hidden_dim, intermediate_dim, bias=False)
        self.up_proj = nn.Linear(hidden_dim, intermediate_dim, bias=False)
        self.down_proj = nn.Linear(intermediate_dim, hidden_dim, bias=False)

    def forward(self, x):
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


class Block(nn.Module):
    def __init__(self, attn_cls, hidden_dim, num_heads, num_kv_heads, intermediate_dim):
        super().__init__()
        self.input_norm = RMSNorm(hidden_dim)
        self.attn = attn_cls(hidden_dim, num_heads, num_kv_heads)
        self.post_attn_norm = RMSNorm(hidden_dim)
        self.ffn = SwiGLUFFN(hidden_dim, intermediate_dim)

    def forward(self, x, cache=None, start_pos=0):
        h = x + self.attn(self.input_norm(x), cache, start_pos)
        return h + self.ffn(self.post_attn_norm(h))


class TinyCachedModel(nn.Module):
    def __init__(self, attn_cls, vocab_size, hidden_dim, num_layers, num_heads, num_kv_heads, intermediate_dim):
        super().__init__()
        self.embed = nn.Embedding(vocab_size, hidden_dim)
        self.layers = nn.ModuleList([Block(attn_cls, hidden_dim, num_heads, num_kv_heads, intermediate_dim)
                                     for _ in range(num_layers)])
        self.norm = RMSNorm(hidden_dim)
        self.lm_head = nn.Linear(hidden_dim, vocab_size, bias=False)
        self.num_layers, self.num_kv_heads, self.head_dim = num_layers, num_kv_heads, hidden_dim // num_heads

    def forward(self, input_ids, caches=None, start_pos=0):
        x = self.embed(input_ids)
        for i, layer in enumerate(self.layers):
            x = layer(x, caches[i] if caches is not None else None, start_pos)
        return self.lm_head(self.norm(x))


def cached_generate(model, input_ids, max_new_tokens, create_caches, temperature=1.0):
    """The loop of ch02/cached_generation.py:208-274 (prefill, sample, then one token per step through the caches),
    without the timers.  `create_caches(batch, max_seq_len)` supplies the per-layer caches."""
    model.eval()
    batch, prompt_len = input_ids.shape
    caches = create_caches(batch, prompt_len + max_new_tokens)
    generated = []
    with torch.no_grad():
        logits = model(input_ids, caches, start_pos=0)
        next_token = torch.multinomial(F.softmax(logits[:, -1, :] / temperature, dim=-1), num_samples=1)
        generated.append(next_token)
        for i in range(max_new_tokens - 1):
            logits = model(next_token, caches, start_pos=prompt_len + i)
            next_token = torch.multinomial(F.softmax(logits[:, -1, :] / temperature, dim=-1), num_samples=1)
            generated.append(next_token)
    return torch.cat([input_ids] + generated, dim=1), logits
