"""`Word2VecModel` with the reference's API and state-dict keys (model/word2vec_model.py:76-102) on the sm_100a
kernels: the text embedder right before the hot path.

    text_embed = model.lang_model(input_ids=tokens, attention_mask=tokens != 0)['pooler_output']   # train/main.py:58-60

    Embedding(V x 300) gather (frozen, as the reference's torch.no_grad lookup) -> fc1 (300 -> 2048) + ReLU -> masked
    max-pool over the 32 words (`-6e4` fill, all-stop-word sentences keep every word) -> fc2 (2048 -> 512)

Forward: tan_embed_gather_bf16, tan_text_pool_fc1 (fc1 GEMM with the pooling fused into its epilogue), tan_linear_bf16
(fc2).  Backward (fc1 / fc2 are trained through `lang_embed`, train/main.py:58-60): one autograd node,
tan_linear_bf16 (d pooled), tan_text_pool_bwd, tan_gemm_tn_bf16 (dW1, dW2), tan_colsum (biases).

The MIL-NCE word2vec weights are not part of either repository (model/readme.md:11-14): construct with the real
vocabulary size and load them with `load_state_dict`, or use the default random initialisation for tests.
`last_hidden_state` (fc2 of every word, never read by TAN) is only computed when `want_last_hidden_state` is set.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._lib import ACT_RELU, TanError
from .tfm_model import _Bf16Cache, _f32

MAX_WORDS = 32          # model/word2vec_model.py:28 (Word2VecTokenizer pads / cuts every sentence to 32 words)
K_PAD = 320             # 300 word2vec dimensions, zero-padded to a multiple of 64 (UMMA K blocks)


class _TextEmbedFn(torch.autograd.Function):
    """pooler_output as ONE autograd node over (fc1.weight, fc1.bias, fc2.weight, fc2.bias)."""

    @staticmethod
    def forward(ctx, module, x_tok, keep_u8, S, w1, b1, w2, b2):
        c = module._cache
        w1p = module._w1_padded()
        pooled, arg = ops.text_pool_fc1(x_tok, w1p, _f32(b1), keep_u8, S, want_argmax=True)
        out = torch.empty(S, w2.shape[0], dtype=torch.float32, device=x_tok.device)
        ops.linear(pooled, c.get(w2), _f32(b2), out_f32=out, tag="text_embed")
        ctx.module, ctx.S = module, S
        ctx.save_for_backward(x_tok, pooled, arg)
        return out

    @staticmethod
    def backward(ctx, g_out):
        module, S = ctx.module, ctx.S
        x_tok, pooled, arg = ctx.saved_tensors
        w1, b1, w2, b2 = module.fc1.weight, module.fc1.bias, module.fc2.weight, module.fc2.bias
        F, dout = w2.shape[1], w2.shape[0]
        dev = x_tok.device
        with torch.no_grad():
            g = g_out.detach().float().contiguous()
            gb = ops.cast_bf16(g) if g.numel() % 8 == 0 else g.to(torch.bfloat16)
            # fc2
            gw2 = torch.zeros(dout, F, dtype=torch.float32, device=dev)
            gb2 = torch.zeros(dout, dtype=torch.float32, device=dev)
            ops.gemm_tn(gb, pooled, gw2, accumulate=False, tag="text_embed")
            ops.colsum(g, gb2, accumulate=False)
            dpool = torch.empty(S, F, dtype=torch.bfloat16, device=dev)
            ops.linear(gb, ops.transpose_bf16(module._cache.get(w2)), out_bf16=dpool, tag="text_embed")   # g @ W2
            # max-pool + ReLU -> the arg-max word of every (sentence, feature)
            dH = ops.text_pool_bwd(dpool, pooled, arg)
            gw1p = torch.zeros(F, K_PAD, dtype=torch.float32, device=dev)
            gb1 = torch.zeros(F, dtype=torch.float32, device=dev)
            ops.gemm_tn(dH, x_tok, gw1p, accumulate=False, tag="text_embed")
            ops.colsum(dH, gb1, accumulate=False)
            gw1 = gw1p[:, :w1.shape[1]].contiguous()
        return (None, None, None, None, gw1.to(w1.dtype), gb1.to(b1.dtype), gw2.to(w2.dtype), gb2.to(b2.dtype))


class Word2VecModel(nn.Module):
    """model/word2vec_model.py:76-102.  Parameters `word_embd.weight [V, 300]`, `fc1.{weight,bias}`,
    `fc2.{weight,bias}` (the reference takes these three modules from the S3D text module, :79-82)."""

    def __init__(self, num_embeddings: int = 66250, word_dim: int = 300, hidden: int = 2048, out_dim: int = 512):
        super().__init__()
        if word_dim > K_PAD or hidden % 128 != 0 or out_dim % 128 != 0:
            raise TanError("Word2VecModel: word_dim <= 320, hidden and out_dim multiples of 128")
        self.word_embd = nn.Embedding(num_embeddings, word_dim)
        self.fc1 = nn.Linear(word_dim, hidden)
        self.fc2 = nn.Linear(hidden, out_dim)
        self.want_last_hidden_state = False
        self._cache = _Bf16Cache()
        self._table = None           # (version, data_ptr, bf16 [V, 320])
        self._w1p = None

    def _table_bf16(self) -> torch.Tensor:
        w = self.word_embd.weight
        ent = self._table
        if ent is None or ent[0] != w._version or ent[1] != w.data_ptr():
            t = torch.zeros(w.shape[0], K_PAD, dtype=torch.bfloat16, device=w.device)
            t[:, :w.shape[1]] = w.detach().to(torch.bfloat16)
            ent = self._table = (w._version, w.data_ptr(), t)
        return ent[2]

    def _w1_padded(self) -> torch.Tensor:
        w = self.fc1.weight
        ent = self._w1p
        if ent is None or ent[0] != w._version or ent[1] != w.data_ptr():
            t = torch.zeros(w.shape[0], K_PAD, dtype=torch.bfloat16, device=w.device)
            t[:, :w.shape[1]] = w.detach().to(torch.bfloat16)
            ent = self._w1p = (w._version, w.data_ptr(), t)
        return ent[2]

    def forward(self, input_ids, attention_mask=None, *args, **kwargs):
        if not input_ids.is_cuda:
            raise TanError("Word2VecModel runs on a CUDA (sm_100a) device only; there is no CPU path")
        if input_ids.dim() != 2 or input_ids.shape[1] > MAX_WORDS:
            raise TanError(f"input_ids must be [sentences, <= {MAX_WORDS} words], got {tuple(input_ids.shape)}")
        S, W = input_ids.shape
        ids = input_ids.long()
        keep = None if attention_mask is None else attention_mask.to(torch.uint8)
        if W < MAX_WORDS:                      # shorter padding length: the missing words are ignored ones
            ids = torch.nn.functional.pad(ids, (0, MAX_WORDS - W))
            keep = torch.nn.functional.pad(keep if keep is not None else torch.ones(S, W, dtype=torch.uint8, device=ids.device),
                                           (0, MAX_WORDS - W))
        ids = ids.contiguous().view(-1)
        keep = None if keep is None else keep.contiguous().view(-1)
        with torch.no_grad():
            x_tok = ops.embed_gather(ids, self._table_bf16())                  # [S*32, 320] bf16 (frozen lookup, :84-85)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in (self.fc1.weight, self.fc1.bias,
                                                                              self.fc2.weight, self.fc2.bias))
        if need_grad:
            pooled_out = _TextEmbedFn.apply(self, x_tok, keep, S, self.fc1.weight, self.fc1.bias, self.fc2.weight,
                                            self.fc2.bias)
        else:
            with torch.no_grad():
                pooled, _ = ops.text_pool_fc1(x_tok, self._w1_padded(), _f32(self.fc1.bias), keep, S)
                pooled_out = torch.empty(S, self.fc2.weight.shape[0], dtype=torch.float32, device=ids.device)
                ops.linear(pooled, self._cache.get(self.fc2.weight), _f32(self.fc2.bias), out_f32=pooled_out,
                           tag="text_embed")
        out = {'pooler_output': pooled_out}
        if self.want_last_hidden_state:       # fc2(relu(fc1(x))) of every word (:98): not read by TAN
            with torch.no_grad():
                h = torch.empty(S * MAX_WORDS, self.fc1.weight.shape[0], dtype=torch.bfloat16, device=ids.device)
                ops.linear(x_tok, self._w1_padded(), _f32(self.fc1.bias), out_bf16=h, act=ACT_RELU, tag="text_embed")
                last = torch.empty(S * MAX_WORDS, self.fc2.weight.shape[0], dtype=torch.float32, device=ids.device)
                ops.linear(h, self._cache.get(self.fc2.weight), _f32(self.fc2.bias), out_f32=last, tag="text_embed")
                out['last_hidden_state'] = last.view(S, MAX_WORDS, -1)[:, :W]
        return out
