"""One small launch of every kernel family of libtan_b200.so, for compute-sanitizer (scripts/gpu_sanitize.sh).
Shapes are tiny (a sanitizer run is 10-100x slower) but exercise every role: multi-tile persistent GEMMs, ragged
tails, masks, both attention-backward kernels, the fused epilogues."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import TemporalAligner, get_loss, ops, synth  # noqa: E402
from temporalalignnet_b200.optim import FusedAdamW  # noqa: E402
from temporalalignnet_b200.word2vec_model import Word2VecModel  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
bf = torch.bfloat16


def r(*s, dtype=torch.float32):
    return torch.randn(*s, device=dev).to(dtype)


# GEMM family
a, w, b = r(300, 128, dtype=bf), r(256, 128, dtype=bf), r(256)
o32, obf = torch.empty(300, 256, device=dev), torch.empty(300, 256, dtype=bf, device=dev)
ops.linear(a, w, b, out_bf16=obf, act=1)
ops.linear(a, w, b, residual=o32.zero_(), out_f32=o32)
pre = torch.empty_like(obf)
ops.linear_dual(a, w, b, obf, pre)
ops.linear_gelu_bwd(a, w, pre, obf)
x = r(300, 512)
ops.linear_res_ln(r(300, 128, dtype=bf), r(512, 128, dtype=bf), r(512), x, r(512), r(512), torch.empty(300, 512, dtype=bf, device=dev))
g = torch.zeros(256, 128, device=dev)
ops.gemm_tn(obf, a, g, accumulate=True)
ops.gemm_tn(r(5000, 512, dtype=bf), r(5000, 512, dtype=bf), torch.zeros(512, 512, device=dev), accumulate=False)
# attention forward + backward (masked, ragged length, two heads)
B, H, L = 2, 2, 100
qkv = r(B * L, 3 * H * 64, dtype=bf)
d = H * 64
kpm = torch.zeros(B, L, dtype=torch.uint8, device=dev)
kpm[1, 90:] = 1
o = torch.empty(B * L, d, dtype=bf, device=dev)
lse = torch.empty(B, H, ops.pad64(L), device=dev)
ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], kpm, o, B, H, L, L, lse=lse)
dq = torch.empty(B * L, 3 * d, dtype=bf, device=dev)
ops.attention_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], o, r(B * L, d, dtype=bf), kpm, dq[:, :d], dq[:, d:2 * d],
                  dq[:, 2 * d:], lse, torch.empty_like(lse), B, H, L, L)
torch.cuda.synchronize()
# the model: inference forward + fused loss, training step (ragged columns, tape, backward), optimizer step
E = D = 1
sd = synth.make_state_dict(E, D)
batch = synth.make_batch(3, 32, 4, pad_video_every=2)
m = TemporalAligner(E, D, random_pos_start=0)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
m = m.to(dev)
video, text = torch.from_numpy(batch["video"]).to(dev), torch.from_numpy(batch["text"]).to(dev)
vpm, tpm = torch.from_numpy(batch["video_padding_mask"]).to(dev), torch.from_numpy(batch["text_padding_mask"]).to(dev)
idata = {"start": batch["start"], "end": batch["end"], "text": batch["text_str"]}
args = types.SimpleNamespace(model="init", sim="cos", learn_agreement=0, loss_threshold=0.0, use_alignability_head=0)
out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm)
print("loss", float(get_loss(idata, video, text, vpm.float(), tpm.float(), out, args, None)["loss"]))
out["logits_joint"].materialize()
args2 = types.SimpleNamespace(model="init", sim="cos", learn_agreement=1, temporal_agreement_type="keep", loss_threshold=0.5,
                              use_alignability_head=0)
print("loss (flags)", float(get_loss(idata, video, text, vpm.float(), tpm.float(), out, args2, None)["loss"]))
m.train()
m.enable_autograd(True)
out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm)
loss = get_loss(idata, video, text, vpm.float(), tpm.float(), out, args, None)["loss"]
loss.backward()
opt = FusedAdamW([p for p in m.parameters() if p.grad is not None], lr=1e-4, clip_grad=3.0)
opt.step()
# text embedder
w2v = Word2VecModel(num_embeddings=200).to(dev)
tok = torch.randint(0, 200, (9, 32), device=dev)
e = w2v(input_ids=tok, attention_mask=tok != 0)["pooler_output"]
e.sum().backward()
torch.cuda.synchronize()
# sliding-window alignment: batched windows, stitching + decision kernels (two batches: accumulate / finalize flags)
import numpy as np  # noqa: E402
from temporalalignnet_b200.align import plan_windows, predicted_frames, sliding_window_alignment  # noqa: E402
sd3 = synth.make_state_dict(1, 3, use_alignability_head=True, seed=12)
m3 = TemporalAligner(1, 3, random_pos_start=0, use_alignability_head=1)
m3.load_state_dict({k: torch.from_numpy(v) for k, v in sd3.items()})
m3 = m3.to(dev)
wins = plan_windows(90, 32, np.linspace(1, 88, 10), np.ones(10, bool))
res = sliding_window_alignment(m3, r(90, 1024), r(10, 512), wins, max_windows_per_batch=3)
print("predicted frames", predicted_frames(res["sim"]).tolist())
torch.cuda.synchronize()
print("sanitize driver done, launches:", ops.launches())
