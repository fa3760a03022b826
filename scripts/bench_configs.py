"""Throughput of the two secondary BASELINE configurations on one B200 (bf16):
   config 4  many-to-many conversion, 512 utterances x 512 frames (9 chunks each incl. the zero chunk) -> convert_chunks
   config 5  AutoVC-style generator fwd+bwd, batch 256 x 128 frames = [512, 80, 64]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch

os.environ.setdefault("DVAE_B200_PRECISION", "bf16")
from autovc_replicate.proposed_autovc import Generator
from model.disentangled_vae import ConvolutionalMulVAE


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
torch.manual_seed(0)
# ---- config 5
g = Generator().cuda()
g.train()
x = torch.rand(512, 80, 64, device="cuda")


def step5():
    for p in g.parameters():
        p.grad = None
    mel, post = g(x)
    t = x.transpose(1, 2).unsqueeze(1)
    (0.5 * ((mel - t).pow(2).sum() + (post - t).pow(2).sum())).backward()


ms = timeit(step5)
res["config5_autovc_fwd_bwd"] = {"ms_per_step": ms, "mel_frames_per_s": 512 * 64 / (ms * 1e-3), "rows": 512}
del g
# ---- config 4
w = ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=8, speaker_size=4, latent_dim=32)
w.model.eval()
U, CH = 512, 9
src = torch.rand(U * CH, 80, 64, device="cuda")
src.view(U, CH, 80, 64)[:, -1] = 0           # chunking_mel appends a zero chunk when T % 64 == 0
trg = torch.rand(U * CH, 80, 64, device="cuda")
utt = torch.arange(U, device="cuda", dtype=torch.int32).repeat_interleave(CH)
ms = timeit(lambda: w.convert_chunks(src, utt, trg, utt, U), n=3, warm=1)
res["config4_conversion"] = {"ms_per_batch": ms, "utterances_per_s": U / (ms * 1e-3), "mel_frames_per_s": U * 512 / (ms * 1e-3),
                             "rows": U * CH}
print(json.dumps(res))
