"""Generate tests/golden/*.pt by EXECUTING THE REFERENCE (authoring container only).

Run:  python oracle/make_golden.py            (needs /root/reference; CPU, ~1 min)

The reference cannot travel to the GPU box, so its outputs on small seeded inputs are frozen
here.  Import recipe = SURVEY.md 8(c): MagicMock stubs for the non-numeric dependencies, a no-op
Tensor.cuda (the model hard-codes .cuda(), model/disentangled_vae.py:224) and externally
supplied reparameterisation noise (replace `_reparameterize` by a queue).
Weights come from oracle.dvae_oracle.synth_state_dict (numpy Philox, independent of torch's
RNG) and are loaded through the reference's own load_state_dict, which also proves the
84 + 33 key inventory.
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference():
    for m in ["librosa", "librosa.display", "librosa.filters", "matplotlib", "matplotlib.pyplot",
              "mpl_toolkits", "mpl_toolkits.axes_grid1", "soundfile", "tensorboardX",
              "wavenet_vocoder", "wavenet_vocoder.builder", "pyworld", "pysptk", "lws"]:
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    import model.disentangled_vae as ref_vae                 # noqa
    import model.utils as ref_utils                          # noqa
    import model.variational_base_vae as ref_base            # noqa
    torch.set_num_threads(os.cpu_count())                    # undo the import side effect (F8)
    return ref_vae, ref_utils, ref_base


def grad_digest(g: torch.Tensor, name: str):
    """Small fingerprint of a gradient tensor: norm, sum, first 8 values, a fixed random projection."""
    import zlib
    flat = g.detach().reshape(-1).double()
    rng = np.random.Generator(np.random.Philox(key=[zlib.crc32(name.encode()) & 0xFFFFFFFF, 77]))
    idx = torch.from_numpy(rng.integers(0, flat.numel(), size=min(64, flat.numel())))
    return {"norm": flat.norm().item(), "sum": flat.sum().item(),
            "head": flat[:8].float().clone(), "idx": idx, "samples": flat[idx].float().clone()}


def main():
    from oracle import dvae_oracle as O
    ref_vae, ref_utils, ref_base = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.manual_seed(0)

    # ---------------- training step, R = 4 ------------------------------------------------
    R = 4
    sd = O.synth_state_dict(seed=0)
    x1, x2, eps = O.synth_inputs(R, seed=1234)
    wrapper = ref_vae.ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=R,
                                          speaker_size=4, device=torch.device("cpu"), latent_dim=32,
                                          beta=0.1, mse_cof=10, kl_cof=10, style_cof=0.1)
    model = wrapper.model
    ref_keys = list(model.state_dict().keys())
    assert ref_keys == list(sd.keys()), "state_dict key inventory / order mismatch"
    model.load_state_dict(sd)
    model.train()
    queue = [e.clone() for e in eps]

    def fake_reparam(mu, logvar, train=True):
        if train:
            e = queue.pop(0)
            assert e.shape == logvar.shape
            return e.mul(logvar.mul(0.5).exp()).add(mu)
        return mu
    model._reparameterize = fake_reparam
    out = model(x1, x2)
    losses = wrapper.loss_functionGVAE2(x1, x2, *out)
    losses[0].backward()
    grads = {k: p.grad for k, p in model.named_parameters()}
    buffers_after = {k: v.clone() for k, v in model.state_dict().items() if "running_" in k or "num_batches" in k}

    # oracle vs reference, same inputs
    osd = O.clone_sd(sd, requires_grad=True)
    o_out, o_losses, o_grads = O.train_step(osd, x1, x2, eps, batch_size=R)
    worst = 0.0
    for a, b in zip(list(out) + list(losses), list(o_out) + list(o_losses)):
        worst = max(worst, (a - b).abs().max().item() / (b.abs().max().item() + 1e-12))
    gworst = 0.0
    for k in grads:
        d = (grads[k] - o_grads[k]).norm().item() / (grads[k].norm().item() + 1e-20)
        gworst = max(gworst, d)
    bworst = max((buffers_after[k].float() - osd[k].float()).abs().max().item() for k in buffers_after)
    print(f"[train R={R}] oracle vs reference: fwd/loss max-rel {worst:.3e}, grad rel-L2 {gworst:.3e}, "
          f"BN buffers max-abs {bworst:.3e}")
    assert worst < 1e-5 and gworst < 1e-4 and bworst < 1e-6

    torch.save({
        "R": R, "weights_seed": 0, "inputs_seed": 1234,
        "forward": [t.detach().clone() for t in out],
        "losses": [t.detach().clone() for t in losses],
        "grad_digest": {k: grad_digest(g, k) for k, g in grads.items()},
        "bn_buffers_after": buffers_after,
        "state_dict_keys": ref_keys,
        "state_dict_shapes": {k: tuple(v.shape) for k, v in model.state_dict().items()},
    }, os.path.join(out_dir, "train_step_R4.pt"))

    # ---------------- forward(train=False) in eval mode (content = mu, style still sampled: F7)
    model.load_state_dict(sd)
    model.eval()
    queue[:] = [eps[2].clone()]
    with torch.no_grad():
        out_eval = model(x1, x2, train=False)
    osd = O.clone_sd(sd)
    with torch.no_grad():
        o_eval = O.forward(osd, x1, x2, eps, training=False, sample_content=False)
    w = max((a - b).abs().max().item() for a, b in zip(out_eval, o_eval))
    print(f"[eval fwd] oracle vs reference max-abs {w:.3e}")
    assert w < 1e-5
    torch.save({"R": R, "forward": [t.clone() for t in out_eval]}, os.path.join(out_dir, "eval_forward_R4.pt"))

    # ---------------- conversion core ----------------------------------------------------
    g = np.random.Generator(np.random.Philox(key=[5, 5]))
    src = g.uniform(0, 1, size=(80, 150)).astype(np.float32)
    trg = g.uniform(0, 1, size=(80, 128)).astype(np.float32)    # T % 64 == 0 -> extra zero chunk
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        src_c = ref_base.chunking_mel(src).float()
        trg_c = ref_base.chunking_mel(trg).float()
    assert np.array_equal(src_c.numpy(), O.chunking_mel(src)) and np.array_equal(trg_c.numpy(), O.chunking_mel(trg))
    with torch.no_grad():   # model/variational_base_vae.py:277-296, verbatim call sequence
        s_mu, _, c_mu, _ = model.encode(src_c)
        t_mu, _, _, _ = model.encode(trg_c)
        s_style = torch.mean(s_mu, axis=0, keepdim=True).repeat(src_c.shape[0], 1)
        t_style = torch.mean(t_mu, axis=0, keepdim=True).repeat(src_c.shape[0], 1)
        rec = model.decode(torch.cat([s_style, c_mu], dim=-1))
        rec = torch.cat([rec[i] for i in range(rec.shape[0])], 1)
        conv = model.decode(torch.cat([t_style, c_mu], dim=-1))
        conv = conv + model.postnet(conv)
        conv = torch.clamp(torch.cat([conv[i] for i in range(conv.shape[0])], 1), min=0, max=1.0)
    o_rec, o_conv = O.convert(O.clone_sd(sd), src_c, trg_c)
    w = max((rec - o_rec).abs().max().item(), (conv - o_conv).abs().max().item())
    print(f"[convert] oracle vs reference max-abs {w:.3e}; chunks src {tuple(src_c.shape)} trg {tuple(trg_c.shape)}")
    assert w < 1e-5
    torch.save({"src": torch.from_numpy(src), "trg": torch.from_numpy(trg), "recons": rec, "converted": conv,
                "chunk_shapes": {T: tuple(O.chunking_mel(np.zeros((80, T), np.float32)).shape) for T in (63, 64, 65, 512)}},
               os.path.join(out_dir, "convert.pt"))

    # ---------------- product-of-Gaussians group utilities (model/utils.py) ---------------
    cases = {}
    patterns = {
        "sorted_equal": np.repeat(np.arange(4), 3),
        "unsorted": np.array([5, 2, 5, 9, 2, 2, 7, 9, 5, 5]),
        "singletons": np.arange(6),
        "one_group": np.zeros(7, dtype=np.int64),
        "noncontig_ids": np.array([1000, 3, 3, 1000, 42, 42, 42, 3]),
    }
    for name, lab in patterns.items():
        gg = np.random.Generator(np.random.Philox(key=[11, len(lab)]))
        mu = gg.standard_normal((len(lab), 8)).astype(np.float32)
        lv = (0.5 * gg.standard_normal((len(lab), 8))).astype(np.float32)
        if name == "unsorted":
            lv[3, 2] = -np.inf          # exact-zero variance -> 1e-6 clamp (model/utils.py:31)
        labels_t = torch.from_numpy(lab.astype(np.int64))
        gm, glv = ref_utils.accumulate_group_evidence(torch.from_numpy(mu.copy()), torch.from_numpy(lv.copy()),
                                                      labels_t, False)
        om, olv = O.accumulate_group_evidence(mu, lv, lab)
        d = max(np.abs(gm.detach().numpy() - om).max(), np.abs(glv.detach().numpy() - olv).max())
        assert d < 2e-6, (name, d)
        gid, counts = O.group_segments(lab)
        cases[name] = {"labels": labels_t, "mu": torch.from_numpy(mu), "logvar": torch.from_numpy(lv),
                       "group_mu": gm.detach().clone(), "group_logvar": glv.detach().clone(),
                       "gid": torch.from_numpy(gid), "counts": torch.from_numpy(counts)}
        print(f"[PoG {name}] oracle vs reference max-abs {d:.3e}")
    torch.save(cases, os.path.join(out_dir, "pog_cases.pt"))
    # ---------------- AutoVC-style generator (autovc_replicate/proposed_autovc.py, BASELINE config 5) ---------------
    import io, contextlib
    from oracle import autovc_oracle as A
    with contextlib.redirect_stdout(io.StringIO()):       # the reference module runs a forward + print at import (:223-227)
        import autovc_replicate.proposed_autovc as ref_avc
    gen = ref_avc.Generator()
    asd = A.synth_state_dict(0)
    assert list(gen.state_dict().keys()) == list(asd.keys()), "autovc state_dict inventory / order mismatch"
    gen.load_state_dict(asd)
    gen.train()
    ax, _, _ = O.synth_inputs(4, seed=4321)
    mel, mel_post = gen(ax)
    loss = A.sq_loss(ax, mel, mel_post)
    loss.backward()
    agr = {k: p.grad for k, p in gen.named_parameters()}
    osd = O.clone_sd(asd, requires_grad=True)
    (o_mel, o_post), o_loss, o_gr = A.train_step(osd, ax)
    w = max((mel - o_mel).abs().max().item(), (mel_post - o_post).abs().max().item())
    gw = max((agr[k] - o_gr[k]).norm().item() / (agr[k].norm().item() + 1e-20) for k in agr)
    print(f"[autovc R=4] oracle vs reference: outputs max-abs {w:.3e}, loss rel {abs(loss.item() - o_loss.item()) / abs(loss.item()):.3e}, grad rel-L2 {gw:.3e}")
    assert w < 1e-5 and gw < 1e-4
    torch.save({"R": 4, "inputs_seed": 4321, "mel": mel.detach().clone(), "mel_postnet": mel_post.detach().clone(),
                "loss": loss.detach().clone(), "grad_digest": {k: grad_digest(g, k) for k, g in agr.items()},
                "state_dict_keys": list(asd.keys()),
                "bn_buffers_after": {k: v.clone() for k, v in gen.state_dict().items() if "running_" in k}},
               os.path.join(out_dir, "autovc_R4.pt"))
    for f in sorted(os.listdir(out_dir)):
        print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == "__main__":
    main()
