"""CPU: the oracle against the golden vectors frozen from the REFERENCE (oracle/make_golden.py ran the reference's own
modules in the authoring container).  This is what pins the oracle; the GPU tests then compare kernels to the oracle."""
import os

import numpy as np
import torch

from oracle import dvae_oracle as O


def test_state_dict_inventory(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "train_step_R4.pt"))
    sd = O.synth_state_dict(0)
    assert list(sd.keys()) == gold["state_dict_keys"]            # 84 parameters + 33 buffers, reference order
    assert {k: tuple(v.shape) for k, v in sd.items()} == gold["state_dict_shapes"]
    assert len(O.param_shapes()) == 84 and len(O.buffer_shapes()) == 33
    assert sum(int(np.prod(s)) for s in O.param_shapes().values()) == 61367680


def test_train_step_matches_reference(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "train_step_R4.pt"))
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.clone_sd(O.synth_state_dict(gold["weights_seed"]), requires_grad=True)
    x1, x2, eps = O.synth_inputs(gold["R"], seed=gold["inputs_seed"])
    out, losses, grads = O.train_step(sd, x1, x2, eps, batch_size=gold["R"])
    for a, b in zip(out, gold["forward"]):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-4)
    for a, b in zip(losses, gold["losses"]):
        assert abs(a.item() - b.item()) <= 1e-5 * abs(b.item()) + 1e-6
    for k, d in gold["grad_digest"].items():
        g = grads[k].detach().reshape(-1).double()
        scale = d["norm"] + 1e-12
        assert abs(g.norm().item() - d["norm"]) <= 1e-3 * scale + 1e-9, k
        assert (g[d["idx"]].float() - d["samples"]).abs().max().item() <= 1e-3 * scale / max(1.0, g.numel() ** 0.5) + 1e-7, k
    for k, v in gold["bn_buffers_after"].items():
        assert torch.allclose(sd[k].float(), v.float(), atol=1e-6), k


def test_eval_forward_and_conversion_match_reference(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "eval_forward_R4.pt"))
    sd = O.synth_state_dict(0)
    x1, x2, eps = O.synth_inputs(gold["R"])
    with torch.no_grad():
        out = O.forward(O.clone_sd(sd), x1, x2, eps, training=False, sample_content=False)
    for a, b in zip(out, gold["forward"]):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-4)
    conv = torch.load(os.path.join(golden_dir, "convert.pt"))
    for T, shape in conv["chunk_shapes"].items():
        assert tuple(O.chunking_mel(np.zeros((80, T), np.float32)).shape) == shape
    assert conv["chunk_shapes"][64] == (2, 80, 64) and conv["chunk_shapes"][63] == (1, 80, 64)   # extra zero chunk at T%64==0
    src = torch.from_numpy(O.chunking_mel(conv["src"].numpy()))
    trg = torch.from_numpy(O.chunking_mel(conv["trg"].numpy()))
    rec, cvt = O.convert(O.clone_sd(sd), src, trg)
    assert torch.allclose(rec, conv["recons"], atol=1e-5) and torch.allclose(cvt, conv["converted"], atol=1e-5)


def test_lstm_builtin_equals_explicit_recurrence():
    sd = O.synth_state_dict(3)
    x = torch.randn(3, 64, 512)
    a = O.lstm(sd, x, "enc_lstm", 2, True)
    b = O.lstm_explicit(sd, x, "enc_lstm", 2, True)
    assert torch.allclose(a, b, atol=1e-5)


def test_product_of_gaussians_matches_reference(golden_dir):
    cases = torch.load(os.path.join(golden_dir, "pog_cases.pt"))
    assert set(cases) == {"sorted_equal", "unsorted", "singletons", "one_group", "noncontig_ids"}
    for name, c in cases.items():
        gm, glv = O.accumulate_group_evidence(c["mu"].numpy(), c["logvar"].numpy(), c["labels"].numpy())
        assert np.allclose(gm, c["group_mu"].numpy(), atol=2e-6), name
        assert np.allclose(glv, c["group_logvar"].numpy(), atol=2e-6), name
        gid, counts = O.group_segments(c["labels"].numpy())
        assert np.array_equal(gid, c["gid"].numpy()) and np.array_equal(counts, c["counts"].numpy())   # bit exact
    u = cases["unsorted"]
    assert np.isneginf(u["logvar"].numpy()[3, 2])          # the exact-zero-variance element took the 1e-6 clamp path
    assert np.isfinite(u["group_logvar"].numpy()).all()
