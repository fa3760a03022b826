"""CPU: the C-ABI library loads and exports every declared symbol, host-side logic (drop-in module surface, sharding,
gradient buckets incl. a 2-rank gloo all-reduce), and the product refuses to run without CUDA."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "disentangle-vae-for-vc_b200")


@pytest.fixture(scope="module")
def built_lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as G
    from dvae_b200.build import build
    return build()


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    header = open(os.path.join(ROOT, "include", "dvae_b200.h")).read()
    names = sorted(set(re.findall(r"\b(dvae_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 36
    for n in names:
        assert hasattr(lib, n), n
    from dvae_b200 import lib as L
    assert set(L.exported_symbols()) <= set(names) | {"dvae_last_error"}
    assert set(names) <= set(L.exported_symbols()), "every header symbol must be bound by the Python host side"
    assert L.version() == 100 and L.lstm_gate_tile(64) in (64, 128, 256) and L.lstm_gate_tile(1024) in (64, 128, 256)


def test_dropin_module_surface(built_lib):
    import inspect
    from model.disentangled_vae import ConvolutionalMulVAE, DisentangledVAE, Postnet
    from model.variational_base_vae import VariationalBaseModelVAE, chunking_mel
    from oracle import dvae_oracle as O
    sig = inspect.signature(DisentangledVAE.__init__)
    assert list(sig.parameters)[1:14] == ["speaker_size", "input_sz", "kernel_szs", "hidden_sz", "latent_sz", "c", "c_delta",
                                          "beta", "beta_delta", "dim_neck", "latent_dim", "dim_pre", "batch_size"]
    sigw = inspect.signature(ConvolutionalMulVAE.__init__)
    assert list(sigw.parameters)[1:] == ["dataset", "width", "height", "latent_sz", "learning_rate", "alpha", "log_interval",
                                         "normalize", "batch_size", "speaker_size", "channels", "device", "latent_dim", "beta",
                                         "mse_cof", "kl_cof", "style_cof"]
    m = DisentangledVAE(4, latent_dim=32)
    sd = O.synth_state_dict(0)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    for name in ("encode", "decode", "forward", "_reparameterize", "update_c", "update_beta"):
        assert callable(getattr(m, name))
    assert isinstance(m.postnet, Postnet) and issubclass(ConvolutionalMulVAE, VariationalBaseModelVAE)
    w = ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=8, speaker_size=4,
                            device=torch.device("cpu"), latent_dim=32)
    assert isinstance(w.optimizer, torch.optim.Adam) and w.kl_cof == 10 and w.mse_cof == 10
    for name in ("step", "train", "run_training", "load_last_model", "estimate_trained_model", "voice_conversion_mel",
                 "loss_functionGVAE2", "update_kl", "set_kl", "update_", "compute_KL_delta_VAE"):
        assert callable(getattr(w, name))
    c = chunking_mel(np.ones((80, 128), np.float32))
    assert tuple(c.shape) == (3, 80, 64) and float(c[2].abs().sum()) == 0.0


def test_no_cpu_fallback(built_lib):
    from model.disentangled_vae import DisentangledVAE
    m = DisentangledVAE(4, latent_dim=32)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.rand(2, 80, 64), torch.rand(2, 80, 64))
    product = []
    for d, _, files in os.walk(PKG):
        product += [os.path.join(d, f) for f in files if f.endswith((".py", ".cu", ".cuh", ".h"))]
    for f in product:
        assert "oracle" not in open(f).read().replace("# oracle", ""), f"{f} must not reference the oracle"


def test_speaker_group_sharding(built_lib):
    from dvae_b200.parallel import assign_groups_to_ranks, group_counts, shard_pairs_by_speaker
    ids = np.array([7, 7, 3, 3, 3, 9, 7, 1, 1, 9, 9, 9])
    gid, counts = group_counts(ids)
    assert gid.tolist() == [0, 0, 1, 1, 1, 2, 0, 3, 3, 2, 2, 2] and counts.tolist() == [3, 3, 4, 2]
    for world in (1, 2, 3, 4):
        shards = [shard_pairs_by_speaker(ids, r, world) for r in range(world)]
        allrows = np.concatenate(shards)
        assert sorted(allrows.tolist()) == list(range(len(ids)))           # a permutation: nothing lost or duplicated
        owners = {}
        for r, s in enumerate(shards):
            for row in s:
                owners.setdefault(int(ids[row]), set()).add(r)
        assert all(len(v) == 1 for v in owners.values())                    # whole speaker groups per rank
        assert all(len(s) > 0 for s in shards)
    assert assign_groups_to_ranks([8] * 32, 8) == [(4 * r, 4 * r + 4) for r in range(8)]   # BASELINE cfg 2: 32 spk x 8 utts


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, PKG)
    from dvae_b200.parallel import GradBuckets
    from oracle import dvae_oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    shapes = [(k, s) for k, s in O.param_shapes().items() if not ("lstm2" in k or "linear" in k)]   # keep it small
    gb = GradBuckets(shapes, "cpu")
    order = [k for k, _ in shapes][::-1]
    for k in order:
        gb.view(k).fill_(float(rank + 1) * (1 + len(k) % 3))
        gb.ready(k)
    gb.finish()
    ok = all(torch.allclose(gb.view(k), torch.full_like(gb.view(k), (1 + len(k) % 3) * (world + 1) / 2.0)) for k in order)
    out.put((rank, ok, gb.payload_bytes()))
    dist.destroy_process_group()


def test_gradient_buckets_allreduce_gloo(built_lib):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def test_bucket_plan_covers_all_parameters(built_lib):
    from dvae_b200.parallel import BUCKET_PREFIXES, GradBuckets
    from oracle import dvae_oracle as O
    shapes = list(O.param_shapes().items())
    gb = GradBuckets(shapes, "cpu")
    assert len(gb.slots) == 84 and gb.payload_bytes() >= 4 * 61367680
    assert all(len(m) > 0 for m in gb.members) and len(gb.members) == len(BUCKET_PREFIXES)
    v = gb.view("postnet.convolutions.0.0.conv.weight")
    assert tuple(v.shape) == (512, 80, 5) and all(off % 64 == 0 for _, off, _ in gb.slots.values())   # 256-byte slots


def test_speaker_group_batch_sampler(built_lib):
    """Per step the ranks' index lists partition the global batch into EQUAL shards of whole speakers, rows come sorted by
    speaker, and all ranks agree on the plan (same seed / epoch)."""
    import numpy as np
    from dvae_b200.data import SpeakerGroupBatchSampler
    ids = np.repeat(np.arange(12), 8)                       # 12 speakers x 8 utterances
    rng = np.random.default_rng(0)
    ids = ids[rng.permutation(len(ids))]
    world, ppr = 4, 8                                       # global batch 32
    samplers = [SpeakerGroupBatchSampler(ids, ppr, r, world, shuffle=True, seed=7) for r in range(world)]
    for s in samplers:
        s.set_epoch(3)
    steps = [list(iter(s)) for s in samplers]
    assert all(len(st) == len(ids) // (world * ppr) == len(samplers[0]) for st in steps)
    seen = []
    for b in range(len(steps[0])):
        union = sorted(i for r in range(world) for i in steps[r][b])
        assert len(set(union)) == len(union) == world * ppr
        seen += union
        owners = {}
        for r in range(world):
            assert len(steps[r][b]) == ppr                 # equal shards, every step
            spk = ids[steps[r][b]]
            assert list(spk) == sorted(spk)
            for sp in set(spk.tolist()):
                assert owners.setdefault(sp, r) == r       # a speaker's rows of this batch are on one rank
    assert len(set(seen)) == len(seen)                      # no index is used twice in an epoch
    again = list(iter(samplers[1]))
    assert again == steps[1]                                # deterministic for a fixed (seed, epoch)
    samplers[1].set_epoch(4)
    assert list(iter(samplers[1])) != steps[1]


def test_speaker_group_batch_sampler_uneven_speakers(built_lib):
    """Speakers with different numbers of utterances (the advisor's scenario: 64 speakers, world 2 and 8): every rank still
    gets exactly pairs_per_rank rows per step, never an empty shard, and a step never splits a speaker."""
    import numpy as np
    from dvae_b200.data import SpeakerGroupBatchSampler
    rng = np.random.default_rng(1)
    counts = rng.integers(8, 20, size=64)
    ids = np.repeat(np.arange(64), counts)
    ids = ids[rng.permutation(len(ids))]
    for world in (2, 8):
        ppr = 32
        samplers = [SpeakerGroupBatchSampler(ids, ppr, r, world, seed=3) for r in range(world)]
        assert samplers[0].group_size == 8
        steps = [list(iter(s)) for s in samplers]
        assert len(steps[0]) >= 1 and len({len(st) for st in steps}) == 1
        for b in range(len(steps[0])):
            owners = {}
            for r in range(world):
                assert len(steps[r][b]) == ppr
                for sp in set(ids[steps[r][b]].tolist()):
                    assert owners.setdefault(sp, r) == r
    with pytest.raises(ValueError, match="distinct speakers"):
        SpeakerGroupBatchSampler(np.repeat(np.arange(3), 8), 8, 0, 4)          # 3 speakers cannot feed 4 ranks


def test_tile_helper(built_lib):
    """model/disentangled_vae.py:35-41 (dead helper): every slice along `dim` repeated n times in place."""
    from model.disentangled_vae import tile
    a = torch.arange(6.0).view(2, 3)
    ref_idx = np.concatenate([2 * np.arange(3) + i for i in range(2)])          # the reference's index construction
    assert torch.equal(tile(a, 0, 3), a.repeat(3, 1)[torch.from_numpy(ref_idx)])
    assert tuple(tile(a, 1, 2).shape) == (2, 6) and torch.equal(tile(a, 1, 2)[:, ::2], a)


def test_reference_train_py_call_sequence(built_lib):
    """The reference's own `train.py` (read from /root/reference when present: this container only) drives the drop-in
    modules unchanged: its argparse defaults feed its literal `ConvolutionalMulVAE(...)` constructor expression
    (train.py:89-93), and its literal `run_training(...)` / `voice_conversion_mel(...)` call expressions (:96-99, :105-108)
    bind to our methods' signatures.  (The calls themselves need a GPU and a dataset tree; the GPU suite runs the same
    sequence on a stand-in dataset, tests/test_model_gpu.py::test_trainer_loop_checkpoint_and_resume.)"""
    import ast
    import inspect
    path = "/root/reference/train.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not present on this box")
    from model.disentangled_vae import ConvolutionalMulVAE
    tree = ast.parse(open(path).read())
    ns = {"argparse": __import__("argparse"), "torch": torch, "os": os}
    get_parse = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_parse")
    exec(compile(ast.Module([get_parse], []), path, "exec"), ns)
    main = next(n for n in tree.body if isinstance(n, ast.If))
    parse = ns["get_parse"]()
    calls = {}
    for node in ast.walk(main):
        if isinstance(node, ast.Expr) and isinstance(node.value, ast.Call) and isinstance(node.value.func, ast.Attribute):
            f = node.value.func
            if f.attr == "add_argument" and getattr(f.value, "id", "") == "parse":
                eval(compile(ast.Expression(node.value), path, "eval"), {"parse": parse, "float": float, "str": str, "bool": bool})
            elif f.attr in ("run_training", "voice_conversion_mel"):
                calls[f.attr] = node.value
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Call) and getattr(node.value.func, "id", "") == "ConvolutionalMulVAE":
            calls["ctor"] = node.value
    assert set(calls) == {"ctor", "run_training", "voice_conversion_mel"}
    args = parse.parse_args(["--train", "true", "--latent-size=32", "--batch-size=8", "--speaker_size=4", "--lr=1e-4"])
    cpu_ctor = lambda *a, **k: ConvolutionalMulVAE(*a, device=torch.device("cpu"), **k)      # no GPU on this box
    vsc = eval(compile(ast.Expression(calls["ctor"]), path, "eval"), {"ConvolutionalMulVAE": cpu_ctor, "args": args})
    assert vsc.batch_size == 8 and vsc.latent_dim == 32 and vsc.model.speaker_size == 4 and vsc.model.latent_dim == 32
    assert vsc.kl_cof == args.kl_cof and vsc.mse_cof == args.mse_cof
    for name in ("run_training", "voice_conversion_mel"):
        call = calls[name]
        env = {"args": args, "train_loader": object()}
        pos = [eval(compile(ast.Expression(a), path, "eval"), env) for a in call.args]
        kw = {k.arg: eval(compile(ast.Expression(k.value), path, "eval"), env) for k in call.keywords}
        inspect.signature(getattr(vsc, name)).bind(*pos, **kw)          # raises TypeError if train.py's call does not fit
