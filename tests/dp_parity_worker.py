"""Worker of tests/test_dp_gpu.py (one process per GPU, launched with torch.distributed.run): data-parallel parity as
SURVEY.md 8(e) defines it.  Every rank runs the fp32 oracle on ITS shard (identical weights, its own inputs / noise) and
the drop-in model with bucketed gradients + overlapped NCCL all-reduce on the same shard.  After the backward:
  * each rank's forward outputs / loss terms match the oracle on that shard (north_star tolerance),
  * the post-all-reduce parameter gradients (identical on every rank) match the MEAN over ranks of the per-shard oracle
    gradients: cosine > 0.999 per tensor at matched branch decisions, and the free-decision number is printed,
  * the gradients really are identical on all ranks (the all-reduce covered every bucket, and the side-stream /
    communication-stream ordering of dvae_b200.engine + dvae_b200.parallel held).
Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))
import torch
import torch.distributed as dist


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    os.environ["DVAE_B200_PRECISION"] = precision
    from dvae_b200.engine import Engine
    from dvae_b200.parallel import GradBuckets
    from model.disentangled_vae import ConvolutionalMulVAE
    from oracle import dvae_oracle as O
    from oracle.parity import FWD_NAMES, LOSS_NAMES, grad_cosines
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = O.synth_state_dict(0)                                   # identical replicas
    x1, x2, eps = O.synth_inputs(R, seed=1234 + rank)            # every rank its own shard
    x1, x2, eps = x1.to(dev), x2.to(dev), [e.to(dev) for e in eps]
    w = ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=R, speaker_size=4, device=dev, latent_dim=32,
                            beta=0.1, mse_cof=10, kl_cof=10, style_cof=0.1)     # per-rank batch_size = shard size
    w.model.load_state_dict(sd)
    w.model.train()
    w.model._engine.buckets = GradBuckets([(n, tuple(p.shape)) for n, p in w.model.named_parameters()], dev)
    queue = list(eps)
    w.model.noise_hook = lambda shape: queue.pop(0)
    w.model._debug_keep_saved = True
    out = w.model(x1, x2)
    losses = w.loss_functionGVAE2(x1, x2, *out)
    losses[0].backward()
    torch.cuda.synchronize(dev)
    mine = {k: p.grad.detach().clone() for k, p in w.model.named_parameters()}
    # ---- per-shard oracle (free and matched decisions), averaged over ranks
    osd = O.clone_sd(sd, requires_grad=True, device=dev)
    o_out, o_losses, o_free = O.train_step(osd, x1, x2, eps, batch_size=R)
    decisions = Engine.discrete_decisions(w.model._last_saved, [t.detach() for t in out], x1, x2)
    osd2 = O.clone_sd(sd, requires_grad=True, device=dev)
    _, _, o_match = O.train_step(osd2, x1, x2, eps, batch_size=R, decisions=decisions)
    for grads in (o_free, o_match):
        for k in grads:
            g = grads[k].contiguous()
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
            grads[k] = g
    fwd = {n: (a - b).norm().item() / b.norm().item() for n, a, b in zip(FWD_NAMES, out, o_out)}
    loss = {n: abs(a.item() - b.item()) / abs(b.item()) for n, a, b in zip(LOSS_NAMES, losses, o_losses)}
    mm, mk, mg, _ = grad_cosines(mine, o_match)
    fm, fk, fg, _ = grad_cosines(mine, o_free)
    # ---- identical on every rank?
    spread = 0.0
    for k, g in mine.items():
        lo, hi = g.clone(), g.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        spread = max(spread, (hi - lo).abs().max().item())
    worst_fwd = torch.tensor([max(v for n, v in fwd.items() if not n.endswith("hat")), max(v for n, v in fwd.items() if n.endswith("hat")),
                              max(list(loss.values())[:5]), max(list(loss.values())[5:])], device=dev)
    dist.all_reduce(worst_fwd, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "precision": precision, "rows_per_call_per_rank": R,
                          "forward_rel_l2_max_8_non_hat": worst_fwd[0].item(), "forward_rel_l2_max_hat": worst_fwd[1].item(),
                          "loss_rel_max": worst_fwd[2].item(), "kl_rel_max": worst_fwd[3].item(),
                          "grad_cosine_matched": {"min": mm, "argmin": mk, "global": mg},
                          "grad_cosine_free": {"min": fm, "argmin": fk, "global": fg},
                          "max_abs_difference_between_ranks": spread}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
