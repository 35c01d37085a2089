"""world_size-2 gloo test of the token-shard reduction logic: reducing the sums
record and the valid-row count across ranks, then normalising, equals the single-process
reference on the concatenated batch.  CPU only (the oracle stands in for the kernels)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kd_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import speech_distill_b200.dist as D

    g = torch.Generator().manual_seed(100)
    B, T, V = 4, 7, 61
    z = torch.randn(B, T, V, generator=g, dtype=torch.float64)
    y = torch.randn(B, T, V, generator=g, dtype=torch.float64)
    lab = torch.randint(0, V, (B, T), generator=g)
    lab[0, :5] = -100  # unequal valid counts per rank
    sl = slice(rank * B // world, (rank + 1) * B // world)
    r = O.closed_form(z[sl].numpy(), lab[sl].numpy(), teacher_logits=y[sl].numpy(), temperature=2.0, alpha=0.5)
    reduce_fn, count_fn = D.make_reduce_fns()
    sums = reduce_fn(torch.from_numpy(r["sums"]))
    n = count_fn(torch.tensor([int(r["sums"][3])], dtype=torch.int32))
    losses = D.losses_from_sums(sums, 2.0, 0.5)
    # local gradient renormalised by the global N, as the kernels do with n_norm
    grad = torch.from_numpy(r["grad"]) * (r["sums"][3] / float(n))
    gw = torch.zeros(B, T, V, dtype=torch.float64)
    gw[sl] = grad
    D.allreduce_grad_(gw, bucket_rows=1)
    if rank == 0:
        out_q.put(([float(x) for x in losses], int(n), gw.numpy()))
    dist.destroy_process_group()


def test_token_shard_reduction_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    losses, n, grad = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(100)
    B, T, V = 4, 7, 61
    z = torch.randn(B, T, V, generator=g, dtype=torch.float64)
    y = torch.randn(B, T, V, generator=g, dtype=torch.float64)
    lab = torch.randint(0, V, (B, T), generator=g)
    lab[0, :5] = -100
    (ref, gref) = O.reference_loss_and_grad(z, lab, teacher_logits=y, temperature=2.0, alpha=0.5)
    assert n == int((lab[:, 1:] != -100).sum())
    np.testing.assert_allclose(losses, [float(x) for x in ref], rtol=1e-12)
    np.testing.assert_allclose(grad, gref.numpy(), rtol=1e-9, atol=1e-15)
