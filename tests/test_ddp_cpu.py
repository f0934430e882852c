"""CPU, world_size 2 over gloo: the data-parallel step (flat gradient buckets,
mean all-reduce, identical AdamW) equals one process on the global batch --
the semantics of Lightning DDP (train_model.py:276-286).  The model here is
the CPU oracle (the product kernels need a GPU); the trainer is product code."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import build_model_case, load_golden

CASE = load_golden("models.pt")["graphlam_dummy_d8"]


def _make(root):
    from oracle import port
    ds, args, batch = build_model_case(CASE["case"], root)
    model = port.GraphLAM(args, None, ds)
    model.load_state_dict(CASE["state_dict"])
    return model, batch


def _worker(rank, world, port_no, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from neural_lam_b200 import train
    r, w, dev = train.init_distributed()
    assert (r, w, dev.type) == (rank, world, "cpu")
    with tempfile.TemporaryDirectory() as root:
        model, batch = _make(root)
    if rank == 1:  # broadcast from rank 0 must overwrite this
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    trainer = train.DataParallelTrainer(model, rank, world)
    shard = tuple(t[rank:rank + 1] for t in batch)  # B=2 global -> 1 per rank
    losses = [trainer.step(shard).item() for _ in range(2)]
    torch.save({"losses": losses, "params": [p.detach().clone() for p in model.parameters()]},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_equals_global_batch():
    from neural_lam_b200 import train
    with tempfile.TemporaryDirectory() as out_dir:
        port_no = 29000 + os.getpid() % 2000
        mp.spawn(_worker, args=(2, port_no, out_dir), nprocs=2, join=True)
        r0 = torch.load(os.path.join(out_dir, "rank0.pt"))
        r1 = torch.load(os.path.join(out_dir, "rank1.pt"))
    # ranks stay in lock-step
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)
    # single process on the whole batch
    torch.set_num_threads(4)
    with tempfile.TemporaryDirectory() as root:
        model, batch = _make(root)
    single = train.DataParallelTrainer(model, 0, 1)
    losses = [single.step(batch).item() for _ in range(2)]
    mean_loss = [(a + b) / 2 for a, b in zip(r0["losses"], r1["losses"])]
    assert mean_loss == pytest.approx(losses, rel=1e-5)
    for a, b in zip(r0["params"], model.parameters()):
        torch.testing.assert_close(a, b.detach(), rtol=1e-4, atol=1e-6)


def test_flat_buckets_layout():
    from neural_lam_b200 import train
    lin = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 2))
    fb = train.FlatGradBuckets([(f"processor.{n}" if n.startswith("1") else n, p)
                                for n, p in lin.named_parameters()], 1, ("processor",))
    assert fb.flat.numel() == sum(p.numel() for p in lin.parameters())
    assert fb.n_early == 4 * 2 + 2  # second Linear first
    lin(torch.randn(5, 3)).sum().backward()
    assert fb.flat.abs().sum() > 0
    for p in lin.parameters():
        assert p.grad.data_ptr() >= fb.flat.data_ptr()  # still views of the flat buffer
    fb.zero()
    assert all(float(p.grad.abs().sum()) == 0 for p in lin.parameters())


def test_early_bucket_waits_for_every_ar_step():
    """An unrolled rollout registers one encoder-output hook per AR step; backward visits
    them last step first, and the decoder / processor gradients of the earlier steps are
    written AFTER the later steps' hooks fire.  The early all-reduce may only start with
    the last hook (= first AR step)."""
    from neural_lam_b200 import train
    lin = torch.nn.Linear(3, 2)
    fb = train.FlatGradBuckets([("processor.w", lin.weight), ("enc.b", lin.bias)], 1, ("processor",))
    launched = []
    fb.overlap = True  # the collective itself is replaced below (needs CUDA + NCCL)
    fb._launch_early = lambda: launched.append(fb._enc_pending)
    fb.zero()
    for _ in range(3):  # forward of a 3-step rollout
        fb.expect_encoder_output()
    for step in range(3):  # backward
        fb.early_ready(None)
        assert len(launched) == (1 if step == 2 else 0)
    assert launched == [0]
    fb.zero()
    fb.expect_encoder_output()
    fb.early_ready(None)
    assert len(launched) == 2
