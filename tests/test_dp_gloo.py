"""CPU, world_size 2, gloo: frames are sharded across ranks, ONE flat all-reduce combines gradients, and the result
equals the single-process gradient over the whole batch (the check the reference never had)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4))


def _data():
    g = torch.Generator().manual_seed(11)
    return torch.randn(6, 16, generator=g), torch.randn(6, 4, generator=g)


def _worker(rank, world, port, q, overlap):
    sys.path.insert(0, ROOT)
    from uninext_b200.dp import FlatGradBucket, shard_frames
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model()
        # overlap: 3 slices (slice_bytes below every tensor size cuts at each parameter boundary) reduced from hooks
        bucket = FlatGradBucket(model.parameters(), overlap=overlap, slice_bytes=600)
        x, y = _data()
        idx = list(shard_frames(x.shape[0], world, rank))
        for step in range(2):                                   # second step: re-armed by zero_()
            bucket.zero_()
            if step == 1:
                model.zero_grad(set_to_none=True)               # breaks the aliasing; must be repaired, not ignored
            loss = ((model(x[idx]) - y[idx]) ** 2).mean()      # per-rank mean over its frames
            loss.backward()
            bucket.finish()
        before = dist.get_world_size()
        # by value (a list), not a tensor: a tensor travels as a file descriptor served by THIS process, which may have
        # exited by the time the parent rebuilds it
        q.put((rank, idx, bucket.flat.tolist(), before, bucket.n_slices))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True])
def test_two_rank_flat_allreduce_equals_single_process_gradient(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * int(overlap)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    got = [(r, idx, torch.tensor(flat, dtype=torch.float32), w, n) for r, idx, flat, w, n in got]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from uninext_b200.dp import FlatGradBucket
    model = _model()
    bucket = FlatGradBucket(model.parameters())
    x, y = _data()
    ((model(x) - y) ** 2).mean().backward()                    # equal shard sizes: mean of means == global mean
    shards = sorted(i for _, idx, _, _, _ in got for i in idx)
    assert shards == list(range(6))                             # every frame on exactly one rank
    for _, _, flat, world, n_slices in got:
        assert world == 2 and n_slices >= 3
        assert torch.allclose(flat, bucket.flat, atol=1e-6)
    assert torch.equal(got[0][2], got[1][2])                    # ranks agree bit for bit after the all-reduce


def test_bucket_views_and_single_process_noop():
    from uninext_b200.dp import FlatGradBucket, shard_frames
    model = _model()
    b = FlatGradBucket(model.parameters())
    assert b.flat.numel() == sum(p.numel() for p in model.parameters())
    x, y = _data()
    ((model(x) - y) ** 2).mean().backward()
    assert all(p.grad.data_ptr() >= b.flat.data_ptr() for p in model.parameters())
    assert b.all_reduce_mean() is None and b.flat.abs().sum() > 0
    assert list(shard_frames(5, 2, 1)) == [1, 3]
    with pytest.raises(ValueError):
        FlatGradBucket([])


def test_broken_aliasing_is_detected_and_repaired():
    """optimizer.zero_grad() / module.zero_grad() default to set_to_none=True, which drops the views into the flat
    buffer; the next backward then allocates fresh gradients.  all_reduce_mean() must reduce THOSE, not stale zeros."""
    from uninext_b200.dp import FlatGradBucket
    model = _model()
    b = FlatGradBucket(model.parameters())
    x, y = _data()
    model.zero_grad(set_to_none=True)
    assert all(p.grad is None for p in model.parameters())
    ((model(x) - y) ** 2).mean().backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    assert b.flat.abs().sum() == 0                              # the buffer knows nothing of the fresh gradients yet
    assert b.check_views() == 4
    assert torch.equal(b.flat, want)
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(b.params, b._views)) and b.check_views() == 0
