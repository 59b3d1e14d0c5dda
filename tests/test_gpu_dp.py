"""2-GPU data-parallel check (skipped on a 1-GPU box): frames sharded over two ranks + ONE flat NCCL all-reduce gives
the same parameter gradients as one GPU over the whole batch, with the MSDeformAttn op on the path."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(device):
    from uninext_b200.modules.deformable_layers import DeformableStack
    from uninext_b200.workloads import CONFIGS, level_tensors
    cfg = CONFIGS["cfg1"]
    torch.manual_seed(5)
    model = DeformableStack(num_layers=2, num_queries=40, d_ffn=256).to(device)
    g = torch.Generator().manual_seed(6)
    src = torch.randn(4, cfg.S, 256, generator=g).to(device)
    pos = torch.randn(4, cfg.S, 256, generator=g).to(device)
    ss, lsi = level_tensors(cfg.shapes, device)
    return cfg, model, src, pos, ss, lsi


def _worker(rank, world, port, q, overlap):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from uninext_b200.dp import FlatGradBucket, shard_frames
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cfg, model, src, pos, ss, lsi = _build(dev)
        # overlap: the flat buffer is all-reduced in 256 KB slices from grad hooks while backward is still running
        bucket = FlatGradBucket(model.parameters(), overlap=overlap, slice_bytes=256 << 10)
        idx = list(shard_frames(src.shape[0], world, rank))
        for _ in range(2):                                      # second step: re-armed by zero_()
            bucket.zero_()
            out = model(src[idx], pos[idx], cfg.shapes, ss, lsi)
            out.square().mean().backward()
            bucket.finish()
        torch.cuda.synchronize()
        assert not overlap or bucket.n_slices > 4
        q.put((rank, bucket.flat.cpu().numpy()))                # by value: a tensor would travel as an fd served by this process
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("overlap", [False, True])
def test_two_gpu_gradients_equal_single_gpu(overlap):
    import torch.multiprocessing as mp
    from uninext_b200.dp import FlatGradBucket
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + 13 * int(overlap)) % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r: torch.from_numpy(a) for r, a in (q.get(timeout=300) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg, model, src, pos, ss, lsi = _build(torch.device("cuda", 0))
    bucket = FlatGradBucket(model.parameters())
    model(src, pos, cfg.shapes, ss, lsi).square().mean().backward()
    ref = bucket.flat.cpu()
    scale = ref.abs().max().item()
    assert torch.equal(got[0], got[1])
    assert (got[0] - ref).abs().max().item() < 2e-4 * scale
