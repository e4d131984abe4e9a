"""Failure detection: a dead / stuck peer becomes a RuntimeError naming the peer (device-side
timeout -> host-mapped status word -> Work / check_health), abort() releases spinning kernels,
and the plan cache survives trimming. (The reference hangs forever in its busy-poll loops,
SURVEY.md §5; its only error path is MPI_CHECK -> finishWorkMPIError,
/root/reference/src/ProcessGroupCGX.cc:120-123, :312-317.)"""
import os
import time

import pytest
import torch
import torch.distributed as dist

import torch_cgx_b200 as cgx
from _dist_utils import spawn

pytestmark = pytest.mark.gpu
C = cgx._C


def dev():
    return torch.device("cuda", 0)


def test_device_timeout_names_the_missing_peer():
    g = C.LocalSraGroup(2, 4, 1 << 20, 200, 256)  # 200 ms device-side timeout
    xs = [torch.randn(10_000, device=dev()) for _ in range(2)]
    torch.cuda.synchronize()
    t0 = time.time()
    g.allreduce_with_absent_rank(xs, [(0, 10_000, 4, 512)], 1)  # virtual rank 1 never shows up
    torch.cuda.synchronize()
    dt = time.time() - t0
    assert 0.15 < dt < 2.0, dt
    with pytest.raises(RuntimeError, match=r"rank 0 waiting for rank 1"):
        g.check()
    with pytest.raises(RuntimeError):  # sticky: the heap is out of step with its peers
        g.check()
    assert "timed out" in g.status(0)


def test_abort_releases_a_spinning_kernel():
    g = C.LocalSraGroup(2, 4, 1 << 20, 120_000, 256)  # would spin for two minutes
    xs = [torch.randn(10_000, device=dev()) for _ in range(2)]
    torch.cuda.synchronize()
    g.allreduce_with_absent_rank(xs, [(0, 10_000, 4, 512)], 0)
    time.sleep(0.2)
    t0 = time.time()
    g.abort_all()
    torch.cuda.synchronize()
    assert time.time() - t0 < 5.0
    assert "aborted" in g.status(1)
    with pytest.raises(RuntimeError, match="aborted"):
        g.check()


def test_plan_cache_trim_keeps_results_correct():
    # > kMaxCachedPlans (1024) distinct layouts on one group: the cache is dropped and rebuilt
    g = C.LocalSraGroup(1, 4, 1 << 20, 5000, 256)
    x0 = torch.randn(40_000, device=dev())
    for i in range(1100):
        n = 1000 + i * 8
        x = x0[:n].clone()
        g.allreduce([x], [(0, n, 4, 512)])
    torch.cuda.synchronize()
    g.check()
    n = 30_000
    x = x0[:n].clone()
    g.allreduce([x], [(0, n, 4, 512)])
    torch.cuda.synchronize()
    assert torch.equal(x.cpu(), cgx.ops.fake_quantize(x0[:n].cpu(), 4, 512))


def _absent_peer(rank, world):
    torch.cuda.set_device(rank)
    dist.init_process_group("cgx", init_method="env://", rank=rank, world_size=world)
    be = cgx.get_backend()
    dev_ = torch.device("cuda", rank)
    os.environ["CGX_COMPRESSION_QUANTIZATION_BITS"] = "4"
    x = torch.randn(1 << 20, device=dev_)
    dist.all_reduce(x)  # both ranks: heap connected, kernels warmed up
    torch.cuda.synchronize()
    if rank == 1:
        time.sleep(3.0)  # skips the next collective
        os._exit(0)
    t0 = time.time()
    work = dist.all_reduce(x, async_op=True)
    with pytest.raises(RuntimeError, match=r"waiting for rank 1"):
        work.wait()  # CGX_BLOCKING_WAIT=1: the host blocks until the kernel gave up
    dt = time.time() - t0
    assert dt < 1.5, dt
    assert not work.is_success()
    assert "rank 1" in be.failure()
    # the group now refuses new work instead of hanging on the dead peer
    w2 = dist.all_reduce(x, async_op=True)
    with pytest.raises(RuntimeError):
        w2.wait()
    with pytest.raises(RuntimeError):
        be.check_health()
    be.abort()
    os._exit(0)


@pytest.mark.multigpu
def test_backend_reports_an_absent_peer_within_a_second():
    spawn(_absent_peer, 2, env={"CGX_TIMEOUT_MS": 200, "CGX_BLOCKING_WAIT": 1}, timeout=120)
