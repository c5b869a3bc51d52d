"""GPU suite: runtime behaviour of the ops -- CUDA-graph capture/replay, side streams, concurrent host
threads (the reference trains under nn.DataParallel, train_stereo.py:139), HostPipeline."""
import threading

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rsm():
    import realtime_stereo_matcher_b200 as m
    m.load_library()
    return m


def _step(rsm, l, r, cost):
    gw = rsm.groupwise_volume(l, r, 4, 12)
    cat = rsm.concat_volume(l, r, 12)
    vol = rsm.inner_product_volume(l, r, 12, mean=True)
    soft, amin, amax = rsm.regress(vol)
    disp = rsm.v4_head(cost, 48, 32, 64)
    return gw, cat, soft, amin, disp


def test_cuda_graph_capture_and_replay(rsm):
    """The ops never synchronise, allocate only through torch and launch on the current stream, so a whole
    step can be captured once and replayed (fixed-shape inference loop)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    l = torch.randn((2, 16, 8, 64), device="cuda", generator=g)
    r = torch.randn((2, 16, 8, 64), device="cuda", generator=g)
    cost = torch.randn((2, 12, 8, 16), device="cuda", generator=g) * 3
    eager = _step(rsm, l, r, cost)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        _step(rsm, l, r, cost)       # warm-up on the capture stream
    side.synchronize()
    with torch.cuda.graph(graph, stream=side):
        captured = _step(rsm, l, r, cost)
    # new inputs, same buffers -> replay must recompute
    l2, r2 = torch.randn_like(l), torch.randn_like(r)
    l.copy_(l2), r.copy_(r2)
    graph.replay()
    torch.cuda.synchronize()
    fresh = _step(rsm, l, r, cost)
    for a, b in zip(captured, fresh):
        assert torch.equal(a, b)
    assert not torch.equal(captured[1], eager[1])


def test_concurrent_host_threads(rsm):
    """Several host threads call the library at once (ctypes releases the GIL); results must match the
    single-threaded ones -- no hidden global state."""
    rng = np.random.default_rng(11)
    inputs = [(rng.standard_normal((1, 8, 6, 70)).astype(np.float32), rng.standard_normal((1, 8, 6, 70)).astype(np.float32))
              for _ in range(6)]
    want = [oracle.inner_product_volume(l, r, 9) for l, r in inputs]
    got = [None] * len(inputs)

    def work(i):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            l, r = torch.from_numpy(inputs[i][0]).cuda(), torch.from_numpy(inputs[i][1]).cuda()
            for _ in range(20):
                out = rsm.inner_product_volume(l, r, 9)
            got[i] = out.cpu().numpy()
        s.synchronize()

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(inputs))]
    [t.start() for t in threads]
    [t.join() for t in threads]
    for a, b in zip(got, want):
        np.testing.assert_allclose(a, b, atol=1e-4)


def test_host_pipeline_matches_serial(rsm):
    g = torch.Generator().manual_seed(2)
    batches = [(torch.randn((2, 8, 8, 64), generator=g).pin_memory(), torch.randn((2, 8, 8, 64), generator=g).pin_memory(),
                (torch.randn((2, 12, 8, 16), generator=g) * 3).pin_memory()) for _ in range(5)]

    def step(l, r, c):
        vol = rsm.inner_product_volume(l, r, 12, mean=True)
        return rsm.soft_argmax(vol).mean() + rsm.v4_head(c, 48, 32, 64)

    outs = [torch.empty((2, 32, 64)).pin_memory() for _ in range(2)]
    pipe = rsm.HostPipeline(step, depth=2)
    got = {}
    for i, buf in pipe.run(iter(batches), outs):
        got[i] = buf.clone()
    assert sorted(got) == list(range(5))
    for i, b in enumerate(batches):
        want = step(*(t.cuda() for t in b)).cpu()
        assert torch.equal(got[i], want)


def test_data_parallel_two_devices(rsm):
    """nn.DataParallel over two GPUs (skipped on a single-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")

    class M(torch.nn.Module):
        def forward(self, l, r):
            return rsm.soft_argmax(rsm.make_correlation_volume(l, r, 8))

    l, r = torch.randn((4, 8, 6, 40), device="cuda:0"), torch.randn((4, 8, 6, 40), device="cuda:0")
    out = torch.nn.DataParallel(M(), device_ids=[0, 1])(l, r)
    assert torch.equal(out, M()(l, r))
