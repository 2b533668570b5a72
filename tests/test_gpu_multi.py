"""GPU: host-resident batches over one or several contexts / devices of ONE process
(pgs_icp_run_batch_multi, LoopCloser.hpp:266-297), handles used across contexts and threads."""
import threading

import numpy as np
import pytest

from oracle import binding as ob
from pgslam_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pm(ctx):
    from pgslam_b200 import pm as _pm
    _pm._DEFAULT_CTX = ctx
    return _pm


def _host(a):
    return np.ascontiguousarray(a.T)


def _pairs(n, seed0=40):
    return [synth.scan_pair(seed0 + i, beams=8 + 2 * (i % 3), az_steps=500 + 20 * i)[:2] for i in range(n)]


def _bits(rec):
    return rec.tobytes()


def _n_devices():
    import torch
    return torch.cuda.device_count()


def test_batch_multi_equals_single_runs_bit_for_bit(pm, ctx):
    """13 ragged pairs from HOST memory through two contexts of the same GPU: every record equals
    pgs_icp_run's for that pair, bit for bit, and the oracle's within the contract."""
    pairs = _pairs(13)
    ctx2 = pm.Context(0)
    icps = []
    for c in (ctx, ctx2):
        i = pm.ICP(c)
        i.loadFromYaml(util.to_yaml(util.C2_COV))
        icps.append(i)
    rd = pm.host_clouds([_host(r) for r, _ in pairs])
    rf = pm.host_clouds([_host(f) for _, f in pairs])
    got = pm.compute_batch_multi(icps, rd, rf)
    assert got.shape == (13,)
    one = pm.ICP(ctx)
    one.loadFromYaml(util.to_yaml(util.C2_COV))
    for k, (r, f) in enumerate(pairs):
        T = one(pm.DataPoints(r, ctx=ctx), pm.DataPoints(f, ctx=ctx))
        assert np.array_equal(got[k]["T"].reshape(4, 4).T, T)
        assert got[k]["iterations"] == one.last["iterations"]
        assert np.array_equal(got[k]["covariance"].reshape(6, 6).T, one.last["covariance"])
        assert got[k]["residual"] == one.last["residual"]
    for k in (0, 7, 12):
        want = ob.icp_run(util.C2_COV, ob.Cloud(pairs[k][0]), ob.Cloud(pairs[k][1]))
        assert got[k]["iterations"] == want["iterations"]
        util.assert_pose_close(got[k]["T"].reshape(4, 4).T, want["T"])
    # one context, pinned-style path flag off vs. two contexts: same records
    again = pm.compute_batch_multi(icps[:1], rd, rf)
    assert _bits(again) == _bits(got)


def test_batch_multi_with_initial_guesses_and_descriptors(pm, ctx, tmp_path):
    rd, rf, truth = synth.scan_pair(3, beams=16, az_steps=700)
    oc = ob.Cloud(rd)
    for it in util.INPUT_FILTERS:
        (name, p), = ob._modlist([it])
        ob.apply_filter(oc, name, **p)
    desc = {k: np.ascontiguousarray(v.T) for k, v in oc.descriptors().items()}
    T0 = synth.pose_matrix(truth[:3, 3] + [0.03, -0.02, 0.0], 0.005, 0.0, 0.0)
    T0[:3, :3] = truth[:3, :3] @ T0[:3, :3]
    icp = pm.ICP(ctx)
    icp.loadFromYaml(util.to_yaml(util.C2))
    n = 9
    got = pm.compute_batch_multi([icp], pm.host_clouds([_host(rd)] * n, [desc] * n), pm.host_clouds([_host(rf)] * n),
                                 T_inits=[T0] * n)
    want = ob.icp_run(util.C2, ob.Cloud(rd, oc.descriptors()), ob.Cloud(rf), T0)
    for k in range(n):
        assert got[k]["iterations"] == want["iterations"]
        util.assert_pose_close(got[k]["T"].reshape(4, 4).T, want["T"])
        assert got[k]["overlap"] == pytest.approx(want["overlap"], abs=1e-12)  # sensor-noise overlap needs the descriptors
        assert _bits(got[k]) == _bits(got[0])


def test_batch_multi_failures_and_argument_checks(pm, ctx):
    icp = pm.ICP(ctx)
    icp.loadFromYaml(util.to_yaml(util.C2))
    a, b = _pairs(1)[0]
    empty = np.ones((0, 4), np.float32)
    rd = pm.host_clouds([_host(a), empty, _host(a), _host(a), _host(a), _host(a), _host(a), _host(a)])
    rf = pm.host_clouds([_host(b)] * 8)
    got = pm.compute_batch_multi([icp], rd, rf)
    assert got[1]["status"] == pm.CONVERGENCE_ERROR and got[0]["status"] == 0 and got[7]["status"] == 0
    assert _bits(got[0]) == _bits(got[7])
    with pytest.raises(pm.PointMatcherError):
        pm.compute_batch_multi([icp, icp], rd, rf)  # two handles on one context
    assert pm.compute_batch_multi([icp], pm.host_clouds([]), pm.host_clouds([])).shape == (0,)


@pytest.mark.skipif("_n_devices() < 2")
def test_batch_multi_two_devices(pm, ctx):
    """The real thing when the box has it: one process, two GPUs, results equal device 0's."""
    pairs = _pairs(10)
    ctx1 = pm.Context(1)
    icps = []
    for c in (ctx, ctx1):
        i = pm.ICP(c)
        i.loadFromYaml(util.to_yaml(util.C2))
        icps.append(i)
    rd = pm.host_clouds([_host(r) for r, _ in pairs])
    rf = pm.host_clouds([_host(f) for _, f in pairs])
    two = pm.compute_batch_multi(icps, rd, rf)
    one = pm.compute_batch_multi(icps[:1], rd, rf)
    assert _bits(two) == _bits(one)
    # a handle of device 1 driven from a thread whose current device is 0 (and vice versa)
    out = {}

    def run(i, dev_icp, dev_ctx):
        r, f = pairs[i]
        out[i] = dev_icp(pm.DataPoints(r, ctx=dev_ctx), pm.DataPoints(f, ctx=dev_ctx))
    ts = [threading.Thread(target=run, args=(0, icps[1], ctx1)), threading.Thread(target=run, args=(1, icps[0], ctx))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert np.array_equal(out[0], one[0]["T"].reshape(4, 4).T) and np.array_equal(out[1], one[1]["T"].reshape(4, 4).T)
    with pytest.raises(pm.PointMatcherError):  # a cloud of device 0 handed to a handle of device 1
        icps[1](pm.DataPoints(pairs[0][0], ctx=ctx), pm.DataPoints(pairs[0][1], ctx=ctx1))


def test_clouds_of_another_context_are_joined_not_raced(pm, ctx):
    """A cloud uploaded (asynchronously, from pinned memory) in context B and consumed at once by
    an ICP handle of context A: the consumer's stream waits for B's upload, and B may free the cloud
    right after the call."""
    import torch
    rd, rf, _ = synth.scan_pair(5, beams=16, az_steps=900)
    want = ob.icp_run(util.C2, ob.Cloud(rd), ob.Cloud(rf))
    ctx_b = pm.Context(0)
    icp = pm.ICP(ctx)
    icp.loadFromYaml(util.to_yaml(util.C2))
    hrd = torch.from_numpy(_host(rd)).pin_memory()
    hrf = torch.from_numpy(_host(rf)).pin_memory()
    for _ in range(5):
        a = pm.DataPoints(ctx=ctx_b, pinned_host_ptr=hrd.data_ptr(), n=hrd.shape[0])
        b = pm.DataPoints(ctx=ctx_b, pinned_host_ptr=hrf.data_ptr(), n=hrf.shape[0])
        T = icp(a, b)
        del a, b  # freed on B's stream, which was made to wait for A's reads
        assert icp.last["iterations"] == want["iterations"]
        util.assert_pose_close(T, want["T"])


def test_handle_used_from_a_thread_with_another_current_device(pm, ctx):
    """PGS_API_BEGIN makes the handle's device current whatever the calling thread had selected."""
    import torch
    rd, rf, _ = synth.scan_pair(6, beams=8, az_steps=600)
    icp = pm.ICP(ctx)
    icp.loadFromYaml(util.to_yaml(util.C1))
    ref = icp(pm.DataPoints(rd, ctx=ctx), pm.DataPoints(rf, ctx=ctx))
    out = []

    def worker():
        if torch.cuda.device_count() > 1:
            torch.cuda.set_device(1)
        out.append(icp(pm.DataPoints(rd, ctx=ctx), pm.DataPoints(rf, ctx=ctx)))
    t = threading.Thread(target=worker)
    t.start()
    t.join()
    assert np.array_equal(out[0], ref)
