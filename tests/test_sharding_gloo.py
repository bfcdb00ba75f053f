"""World-size-2 gloo test of the N>1 path's host logic on CPU: row sharding, the Reducer that backs the
library's all-reduce hook and the preprocessing statistics, and the identity the sample-sharded design rests on
(sum over shards of the per-shard partial moments == the unsharded moments)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import corex_oracle as oc
        from linearcorex_b200 import Reducer, shard_rows
        red = Reducer(True)
        assert red.world == world and red.rank == rank and red.backend == "gloo"
        N, n, m = 1001, 40, 4
        x = oc.latent_factor_data(N, n, m, seed=5, snr=1.5, snr_spread=0.3).astype(np.float64)
        x[::13, 3] = -1e6  # missing marker
        lo, hi = shard_rows(N, rank, world)
        xs = x[lo:hi]
        assert red.sum_scalar(hi - lo) == N
        # --- preprocessing statistics exactly as Corex.preprocess combines them --------------------
        obs = xs != -1e6
        ssum = torch.from_numpy(np.where(obs, xs, 0.0).sum(0))
        cnt = torch.from_numpy(obs.sum(0).astype(np.float64))
        red.sum_(ssum)
        red.sum_(cnt)
        mean = (ssum / cnt).numpy()
        sq = torch.from_numpy((np.where(obs, xs - mean, 0.0) ** 2).sum(0))
        red.sum_(sq)
        sd = np.sqrt(sq.numpy() / cnt.numpy()).clip(1e-10)
        xt_full, theta, n_obs = oc.standardize(x, 'standard', -1e6)
        np.testing.assert_allclose(mean, theta[0], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(sd, theta[1], rtol=1e-12)
        np.testing.assert_array_equal(cnt.numpy().astype(np.int64), n_obs)
        xt = np.where(obs, (xs - mean) / sd, 0.0)
        np.testing.assert_allclose(xt, xt_full[lo:hi], rtol=1e-10, atol=1e-12)
        # --- the per-pair exchange: [D | s] partials summed over ranks ------------------------------
        rng = np.random.RandomState(0)
        w = rng.randn(m, n) / 20
        y = xt.dot(w.T)
        buf = torch.from_numpy(np.concatenate([(xt.T.dot(y)).T.ravel(), (y * y).sum(0)]))
        red.sum_(buf)
        d_all = buf.numpy()[:m * n].reshape(m, n)
        s_all = buf.numpy()[m * n:]
        y_full = xt_full.dot(w.T)
        np.testing.assert_allclose(d_all, xt_full.T.dot(y_full).T, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(s_all, (y_full * y_full).sum(0), rtol=1e-12)
        # every rank holds identical bytes after the reduction -> replicated control flow stays in lock step
        digest = torch.tensor([float(np.frombuffer(buf.numpy().tobytes(), dtype=np.uint8).sum())], dtype=torch.float64)
        lo_d, hi_d = digest.clone(), digest.clone()
        dist.all_reduce(lo_d, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_d, op=dist.ReduceOp.MAX)
        assert lo_d.item() == hi_d.item()
        eps = 0.36
        rho = (1 - eps ** 2) * d_all / N + eps ** 2 * w
        want = oc.moments_ns(xt_full, w, eps, quick=True)
        np.testing.assert_allclose(rho, want["rho"], rtol=1e-11, atol=1e-13)
        uj = (1 - eps ** 2) * s_all / N + eps ** 2 * (w ** 2).sum(1)
        np.testing.assert_allclose(uj, want["uj"], rtol=1e-12)
        # --- the Gram route's one-off exchange: per-shard X~^T X~ / N summed over ranks; afterwards every rank evaluates
        #     _sig (linearcorex.py:196-213) and sum Y^2 / N from the same matrix with no further exchange ----------------
        g = torch.from_numpy(xt.T.dot(xt) / N)
        red.sum_(g)
        g = g.numpy()
        np.testing.assert_allclose(g, xt_full.T.dot(xt_full) / N, rtol=1e-11, atol=1e-13)
        d_gram = g.dot(w.T).T
        np.testing.assert_allclose(d_gram, d_all / N, rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose((w * d_gram).sum(1), s_all / N, rtol=1e-10)
        # collective decisions (streamed preparation, Gram route) are taken with max / min over the ranks
        assert red.max_scalar(rank) == world - 1 and red.min_scalar(rank + 1) == 1
        ret[rank] = "ok"
    except Exception as exc:  # surface the failure in the parent
        ret[rank] = "FAIL: %r" % (exc,)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)


def test_reducer_identity_single_rank():
    sys.path.insert(0, ROOT)
    from linearcorex_b200 import Reducer
    r = Reducer(None)
    t = torch.arange(4, dtype=torch.float64)
    assert r.world == 1 and r.sum_(t) is t and r.sum_scalar(5) == 5 and r.min_scalar(3) == 3 and r.max_scalar(3) == 3
