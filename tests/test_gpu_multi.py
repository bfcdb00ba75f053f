"""Two-rank NCCL run of the sample-sharded fit on a 2-GPU box (skipped on one GPU): each rank fits its own
row block; results must match the single-process reference golden at 1e-9 and be bit-identical across ranks."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret, peer_mode):
    os.environ["LCX_PEER_ALLREDUCE"] = peer_mode
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from conftest import load_golden
        from linearcorex_b200 import Corex, shard_rows
        out = {}
        for name, prec, algo in (("syn_400x300x10_f64", "fp64", "stream"), ("standard_missing_f64", "fp64", "stream"),
                                 ("syn_400x300x10_f64", "fp64_split", "stream"), ("syn_400x300x10_f64", "fp64_split", "gram"),
                                 ("standard_missing_f64", "fp64_split", "gram"), ("syn_4000x2000x20_f64", "fp64_split", "gram")):
            z, kw, x = load_golden(name)
            kw = dict(kw, precision=prec, algorithm=algo)
            lo, hi = shard_rows(x.shape[0], rank, world)
            mdl = Corex(comm=True, **kw).fit(x[lo:hi])
            assert mdl.n_samples == x.shape[0] and mdl.algorithm_used == algo
            # Gram route: the ranks' partial X~^T X~ / N are summed once (NCCL); with NVLink peers the per-iteration product is
            # sharded over the ranks' row tiles of the matrix and gathered in place, without them it is replicated
            assert (mdl._sess._peer_buf is not None) == (peer_mode == "require")
            assert len(mdl.history["TC"]) == len(z["history_TC"])
            err = np.abs(mdl.ws - z["ws"]).max() / np.abs(z["ws"]).max()
            err_tc = np.abs(mdl.moments["TCs"] - z["m_TCs"]).max() / np.abs(z["m_TCs"]).max()
            assert err < 1e-9 and err_tc < 1e-9, (name, err, err_tc)
            assert (mdl.clusters() == z["clusters"]).all()
            y = mdl.transform(x[lo:hi])
            assert np.abs(y - z["transform"][lo:hi]).max() < 1e-9 * np.abs(z["transform"]).max()
            # replicated state must be bit-identical on every rank
            t = torch.from_numpy(mdl.ws.copy()).cuda()
            lo_t, hi_t = t.clone(), t.clone()
            dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
            assert torch.equal(lo_t, hi_t)
            out[name + "/" + prec + "/" + algo] = float(err)
        ret[rank] = "ok %r" % (out,)
    except Exception as exc:
        import traceback
        ret[rank] = "FAIL: %s" % traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("peer_mode", ["require", "0"])
def test_two_rank_nccl_fit_matches_golden(peer_mode):
    """peer_mode 'require': the fused split-K-combine + all-reduce kernel over NVLink peer memory; '0': NCCL hook."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29700 + os.getpid() % 1000 + (7 if peer_mode == "0" else 0), ret, peer_mode), nprocs=2, join=True)
    assert all(str(v).startswith("ok") for v in ret.values()) and len(ret) == 2, dict(ret)
