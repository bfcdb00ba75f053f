"""BASELINE.json configs[4]: hierarchical layers = 30,5,1 on synthetic N = 1M x n = 20k, rows sharded over the ranks.
Runs linearcorex_b200.fit_layers with a fixed iteration budget per annealing stage and prints one JSON line (rank 0).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_config5.py [--rows 1000000] [--max-iter 3]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (DeviceRows: block-seeded synthetic rows drawn on the device)


def main():
    import torch
    import torch.distributed as dist
    from linearcorex_b200 import fit_layers, shard_rows
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1000000)
    ap.add_argument("--vars", type=int, default=20000)
    ap.add_argument("--groups", type=int, default=30)
    ap.add_argument("--max-iter", type=int, default=3)
    ap.add_argument("--precision", default="fp64_split")
    ap.add_argument("--algorithm", default="auto", choices=["auto", "stream", "gram"])
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lo, hi = shard_rows(args.rows, rank, world)
    x = bench.DeviceRows(args.rows, args.vars, args.groups, lo, hi)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    models = fit_layers(x, [30, 5, 1], seed=0, max_iter=args.max_iter, tol=1e-12, precision=args.precision,
                        comm=True if world > 1 else None, stream_rows=32768, algorithm=args.algorithm)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        out = {"config": "layers=30,5,1 on synthetic N=%d x n=%d (%d planted groups), %d ranks, precision=%s, algorithm=%s, max_iter=%d per stage"
                         % (args.rows, args.vars, args.groups, world, args.precision, args.algorithm, args.max_iter),
               "seconds_total": dt,
               "layers": [{"n": int(m.nv), "m": int(m.m), "iterations": len(m.history["TC"]), "TC": float(m.tc),
                           "algorithm_used": m.algorithm_used, "precision_used": m.precision_used,
                           "timings": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in m.timings.items()},
                           "pure_clusters": bool(all(len(set(m.clusters()[g::args.groups])) == 1 for g in range(args.groups)))
                           if m.nv == args.vars else None} for m in models]}
        os.write(1, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
