"""
What the host side of one node can move: every rank copies pinned host buffers to / from its GPU at the same time -- H2D alone,
D2H alone, both at once on two streams (the shape of bench.py's `e2e`: 5.1 GB in, 10.4 GB out per step and GPU).  Names the limiter
of the end-to-end scaling curve independently of the scoring pipeline.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/bench_pcie.py [--numa 1]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numa", type=int, default=1)
    ap.add_argument("--gb", type=float, default=2.0)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import bench

    numa = bench.bind_to_gpu_numa_node(local) if a.numa else {"bound": False}
    import torch

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(a.gb * (1 << 30)) // 4
    hi, ho = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(2 * n, dtype=torch.float32).pin_memory()
    hi.fill_(1.0); ho.fill_(0.0)  # touch: pages exist on this rank's NUMA node
    di, do = torch.empty(n, dtype=torch.float32, device="cuda"), torch.ones(2 * n, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        if dist is not None:
            dist.barrier()
        return dt

    def h2d():
        with torch.cuda.stream(s1):
            di.copy_(hi, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)

    def both():
        h2d(); d2h()

    res = {"h2d_alone_gbs": n * 4 / timed(h2d) / 1e9, "d2h_alone_gbs": 2 * n * 4 / timed(d2h) / 1e9}
    dt = timed(both)
    res["both_h2d_gbs"], res["both_d2h_gbs"] = n * 4 / dt / 1e9, 2 * n * 4 / dt / 1e9
    t = torch.tensor([res[k] for k in sorted(res)], device="cuda", dtype=torch.float64)
    if dist is not None:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
    else:
        allv = [t]
    if rank == 0:
        per_rank = [[round(float(v), 1) for v in r] for r in allv]
        keys = sorted(res)
        print(json.dumps({"n_gpus": world, "numa_rank0": numa, "keys": keys, "per_rank_gbs": per_rank,
                          "sum_gbs": {k: round(sum(r[i] for r in per_rank), 1) for i, k in enumerate(keys)},
                          "note": "1:2 H2D:D2H byte ratio as in bench.py's e2e; all ranks copy at the same time"}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
