"""BASELINE.json's named configurations C3 / C4 / C5 at FULL size with the truncated-run parity of
BASELINE.md section 5 against the UNMODIFIED reference on this box's host (C3: 20 CG steps, C5:
maxiter=50 at the fp32 tolerance 1e-4; C4: see tools/run_c4_reference.py), one process per GPU.

  python tools/run_configs_parity.py c5 c3 [--steps-c3 20 --steps-c5 50]            (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29631 tools/run_configs_parity.py c3                             (8 GPUs)
Writes one JSON document to stdout (rank 0)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--steps-c3", type=int, default=20)
    ap.add_argument("--steps-c5", type=int, default=50)
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    import torch
    import bench
    import bench_configs
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist = None
    rank = 0
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        rank = dist.get_rank()
    peak, src = bench.peaks()
    out = {"n_gpus": world, "peak_gbs": peak, "peak_source": src}
    steps = {"c3": a.steps_c3, "c5": a.steps_c5, "c4": 0}
    for c in a.configs:
        out[c] = bench_configs.run_device(c, peak, dist=dist, rank=rank, world=world,
                                          ref_steps=0 if a.no_parity else steps[c],
                                          log=lambda *m: sys.stderr.write(" ".join(map(str, m)) + "\n"))
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(out, indent=1))
    if dist is not None:
        torch.cuda.synchronize()
        dist.barrier()
        from krypy_b200 import dist as kd
        kd.shutdown()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
