"""Config C4 with the time axis split across the GPUs of one box: N = 1e6, SHO J = 30 (rank 60).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/scan_multi_gpu.py
Every rank holds the series, folds its own range, one NCCL all-gather exchanges the G composites (97 KB each), every rank
re-filters its range, one NCCL all-reduce adds the two partial sums.  Prints one JSON line (rank 0): device time of an
evaluation (library events from the start of the fold to the end of the re-filter, max over ranks) and the value against
the single-GPU scan."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
import torch.distributed as dist
import pioran_b200 as pb
from pioran_b200.parallel import scan_logl_sharded, torch_collectives
import workloads as wl

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = pb.Context(local)
t, y, s2, f_min, f_max = wl.make_series_fast(N, seed=4)
spec = pb.make_spec("SingleBendingPowerLaw", f_min, f_max, 30)
a, b, c, d = ctx.approx_coeffs(spec, np.array([[0.82, 0.01, 3.3, float(np.var(y))]]))
ser = ctx.upload_series(t, y, s2)
single = ctx.celerite_logl_scan(ser, a, b, c, d)[0] if rank == 0 else None
ag, ar = torch_collectives(device=torch.device("cuda", local)) if world > 1 else (None, None)
begin = lambda lo, hi: ctx.scan_range_begin(ser, a, b, c, d, lo, hi, max_prev=world)
ms, wall = [], []
for rep in range(5):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    val = scan_logl_sharded(begin, ctx.scan_range_end, N, rank=rank, world=world, all_gather=ag, all_reduce_sum=ar,
                            range_check=ctx.scan_range_check)
    torch.cuda.synchronize()
    wall.append((time.perf_counter() - t0) * 1e3)
    ms.append(ctx.last_kernel_ms())
dev = torch.tensor([float(np.median(ms[1:])), float(np.median(wall[1:]))], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dev, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": f"C4: one series of N={N}, SHO J=30 (rank 60), time axis split over {world} GPU(s)", "n_gpus": world,
                      "device_ms_max_over_ranks": float(dev[0]), "wall_ms_max_over_ranks": float(dev[1]), "logL": val,
                      "logL_single_gpu_scan": float(single), "rel_diff": abs(val - single) / abs(single)}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
