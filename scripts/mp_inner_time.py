"""Inner-solve time on several GPUs (torchrun): config C2 (or --dim 3), every rank times the same sequence."""
import json, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthetic import problems
from control_b200 import MultiBlockSystem
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2
q = problems.heat_problem(nx, 64, True) if dim == 2 else problems.heat_problem_3d(nx, 32, False)
s = MultiBlockSystem(q["M"], q["K"], n_t=q["n_t"], beta=q["beta"], CN=(dim == 2), time_interval=q["time_interval"],
                     bc_dofs=q["bdofs"], device=lr, rank=rank, world=world)
if world > 1:
    s.init_comm(dist)
t0 = time.perf_counter()
s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"])
setup = time.perf_counter() - t0
a = s.micro_benchmarks(reps=30, flush_l2=False)
if rank == 0:
    print(json.dumps({"world": world, "nx": nx, "dim": dim, "setup_s": round(setup, 2), "inner_solve_ms": round(a["inner_solve_ms"], 4),
                      "kernels": a["inner_solve_kernels"], "cheb_warm_us": round(a["cheb_ms"] * 1e3, 2),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("CTL_")}}), flush=True)
s.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
