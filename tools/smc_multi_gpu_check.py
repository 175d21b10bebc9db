"""The whole SMC loop (smc.run_smc_anneal_data) under torchrun with the NCCL backend: every rank ends with the same
particles, and they are the single-GPU run's particles (GPU box, N >= 2):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/smc_multi_gpu_check.py
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import autogp.jl_b200 as agp  # noqa: E402
from autogp.jl_b200 import smc, tree_moves as tm  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, P = 300, 24
rng = np.random.default_rng(4)
ts = rng.permutation(np.arange(n) / (n - 1))
xs = 0.8 * np.sin(2 * np.pi * ts / 0.25) + 0.5 * ts + 0.05 * rng.standard_normal(n)
eng = agp.Engine(local)
kw = dict(config=tm.GPConfig(max_depth=3), n_particles=P, n_mcmc=4, n_hmc=2, schedule=smc.linear_schedule(n, 0.25), seed=5, engine=eng)
state = smc.run_smc_anneal_data(ts, xs, **kw)
digest = hashlib.sha256((repr(state.nodes) + repr(state.noises) + repr(state.scores.tolist()) + repr(state.log_weights.tolist())).encode()).hexdigest()
if world > 1:
    got = [None] * world
    dist.all_gather_object(got, digest)
    assert len(set(got)) == 1, got
    dist.barrier()
    dist.destroy_process_group()          # the single-GPU run below must not shard
    os.environ["WORLD_SIZE"] = "1"
if rank == 0:
    single = smc.run_smc_anneal_data(ts, xs, **kw)
    d1 = hashlib.sha256((repr(single.nodes) + repr(single.noises) + repr(single.scores.tolist()) + repr(single.log_weights.tolist())).encode()).hexdigest()
    same = d1 == digest
    dl = float(np.max(np.abs(single.scores - state.scores)))
    print(f"world {world}: ranks agree; against the single-GPU run: bitwise {'equal' if same else 'DIFFERENT'}, max |dLML| {dl:.3e}, "
          f"kernels equal: {repr(single.nodes) == repr(state.nodes)}", flush=True)
