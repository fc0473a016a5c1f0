"""Multi-GPU check (torchrun): one cosmology sharded over ranks — strided ODE modes, all-reduce gather of the sources,
contiguous fine-k LOS slices, NCCL all-reduce of partial C_l — must reproduce the single-GPU result."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import symboltz.jl_b200 as sb
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
M = sb.ΛCDM(lmax=10)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
bg = sb.solvebg(prob)
ls = np.array([2, 3, 5, 10] + list(range(20, 2501, 20)))
jl = sb.SphericalBesselCache(ls, xcut=2e3 * bg.tau0 * 1.001)
ref = {d: sb.spectrum_cmb(["TT", "EE", "TE"], prob, jl, bgsol=bg, direct=d) for d in (False, True)}  # single-GPU (process group not yet initialised)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for d in (False, True):
    for rep in range(2):
        dist.barrier(); torch.cuda.synchronize(); t = time.time()
        Cl = sb.spectrum_cmb(["TT", "EE", "TE"], prob, jl, bgsol=bg, direct=d)
        torch.cuda.synchronize(); dt = time.time() - t
    err = np.abs(Cl / ref[d] - 1).max()
    if rank == 0:
        print(f"world={world} direct={d}: sharded vs single-GPU max rel diff {err:.2e}; wall {dt*1e3:.1f} ms")
    assert err < 1e-12, err
dist.destroy_process_group()
