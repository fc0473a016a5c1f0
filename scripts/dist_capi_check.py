"""torchrun check of the library-owned NCCL communicator (libsbc.so, `sbc_*`): the sharded single-cosmology C_l and the sharded
parameter sweep driven through `group=Communicator` must equal the torch.distributed path bit for bit.  The unique id travels over the
torch process group here; a Julia host would pass it through a file or MPI.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_capi_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import symboltz.jl_b200 as sb
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ids = [sb.Communicator.unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = sb.Communicator(rank, world, ids[0])
M = sb.ΛCDM(lmax=10)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
bg = sb.solvebg(prob)
ls = np.array([2, 3, 5, 10] + list(range(20, 2501, 20)))
jl = sb.SphericalBesselCache(ls, xcut=2e3 * bg.tau0 * 1.02)
out = {}
for name, group in (("torch.distributed", None), ("library communicator", comm)):
    for direct in (False, True):
        sb.spectrum_cmb(["TT", "EE", "TE"], prob, jl, bgsol=bg, direct=direct, group=group)
        dist.barrier(); torch.cuda.synchronize(); t = time.perf_counter()
        out[name, direct] = sb.spectrum_cmb(["TT", "EE", "TE"], prob, jl, bgsol=bg, direct=direct, group=group)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        if rank == 0:
            print(f"{name:22s} direct={direct}: {1e3 * dt:.1f} ms", flush=True)
same = all(np.array_equal(out["torch.distributed", d], out["library communicator", d]) for d in (False, True))
Mw = sb.w0waCDM(lmax=10)
pw = sb.CosmologyProblem(Mw, sb.parameters_Planck18(Mw))
rng = np.random.default_rng(0)
th = np.array([0.6736, 0.2645, -0.9, 0.1]) * (1 + 0.05 * (rng.random((8, 4)) - 0.5))
ks = sb.loggrid(1e-4, 1.0, length=32) / sb.k0
Pa = sb.spectrum_matter_sweep(pw, ["h", "Omega_c", "w0", "wa"], th, ks, chunk=4)
Pb = sb.spectrum_matter_sweep(pw, ["h", "Omega_c", "w0", "wa"], th, ks, chunk=4, group=comm)
if rank == 0:
    print(f"world={world}: C_l identical {same}; sweep identical {np.array_equal(Pa, Pb)}", flush=True)
assert same and np.array_equal(Pa, Pb)
comm.close()
dist.destroy_process_group()
