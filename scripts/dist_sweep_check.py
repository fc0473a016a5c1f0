"""Multi-GPU check (torchrun) of the sharded parameter sweep (config 4): cosmologies strided over ranks, one all-reduce gather of P(k);
must reproduce the single-process sweep bit for bit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import symboltz.jl_b200 as sb
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = sb.w0waCDM(lmax=10)
prob = sb.CosmologyProblem(M, sb.parameters_Planck18(M))
names = ["h", "Omega_c", "Omega_b", "ln_As1e10", "ns", "w0", "wa"]
lo = np.array([0.6, 0.10, 0.020, 2.9, 0.92, -1.2, -0.3]); hi = np.array([0.8, 0.14, 0.025, 3.2, 1.0, -0.8, 0.3])
rng = np.random.default_rng(0)
u = (rng.permuted(np.tile(np.arange(nc), (7, 1)), axis=1).T + rng.random((nc, 7))) / nc
th = lo + (hi - lo) * u
th[:, 1] /= th[:, 0] ** 2; th[:, 2] /= th[:, 0] ** 2
ks = sb.loggrid(1e-4, 1.0, length=256) / sb.k0
sb.spectrum_matter_sweep(prob, names, th[:4], ks)
torch.cuda.synchronize(); t = time.time(); ref, iref = sb.spectrum_matter_sweep(prob, names, th, ks, chunk=32, return_info=True); t1 = time.time() - t  # every rank alone
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
sb.spectrum_matter_sweep(prob, names, th[:2 * world], ks)
dist.barrier(); torch.cuda.synchronize(); t = time.time()
P, info = sb.spectrum_matter_sweep(prob, names, th, ks, chunk=32, return_info=True)
torch.cuda.synchronize(); dist.barrier(); tn = time.time() - t
same = np.array_equal(P, ref, equal_nan=True)
if rank == 0:
    print(f"world={world}: {nc} cosmologies x 256 modes: single process {t1:.2f} s ({nc*256/t1:.0f} modes/s), sharded {tn:.2f} s ({nc*256/tn:.0f} modes/s); identical {same}; info {info} vs {iref}")
assert same and info["mode_failures"] == iref["mode_failures"]
dist.destroy_process_group()
