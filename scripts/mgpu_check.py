"""Multi-GPU correctness check (run under torchrun on N GPUs): the sharded solve (captures
split over ranks, one NCCL allreduce per linearisation) must walk the same LM trajectory
as the single-GPU solve of the whole problem."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ar_slam_b200 as ar  # noqa: E402
import bench  # noqa: E402
from ar_slam_b200 import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def fresh_uid():
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(ar.Solver.comm_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    return bytes(uid.cpu().numpy().tobytes())


ok = True
# "few-captures": every rank misses some of the tags, which must move all the same
# "radial-*": the three-intrinsics model (f, l1, l2); its border columns travel in the same allreduce
for name, ls, (nc, nt) in (("dense", ar.LINSOLVE_DENSE, (3000, 400)), ("pcg", ar.LINSOLVE_PCG, (20000, 1500)),
                           ("few-captures", ar.LINSOLVE_DENSE, (40 * world, 200)),
                           ("radial-dense", ar.LINSOLVE_DENSE, (3000, 400)), ("radial-pcg", ar.LINSOLVE_PCG, (8000, 900))):
    radial = name.startswith("radial")
    m = synth.make_map(nc, nt, seed=31, distortion=(-0.05, 0.01) if radial else (0.0, 0.0))
    opts = ar.default_options(linear_solver=ls, pcg_tolerance=1e-10, pcg_max_iterations=3000, num_intrinsics=3 if radial else 1)
    s = ar.Solver(device=local, options=opts)
    s.comm_init(rank, world, fresh_uid())
    ci, ti, ob = bench.shard(m, rank, world)
    s.set_problem(m.n_cap, m.n_tag, ci, ti, ob)
    s.set_params(m.cam0, m.cap0, m.tag0)
    summ, log = s.solve()
    cam, cap, tag = s.get_params()   # cap: only this rank's capture range is filled, zeros elsewhere
    s.close()
    cap_all = torch.from_numpy(cap).cuda()
    dist.all_reduce(cap_all)           # disjoint ranges: the sum is the union
    cap = cap_all.cpu().numpy()
    if rank == 0:
        s1 = ar.Solver(device=local, options=opts)
        s1.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s1.set_params(m.cam0, m.cap0, m.tag0)
        summ1, log1 = s1.solve()
        cam1, cap1, tag1 = s1.get_params()
        s1.close()
        same_it = summ["iterations"] == summ1["iterations"]
        dcost = float(np.abs(log[:, 0] - log1[:len(log), 0]).max() / log1[0, 0]) if same_it else float("nan")
        dp = max(np.abs(cap - cap1).max(), np.abs(tag - tag1).max(), abs(cam[0] - cam1[0]) / cam1[0])
        good = same_it and dcost < 1e-9 and dp < 1e-6 and abs(summ["final_cost"] - summ1["final_cost"]) < 1e-8 * summ1["final_cost"]
        ok = ok and good
        print("%s: world %d  iterations %d vs %d  final cost %.9g vs %.9g  max trajectory diff %.2e  max param diff %.2e  %s"
              % (name, world, summ["iterations"], summ1["iterations"], summ["final_cost"], summ1["final_cost"], dcost, dp,
                 "OK" if good else "MISMATCH"), flush=True)
# ---- rank-consistent failure: overlapping capture ranges and a rank-local bad index must fail on EVERY rank
# (round 1 hung the healthy ranks in the next collective)
m = synth.make_map(2000, 200, seed=32)
for case in ("overlap", "bad_index", "no_params"):
    s = ar.Solver(device=local, options=ar.default_options())
    s.comm_init(rank, world, fresh_uid())
    ci, ti, ob = bench.shard(m, rank, world)
    if case == "overlap" and rank == 1:      # rank 1 also claims the first capture of rank 0
        ci, ti, ob = np.concatenate([[0], ci]).astype(np.int32), np.concatenate([[ti[0]], ti]).astype(np.int32), np.concatenate([ob[:1], ob])
    if case == "bad_index" and rank == world - 1:
        ti = ti.copy()
        ti[3] = m.n_tag
    failed = False
    try:
        s.set_problem(m.n_cap, m.n_tag, ci, ti, ob)
        if case != "no_params" or rank != 0:
            s.set_params(m.cam0, m.cap0, m.tag0)
        s.solve()
    except ar.ArslamError as e:
        failed = True
        msg = str(e)
    s.close()
    flags = torch.tensor([1.0 if failed else 0.0], device="cuda")
    dist.all_reduce(flags)
    good = int(flags.item()) == world
    ok = ok and good
    if rank == 0:
        print("%s: %d of %d ranks raised (%s)  %s" % (case, int(flags.item()), world, msg if failed else "-", "OK" if good else "MISMATCH"), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
