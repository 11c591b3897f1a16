"""NCCL-side check of the evaluation collectives (run under torchrun on >= 2 GPUs):
parcel-sharded Pearson (all-to-all re-lay of (windows, parcels, TRs) predictions), the statistics all-reduce variant, the
metric classes' distributed compute, ensemble averaging (one member per rank) and the retrieval metric's rank gather —
each against the same quantity computed on ONE GPU from the concatenated data (oracle for r: float64 numpy)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import algonauts2025_b200  # noqa: E402
from algonauts2025_b200 import parallel  # noqa: E402
from algonauts2025_b200.metrics import GroupedMetric, MultidimPearsonCorrCoef, TopkAcc  # noqa: E402
from oracle import aux_oracle as A  # noqa: E402
from oracle import tribe_oracle as O  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
g = torch.Generator().manual_seed(0)
n_win, Onum, T = 9 * world + 3, 1000, 100           # ragged window shards
true = torch.randn(n_win, Onum, T, generator=g)
pred = 0.25 * true + torch.randn(n_win, Onum, T, generator=g)
bounds = np.linspace(0, n_win, world + 1).astype(int)
bounds[1] += 2                                        # uneven on purpose
lo, hi = int(bounds[rank]), int(bounds[rank + 1])
p_loc, t_loc = pred[lo:hi].cuda(), true[lo:hi].cuda()
ref = O.pearson_columns_f64(O.flatten_bdt(pred).numpy(), O.flatten_bdt(true).numpy())

r1 = parallel.sharded_pearson(p_loc, t_loc).cpu().numpy()                      # (n, O, T): all-to-all + bdt kernel
assert np.abs(r1 - ref).max() < 1e-5, np.abs(r1 - ref).max()
r2 = parallel.sharded_pearson(O.flatten_bdt(p_loc).contiguous(), O.flatten_bdt(t_loc).contiguous()).cpu().numpy()  # (rows, O)
assert np.abs(r2 - ref).max() < 1e-5
r3 = parallel.allreduced_pearson(p_loc, t_loc).cpu().numpy()
assert np.abs(r3 - ref).max() < 1e-5
m = MultidimPearsonCorrCoef(num_outputs=Onum)
m.update_bdt(p_loc, t_loc)
assert abs(float(m.compute()) - float(ref.mean())) < 1e-5
subj = torch.arange(n_win) % 3
gm = GroupedMetric("MultidimPearsonCorrCoef", {"num_outputs": Onum})
gm.update_bdt(p_loc, t_loc, groups=subj[lo:hi].cuda())
got = gm.compute()
for s in range(3):
    sel = subj == s
    want = O.pearson_columns_f64(O.flatten_bdt(pred[sel]).numpy(), O.flatten_bdt(true[sel]).numpy()).mean()
    assert abs(got[str(s)] - want) < 1e-5, (s, got[str(s)], want)
# ensemble: every rank is a member predicting the SAME windows
member = (pred + 0.1 * (rank + 1) * torch.randn(pred.shape, generator=torch.Generator().manual_seed(10 + rank))).cuda()
from algonauts2025_b200 import ops  # noqa: E402

r_member = ops.pearson_r(member, true.cuda(), layout="bdt")[0]
ens = parallel.ensemble_average(member, r_member, temperature=0.3)
members = [torch.empty_like(member) for _ in range(world)]
rs = [torch.empty_like(r_member) for _ in range(world)]
dist.all_gather(members, member)
dist.all_gather(rs, r_member)
want = A.average_members(np.stack([x.cpu().numpy().transpose(0, 2, 1).reshape(-1, Onum) for x in members]), np.stack([x.cpu().numpy() for x in rs]),
                         weigh_by_score=True, per_voxel_weights=True, temperature=0.3)
got_e = ens.cpu().numpy().transpose(0, 2, 1).reshape(-1, Onum)
assert np.abs(got_e - want).max() < 1e-4 * np.abs(want).max() + 1e-6
# retrieval metric: ranks concatenated over ranks
tk = TopkAcc(topk=1)
tk.update_bdt(p_loc, t_loc)
acc = float(tk.compute())
parts = [A.retrieval_ranks(pred[int(bounds[r]):int(bounds[r + 1])].mean(-1).numpy().astype(np.float64), true[int(bounds[r]):int(bounds[r + 1])].mean(-1).numpy().astype(np.float64))
         for r in range(world)]
assert abs(acc - A.topk_acc(np.concatenate(parts), 1)) < 1e-6
dist.barrier()
if rank == 0:
    print(f"multi-GPU evaluation collectives OK on {world} GPUs (max |dr| {np.abs(r1 - ref).max():.1e})", flush=True)
dist.destroy_process_group()
