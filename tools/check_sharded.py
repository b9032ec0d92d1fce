"""2+ ranks (torchrun, NCCL): sharded class_statistics == unsharded, and pair-sharded fit loss == replicated."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
from conftest import make_class_data, rel_err
from sqfa_b200 import statistics as S
from sqfa_b200.model import SQFA

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X, y = make_class_data(8000, 300, 12, seed=3)
X = X / (X.std() * 300**0.5)
lo, hi = rank * 8000 // world, (rank + 1) * 8000 // world
full = S.class_statistics(X.cuda(), y.cuda())
part = S.class_statistics(X[lo:hi].cuda(), y[lo:hi].cuda(), group=dist.group.WORLD)
errs = {k: rel_err(part[k], full[k]) for k in full}
F0 = torch.randn(4, 300, generator=torch.Generator().manual_seed(0))
m1 = SQFA(n_dim=300, feature_noise=0.01, n_filters=4, filters=F0.clone())
l1, _ = m1.fit(data_statistics=full, max_epochs=3, show_progress=False, return_loss=True)
m2 = SQFA(n_dim=300, feature_noise=0.01, n_filters=4, filters=F0.clone())
l2, _ = m2.fit(data_statistics=full, max_epochs=3, show_progress=False, return_loss=True, process_group=dist.group.WORLD)
if rank == 0:
    print("sharded stats rel err", errs)
    print("fit losses replicated", l1.tolist(), "pair-sharded", l2.tolist())
    assert all(v < 1e-5 for v in errs.values())
    assert abs(l1[-1] - l2[-1]) < 1e-4 * abs(l1[-1])
    print("SHARDED OK")
dist.barrier(); dist.destroy_process_group()
