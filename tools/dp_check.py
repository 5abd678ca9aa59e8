"""2-rank check of the data-parallel step (torchrun --nproc-per-node 2 tools/dp_check.py MODE):
MODE = eager | graph   with MTBC_DP_OVERLAP=0/1.  Prints the loss of each step and a parameter checksum on rank 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from oracle import torch_oracle as O
from multi_task_breast_cancer_b200 import models as M
from multi_task_breast_cancer_b200.train import TrainStep
mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(1993)
model = M.MTnnUNet(1, 1, 3).to(dev)
ts = TrainStep(model, (2, 1, 64, 64), process_group=dist.group.WORLD, use_graph=(mode == "graph"))
img, mask, onehot, _ = O.synthetic_batch(2, 64, 64, seed=1993 + rank)
ts.load_batch(img.to(dev), mask.to(dev), onehot.to(dev))
out = []
for i in range(4):
    ts.step()
    torch.cuda.synchronize()
    out.append(ts.losses()[0].item())
chk = ts.flat_p.double().abs().sum().item()
allc = [None, None]
dist.all_gather_object(allc, chk)
if rank == 0:
    print(f"mode {mode} overlap {os.environ.get('MTBC_DP_OVERLAP', '1')}: losses {['%.5f' % v for v in out]} param checksum per rank {allc} "
          f"(ranks identical: {abs(allc[0] - allc[1]) < 1e-9 * abs(allc[0])})", flush=True)
dist.barrier()
dist.destroy_process_group()
