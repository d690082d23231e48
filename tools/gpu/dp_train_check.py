"""Data-parallel train step on N GPUs (run under torchrun): every rank steps on its own shard with per-rank BatchNorm
statistics, ONE all-reduce of the trainer's reduce buffer; replicas must stay bit-identical and the reported loss must
be the global mean."""
import copy, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from common import synth_batch
from noise_flow_b200 import NoiseFlow, hps_loader, load_checkpoint
from noise_flow_b200.train import DeviceTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
g = os.path.join(ROOT, "tests", "golden", "NoiseFlow")
hps = hps_loader(os.path.join(g, "hps.txt")); ck = load_checkpoint(os.path.join(g, "ckpt", "model.ckpt.best"))
nf = NoiseFlow([32, 32, 4], True, copy.copy(hps), variables=ck, device=dev, first_call="inverse")
tr = DeviceTrainer(nf, learning_rate=1e-4, max_batch=16)
ok = True
for s in range(3):
    x, y = synth_batch(16 * world, seed=500 + s)
    xs, ys = x[rank * 16:(rank + 1) * 16], y[rank * 16:(rank + 1) * 16]
    # per-rank loss before the all-reduce (separate evaluation on a throw-away trainer state is not needed: loss_and_grad
    # does not modify variables)
    tr.loss_and_grad(xs, ys, iso=[100.0], cam=[2.0])
    local_loss = tr.loss()[0]
    loss, sd = tr.step(xs, ys, iso=[100.0], cam=[2.0])
    t = torch.tensor([local_loss], device=dev, dtype=torch.float64)
    dist.all_reduce(t)
    ok &= abs(float(t[0]) / world - loss) < 1e-6 * abs(loss)
    if rank == 0:
        print("step %d: global mean loss/dim %.6f (mean of rank losses %.6f), sd_z %.4f" % (s, loss / 4096, float(t[0]) / world / 4096, sd))
flat = np.concatenate([v.reshape(-1) for v in tr.variables().values()])
t = torch.as_tensor(flat, device=dev)
lo, hi = t.clone(), t.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
same = bool(torch.equal(lo, hi))
if rank == 0:
    print("replicas bit-identical after 3 steps:", same, "| loss = global mean:", ok)
    print("DP_TRAIN_CHECK", "OK" if (same and ok) else "FAILED")
# the host-synchronous path (train.train_step): gradients AND batch statistics ride in one all-reduce, so the BatchNorm
# moving averages -- and therefore a checkpoint written by any rank -- stay identical across ranks
from noise_flow_b200.train import AdamOptimizer, train_step
nf2 = NoiseFlow([32, 32, 4], True, copy.copy(hps), variables={k: v.copy() for k, v in ck.items()}, device=dev, first_call="inverse")
opt = AdamOptimizer(learning_rate=1e-4)
for s in range(2):
    x, y = synth_batch(8 * world, seed=600 + s)
    train_step(nf2, opt, x[rank * 8:(rank + 1) * 8], y[rank * 8:(rank + 1) * 8], iso=[100.0], cam=[2.0])
flat2 = np.concatenate([v.reshape(-1) for k, v in sorted(nf2.variables.items())])
t2 = torch.as_tensor(flat2, device=dev)
lo2, hi2 = t2.clone(), t2.clone()
dist.all_reduce(lo2, op=dist.ReduceOp.MIN); dist.all_reduce(hi2, op=dist.ReduceOp.MAX)
if rank == 0:
    print("host path: variables incl. BatchNorm moving statistics identical on all ranks after 2 steps:", bool(torch.equal(lo2, hi2)))
    print("DP_HOST_TRAIN_CHECK", "OK" if bool(torch.equal(lo2, hi2)) else "FAILED")
dist.destroy_process_group()
