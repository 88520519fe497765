"""Timeline of the shared IB chain inside the C5 step on N GPUs (VSB_SHARD_TRACE=1, torchrun): per rank, for the last
replayed steps, when each flag barrier was entered / left relative to the first barrier of the step."""
import os, sys, json
os.environ["VSB_SHARD_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from vivsim_b200 import configs
from vivsim_b200.multidevice import SlabStepper
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
spec, body = configs.oscillating_cylinder_3d()
st = SlabStepper(spec, body=dict(body), dyn_mode="device", halo="peer", ib="shard", follow=2)
st.set_f_local(configs.uniform_state(dict(spec, shape=st.slab.local_shape), noise=1e-3))
st.step(3)
loop = bench.GraphLoop([st], 2)
loop.run(6)
torch.cuda.synchronize(); dist.barrier()
n_bar = int(st.ib_shard.counter[0].item())
tr = st.ib_shard.trace.cpu().numpy().reshape(4096, 2)
per_step = st.stepper.n_iter + 1
rows = []
for k in range(n_bar - 3 * per_step + 1, n_bar + 1):      # the last three steps
    rows.append((k, int(tr[k % 4096, 0]), int(tr[k % 4096, 1])))
t0 = rows[0][1]
line = {"rank": rank, "barriers_us(enter,exit)": [(k, round((a - t0) / 1e3, 1), round((b - t0) / 1e3, 1)) for k, a, b in rows]}
out = [None] * world
dist.all_gather_object(out, line)
if rank == 0:
    for o in out:
        print(json.dumps(o))
dist.barrier(); dist.destroy_process_group()
