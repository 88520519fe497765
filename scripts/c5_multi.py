"""BASELINE config 5 (or 4) on all GPUs of the node through bench.multi_gpu_config (CUDA-graph loop); torchrun."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
for name in sys.argv[1:] or ["c5"]:
    out = bench.multi_gpu_config(name, world, rank, bench.peaks()[0])
    if rank == 0:
        print(json.dumps(out), flush=True)
dist.barrier(); dist.destroy_process_group()
