"""C5 on N GPUs with the shared IB chain: whole step, chain alone, fluid alone (torchrun; VSB_SHARD_STOP=k truncates the chain)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from vivsim_b200 import configs
from vivsim_b200.multidevice import SlabStepper
import ctypes as C

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
spec, body = configs.oscillating_cylinder_3d(nx=int(1024 * scale), ny=int(512 * scale), nz=int(512 * scale))
st = SlabStepper(spec, body=dict(body), dyn_mode="device", halo="peer", ib="shard", follow=2)
st.set_f_local(configs.uniform_state(dict(spec, shape=st.slab.local_shape), noise=1e-3))
st.step(4)

def sync():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

def timeit(fn, n):
    fn(); sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); sync()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    lo = t.clone(); dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return float(t), float(lo)

inner = st.stepper
full = timeit(lambda: st.advance_raw(2), 6)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def chain():
    inner._ib_part(stream)
    inner._parity ^= 1
chain_t = timeit(lambda: (chain(), chain()), 6)
a = inner._args
def fluid():
    import vivsim_b200._lib as L
    a.f_in, a.f_out = inner._bufs[0].data_ptr(), inner._bufs[1].data_ptr()
    a.do_stream, a.do_collide, a.band, a.edges = 1, 1, 0, 2 if (a.n_post > 0 and inner.edge_fused) else 0
    a.sub_begin, a.sub_end, a.edge_rows_only = 0, 0, 0
    L.check(L.lib().vsb_step(C.byref(a), stream))
fluid_t = timeit(lambda: (fluid(), fluid()), 6)
if rank == 0:
    print(json.dumps({"world": world, "stop": os.environ.get("VSB_SHARD_STOP"), "ms_two_steps_max_min": full,
                      "ms_two_chains_max_min": chain_t, "ms_two_fluid_passes_max_min": fluid_t,
                      "need_boxes": [st.ib_shard.plan["need_lo"].tolist(), st.ib_shard.plan["need_hi"].tolist()],
                      "window": [list(spec["ib"]["window"][0]), list(spec["ib"]["window"][1])]}), flush=True)
dist.barrier(); dist.destroy_process_group()
