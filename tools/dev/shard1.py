import os, sys
sys.path.insert(0, "/root/repo")
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
import numpy as np, torch, torch.distributed as dist
from yadcc_b200 import TaskDispatcher, streams as S
from yadcc_b200.sharded import RangeShardedDispatcher
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
w = S.config2(variant="mod")
outs = []
for mode in ("sharded", "single"):
    d = TaskDispatcher()
    w.register(d, now=0.0, expires_in=3600.0)
    reqs = w.build_requests(d)
    print("=====", mode, file=sys.stderr, flush=True)
    if mode == "sharded":
        sd = RangeShardedDispatcher(d, 0, 1, device=torch.device("cuda", 0))
        g = sd.wait_for_starting_new_tasks(reqs, 0.001).copy()
    else:
        g = d.wait_for_starting_new_tasks(reqs, 0.001).copy()
    outs.append(g)
print("mismatch servant:", int((outs[0]["servant_index"] != outs[1]["servant_index"]).sum()))
