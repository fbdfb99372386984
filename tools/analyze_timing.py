import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from cholesky_b200.distributed import make_partitioned
dist.init_process_group("gloo")
t = time.time()
ch = make_partitioned(grid=(128,128,128,7,0))
dist.barrier()
if dist.get_rank() == 0: print("share", os.environ.get("CHOL_SHARE_ANALYSIS","1"), "analyze_s", round(time.time()-t, 2), flush=True)
dist.destroy_process_group()
