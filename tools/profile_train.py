"""torch.profiler tables of one training update and one rollout (16,384 episodes): python tools/profile_train.py"""
import sys, os; sys.path.insert(0, os.getcwd())
import torch, time
from torch.profiler import profile, ProfilerActivity
from azul_deep_reinforcement_learning_b200.train import SelfPlayTrainer
tr = SelfPlayTrainer(16384, seed=0)
for i in range(24):
    t0=time.perf_counter(); b=tr.rollout(); torch.cuda.synchronize(); t1=time.perf_counter(); tr.update(b); torch.cuda.synchronize(); t2=time.perf_counter()
    print("iter", i, "rollout ms %.1f update ms %.1f" % (1e3*(t1-t0), 1e3*(t2-t1)), "T", b["active"].shape[0])
torch.cuda.synchronize()
b = tr.rollout(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    t0=time.perf_counter(); tr.update(b); torch.cuda.synchronize(); print("update s", time.perf_counter()-t0)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    t0=time.perf_counter(); b = tr.rollout(); torch.cuda.synchronize(); print("rollout s", time.perf_counter()-t0)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
