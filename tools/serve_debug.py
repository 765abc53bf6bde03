import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
print("CUDA_LAUNCH_BLOCKING", os.environ.get("CUDA_LAUNCH_BLOCKING"), "MAX_CONNECTIONS", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s1):
    torch.cuda._sleep(int(1.9e9 * 0.2))
with torch.cuda.stream(s2):
    torch.cuda._sleep(int(1.9e9 * 0.2))
torch.cuda.synchronize()
print("two 0.2 s sleeps on two streams:", time.perf_counter() - t0, "s (0.2 = concurrent, 0.4 = serialised)")
from oracle.gen_golden import CONFIG2_LAYOUT
from tests.test_gpu_gridworld import _vec_kwargs
from rl_envs_forge_b200 import VecGridWorld, _lib
cfg = _vec_kwargs(dict(CONFIG2_LAYOUT, p_success=0.8))
env = VecGridWorld(4096, seed=1, **cfg)
torch.cuda.synchronize()
st = env.serve_begin(5)
time.sleep(0.5)
print("after 0.5 s: act_seq", st["act_seq"][:4].tolist() if False else "(not read: reading would sync)")
prod = torch.cuda.Stream()
with torch.cuda.stream(prod):
    rc = env._lib.ef_policy_demo_serve(env._obs.data_ptr(), st["action"].data_ptr(), st["obs_seq"].data_ptr(), st["act_seq"].data_ptr(), 5,
                                       st["status"].data_ptr(), env.num_envs, prod.cuda_stream)
print("producer rc", rc)
t0 = time.perf_counter()
prod.synchronize()
print("producer done after", time.perf_counter() - t0)
t0 = time.perf_counter()
st["stream"].synchronize()
print("server done after", time.perf_counter() - t0, "status", st["status"].item(), "act_seq", st["act_seq"][:4].tolist(), "obs_seq", st["obs_seq"][:4].tolist())
