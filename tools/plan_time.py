import importlib, time, numpy as np, sys, os, ctypes as C
sys.path.insert(0,'.')
abi = importlib.import_module("jsvm-code-notes_b200.abi"); synth = importlib.import_module("jsvm-code-notes_b200.synth")
cfg = synth.CONFIGS["C4"]; W,H=cfg["width"],cfg["height"]
jobs = synth.make_jobs(W,H,cfg["parts"],1,0,qp=35,field="random",seed=1)[:30006]
lib=abi.load_library(); gsz=lib.jsvm_me_group_size()
groups=np.empty(len(jobs)*gsz,np.uint8); info=abi.PlanInfo()
t0=time.perf_counter()
for r in range(20): lib.jsvm_me_plan_host(W,H,2,64,C.c_void_p(jobs.ctypes.data),len(jobs),C.c_void_p(groups.ctypes.data),C.byref(info))
print(f"single-thread plan {1e9*(time.perf_counter()-t0)/20/len(jobs):.1f} ns/job, cpus {os.cpu_count()}")
