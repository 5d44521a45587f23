"""Host-side cost of issuing one training step (CIFAR bs=128): wall time of the three C-ABI calls without any synchronisation,
against the device time of the same steps.  If the two are close the step is launch-bound, not kernel-bound."""
import os, sys, time
here = os.path.dirname(os.path.abspath(__file__))
sys.argv = [sys.argv[0], "train", "128"]
src = open(os.path.join(here, "profile_step.py")).read().split("for i in range(2):")[0]
exec(src)
for i in range(5): step(i)
torch.cuda.synchronize()
N = 50
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); t0 = time.perf_counter()
per = []
for i in range(N):
    a = time.perf_counter(); step(10 + i); per.append(time.perf_counter() - a)
t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
per.sort()
print("host enqueue ms/step: mean %.3f  median %.3f  min %.3f" % ((t1 - t0) / N * 1e3, per[N // 2] * 1e3, per[0] * 1e3))
print("device ms/step: %.3f   host tail wait after last enqueue: %.3f ms" % (e0.elapsed_time(e1) / N, (t2 - t1) * 1e3))
