"""Host CPU time per e2e step (time.process_time / thread_time) against the wall time of the same steps, per_call and stream form:
how close the e2e leg is to being host-bound.  Tuning aid."""
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402

wl = bench.Cascade(0, 0, True)
for _ in range(3):
    wl.step_e2e()
torch.cuda.synchronize()
for name, run in (("per_call", lambda n: [wl.step_e2e() for _ in range(n)]), ("stream", lambda n: [None for _ in wl.stream_e2e(n)])):
    run(2)
    torch.cuda.synchronize()
    w0, c0, t0 = time.perf_counter(), time.process_time(), time.thread_time()
    n = 10
    run(n)
    torch.cuda.synchronize()
    w1, c1, t1 = time.perf_counter(), time.process_time(), time.thread_time()
    print(f"{name}: wall {1e3 * (w1 - w0) / n:.1f} ms/step, process CPU {1e3 * (c1 - c0) / n:.1f} ms/step, main thread CPU {1e3 * (t1 - t0) / n:.1f} ms/step")
