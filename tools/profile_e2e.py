"""Host-side profile of the e2e leg (OcrSystemTask.predict_pages on the bench workload): cProfile over a few steps.  Tuning aid."""
import cProfile
import pstats
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402

wl = bench.Cascade(0, 0, True)
for _ in range(3):
    wl.step_e2e()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    wl.step_e2e()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
