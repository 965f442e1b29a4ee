"""cProfile over the stream form of the e2e leg (OcrSystemTask.predict_stream on the bench workload).  Tuning aid."""
import cProfile
import pstats
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402

wl = bench.Cascade(0, 0, True)
for _ in wl.stream_e2e(4):
    pass
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in wl.stream_e2e(10):
    pass
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.print_callers("engine.py:44")
st.print_callers("method 'to' of")
