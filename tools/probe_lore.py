"""GPU-box diagnostic: per-tensor error statistics of the Lore detector against the oracle (not a test)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import lore_net_ref  # noqa: E402
from pdf_table_b200 import synth, weights  # noqa: E402
from pdf_table_b200.engine import Engine  # noqa: E402

sd = synth.lore_dla34_state_dict(0)
eng = Engine("lore_dla34", weights.pack_lore_dla34(sd))
rng = np.random.default_rng(21)
x = torch.from_numpy(rng.standard_normal((2, 3, 96, 160)).astype(np.float32))
maps = eng.lore_detect_forward(x.cuda())
eng.sync()


def stats(name, got, want):
    d = np.abs(got - want)
    i = np.unravel_index(d.argmax(), d.shape)
    print(f"{name:10s} shape {tuple(want.shape)} max|x| {np.abs(want).max():.3f} err max {d.max():.3e} mean {d.mean():.3e} "
          f"p99 {np.percentile(d, 99):.3e} argmax {i}")


base = lore_net_ref.dla34_base(sd, x)
for lvl in range(6):
    stats(f"level{lvl}", eng.debug_tensor(f"level{lvl}").cpu().numpy(), base[lvl].numpy())
out = lore_net_ref.lore_dla34_forward(sd, x)
stats("feat", eng.debug_tensor("feat").cpu().numpy(), out["feat"].numpy())
m = maps.cpu().numpy().transpose(0, 3, 1, 2)
stats("hm", m[:, 0:2], torch.sigmoid(out["hm"]).numpy())
stats("reg", m[:, 2:4], out["reg"].numpy())
stats("wh", m[:, 4:12], out["wh"].numpy())
stats("st", m[:, 12:20], out["st"].numpy())
for name in sys.argv[1:]:
    pass
