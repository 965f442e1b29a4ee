#!/usr/bin/env python
"""The Blackwell bar: the reference's own GPU path is its torch modules in fp16 on cuDNN /
cuBLAS (+ torchvision.ops.deform_conv2d for Lore's DCN) -- base_infer_task.py:56-57, utils/deploy_utils.py:226-240,
lore/dcnv2.py:77-84.  This tool times that path (the oracle restatements of the modules, `.half().cuda()`, torch kernels) beside
the engine on the SAME inputs on the SAME box, stage by stage, and prints engine / torch-fp16 per stage.  Informational: not the
driver's bench, not a product path (it imports oracle/, like a test)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import convnextvit_ref, dbnet_ref, lore_net_ref, picodet_net_ref, pp_rec_ref  # noqa: E402
from pdf_table_b200 import picodet_graph, pp_rec_graph, synth, weights  # noqa: E402
from pdf_table_b200.engine import Engine  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def half_sd(sd):
    return {k: torch.from_numpy(np.asarray(v)).to(dev).half() if np.asarray(v).dtype.kind == "f" else torch.from_numpy(np.asarray(v)).to(dev) for k, v in sd.items()}


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


out = {}
rng = np.random.default_rng(0)
with torch.no_grad():
    # ---- DBNet-R18, 32 pages 960x960 (configs[1] detector)
    sd = synth.dbnet_r18_state_dict(0)
    x = torch.from_numpy(rng.standard_normal((32, 3, 960, 960)).astype(np.float32)).to(dev)
    eng = Engine("dbnet_r18", weights.pack_dbnet_r18(sd))
    hs, xh = half_sd(sd), x.half()
    out["dbnet_r18 32x960x960"] = {"engine_ms": timeit(lambda: eng.dbnet_forward(x)), "torch_fp16_ms": timeit(lambda: dbnet_ref.dbnet_r18_forward(hs, xh))}
    eng.close()
    del x, xh
    # ---- ConvNextViT, 384 crops 32x320 (configs[3] recogniser, one pass)
    sd = synth.convnext_vit_state_dict(0)
    crops = rng.integers(0, 256, (384, 32, 320, 3), dtype=np.uint8)
    chunks = convnextvit_ref.preprocess(list(crops)).to(dev)
    eng = Engine("convnext_vit", weights.pack_convnext_vit(sd))
    cu = torch.from_numpy(crops).to(dev)
    hs, ch = convnextvit_ref.to_torch(sd, "cuda", torch.float16), chunks.half()
    out["convnextvit 384 crops"] = {"engine_ms": timeit(lambda: eng.convnextvit_forward_u8(cu)),
                                    "torch_fp16_ms": timeit(lambda: convnextvit_ref.convnextvit_forward(hs, ch).argmax(-1))}
    eng.close()
    # ---- PP-OCRv4 rec, 1024 crops 48x320
    sd = synth.pp_ocrv4_rec_state_dict(0, 97)
    x = torch.from_numpy(rng.standard_normal((1024, 3, 48, 320)).astype(np.float32)).to(dev)
    eng = Engine("pp_rec", pp_rec_graph.pack_pp_rec(sd))
    hs, xh = half_sd(sd), x.half()
    out["pp_ocrv4_rec 1024 crops"] = {"engine_ms": timeit(lambda: eng.rec_forward(x)), "torch_fp16_ms": timeit(lambda: pp_rec_ref.pp_rec_forward(hs, xh))}
    eng.close()
    # ---- PicoDet, 32 pages 800x608
    bb, nk, hd = synth.picodet_state_dicts(0, 5)
    x = torch.from_numpy(rng.standard_normal((32, 3, 800, 608)).astype(np.float32)).to(dev)
    eng = Engine("picodet", picodet_graph.pack_picodet(bb, nk, hd, 5))
    hb, hn, hh, xh = half_sd(bb), half_sd(nk), half_sd(hd), x.half()
    out["picodet 32x800x608"] = {"engine_ms": timeit(lambda: eng.picodet_forward(x)), "torch_fp16_ms": timeit(lambda: picodet_net_ref.picodet_forward(hb, hn, hh, xh, 5))}
    eng.close()
    # ---- Lore DLA-34 + DCNv2 detector, 16 x 1024x1024 (configs[2]); the reference evaluates all six heads densely
    sd = synth.lore_dla34_state_dict(0)
    x = torch.from_numpy(rng.standard_normal((16, 3, 1024, 1024)).astype(np.float32)).to(dev)
    eng = Engine("lore_dla34", weights.pack_lore_dla34(sd))
    hs, xh = half_sd(sd), x.half()
    try:
        t_ref = timeit(lambda: lore_net_ref.lore_dla34_forward(hs, xh, use_torchvision=True), reps=3, warm=1)
    except Exception as ex:  # torchvision's CUDA deform_conv2d missing on this box
        t_ref = None
        out["lore_error"] = repr(ex)[:300]
    out["lore_dla34_dcn 16x1024x1024"] = {"engine_ms": timeit(lambda: eng.lore_detect_forward(x)), "torch_fp16_ms": t_ref}
    eng.close()
for k, v in out.items():
    if isinstance(v, dict) and v.get("torch_fp16_ms"):
        v["engine_over_torch_fp16"] = v["torch_fp16_ms"] / v["engine_ms"]
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/torch_fp16_bar.json", "w"), indent=1)
