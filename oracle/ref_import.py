"""Import recipe for the reference's modules (BUILD CONTAINER ONLY -- /root/reference does not exist on
the GPU box; nothing in `-m gpu` tests, smoke() or bench.py may call this).

The reference package cannot be imported whole here (pdfminer / onnx / pyclipper / shapely missing,
transformers>=5 broke `transformers.onnx`, `attribute_map` config defaults and `ViTModel.get_head_mask`),
so stub parent packages are registered with __path__ set and sub-module *files* are imported without
running any of the reference's __init__.py.
"""
from __future__ import annotations

import logging
import os
import sys
import types

REF_ROOT = os.environ.get("PDFTABLE_REFERENCE", "/root/reference")
R = os.path.join(REF_ROOT, "src", "pdftable")


def available() -> bool:
    return os.path.isdir(R)


def _stub(name, path=None):
    m = types.ModuleType(name)
    if path:
        m.__path__ = [path]
    sys.modules[name] = m
    return m


_done = False


def setup() -> None:
    global _done
    if _done:
        return
    if not available():
        raise RuntimeError(f"reference not found at {R}")
    import torch
    import transformers

    onnx_mod = _stub("transformers.onnx")
    onnx_mod.OnnxConfig = type("OnnxConfig", (), {"__init__": lambda s, *a, **k: None})
    transformers.onnx = onnx_mod
    for n in ("shapely", "shapely.geometry", "pyclipper"):
        try:
            __import__(n)
        except ImportError:
            m = _stub(n)
            if n == "shapely.geometry":
                m.MultiPoint = m.Point = m.Polygon = object
    _stub("pdftable", R)
    _stub("pdftable.model", R + "/model")
    _stub("pdftable.loss", R + "/loss")
    for d in os.listdir(R + "/model"):
        if os.path.isdir(f"{R}/model/{d}"):
            _stub(f"pdftable.model.{d}", f"{R}/model/{d}")
    u = _stub("pdftable.utils", R + "/utils")
    u.logger = logging.getLogger("ref")
    u.FileUtils = type("FU", (), {"check_file_exists": staticmethod(os.path.exists)})
    u.CommonUtils = type("CU", (), {"get_torch_device": staticmethod(lambda use_gpu=True: torch.device("cpu"))})
    u.TimeUtils = type("TU", (), {})
    _done = True


def dbmodel():
    setup()
    from pdftable.model.db_net.dbnet import DBModel
    return DBModel


def ctc_label_decode():
    setup()
    from pdftable.model.ocr_rec_pp.rec_postprocess import CTCLabelDecode
    return CTCLabelDecode


def convnext_vit():
    setup()
    from pdftable.model.convnext_vit.modeling_convnext_vit import ConvNextViT
    m = ConvNextViT().eval()
    m.vitstr.vit.get_head_mask = lambda mask, n: None  # API removed in transformers>=5
    return m
