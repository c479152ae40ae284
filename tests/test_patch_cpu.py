"""nuhtc_b200.patch rebinds the names the reference modules captured at import time.  mmcv / mmdet / nuhtc are not in this
image, so stand-in modules with the reference's module paths and attribute names take their place."""
import sys
import types

import pytest


@pytest.fixture
def fake_reference_modules(monkeypatch):
    names = ["mmcv", "mmcv.ops", "mmcv.ops.nms", "nuhtc", "nuhtc.models", "nuhtc.models.bbox_head", "nuhtc.core",
             "nuhtc.core.post_processing", "nuhtc.core.post_processing.bbox_nms", "mmdet", "mmdet.core",
             "mmdet.core.post_processing", "mmdet.core.post_processing.bbox_nms", "mmdet.models", "mmdet.models.dense_heads",
             "mmdet.models.dense_heads.rpn_head", "mmdet.models.roi_heads", "mmdet.models.roi_heads.mask_heads",
             "mmdet.models.roi_heads.mask_heads.fcn_mask_head"]
    mods = {}
    for n in names:
        m = types.ModuleType(n)
        m.__path__ = []
        mods[n] = m
        monkeypatch.setitem(sys.modules, n, m)
        if "." in n:
            parent, leaf = n.rsplit(".", 1)
            setattr(mods[parent], leaf, m)
    sentinel = object()
    for n in ("mmcv.ops", "mmcv.ops.nms"):
        mods[n].nms = mods[n].batched_nms = sentinel
    mods["mmcv.ops"].RoIAlign = mods["mmcv.ops"].roi_align = sentinel
    for n in ("nuhtc.models.bbox_head", "nuhtc.core.post_processing.bbox_nms", "mmdet.core.post_processing.bbox_nms",
              "mmdet.models.dense_heads.rpn_head"):
        mods[n].batched_nms = sentinel
    mods["mmdet.models.roi_heads.mask_heads.fcn_mask_head"]._do_paste_mask = sentinel
    return mods, sentinel


def test_patch_mmcv_rebinds_every_captured_name(fake_reference_modules):
    import nuhtc_b200 as nb
    from nuhtc_b200 import patch
    mods, sentinel = fake_reference_modules
    done = patch.patch_mmcv()
    assert mods["mmcv.ops"].RoIAlign is nb.RoIAlign and mods["mmcv.ops"].roi_align is nb.roi_align
    for n in ("mmcv.ops", "mmcv.ops.nms"):
        assert mods[n].nms is nb.nms and mods[n].batched_nms is nb.batched_nms
    for n in ("nuhtc.models.bbox_head", "nuhtc.core.post_processing.bbox_nms", "mmdet.core.post_processing.bbox_nms",
              "mmdet.models.dense_heads.rpn_head"):
        assert mods[n].batched_nms is nb.batched_nms
    assert mods["mmdet.models.roi_heads.mask_heads.fcn_mask_head"]._do_paste_mask is nb._do_paste_mask
    assert len(done) == 11 and all(isinstance(d, str) for d in done)
    for m in mods.values():
        assert all(v is not sentinel for v in vars(m).values())


def test_patch_wsi_tools():
    import nuhtc_b200 as nb
    from nuhtc_b200 import patch
    infer, merge = types.SimpleNamespace(mask_nms=None, mask2inst=None), types.SimpleNamespace(merge_overlap=None)
    assert patch.patch_wsi_tools(infer, merge) == ["infer_wsi.mask_nms", "infer_wsi.mask2inst", "nuclei_merge.merge_overlap"]
    assert infer.mask_nms is nb.mask_nms and infer.mask2inst is nb.mask2inst and merge.merge_overlap is nb.merge_overlap


def test_patch_without_optional_modules(monkeypatch):
    """Only mmcv present: the nuhtc / mmdet rebinds are skipped, not fatal."""
    from nuhtc_b200 import patch
    for n in ("mmcv", "mmcv.ops", "mmcv.ops.nms"):
        m = types.ModuleType(n)
        m.__path__ = []
        monkeypatch.setitem(sys.modules, n, m)
    sys.modules["mmcv"].ops = sys.modules["mmcv.ops"]
    sys.modules["mmcv.ops"].nms = sys.modules["mmcv.ops.nms"]
    for n in ("nuhtc", "mmdet"):
        monkeypatch.setitem(sys.modules, n, None)      # import_module raises ImportError
    assert len(patch.patch_mmcv()) == 6
