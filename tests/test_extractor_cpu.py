"""Host logic of the encoder call-pattern mirrors (extractor.extract_vsc_feat / single_infer) with a stand-in module on
CPU tensors; against the reference's own extract_vsc_feat where /root/reference is present."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import refload


class TinyEncoder(torch.nn.Module):
    def __init__(self, tokens=False):
        super().__init__()
        self.proj = torch.nn.Linear(3 * 4 * 4, 8)
        self.tokens = tokens

    def forward(self, x):
        y = self.proj(x.reshape(x.shape[0], -1))
        return torch.stack([y, y + 1], 1) if self.tokens else y


def batches(seed=0):
    """What D_vsc.collate_fn yields (dataset.py:149-155): zero-padded frames, mask = frame has any non-zero sample."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for b, lens in enumerate([(3, 5), (1, 1), (4, 2)]):
        S = max(lens)
        frames = torch.zeros(len(lens), S, 3, 4, 4)
        for i, n in enumerate(lens):
            frames[i, :n] = torch.randn(n, 3, 4, 4, generator=g)
        mask = (frames.reshape(len(lens), S, -1).sum(-1) != 0).long()
        out.append((frames, mask, tuple(f"Q{b}{i}" for i in range(len(lens)))))
    return out


def test_extract_vsc_feat_layout():
    from vsc22_submission_b200.extractor import extract_vsc_feat
    torch.manual_seed(0)
    model = TinyEncoder().eval()
    vids, feat, ts = extract_vsc_feat(model, batches(), "cpu")
    assert vids == ["Q00"] * 3 + ["Q01"] * 5 + ["Q10", "Q11"] + ["Q20"] * 4 + ["Q21"] * 2
    assert ts.tolist() == [0, 1, 2, 0, 1, 2, 3, 4, 0, 0, 0, 1, 2, 3, 0, 1] and feat.shape == (16, 8)
    fr, mask, _ = batches()[0]
    with torch.no_grad():
        np.testing.assert_array_equal(feat[:8], model(fr[mask.bool()]).numpy())
    with pytest.raises(ValueError):
        extract_vsc_feat(model, [], "cpu")


def test_single_infer_chunks_and_token_outputs():
    from vsc22_submission_b200.extractor import single_infer
    torch.manual_seed(1)
    x = torch.randn(11, 3, 4, 4)
    for tokens in (False, True):
        model = TinyEncoder(tokens).eval()
        with torch.no_grad():
            want = model(x)
            want = (want[:, 0] if tokens else want).numpy()
        np.testing.assert_allclose(single_infer(model, x, len_threshold=4), want, atol=1e-6)
        model.max_frames = 64                       # plan-style module: one call for the whole video
        np.testing.assert_allclose(single_infer(model, x), want, atol=1e-6)


@pytest.mark.skipif(not refload.available(), reason="/root/reference not present")
def test_extract_vsc_feat_equals_reference():
    from vsc22_submission_b200.extractor import extract_vsc_feat
    refload.vsc_package("D_infer")
    try:
        spec = importlib.util.spec_from_file_location("_ref_extractor", os.path.join(refload.D, "infer/src/extractor.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        torch.manual_seed(0)
        model = TinyEncoder().eval()
        rv, rf, rt = ref.extract_vsc_feat(model, batches(3), "cpu")
        gv, gf, gt = extract_vsc_feat(model, batches(3), "cpu")
        assert rv == gv
        np.testing.assert_array_equal(rf, gf)
        np.testing.assert_array_equal(rt, gt)
    finally:
        refload.unload_vsc()
