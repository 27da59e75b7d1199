"""ops.scale_boxes + clip_boxes (reference utils/ops.py:66-98, 276-295): the host helper and the `yl_scale_boxes`
kernel against tests/golden/scale_boxes.npz, written by the unmodified reference on seeded boxes (letterbox pad
rounded with round(x - 0.1), explicit validator ratio_pad, boxes straddling every image edge).  Bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch


def _cases(g):
    for i in range(int(g["n_cases"])):
        rp = None
        if bool(g[f"c{i}.has_rp"]):
            r, px, py = (float(v) for v in g[f"c{i}.ratio_pad"])
            rp = ((r, r), (px, py))
        yield (tuple(int(v) for v in g[f"c{i}.img1"]), tuple(int(v) for v in g[f"c{i}.img0"]), rp,
               g[f"c{i}.boxes"], g[f"c{i}.out"])


def test_host_scale_boxes_matches_reference(golden):
    from yololite.utils import ops

    n = 0
    for s1, s0, rp, boxes, want in _cases(golden("scale_boxes.npz")):
        got = ops.scale_boxes(s1, torch.from_numpy(boxes.copy()), s0, ratio_pad=rp).numpy()
        np.testing.assert_array_equal(got, want)
        n += 1
    assert n >= 12


@pytest.mark.gpu
def test_scale_boxes_kernel_matches_reference(golden):
    from yololite import _C
    from yololite.utils import ops

    lib = _C.init(torch.device("cuda", 0))
    cases = list(_cases(golden("scale_boxes.npz")))
    B, max_det = len(cases), 64
    dets = np.zeros((B, max_det, 6), np.float32)
    params = np.zeros((B, 5), np.float32)
    counts = np.zeros((B,), np.int32)
    for i, (s1, s0, rp, boxes, _) in enumerate(cases):
        n = len(boxes)
        dets[i, :n, :4] = boxes
        dets[i, :n, 4] = np.linspace(0.9, 0.1, n)
        dets[i, :n, 5] = np.arange(n) % 7
        dets[i, n:, :4] = 12345.0                       # rows beyond the count must stay untouched
        counts[i] = n
        gain, pad = ops.letterbox_params(s1, s0, rp)
        params[i] = (gain, pad[0], pad[1], s0[1], s0[0])
    d = torch.from_numpy(dets).cuda()
    c = torch.from_numpy(counts).cuda()
    p = torch.from_numpy(params).cuda()
    _C.check(lib.yl_scale_boxes(d.data_ptr(), c.data_ptr(), B, max_det, p.data_ptr(), _C.stream_ptr()), "yl_scale_boxes")
    got = d.cpu().numpy()
    for i, (_, _, _, boxes, want) in enumerate(cases):
        n = len(boxes)
        np.testing.assert_array_equal(got[i, :n, :4], want, err_msg=f"case {i}")
        np.testing.assert_array_equal(got[i, :n, 4:], dets[i, :n, 4:])
        np.testing.assert_array_equal(got[i, n:], dets[i, n:])
