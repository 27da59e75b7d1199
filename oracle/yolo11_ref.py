"""oracle/yolo11_ref.py — TEST INFRASTRUCTURE, not product code.

Plain PyTorch fp32 CPU restatement of the reference's YOLO11 detection forward, written functionally over a
state_dict with the reference's key names.  It reproduces the reference's *unfused* arithmetic
(conv -> BatchNorm2d(eps=1e-3) -> SiLU, reference yololite/nn/modules/conv.py:47-49,
utils/torch_utils.py:248-250) because the reference never folds BN at inference.
Each function cites the reference lines it follows.

Parity pin: tests/golden/model_*.npz were produced by the reference itself (oracle/gen_golden.py, run in the
build container where /root/reference exists) with the name-keyed weights of oracle/weights.py;
tests/test_oracle.py checks this restatement against them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# cfg/yolo11.yaml:15-47 — (from, type); every top-level Conv is k=3, s=2; topology is scale independent.
LAYERS = [
    (-1, "Conv"), (-1, "Conv"), (-1, "C3k2"), (-1, "Conv"), (-1, "C3k2"), (-1, "Conv"), (-1, "C3k2"),
    (-1, "Conv"), (-1, "C3k2"), (-1, "SPPF"), (-1, "C2PSA"),
    (-1, "Upsample"), ([-1, 6], "Concat"), (-1, "C3k2"),
    (-1, "Upsample"), ([-1, 4], "Concat"), (-1, "C3k2"),
    (-1, "Conv"), ([-1, 13], "Concat"), (-1, "C3k2"),
    (-1, "Conv"), ([-1, 10], "Concat"), (-1, "C3k2"),
    ([16, 19, 22], "Detect"),
]
STRIDES = (8.0, 16.0, 32.0)  # nn/tasks.py:258-268 probe result for this topology
REG_MAX = 16                 # head.py:34
BN_EPS = 1e-3                # utils/torch_utils.py:248-250


def conv(sd, p, x, s=1, act=True, eps=BN_EPS):
    """Conv.forward (conv.py:35-53): act(bn(conv(x))), pad = k // 2 (autopad conv.py:26-32), bias-free.

    eps: 1e-3 inside a DetectionModel (initialize_weights, torch_utils.py:248-250); a standalone reference
    module keeps nn.BatchNorm2d's default 1e-5."""
    w = sd[p + ".conv.weight"]
    k = w.shape[-1]
    groups = x.shape[1] // w.shape[1]
    y = F.conv2d(x, w, None, s, k // 2, 1, groups)
    y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"],
                     sd[p + ".bn.bias"], False, 0.0, eps)
    return F.silu(y) if act else y


def bottleneck(sd, p, x):
    """Bottleneck.forward (block.py:330-343), shortcut and c1 == c2 in every yolo11 use."""
    return x + conv(sd, p + ".cv2", conv(sd, p + ".cv1", x))


def _count(sd, prefix):
    n = 0
    while any(k.startswith(f"{prefix}.{n}.") for k in sd):
        n += 1
    return n


def c3k(sd, p, x):
    """C3.forward via C3k (block.py:257-259, 731-739): cv3(cat(m(cv1(x)), cv2(x)))."""
    y = conv(sd, p + ".cv1", x)
    for i in range(_count(sd, p + ".m")):
        y = bottleneck(sd, f"{p}.m.{i}", y)
    return conv(sd, p + ".cv3", torch.cat((y, conv(sd, p + ".cv2", x)), 1))


def c3k2(sd, p, x):
    """C2f.forward as inherited by C3k2 (block.py:231-235, 720-728)."""
    y = list(conv(sd, p + ".cv1", x).chunk(2, 1))
    for i in range(_count(sd, p + ".m")):
        q = f"{p}.m.{i}"
        y.append(c3k(sd, q, y[-1]) if (q + ".cv3.conv.weight") in sd else bottleneck(sd, q, y[-1]))
    return conv(sd, p + ".cv2", torch.cat(y, 1))


def sppf(sd, p, x):
    """SPPF.forward (block.py:165-184): three chained MaxPool2d(5, 1, 2)."""
    y = [conv(sd, p + ".cv1", x)]
    for _ in range(3):
        y.append(F.max_pool2d(y[-1], 5, 1, 2))
    return conv(sd, p + ".cv2", torch.cat(y, 1))


def attention(sd, p, x):
    """Attention.forward (block.py:863-916); head_dim 64, key_dim 32 (attn_ratio 0.5)."""
    B, C, H, W = x.shape
    N = H * W
    heads = C // 64                        # PSABlock(num_heads=c // 64), block.py:1034
    hd = C // heads
    kd = int(hd * 0.5)
    qkv = conv(sd, p + ".qkv", x, act=False)
    q, k, v = qkv.view(B, heads, kd * 2 + hd, N).split([kd, kd, hd], dim=2)
    attn = (q.transpose(-2, -1) @ k) * (kd ** -0.5)
    attn = attn.softmax(dim=-1)
    y = (v @ attn.transpose(-2, -1)).view(B, C, H, W) + conv(sd, p + ".pe", v.reshape(B, C, H, W), act=False)
    return conv(sd, p + ".proj", y, act=False)


def psablock(sd, p, x):
    """PSABlock.forward (block.py:919-953)."""
    x = x + attention(sd, p + ".attn", x)
    return x + conv(sd, p + ".ffn.1", conv(sd, p + ".ffn.0", x), act=False)


def c2psa(sd, p, x):
    """C2PSA.forward (block.py:999-1038)."""
    a, b = conv(sd, p + ".cv1", x).chunk(2, 1)
    for i in range(_count(sd, p + ".m")):
        b = psablock(sd, f"{p}.m.{i}", b)
    return conv(sd, p + ".cv2", torch.cat((a, b), 1))


def detect_raw(sd, p, feats):
    """Detect.forward up to the per-level cat (head.py:59-65): cv2 box branch, cv3 DW-separable cls branch."""
    out = []
    for i, x in enumerate(feats):
        b = conv(sd, f"{p}.cv2.{i}.1", conv(sd, f"{p}.cv2.{i}.0", x))
        b = F.conv2d(b, sd[f"{p}.cv2.{i}.2.weight"], sd[f"{p}.cv2.{i}.2.bias"])
        c = conv(sd, f"{p}.cv3.{i}.0.1", conv(sd, f"{p}.cv3.{i}.0.0", x))
        c = conv(sd, f"{p}.cv3.{i}.1.1", conv(sd, f"{p}.cv3.{i}.1.0", c))
        c = F.conv2d(c, sd[f"{p}.cv3.{i}.2.weight"], sd[f"{p}.cv3.{i}.2.bias"])
        out.append(torch.cat((b, c), 1))
    return out


def decode(raw, strides=STRIDES, reg_max=REG_MAX):
    """Detect._inference + DFL + make_anchors + dist2bbox (head.py:95-126, block.py:66-69, tal.py:326-350)."""
    B = raw[0].shape[0]
    no = raw[0].shape[1]
    x_cat = torch.cat([r.reshape(B, no, -1) for r in raw], 2)
    anchors, svec = [], []
    for r, s in zip(raw, strides):
        h, w = r.shape[2:]
        sx = torch.arange(w, dtype=torch.float32) + 0.5
        sy = torch.arange(h, dtype=torch.float32) + 0.5
        gy, gx = torch.meshgrid(sy, sx, indexing="ij")
        anchors.append(torch.stack((gx, gy), -1).view(-1, 2))
        svec.append(torch.full((h * w, 1), float(s)))
    dev = raw[0].device                       # CPU in every test; tools/eager_gpu_baseline.py runs it on CUDA
    anchors = torch.cat(anchors).t().to(dev)  # (2, A)
    svec = torch.cat(svec).t().to(dev)        # (1, A)
    box, cls = x_cat.split((reg_max * 4, no - reg_max * 4), 1)
    a = box.shape[-1]
    proj = torch.arange(reg_max, dtype=torch.float32, device=dev)
    dist = (box.view(B, 4, reg_max, a).transpose(2, 1).softmax(1) * proj.view(1, reg_max, 1, 1)).sum(1)
    lt, rb = dist.chunk(2, 1)
    x1y1 = anchors.unsqueeze(0) - lt
    x2y2 = anchors.unsqueeze(0) + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * svec
    return torch.cat((dbox, cls.sigmoid()), 1)


@torch.no_grad()
def forward(sd, x, return_features=False):
    """BaseModel._predict_once (nn/tasks.py:118-145) over LAYERS. Returns (y (B,84,A), [raw maps])."""
    sd = {k: v.float() for k, v in sd.items()}
    x = x.float()
    ys = []
    for i, (f, t) in enumerate(LAYERS):
        p = f"model.{i}"
        if f != -1:
            x = [ys[j] if j != -1 else x for j in f] if isinstance(f, list) else ys[f]
        if t == "Conv":
            x = conv(sd, p, x, s=2)
        elif t == "C3k2":
            x = c3k2(sd, p, x)
        elif t == "SPPF":
            x = sppf(sd, p, x)
        elif t == "C2PSA":
            x = c2psa(sd, p, x)
        elif t == "Upsample":
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        elif t == "Concat":
            x = torch.cat(x, 1)
        elif t == "Detect":
            raw = detect_raw(sd, p, x)
            y = decode(raw)
            return (y, raw, ys) if return_features else (y, raw)
        ys.append(x)
    raise AssertionError("unreachable")
