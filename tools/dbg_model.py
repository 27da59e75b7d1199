"""Debug helper (GPU box): print error statistics of modules/models vs the goldens and the oracle."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT), str(ROOT / "tests")]
from conftest import model_state_dict  # noqa: E402
from oracle import yolo11_ref  # noqa: E402
from oracle.weights import fill_state_dict_  # noqa: E402

from yololite.nn.modules import (C2PSA, C3k, C3k2, DFL, SPPF, Attention, Bottleneck, Conv, Detect, DWConv,  # noqa: E402
                                 PSABlock)
from yololite.nn.tasks import DetectionModel  # noqa: E402

G = ROOT / "tests" / "golden"


def stats(tag, got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    d = np.abs(got - ref)
    print(f"{tag:28s} max|d|={d.max():.4e} mean|d|={d.mean():.3e} ref_absmax={np.abs(ref).max():.3f} "
          f"rel={d.max() / (np.abs(ref).max() + 1e-12):.3e}")


def modules():
    g = np.load(G / "modules.npz")
    specs = {
        "conv_k3s2": lambda: Conv(16, 32, 3, 2), "conv_k1": lambda: Conv(48, 64, 1, 1),
        "conv_k3s1_noact": lambda: Conv(32, 16, 3, 1, act=False), "conv_stem": lambda: Conv(3, 16, 3, 2),
        "dwconv": lambda: DWConv(64, 64, 3), "bottleneck": lambda: Bottleneck(32, 32, True),
        "c3k": lambda: C3k(64, 64, 2), "c3k2_plain": lambda: C3k2(64, 128, 1, False, 0.25),
        "c3k2_c3k": lambda: C3k2(128, 128, 1, True), "sppf": lambda: SPPF(128, 128, 5),
        "attention": lambda: Attention(128, num_heads=2, attn_ratio=0.5), "psablock": lambda: PSABlock(128, 0.5, 2),
        "c2psa": lambda: C2PSA(256, 256, 1),
    }
    for tag, mk in specs.items():
        m = fill_state_dict_(mk()).cuda().eval()
        y = m(torch.from_numpy(g[f"{tag}.x"]).cuda())
        stats(tag, y.cpu().numpy(), g[f"{tag}.y"])
    d = DFL(16).cuda()
    stats("dfl", d(torch.from_numpy(g["dfl.x"]).cuda()).cpu().numpy(), g["dfl.y"])
    det = Detect(80, (64, 128, 256))
    det.stride = torch.tensor([8.0, 16.0, 32.0])
    fill_state_dict_(det).cuda().eval()
    y, raw = det([torch.from_numpy(g[f"detect.x{i}"]).cuda() for i in range(3)])
    for i in range(3):
        stats(f"detect.raw{i}", raw[i].cpu().numpy(), g[f"detect.raw{i}"])
    stats("detect.y.box", y[:, :4].cpu().numpy(), g["detect.y"][:, :4])
    stats("detect.y.cls", y[:, 4:].cpu().numpy(), g["detect.y"][:, 4:])


def models():
    for scale in "nsm":
        g = np.load(G / f"model_yolo11{scale}.npz")
        m = DetectionModel(f"yolo11{scale}.yaml", verbose=False)
        m.load_state_dict(model_state_dict(g))
        m = m.cuda().eval()
        y, raw = m(torch.from_numpy(g["x"]).cuda())
        for i in range(3):
            stats(f"{scale}.raw{i}", raw[i].cpu().numpy(), g[f"raw{i}"])
        stats(f"{scale}.y.box", y[:, :4].cpu().numpy(), g["y"][:, :4])
        stats(f"{scale}.y.cls", y[:, 4:].cpu().numpy(), g["y"][:, 4:])
    # 640x640 vs the oracle
    g = np.load(G / "model_yolo11n.npz")
    sd = model_state_dict(g)
    m = DetectionModel("yolo11n.yaml", verbose=False)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = torch.rand(2, 3, 640, 640, generator=torch.Generator().manual_seed(0))
    t0 = time.time()
    yr, rawr = yolo11_ref.forward(sd, x)
    print("oracle 640 forward", time.time() - t0, "s")
    y, raw = m(x.cuda())
    y = y.cpu().numpy()
    yr = yr.numpy()
    for name, lo, hi in (("P3", 0, 6400), ("P4", 6400, 8000), ("P5", 8000, 8400)):
        stats(f"640.box.{name}", y[:, :4, lo:hi], yr[:, :4, lo:hi])
    stats("640.cls", y[:, 4:], yr[:, 4:])
    # timing eager vs graph
    xc = x.cuda()
    for _ in range(3):
        m.infer(xc)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(20):
        m.infer(xc)
    torch.cuda.synchronize()
    print("graph replay bs2 ms/iter", (time.time() - t0) / 20 * 1e3)


if __name__ == "__main__":
    torch.manual_seed(0)
    what = sys.argv[1:] or ["modules", "models"]
    with torch.no_grad():
        if "modules" in what:
            modules()
        if "models" in what:
            models()
