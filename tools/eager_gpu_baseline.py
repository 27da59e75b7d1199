"""The "honest GPU bar" of SURVEY §8d: the reference's architecture run with stock PyTorch kernels on the same B200 —
eager ATen/cuDNN forward (fp32 NCHW, and bf16 channels_last autocast) + torchvision's CUDA nms per image, i.e. what a
user gets by just moving the reference to the GPU.  A measurement tool (tools/ is not product code): it drives the oracle
restatement of the forward (oracle/yolo11_ref.py) because /root/reference does not exist on the GPU box.

    python tools/eager_gpu_baseline.py [batch] > gpurun_out/<tag>/eager_gpu.json
"""
import json
import sys
import time
from pathlib import Path

import torch
import torchvision

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "yolo-lite_b200"), str(ROOT)]
from bench import randomise_model_  # noqa: E402
from oracle import yolo11_ref  # noqa: E402
from yololite.nn.tasks import DetectionModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.backends.cudnn.benchmark = True
m = randomise_model_(DetectionModel("yolo11n.yaml", verbose=False)).eval()
sd = {k: v.detach().cuda() for k, v in m.state_dict().items()}
x = torch.rand(B, 3, 640, 640, device="cuda")


def nms_torchvision(y, conf=0.25, iou=0.7, max_det=300):
    """Per-image loop of utils/ops.py:199-273 on the GPU (best class, conf filter, class-offset nms)."""
    out = []
    for p in y.transpose(1, 2):
        box, cls = p[:, :4], p[:, 4:]
        c, j = cls.max(1)
        k = c > conf
        box, c, j = box[k], c[k], j[k]
        xyxy = torch.cat([box[:, :2] - box[:, 2:] / 2, box[:, :2] + box[:, 2:] / 2], 1)
        keep = torchvision.ops.nms(xyxy + j[:, None].float() * 7680, c, iou)[:max_det]
        out.append(torch.cat([xyxy[keep], c[keep, None], j[keep, None].float()], 1))
    return out


def run(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return B * n / (time.perf_counter() - t0)


res = {}
with torch.inference_mode():
    res["eager_fp32_forward_only"] = run(lambda: yolo11_ref.forward(sd, x))
    res["eager_fp32_forward_nms"] = run(lambda: nms_torchvision(yolo11_ref.forward(sd, x)[0]))
    xb = x.to(memory_format=torch.channels_last)
    sdb = {k: (v.to(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}

    def bf16():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return yolo11_ref.forward(sdb, xb)

    res["eager_bf16_channels_last_forward_only"] = run(bf16)
    res["eager_bf16_channels_last_forward_nms"] = run(lambda: nms_torchvision(bf16()[0].float()))
print(json.dumps({"workload": f"yolo11n 640x640 bs={B}, unfused conv/BN/SiLU as in the reference, torch "
                              f"{torch.__version__}, torchvision {torchvision.__version__}",
                  "images_per_s": {k: round(v, 1) for k, v in res.items()}}))
