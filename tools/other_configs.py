"""Step time of the other BASELINE configs (parity-test cases, not bench lines): eager and graph."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics
from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
from pytorch_ddp_resnet_b200.utils.graph_util import GraphedTrainStep
from pytorch_ddp_resnet_b200.utils.optim_util import get_optimizer

CONFIGS = {
    "resnet-v1-20": ("c3,16,3,1,1 n a r3 r3 r3 ap8,1,0 fc64,10", False, False, 0.0, 128, 32, 10),
    "resnet-v2-164": ("c3,64,3,1,1 b18 b18 b18 n a ap8,1,0 fc256,10", True, True, 0.0, 128, 32, 10),
    "wrn-50-2-like-imagenet": ("c3,512,7,2,3 n a mp3,2,1 b3 b4 b6 b3 ap7,1,0 fc4096,1000", False, True, 0.0, 32, 224, 1000),
}
which = sys.argv[1:] or list(CONFIGS)
for name in which:
    spec, preact, proj, p, B, hw, ncls = CONFIGS[name]
    torch.manual_seed(0)
    m = ResNet(spec, preact, proj, p).cuda().train()
    opt = get_optimizer("SGD", m, dict(lr=0.05, momentum=0.9, weight_decay=1e-4))
    x = torch.randn(B, 3, hw, hw, device="cuda"); y = torch.randint(0, ncls, (B,), device="cuda")
    step = GraphedTrainStep(m, opt, x, y)
    for _ in range(3): step(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n): out = step(x, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"config": name, "batch": B, "ms_per_step": ms, "img_per_s": B / ms * 1e3,
                      "launches_per_step": step.launches_per_step, "loss": out["loss"].item(),
                      "params": sum(q.numel() for q in m.parameters())}), flush=True)
    del m, opt, step
    torch.cuda.empty_cache()
