"""2-GPU diagnostic: per-parameter difference graph-DDP vs eager-DDP vs eager-DDP (noise floor) after k steps."""
import os, sys
import torch, torch.distributed as dist, torch.multiprocessing as mp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SPEC = "c3,32,3,1,1 r1 r1 n a ap16,1,0 fc64,10"
SGD = dict(lr=0.1, momentum=0.9, dampening=0.0, nesterov=True, weight_decay=5e-4)

def worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    from pytorch_ddp_resnet_b200.utils.ddp_util import prepare_env_for_graphs, wrap_ddp
    prepare_env_for_graphs()
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    from pytorch_ddp_resnet_b200.utils.graph_util import GraphedTrainStep
    from pytorch_ddp_resnet_b200.utils.optim_util import get_optimizer
    torch.manual_seed(0)
    base = ResNet(SPEC, True, True, 0.0).cuda()
    init = {k: v.clone() for k, v in base.state_dict().items()}
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(8, 3, 32, 32, generator=g).cuda(); y = torch.randint(0, 10, (8,), generator=g).cuda()
    def eager(steps):
        m = ResNet(SPEC, True, True, 0.0).cuda().train(); m.load_state_dict(init)
        d = wrap_ddp(m, torch.device("cuda", rank)); o = get_optimizer("SGD", d, dict(SGD))
        for _ in range(steps):
            compute_losses_and_metrics(logits=d(x), labels=y)["loss"].backward(); o.step(); o.zero_grad(set_to_none=True)
        return m
    def graph(steps):
        m = ResNet(SPEC, True, True, 0.0).cuda().train(); m.load_state_dict(init)
        d = wrap_ddp(m, torch.device("cuda", rank)); o = get_optimizer("SGD", d, dict(SGD))
        st = GraphedTrainStep(d, o, x, y, bucket_bytes=64 << 10)
        for _ in range(steps):
            st(x, y)
        torch.cuda.synchronize(); st.close()
        return m
    for steps in (1, 3):
        e1, e2, gr = eager(steps), eager(steps), graph(steps)
        worst = []
        for (n, a), (_, b), (_, c) in zip(e1.named_parameters(), e2.named_parameters(), gr.named_parameters()):
            den = a.float().norm().clamp_min(1e-12)
            worst.append((((c.float() - a.float()).norm() / den).item(), ((b.float() - a.float()).norm() / den).item(), n, den.item()))
        worst.sort(reverse=True)
        if True:
            print(f"rank {rank} steps {steps}: worst graph-vs-eager / eager-vs-eager / name / |p|:", [(f"{w[0]:.1e}", f"{w[1]:.1e}", w[2], f"{w[3]:.1e}") for w in worst[:4]], flush=True)
    dist.barrier(); os._exit(0)

if __name__ == "__main__":
    mp.spawn(worker, args=(2, 29517), nprocs=2, join=True)
