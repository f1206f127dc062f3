"""
Launcher with the reference's command line (script.py:20-136 of lucaslingle/pytorch_ddp_resnet):

    python script.py --mode train --models_dir models_dir --run_name wrn-28-10-dropout_cifar10 --data_dir data_dir

Reads models_dir/<run_name>/config.yaml (same keys as the reference), spawns `world_size` processes, builds
ResNet -> DistributedDataParallel -> FusedSGD -> scheduler -> checkpoint strategy, resumes from the
newest aligned checkpoint and runs training_loop / evaluation_loop on the sm_100a kernels.
Extra optional config keys: cuda_graph (bool, capture the whole step), synthetic_train_size/_test_size.
"""
import argparse
import os

import torch as tc

from pytorch_ddp_resnet_b200.algos.evaluation import evaluation_loop
from pytorch_ddp_resnet_b200.algos.training import training_loop
from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
from pytorch_ddp_resnet_b200.utils.checkpoint_util import get_checkpoint_strategy, maybe_load_checkpoints
from pytorch_ddp_resnet_b200.utils.config_util import ConfigParser
from pytorch_ddp_resnet_b200.utils.data_util import get_datasets, get_dataloaders, get_samplers
from pytorch_ddp_resnet_b200.utils.ddp_util import prepare_env_for_graphs, wrap_ddp
from pytorch_ddp_resnet_b200.utils.optim_util import get_optimizer, get_scheduler


def create_argparser():
    parser = argparse.ArgumentParser(
        description="Deep Residual Networks with Distributed Data Parallel on B200-native kernels.")
    parser.add_argument("--mode", choices=['train', 'eval'], default='train')
    parser.add_argument("--models_dir", type=str, default='models_dir')
    parser.add_argument("--run_name", type=str, default='wrn-28-10-dropout_cifar10')
    parser.add_argument("--data_dir", type=str, default='data_dir')
    return parser


def get_config(args) -> ConfigParser:
    base = os.path.join(args.models_dir, args.run_name)
    config = ConfigParser(defaults={
        'mode': args.mode,
        'data_dir': args.data_dir,
        'checkpoint_dir': os.path.join(base, 'checkpoints'),
        'log_dir': os.path.join(base, 'tensorboard_logs'),
    })
    config.read(os.path.join(base, 'config.yaml'), verbose=True)
    return config


def setup(rank: int, config: ConfigParser) -> dict:
    os.environ['MASTER_ADDR'] = str(config.get('master_addr'))
    os.environ['MASTER_PORT'] = str(config.get('master_port'))
    if not tc.cuda.is_available():
        raise RuntimeError("pytorch_ddp_resnet_b200 needs a CUDA sm_100 device; there is no CPU path")
    tc.cuda.set_device(rank)
    device = tc.device("cuda", rank)
    prepare_env_for_graphs()
    tc.distributed.init_process_group(backend=config.get('backend'), world_size=config.get('world_size'),
                                      rank=rank, device_id=device)
    datasets = get_datasets(**config)
    samplers = get_samplers(rank, **config, **datasets)
    dataloaders = get_dataloaders(**config, **datasets, **samplers)

    model = ResNet(architecture_spec=config.get('architecture_spec'), preact=config.get('preact'),
                   use_proj=config.get('use_proj'), dropout_prob=config.get('dropout_prob')).to(device)
    classifier = wrap_ddp(model, device)
    optimizer = get_optimizer(model=classifier, optimizer_cls_name=config.get('optimizer_cls_name'),
                              optimizer_args=config.get('optimizer_args'))
    scheduler = get_scheduler(optimizer=optimizer, scheduler_cls_name=config.get('scheduler_cls_name'),
                              scheduler_args=config.get('scheduler_args'))
    checkpoint_strategy = get_checkpoint_strategy(
        checkpoint_strategy_cls_name=config.get('checkpoint_strategy_cls_name'),
        checkpoint_strategy_args=config.get('checkpoint_strategy_args'))
    scaler = None  # bf16 kernels need no loss scaling (the reference's fp16 autocast does)
    global_step = maybe_load_checkpoints(
        checkpoint_dir=config.get('checkpoint_dir'),
        checkpointables={'checkpoint_strategy': checkpoint_strategy, 'classifier': classifier,
                         'optimizer': optimizer, 'scheduler': scheduler, 'scaler': scaler},
        map_location=device, steps=None)
    return {'device': device, **samplers, **dataloaders, 'classifier': classifier, 'optimizer': optimizer,
            'scheduler': scheduler, 'scaler': scaler, 'checkpoint_strategy': checkpoint_strategy,
            'global_step': global_step}


def train(rank, config):
    system = setup(rank, config)
    training_loop(rank, **config, **system)
    tc.distributed.destroy_process_group()


def evaluate(rank, config):
    system = setup(rank, config)
    metrics = evaluation_loop(**config, **system)
    if rank == 0:
        print(f"Test metrics: {metrics}")
    tc.distributed.destroy_process_group()


if __name__ == '__main__':
    cfg = get_config(create_argparser().parse_args())
    tc.multiprocessing.spawn(train if cfg.get('mode') == 'train' else evaluate, args=(cfg,),
                             nprocs=cfg.get('world_size'), join=True)
