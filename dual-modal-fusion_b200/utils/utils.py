"""Optimizer / loss / scheduler / checkpoint factories, same names and file formats as the
reference's utils/utils.py:8-122.  ADAM is dmf.FusedAdam (torch.optim.Adam's update in one sm_100a kernel over the
flattened model, same defaults and state names); the other factories are the plain torch objects."""
import os
import random

import numpy as np
import torch
import torch.nn as nn
import torch.optim.lr_scheduler as lr_scheduler


def make_optimizer(cfg, params):
    sch = cfg['schedule']
    kind = sch['optimizer']
    if kind == "ADAM":
        import dmf
        return dmf.FusedAdam(params, lr=sch['lr'])
    if kind == "SGD":
        return torch.optim.SGD(params, lr=sch['lr'], momentum=sch['momentum'])
    if kind == "RMSprop":
        return torch.optim.RMSprop(params, lr=sch['lr'], alpha=sch['alpha'])
    raise ValueError(kind)


def make_loss(loss_type, cfg):
    table = {"MSE": lambda: nn.MSELoss(reduction='mean'), "L1": lambda: nn.L1Loss(reduction='mean'),
             "Criterion": nn.CrossEntropyLoss, "KL": lambda: nn.KLDivLoss(reduction='batchmean')}
    if loss_type not in table:
        raise ValueError(loss_type)     # qua_loss belongs to the out-of-scope two-stage (dqtl) solver
    return table[loss_type]()


def make_scheduler(optimizer, cfg):
    sch = cfg['schedule']
    if not sch['if_scheduler']:
        return None
    kind, lr, base = sch['scheduler'], sch['lr'], sch['base_lr']
    if kind == "StepLR":
        return lr_scheduler.StepLR(optimizer, step_size=50, gamma=base / lr)
    if kind == "LinearLR":
        return lr_scheduler.LinearLR(optimizer, start_factor=0.1, end_factor=1, total_iters=10)
    if kind == "CosineAnnealingLR":
        return lr_scheduler.CosineAnnealingLR(optimizer, 50, base)
    if kind == "CyclicLR":
        return lr_scheduler.CyclicLR(optimizer, base_lr=base, max_lr=lr, step_size_up=10, step_size_down=40, cycle_momentum=False)
    if kind == "OneCycleLR":
        return lr_scheduler.OneCycleLR(optimizer, max_lr=lr, pct_start=0.5, total_steps=cfg['epoch'],
                                       div_factor=lr / base, final_div_factor=lr / base)
    if kind == "ConstantLR":
        return lr_scheduler.ConstantLR(optimizer, factor=base / lr, total_iters=10)
    if kind == "ChainedScheduler":
        return lr_scheduler.ChainedScheduler([lr_scheduler.LinearLR(optimizer, start_factor=0.1, end_factor=1, total_iters=10),
                                              lr_scheduler.ExponentialLR(optimizer, gamma=0.98)])
    if kind == "ExponentialLR":
        return lr_scheduler.ExponentialLR(optimizer=optimizer, gamma=0.98)
    raise ValueError(kind)


def save_checkpoint(model, optimizer, filename="my_checkpoint.pth.tar"):
    torch.save({"state_dict": model.state_dict(), "optimizer": optimizer.state_dict()}, filename)


def load_checkpoint(checkpoint_file, model, optimizer, lr, device):
    ckpt = torch.load(checkpoint_file, map_location=device)
    model.load_state_dict(ckpt["state_dict"], strict=False)
    optimizer.load_state_dict(ckpt["optimizer"])
    for group in optimizer.param_groups:
        group["lr"] = lr


def load_model(checkpoint_file, model, device):
    ckpt = torch.load(checkpoint_file, map_location=device)
    model.load_state_dict(ckpt["state_dict"], strict=False)


def seed_everything(seed=42):
    os.environ["PYTHONHASHSEED"] = str(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
