"""ORACLE (test infrastructure, not product code): the fp32 PyTorch definition of GMFNet.

The reference never committed its network: solver/mainsolver.py:30-34 does
``importlib.import_module('model.' + cfg['model_name'].lower()).Net(args=cfg)`` with
``model_name: gmfnet`` (config.yml:6) but there is no ``model/`` directory.  The reference
therefore pins only the I/O contract (solver/mainsolver.py:52,109,169):

    Net(args=cfg).forward(ms[B,4,p,p] f32, pan[B,1,4p,4p] f32) -> logits[B,C] f32

PARITY UNPINNED for the network arithmetic: this file *is* the definition that the CUDA
implementation (dual-modal-fusion_b200/csrc) is checked against.  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import it.

Architecture (all convs 3x3 / pad 1 / bias, BatchNorm2d after each conv, ReLU):

    MS  branch : conv 4->64 . conv 64->128 . maxpool2                         -> [128, p/2, p/2]
    PAN branch : conv 1->32 . maxpool2 . conv 32->64 . maxpool2 .
                 conv 64->128 . maxpool2                                      -> [128, p/2, p/2]
    fusion     : concat(256) . conv1x1 256->128 . BN . ReLU . global-avg-pool -> [128]
    head       : Linear 128->64 . ReLU . Linear 64->C                          -> logits
"""
import torch
import torch.nn as nn

MS_BANDS = 4
C_MS1, C_MS2 = 64, 128
C_PAN1, C_PAN2, C_PAN3 = 32, 64, 128
C_FUSE, C_HID = 128, 64


def _block(cin, cout, k=3):
    return nn.Sequential(nn.Conv2d(cin, cout, k, padding=k // 2, bias=True), nn.BatchNorm2d(cout))


class Net(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.num_classes = int(args['Categories_Number'])
        self.patch = int(args['patch_size'])
        act = str(args.get('schedule', {}).get('activate', 'Relu')).lower()
        if act != 'relu':
            raise ValueError("gmfnet oracle: only schedule.activate == Relu is defined")
        self.ms1 = _block(MS_BANDS, C_MS1)
        self.ms2 = _block(C_MS1, C_MS2)
        self.pan1 = _block(1, C_PAN1)
        self.pan2 = _block(C_PAN1, C_PAN2)
        self.pan3 = _block(C_PAN2, C_PAN3)
        self.fuse = _block(C_MS2 + C_PAN3, C_FUSE, k=1)
        self.fc1 = nn.Linear(C_FUSE, C_HID)
        self.fc2 = nn.Linear(C_HID, self.num_classes)

    def features(self, ms, pan):
        r, mp = torch.relu, nn.functional.max_pool2d
        m = r(self.ms1(ms))
        m = mp(r(self.ms2(m)), 2)
        q = mp(r(self.pan1(pan)), 2)
        q = mp(r(self.pan2(q)), 2)
        q = mp(r(self.pan3(q)), 2)
        f = r(self.fuse(torch.cat([m, q], dim=1)))
        return f.mean(dim=(2, 3))

    def forward(self, ms, pan):
        g = self.features(ms, pan)
        return self.fc2(torch.relu(self.fc1(g)))


def flops_per_patch(p, num_classes):
    """Algorithmic FLOPs (2*MAC) of one forward, no credit for padding."""
    def conv(cin, cout, k, h):
        return 2 * cin * cout * k * k * h * h
    f = conv(4, C_MS1, 3, p) + conv(C_MS1, C_MS2, 3, p)
    f += conv(1, C_PAN1, 3, 4 * p) + conv(C_PAN1, C_PAN2, 3, 2 * p) + conv(C_PAN2, C_PAN3, 3, p)
    f += conv(C_MS2 + C_PAN3, C_FUSE, 1, p // 2)
    f += 2 * C_FUSE * C_HID + 2 * C_HID * num_classes
    return f
