"""ORACLE (test infrastructure): networks whose predictions VARY, for parity evidence that means something.

The default-initialised GMFNet predicts one class for every pixel (its logits are dominated by the fc2 bias), so an
"argmax agreement" measured on it is trivially 100 % and the confusion matrix has a single non-zero row.  The fitted nets
keep the seed-3407 default-initialised convolutions of oracle/gmfnet_ref.py (torch.manual_seed(3407), test.py:8 of the
reference) and replace
  * the BatchNorm affine parameters (seeded random gamma / beta) and running statistics (calibrated on a batch of patches of
    the workload's own structured synthetic scene, so that every layer sees O(1) activations as in a trained network), and
  * the head (fc1, fc2), fitted in fp32 on the pooled features of sampled labelled pixels of that scene
with the values stored in tests/golden/fitted_nets.npz (produced by tests/golden/make_fitted_nets.py in the authoring
container).  The result classifies the structured scene with >= 5 predicted classes and Kappa > 0.1 — the network the
north-star's ">= 99.9 % argmax agreement" is asserted on (tests/test_gpu_parity_fitted.py, bench.py, smoke()).
"""
import os

import numpy as np
import torch

from .gmfnet_ref import Net

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'fitted_nets.npz')

# workload tag -> (H, W, classes without background, patch size)
WORKLOADS = {'c1': (128, 128, 7, 16), 'c2': (1000, 1000, 12, 16), 'c3': (2001, 2101, 11, 16), 'smoke': (40, 36, 7, 16),
             'p8': (96, 100, 7, 8), 'p32': (96, 104, 7, 32)}          # the other patch sizes of BASELINE.json configs[4]
# (scene seed, label seed, region size in MS pixels): land-cover regions a few patches wide, so that most patches are not mixtures
SCENES = {'c1': (0, 1, 32), 'c2': (0, 1, 64), 'c3': (0, 1, 64), 'smoke': (2, 3, 20), 'p8': (4, 5, 24), 'p32': (6, 7, 32)}


def scene(tag):
    """(ms uint16 [H,W,4], pan uint16 [4H,4W], label uint8 [H,W]) of a workload: the structured synthetic scene."""
    from . import dmf_oracle as orc
    H, W, ncls, _ = WORKLOADS[tag]
    seed, label_seed, cell = SCENES[tag]
    return orc.synthetic_scene_structured(H, W, ncls, seed=seed, label_seed=label_seed, cell=cell)


def cfg_for(tag):
    H, W, ncls, p = WORKLOADS[tag]
    return {'Categories_Number': ncls + 1, 'patch_size': p, 'schedule': {'activate': 'Relu'}}


def base_net(tag, seed=3407):
    """The seed-determined part: default PyTorch initialisation under torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    return Net(cfg_for(tag)).eval()


def fitted_state(tag):
    """state_dict of the fitted net for a workload: seed-determined convolutions + the stored BN / head tensors."""
    net = base_net(tag)
    sd = net.state_dict()
    with np.load(GOLDEN, allow_pickle=False) as z:
        prefix = tag + '/'
        for k in z.files:
            if k.startswith(prefix):
                name = k[len(prefix):]
                assert name in sd and tuple(sd[name].shape) == z[k].shape, 'fitted_nets.npz: bad entry %s' % k
                sd[name] = torch.from_numpy(z[k].copy())
    return sd


def fitted_net(tag):
    net = base_net(tag)
    net.load_state_dict(fitted_state(tag))
    return net.eval()
