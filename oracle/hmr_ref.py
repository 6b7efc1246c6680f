"""CPU restatement of the HMR ResNet-50 feature extractor -- TEST INFRASTRUCTURE (only tests/, smoke() and bench.py's CPU legs may
import it; the product never does).  Plain torch.nn.functional ops in fp32 on a state_dict, eval-mode BatchNorm, following
lib/models/spin.py:16-56 (Bottleneck.forward), :59-125 (layer structure: [3, 4, 6, 3] bottlenecks, stride on conv2 and on the
1x1 downsample of each stage's first block) and :127-141 (feature_extractor).  Pinned against tests/golden/hmr_N2.npz, i.e. the
outputs of the unmodified reference class (tests/test_hmr.py::test_hmr_oracle_restatement_matches_the_reference_golden)."""
import torch
import torch.nn.functional as F

LAYERS = (3, 4, 6, 3)          # lib/models/spin.py:300 -- HMR(Bottleneck, [3, 4, 6, 3], ...)


def _bn(sd, prefix, x):
    # nn.BatchNorm2d in eval mode (the reference calls .eval() before extracting features: demo.py:121)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"], sd[prefix + ".bias"],
                        training=False, eps=1e-5)


def _bottleneck(sd, p, x, stride, has_down):
    # lib/models/spin.py:35-56
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    out = F.relu(_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1)))
    out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    residual = x
    if has_down:                                  # lib/models/spin.py:108-115
        residual = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(out + residual)


def feature_extractor(sd, x):
    """sd: HMR state_dict (torch tensors, fp32), x [N,3,224,224] -> [N,2048] (lib/models/spin.py:127-141)."""
    x = F.relu(_bn(sd, "bn1", F.conv2d(x, sd["conv1.weight"], stride=2, padding=3)))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, blocks in enumerate(LAYERS, start=1):
        for b in range(blocks):
            stride = 2 if (b == 0 and li > 1) else 1
            x = _bottleneck(sd, f"layer{li}.{b}", x, stride, has_down=(b == 0))
    x = F.avg_pool2d(x, 7, stride=1)
    return x.view(x.size(0), -1)
