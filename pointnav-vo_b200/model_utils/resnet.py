"""Parameter containers with the reference's module tree (model_utils/visual_encoders/resnet.py:29-286),
so that state_dict keys, shapes and default initialisation match checkpoint-for-checkpoint:

    conv1.{0,1}; layer{1..4}.{b}.convs.{0,1,3,4[,6,7]}; layer{2..4}.0.downsample.{0,1}

These modules hold weights only.  The arithmetic is executed by libpnvo through pointnav_vo_b200.engine
(conv -> GroupNorm -> ReLU chains fused there); calling .forward on a container raises.
"""
import torch.nn as nn

LAYERS = {"resnet18": ("basic", [2, 2, 2, 2]), "resnet50": ("bottleneck", [3, 4, 6, 3]),
          "resnet101": ("bottleneck", [3, 4, 23, 3])}


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("parameter container: the forward pass runs in libpnvo (pointnav_vo_b200.engine)")


def _conv(cin, cout, k, stride, pad):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=pad, bias=False)


class Block(_Holder):
    def __init__(self, kind, inplanes, planes, ngroups, stride, with_downsample):
        super().__init__()
        if kind == "basic":
            self.expansion = 1
            mods = [_conv(inplanes, planes, 3, stride, 1), nn.GroupNorm(ngroups, planes), nn.ReLU(True),
                    _conv(planes, planes, 3, 1, 1), nn.GroupNorm(ngroups, planes)]
        else:
            self.expansion = 4
            mods = [_conv(inplanes, planes, 1, 1, 0), nn.GroupNorm(ngroups, planes), nn.ReLU(True),
                    _conv(planes, planes, 3, stride, 1), nn.GroupNorm(ngroups, planes), nn.ReLU(True),
                    _conv(planes, planes * 4, 1, 1, 0), nn.GroupNorm(ngroups, planes * 4)]
        self.convs = nn.Sequential(*mods)
        self.downsample = None
        if with_downsample:
            out = planes * self.expansion
            self.downsample = nn.Sequential(_conv(inplanes, out, 1, stride, 0), nn.GroupNorm(ngroups, out))
        self.relu = nn.ReLU(True)


class ResNet(_Holder):
    def __init__(self, in_channels, base_planes, ngroups, backbone):
        super().__init__()
        kind, layers = LAYERS[backbone]
        exp = 1 if kind == "basic" else 4
        self.conv1 = nn.Sequential(_conv(in_channels, base_planes, 7, 2, 3), nn.GroupNorm(ngroups, base_planes),
                                   nn.ReLU(True))
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        inplanes = base_planes
        for li, nb in enumerate(layers, start=1):
            planes = base_planes * 2 ** (li - 1)
            stride = 1 if li == 1 else 2
            blocks = []
            for b in range(nb):
                s = stride if b == 0 else 1
                down = b == 0 and (s != 1 or inplanes != planes * exp)
                blocks.append(Block(kind, inplanes, planes, ngroups, s, down))
                inplanes = planes * exp
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        self.final_channels = inplanes
        self.final_spatial_compress = 1.0 / 32


def make_backbone(name):
    if name not in LAYERS:
        raise NotImplementedError(f"backbone {name!r}: the B200 path implements {sorted(LAYERS)} "
                                  "(the SE / ResNeXt factories of the reference are not used by any VO model or config)")
    return lambda in_channels, base_planes, ngroups: ResNet(in_channels, base_planes, ngroups, name)
