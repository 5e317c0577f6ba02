"""Stock-torch encoder shared by the synthetic loops: a torchvision ResNet trunk (random init) + 1x1 neck convs to 128
channels at strides 4/8/16/32 (reference dmm/modules/base.py:35-54, dmm/modules/vision.py:6-55).  The north-star leaves
the backbone on stock torch convs; only its output shapes matter to the matching path."""
import torch.nn as nn


class Encoder(nn.Module):
    def __init__(self, arch="resnet50"):
        super().__init__()
        import torchvision
        net = getattr(torchvision.models, arch)(weights=None)
        self.stem = nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool)
        self.layers = nn.ModuleList([net.layer1, net.layer2, net.layer3, net.layer4])
        chans = [64, 128, 256, 512] if arch in ("resnet18", "resnet34") else [256, 512, 1024, 2048]
        self.neck = nn.ModuleList([nn.Conv2d(c, 128, 1) for c in chans])

    def forward(self, x):
        x = self.stem(x)
        outs = []
        for layer, neck in zip(self.layers, self.neck):
            x = layer(x)
            outs.append(neck(x))
        return tuple(outs)                                           # strides 4, 8, 16, 32
