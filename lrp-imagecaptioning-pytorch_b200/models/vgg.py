"""VGG encoder definition (state_dict-compatible with the reference's models/vgg.py:23-83).

Only the layer list / shapes matter to the LRP path: the encoder the captioning
models use is ``vgg16().features[0:-1]`` (13 conv3x3 + ReLU, 4 max-pools -> 512 x H/16 x W/16).
No pretrained download exists here (offline); ``pretrained`` is accepted and ignored with random init.
"""
import torch.nn as nn

cfgs = {
    'A': [64, 'M', 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'B': [64, 64, 'M', 128, 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'D': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M'],
    'E': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512, 'M'],
}


def make_layers(cfg, batch_norm=False, in_channels=3):
    layers = []
    for v in cfg:
        if v == 'M':
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            continue
        layers.append(nn.Conv2d(in_channels, v, kernel_size=3, padding=1))
        if batch_norm:
            layers.append(nn.BatchNorm2d(v))
        layers.append(nn.ReLU(inplace=True))
        in_channels = v
    return nn.Sequential(*layers)


class VGG(nn.Module):
    def __init__(self, features, num_classes=1000, init_weights=True, with_classifier=False):
        super().__init__()
        self.features = features
        self.feat_dim = 512
        self.avgpool = nn.AdaptiveAvgPool2d((7, 7))
        # the classifier head is never used by the captioning encoders; built only on request
        self.classifier = nn.Sequential(nn.Linear(512 * 7 * 7, 4096), nn.ReLU(True), nn.Dropout(),
                                        nn.Linear(4096, 4096), nn.ReLU(True), nn.Dropout(),
                                        nn.Linear(4096, num_classes)) if with_classifier else None
        if init_weights:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                    if m.bias is not None:
                        nn.init.constant_(m.bias, 0)
                elif isinstance(m, nn.Linear):
                    nn.init.normal_(m.weight, 0, 0.01)
                    nn.init.constant_(m.bias, 0)

    def forward(self, x):
        x = self.features(x)
        if self.classifier is None:
            return x
        x = self.avgpool(x)
        return self.classifier(x.flatten(1))


def _vgg(cfg, batch_norm, pretrained=False, **kwargs):
    return VGG(make_layers(cfgs[cfg], batch_norm=batch_norm), **kwargs)


def vgg11(pretrained=False, **kw): return _vgg('A', False, pretrained, **kw)
def vgg13(pretrained=False, **kw): return _vgg('B', False, pretrained, **kw)
def vgg16(pretrained=False, **kw): return _vgg('D', False, pretrained, **kw)
def vgg19(pretrained=False, **kw): return _vgg('E', False, pretrained, **kw)
def vgg16_bn(pretrained=False, **kw): return _vgg('D', True, pretrained, **kw)
